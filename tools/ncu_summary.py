#!/usr/bin/env python3
"""Summarise an .ncu-rep (first kernel) for profiles/: time, pipe utilisation, stall reasons,
shared-memory wavefronts/conflicts, DRAM bytes.  Usage: ncu_summary.py file.ncu-rep
       ncu_summary.py file.ncu-rep --traffic <kernel key> <codewords in the captured launch>
           also records dram bytes per codeword of that capture in profiles/ncu_traffic.json,
           which bench.py reads for roofline.traffic (measured, not a literal)."""
import csv
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_st.sum"]


def main(path, traffic=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in WANT:
            if k in d:
                print("  %-62s %s %s" % (k, d[k], u[k]))
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls.append((float(d[h]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("  warp stalls per issued instruction (top):",
              ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:8]))
        if traffic:
            key, n_cw = traffic[0], int(traffic[1])
            def to_bytes(name):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[name]]
                return float(d[name]) * mult
            total = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            jp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
            try:
                with open(jp) as f:
                    doc = json.load(f)
            except OSError:
                doc = {"kernels": {}}
            doc["kernels"][key] = {"dram_bytes_per_codeword": total / n_cw, "dram_bytes": total, "codewords": n_cw,
                                   "source": "ncu --set full capture %s (dram__bytes_read.sum + dram__bytes_write.sum)"
                                             % os.path.basename(path)}
            with open(jp, "w") as f:
                json.dump(doc, f, indent=1)
            print("  -> %s: %.1f DRAM bytes per codeword" % (jp, total / n_cw))
            traffic = None


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3:5] if len(sys.argv) >= 5 and sys.argv[2] == "--traffic" else None)
