#!/usr/bin/env python3
"""bench.py -- decoded information Gbit/s of the LDPC decode hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2]): the reference's shipped 32x64 code, synthetic
AWGN-BPSK codewords at Eb/N0 = 2 dB (reference convention sigma^2 = 10^(-EbN0/10),
apps/ldpc_lapack.cpp:635-642), sum-product, FIXED 50 iterations with early stop off;
10 M codewords per GPU.  A "step" is one decode pass over the whole batch.  Codewords are
independent: each rank decodes its own shard, no data-path collective (weak scaling);
torch.distributed is only the barrier and the max-over-ranks of the device time.

  value     device-resident: symbols already in HBM, one kernel launch per step, timed with
            CUDA events on the launching stream.
  e2e       the call a block's work() makes: ldpc535_decode_batch on PINNED HOST buffers,
            H2D of the symbols and D2H of bytes / syndrome weights / iteration counts inside
            the timed region.
  roofline  the roofline that governs a 50-iteration sum-product: executed MUFU operations per
            second against the SM special-function pipe's ceiling MEASURED IN THE SAME RUN
            (ldpc535_probe_pipe_peak); the HBM figure (algorithmic bytes against the measured copy
            bandwidth, ~1 %) sits beside it under `hbm` (DESIGN.md "Rooflines").
  configs   (one GPU only) the other BASELINE configs and kernels, each with its own roofline:
            the shipped code at the reference block's settings (5 iterations, early stop) at
            2 / 4 / 6 dB, min-sum, hard decision, both encoders, and configs[3] -- the
            (3,6)-regular n = 8192 code with early termination over batches of 1 k ... 1 M.
  inputs    counter-based: Philox4x32-10, seed 535, counter = GLOBAL codeword index, so the
            shard of rank r (codewords r*n .. (r+1)*n) is the same at any GPU count.
  cpu_baseline  oracle/_ref -- the reference's own decoder sources compiled against dependency
            stand-ins (`kind: reference`) -- on all host cores, on a bounded sample of the same
            workload; beside it (`port_same_work`) the oracle port running exactly the GPU arm's
            work (the reference cannot switch its early exit off).  Falls back to the port alone
            (`kind: port`) where oracle/_ref was not built.

`--impl reference` times only that CPU arm (rank 0), same metric/config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "decoded info Gbit/s (sum-product, fixed iterations, early stop off)"
EBN0_DB = 2.0
MAX_ITERS = 50
K_INFO, N_SYM, M_CHK, E_EDGES = 32, 64, 32, 168
BYTES_PER_CW = 8 * N_SYM + K_INFO // 8 + 2          # complex64 in + packed out + 2
XU_PER_EDGE_ITER = 4                                # SURVEY 8d algorithmic figure (tanh, log, div)
MUFU_PER_EDGE_ITER = 3                              # what the kernels execute: 1 ex2 + 2 lg2 (spa_math.cuh)
SM_COUNT, XU_LANES = 148, 16
# (the XU-pipe ceiling is measured in every run by ldpc535_probe_pipe_peak: 1 ex2 : 2 lg2 mix)
# Min-sum (fp64, decode_warp_kernel<0,6,3>): warp instructions per codeword and iteration, counted in
# the SASS of the iteration loop (tools/sass_count.py): 164 in all, 30 of them on the fp64 pipe
# (round 1: 239 / 51; the check update now uses prefix/suffix minima and sign bits, spa_math.cuh).
MINSUM_INSTR_PER_CW_ITER = 164
MINSUM_FP64_PER_CW_ITER = 30


def config(n_cw, n_gpus):
    return {"workload": "configs[2]: shipped 32x64 code, synthetic AWGN-BPSK at Eb/N0=2 dB, "
                        "sum-product, fixed 50 iterations, early stop off",
            "code": "32x64 (lib/ldpc_decoder_cb_impl.cc:63-96)", "codewords_per_gpu": n_cw,
            "codewords_total": n_cw * n_gpus, "max_iters": MAX_ITERS, "early_stop": 0,
            "ebn0_db": EBN0_DB, "sharding": "contiguous codeword shards, no collective",
            "l2": "inputs (%.2f GB per pass) exceed the 126 MB L2" % (n_cw * 512 / 1e9)}


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm (oracle): bench.py's cpu_baseline leg and --impl reference
# --------------------------------------------------------------------------------------
def cpu_frames(n, seed):
    """Same workload, generated on the host: oracle-encoded random words + AWGN."""
    import numpy as np
    from oracle import oracle as O
    codes = O.load_ref_codes()
    Hp, Lm, Um, _ = O.reorder_h(codes["shipped"]["H"])
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, 4 * n).astype(np.uint8)
    sym, _ = O.encoder_work(Hp, Lm, Um, data, n * N_SYM)
    sigma = np.float32(np.sqrt(10.0 ** (-EBN0_DB / 10.0)))
    sym = sym.copy()
    sym.real += rng.standard_normal(sym.shape, dtype=np.float32) * sigma
    return Hp, sym


def cpu_pass(Hp, sym, cores):
    """The oracle port: exactly the GPU arm's work (50 iterations, no early exit)."""
    from oracle import oracle as O
    _, _, _, dt = O.decode_frames(sym, Hp, method=1, iterations=MAX_ITERS, early_stop=False,
                                  threads=cores, pin=True)
    return dt


def ref_available():
    from oracle import ref as R
    return R.available()


def ref_pass(sym, cores):
    """oracle/_ref: the reference's own decodeSumProductSoft (lib/ldpc_decoder_cb_impl.cc:478-557,
    compiled from the reference's sources) with iterations = 50, one decoder block instance per
    host thread.  Its exit on a zero syndrome (:535) cannot be switched off."""
    from oracle import ref as R
    _, _, dt = R.decode_frames(sym, method=1, iterations=MAX_ITERS, threads=cores, pin=True)
    return dt


REF_NOTE = ("the reference's own decodeSumProductSoft compiled from its sources (oracle/_ref: "
            "lib/ldpc_decoder_cb_impl.cc against the uBLAS/GNU Radio stand-ins of oracle/refshim, g++ -O3), "
            "iterations = 50; the reference cannot switch off its exit on a zero syndrome, so it runs "
            "fewer iterations than the GPU arm on frames that converge")
PORT_NOTE = ("dense fp64 restatement of decodeSumProductSoft (oracle/ldpc_oracle.c, gcc -O3), fixed 50 "
             "iterations without early exit = exactly the GPU arm's work")


def cpu_baseline(per_core):
    cores = len(os.sched_getaffinity(0))
    n = cores * per_core
    Hp, sym = cpu_frames(n, 535)
    cpu_pass(Hp, sym[:cores * 8 * N_SYM], cores)                  # touch code + data once
    dt_port = cpu_pass(Hp, sym, cores)
    port = {"value": n * K_INFO / dt_port / 1e9, "unit": "Gbit/s", "codewords_per_s": n / dt_port,
            "seconds": dt_port, "what": PORT_NOTE}
    if not ref_available():
        return {"value": port["value"], "unit": "Gbit/s", "cores": cores, "kind": "port",
                "sample": "%d codewords (%d per core, one pinned thread per core) of the same workload, %s, "
                          "%.1f s" % (n, per_core, PORT_NOTE, dt_port),
                "codewords_per_s": n / dt_port}
    ref_pass(sym[:cores * 8 * N_SYM], cores)
    dt = ref_pass(sym, cores)
    return {"value": n * K_INFO / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "reference",
            "sample": "%d codewords (%d per core, one pinned thread per core) of the same workload, %s, "
                      "%.1f s" % (n, per_core, REF_NOTE, dt),
            "codewords_per_s": n / dt, "port_same_work": port, "encoder": ref_encoder_baseline(cores, per_core)}


def ref_encoder_baseline(cores, per_core):
    """The reference's encoder block (lib/ldpc_encoder_bc_impl.cc:118-178: general_work ->
    makeParityCheck :275-294 -> solve/dgesv :180-223, compiled from its sources in oracle/_ref),
    one block instance per pinned host thread, on per_core frames each."""
    import numpy as np
    from oracle import ref as R
    rng = np.random.default_rng(535)
    data = rng.integers(0, 256, (cores, per_core * 4)).astype(np.uint8)
    encs = [R.RefEncoder() for _ in range(cores)]
    cpus = sorted(os.sched_getaffinity(0))

    def worker(t):
        try:
            os.sched_setaffinity(0, {cpus[t % len(cpus)]})
        except OSError:
            pass
        done = 0
        while done < per_core:                        # general_work-sized calls, like a scheduler would make
            n = min(4096, per_core - done)
            encs[t].work(data[t, done * 4:(done + n) * 4], n * N_SYM)
            done += n

    worker_threads = [threading.Thread(target=worker, args=(t,)) for t in range(cores)]
    t0 = time.perf_counter()
    for th in worker_threads:
        th.start()
    for th in worker_threads:
        th.join()
    dt = time.perf_counter() - t0
    for e in encs:
        e.close()
    n = cores * per_core
    return {"value": n * K_INFO / dt / 1e9, "unit": "Gbit/s info", "frames_per_s": n / dt, "cores": cores,
            "kind": "reference",
            "sample": "%d frames (%d per core) through the reference's ldpc_encoder_bc::general_work "
                      "(oracle/_ref), %.2f s" % (n, per_core, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    per_core = 400                                              # ~1 s of CPU work per step
    n = cores * per_core
    Hp, sym = cpu_frames(n, 535)
    use_ref = ref_available()
    one = (lambda s: ref_pass(s, cores)) if use_ref else (lambda s: cpu_pass(Hp, s, cores))
    for _ in range(args.warmup):
        one(sym[:cores * 40 * N_SYM])
    t = 0.0
    for _ in range(args.steps):
        t += one(sym)
    v = n * args.steps * K_INFO / t / 1e9
    sample = ("each step = %d codewords (%d per core) of the same workload on %d host cores (one pinned "
              "thread per core), %s" % (n, per_core, cores, REF_NOTE if use_ref else PORT_NOTE))
    base = {"value": v, "unit": "Gbit/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": sample}
    if use_ref:                                                 # the same-work figure beside it
        dtp = cpu_pass(Hp, sym, cores)
        base["port_same_work"] = {"value": n * K_INFO / dtp / 1e9, "unit": "Gbit/s", "what": PORT_NOTE}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Gbit/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config(args.codewords, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
SEED = 535                                   # Philox, seed 535, counter = global codeword index


def pinned_array(L, nbytes, dtype, shape):
    import numpy as np
    from ldpc_ece535a import _abi
    p = C.c_void_p()
    _abi.check(_abi.lib().ldpc535_host_alloc(nbytes, C.byref(p)), "host_alloc")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def bind_near_gpu(local):
    """Best effort: restrict this rank to the host cores local to its GPU (sysfs local_cpulist of
    the GPU's PCI function), so pinned buffers are first-touched on, and the packing team runs on,
    the GPU's own NUMA node.  -> description for the JSON line."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        have = os.sched_getaffinity(0)
        near = cpus & have
        if near and len(near) < len(have):
            os.sched_setaffinity(0, near)
            return "bound to the %d cores local to GPU %s (of %d)" % (len(near), bdf, len(have))
        return "all %d cores are local to GPU %s" % (len(have), bdf)
    except Exception as e:            # no sysfs entry in a VM, older torch: leave the affinity alone
        return "not bound (%s)" % type(e).__name__


def synth_shard(code, first_frame, n, ebn0_db, stream_ptr, d_data=None, d_sym=None):
    """Frames first_frame .. first_frame + n of the seeded stream, on the device: Philox data bytes ->
    the library's encoder -> AWGN on the real axis (sigma^2 = 10^(-EbN0/10))."""
    import numpy as np
    import torch
    if d_data is None:
        d_data = torch.empty((n, code.nbytes), dtype=torch.uint8, device="cuda")
    if d_sym is None:
        d_sym = torch.empty((n, code.N, 2), dtype=torch.float32, device="cuda")
    code.synth_bytes_dev(SEED, first_frame, n, d_data.data_ptr(), stream=stream_ptr)
    code.encode_dev(d_data.data_ptr(), n, d_sym.data_ptr(), stream=stream_ptr)
    if ebn0_db is not None:
        sigma = float(np.sqrt(10.0 ** (-ebn0_db / 10.0)))
        code.synth_awgn_dev(SEED, first_frame, n, sigma, d_sym.data_ptr(), stream=stream_ptr)
    return d_data, d_sym


def time_launches(stream, fn, steps, warmup):
    """ms per call of fn(), CUDA events on the launching stream, after warm-up."""
    import torch
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(steps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def traffic_from_profiles(kernel):
    """DRAM bytes per codeword of `kernel` from the committed ncu --set full summary
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the capture named in it)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        e = t["kernels"].get(kernel)
        return (e["dram_bytes_per_codeword"], e.get("source")) if e else (None, None)
    except (OSError, KeyError, ValueError):
        return None, None


def extra_configs(L, torch, np, c4, stream, sp, d_data, d_sym, n_cw, first_frame, mufu_peak, fp64_peak, hbm_peak, f_max):
    """The other BASELINE configs and kernels, device-resident, same timing method as `value`
    (CUDA events on the launching stream, warm-up first, inputs larger than L2 except where the
    batch size itself is the point).  One GPU only."""
    out = {}
    d_bytes = torch.empty((n_cw, 4), dtype=torch.uint8, device="cuda")
    d_synd = torch.empty(n_cw, dtype=torch.uint8, device="cuda")
    d_iters = torch.empty(n_cw, dtype=torch.uint8, device="cuda")

    # ---- configs[0]: the shipped code at the reference block's own settings (5 iterations, early stop) ----
    rows = []
    for ebn0 in (2.0, 4.0, 6.0):
        synth_shard(c4, first_frame, n_cw, ebn0, sp, d_data, d_sym)
        for method, name in ((L.METHOD_SUMPRODUCT, "sum-product"), (L.METHOD_LOGDOMAIN, "min-sum (GRC default)")):
            if method == L.METHOD_LOGDOMAIN and ebn0 != 2.0:
                continue

            def one():
                c4.decode_dev(d_sym.data_ptr(), n_cw * N_SYM, n_cw, d_bytes.data_ptr(), d_synd.data_ptr(),
                              d_iters.data_ptr(), method=method, max_iters=5, early_stop=True, stream=sp)
            ms = time_launches(stream, one, 3, 2)
            it_sum = float(d_iters.sum(dtype=torch.float64).item())
            row = {"method": name, "ebn0_db": ebn0, "max_iters": 5, "early_stop": 1, "codewords": n_cw,
                   "kernel": c4.kernel_for(method, True, n_cw), "ms": ms, "gbit_s": n_cw * K_INFO / ms / 1e6,
                   "mean_iters": it_sum / n_cw,
                   "frames_recovered": float((d_bytes == d_data).all(dim=1).float().mean().item())}
            eit = it_sum * E_EDGES / (ms * 1e-3)
            row["edge_iterations_per_s"] = eit
            if method == L.METHOD_SUMPRODUCT:
                # + the start-up exponentials (one ex2 per bit) each codeword pays once
                mufu = (MUFU_PER_EDGE_ITER * it_sum * E_EDGES + n_cw * N_SYM) / (ms * 1e-3)
                row["roofline"] = {"bound": "sm_xu", "achieved": mufu / 1e12, "peak": mufu_peak / 1e12,
                                   "unit": "T MUFU/s", "frac": mufu / mufu_peak}
            else:
                # min-sum runs in fp64; its loop is 164 warp instructions per codeword and iteration, so
                # the issue slots (1 warp instruction per clock and SM sub-partition) bound it before the
                # fp64 pipe does (30 of the 164, 2 clocks each at the measured fp64 rate)
                wi = MINSUM_INSTR_PER_CW_ITER * it_sum / (ms * 1e-3)
                issue_peak = SM_COUNT * 4 * f_max
                fp64_ops = MINSUM_FP64_PER_CW_ITER * 32 * it_sum / (ms * 1e-3)
                row["roofline"] = {"bound": "sm_issue", "achieved": wi / 1e12, "peak": issue_peak / 1e12,
                                   "unit": "T warp-instr/s", "frac": wi / issue_peak,
                                   "fp64_pipe": {"achieved": fp64_ops / 1e12, "peak": fp64_peak / 1e12,
                                                 "unit": "T thread-instr/s", "frac": fp64_ops / fp64_peak},
                                   "def": "SASS count of the kernel's iteration loop (tools/sass_count.py): %d warp "
                                          "instructions per codeword and iteration, %d on the fp64 pipe; issue peak = "
                                          "148 SMs x 4 schedulers x max SM clock; fp64 peak measured by "
                                          "ldpc535_probe_pipe_peak (2 DADD : 1 DSETP)"
                                          % (MINSUM_INSTR_PER_CW_ITER, MINSUM_FP64_PER_CW_ITER)}
            rows.append(row)
    out["configs[0] shipped code, reference block settings"] = rows

    # ---- hard decision (method 3): HBM-bound ----
    def hard():
        c4.decode_dev(d_sym.data_ptr(), n_cw * N_SYM, n_cw, d_bytes.data_ptr(), d_synd.data_ptr(),
                      d_iters.data_ptr(), method=L.METHOD_HARD, max_iters=5, early_stop=True, stream=sp)
    ms = time_launches(stream, hard, 5, 2)
    gbs = n_cw * BYTES_PER_CW / ms / 1e6
    out["hard decision (method 3), shipped code"] = {
        "codewords": n_cw, "kernel": c4.kernel_name(L.METHOD_HARD), "ms": ms, "gbit_s": n_cw * K_INFO / ms / 1e6,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak}}

    # ---- encoder, shipped code ----
    def enc4():
        c4.encode_dev(d_data.data_ptr(), n_cw, d_sym.data_ptr(), stream=sp)
    ms = time_launches(stream, enc4, 5, 2)
    gbs = n_cw * (4 + 8 * N_SYM) / ms / 1e6
    out["encoder, shipped code"] = {
        "frames": n_cw, "ms": ms, "gbit_s": n_cw * K_INFO / ms / 1e6,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak}}
    del d_bytes, d_synd, d_iters

    # ---- configs[3]: (3,6)-regular rate-1/2 n = 8192, early termination on, batch sweep ----
    c8, seed8 = L.codes.first_invertible(n=8192, seed=SEED, device=c4.device)
    M8, N8, K8, E8 = c8.M, c8.N, c8.K, c8.E
    nmax = 1_000_000
    free, _ = torch.cuda.mem_get_info()
    while nmax > 1000 and nmax * (N8 * 8 + 2 * c8.nbytes + 2) > 0.8 * free:
        nmax //= 10
    d8_data, d8_sym = synth_shard(c8, 0, nmax, EBN0_DB, sp)
    d8_bytes = torch.empty((nmax, c8.nbytes), dtype=torch.uint8, device="cuda")
    d8_synd = torch.empty(nmax, dtype=torch.uint8, device="cuda")
    d8_iters = torch.empty(nmax, dtype=torch.uint8, device="cuda")
    rows = []
    for nb in (1000, 10_000, 100_000, 1_000_000):
        if nb > nmax:
            continue

        def one():
            c8.decode_dev(d8_sym.data_ptr(), nb * N8, nb, d8_bytes.data_ptr(), d8_synd.data_ptr(),
                          d8_iters.data_ptr(), method=L.METHOD_SUMPRODUCT, max_iters=MAX_ITERS,
                          early_stop=True, stream=sp)
        ms = time_launches(stream, one, 2 if nb >= 1_000_000 else 4, 1 if nb >= 100_000 else 3)
        it_sum = float(d8_iters[:nb].sum(dtype=torch.float64).item())
        mufu = (MUFU_PER_EDGE_ITER * it_sum * E8 + nb * N8) / (ms * 1e-3)
        rows.append({"batch": nb, "ms": ms, "gbit_s": nb * K8 / ms / 1e6, "mean_iters": it_sum / nb,
                     "edge_iterations_per_s": it_sum * E8 / (ms * 1e-3),
                     "frames_recovered": float((d8_bytes[:nb] == d8_data[:nb]).all(dim=1).float().mean().item()),
                     "roofline": {"bound": "sm_xu", "achieved": mufu / 1e12, "peak": mufu_peak / 1e12,
                                  "unit": "T MUFU/s", "frac": mufu / mufu_peak}})
    nfix = min(nmax, 20_000)

    def fixed():
        c8.decode_dev(d8_sym.data_ptr(), nfix * N8, nfix, d8_bytes.data_ptr(), d8_synd.data_ptr(),
                      d8_iters.data_ptr(), method=L.METHOD_SUMPRODUCT, max_iters=20, early_stop=False, stream=sp)
    ms = time_launches(stream, fixed, 3, 1)
    mufu = (MUFU_PER_EDGE_ITER * 20.0 * nfix * E8 + nfix * N8) / (ms * 1e-3)
    out["configs[3] (3,6)-regular n=8192, seed %d, sum-product, 50 iterations max, early stop on, 2 dB" % seed8] = {
        "kernel": c8.kernel_name(L.METHOD_SUMPRODUCT), "batches": rows,
        "fixed_20_iterations": {"codewords": nfix, "ms": ms, "edge_iterations_per_s": 20.0 * nfix * E8 / (ms * 1e-3),
                                "roofline": {"bound": "sm_xu", "achieved": mufu / 1e12, "peak": mufu_peak / 1e12,
                                             "unit": "T MUFU/s", "frac": mufu / mufu_peak}}}
    del d8_bytes, d8_synd, d8_iters

    # ---- encoder, n = 8192 code (table look-up kernel) ----
    rows = []
    for nf in (2_000, 20_000, 400_000):
        if nf > nmax:
            continue

        def enc8():
            c8.encode_dev(d8_data.data_ptr(), nf, d8_sym.data_ptr(), stream=sp)
        ms = time_launches(stream, enc8, 4, 2)
        gbs = nf * (c8.nbytes + 8 * N8) / ms / 1e6
        rows.append({"frames": nf, "ms": ms, "gbit_s": nf * K8 / ms / 1e6,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": gbs / hbm_peak}})
    out["encoder, n=8192 code"] = rows
    c8.close()
    return out


def run_gpu(args):
    import numpy as np
    import torch
    import ldpc_ece535a as L
    from ldpc_ece535a import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    cores_host = len(os.sched_getaffinity(0))
    numa_note = bind_near_gpu(local) if world > 1 else "single rank: affinity left alone"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_cw = args.codewords
    first_frame = rank * n_cw                # this rank's contiguous shard of the global stream
    code = L.Code(None, device=local)
    if args.kernel:
        code.set_kernel(args.kernel)
    # the launcher knows how many ranks share the host: give each its share of the cores and let
    # the handle decide pack-or-raw chunk by chunk from its own measurements -- unless told otherwise
    share = max(1, min(cores_host // local_world, len(os.sched_getaffinity(0))))
    code.set_host_path(args.pack_pinned, args.pack_threads or share)
    stream = torch.cuda.Stream()            # a real stream: NULL would mean "the handle's own"
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert sp.value, "expected a non-default stream"

    d_data, d_sym = synth_shard(code, first_frame, n_cw, EBN0_DB, sp)
    d_bytes = torch.empty((n_cw, 4), dtype=torch.uint8, device="cuda")
    d_synd = torch.empty(n_cw, dtype=torch.uint8, device="cuda")
    d_iters = torch.empty(n_cw, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def dev_pass():
        code.decode_dev(d_sym.data_ptr(), n_cw * N_SYM, n_cw, d_bytes.data_ptr(), d_synd.data_ptr(),
                        d_iters.data_ptr(), method=L.METHOD_SUMPRODUCT, max_iters=MAX_ITERS,
                        early_stop=False, stream=sp)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not dist:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        dev_pass()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    l0 = code.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for k in range(args.steps):
        dev_pass()
        evs[k + 1].record(stream)
    barrier()
    launches = code.launch_count() - l0
    total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]))
    kern_ms = sum(evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = n_cw * world * args.steps * K_INFO / (total_ms * 1e-3) / 1e9

    # correctness guard on the timed output: the decoded shard must mostly equal the data sent
    frame_ok = float((d_bytes == d_data).all(dim=1).float().mean().item())
    bad_iters = int((d_iters != MAX_ITERS).sum().item())
    # a digest of the shard's data bytes: equal for the same global indices at any GPU count
    shard_digest = int(d_data.view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF

    # ---- end to end through the host-buffer C ABI ----
    # The same shard, from pinned host memory.  If the box cannot pin a whole shard (5.1 GB per
    # rank) the leg falls back to the largest prefix it can pin, agreed across ranks, and says so.
    e2e_cw = n_cw
    bufs = None
    while True:
        try:
            bufs = [pinned_array(L, e2e_cw * 512, np.float32, (e2e_cw * N_SYM * 2,)),
                    pinned_array(L, e2e_cw * 4, np.uint8, (e2e_cw, 4)),
                    pinned_array(L, e2e_cw, np.uint8, (e2e_cw,)),
                    pinned_array(L, e2e_cw, np.uint8, (e2e_cw,))]
            ok = 1
        except L.Ldpc535Error:
            ok = 0
        if dist:
            t = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            all_ok = int(t.item())
        else:
            all_ok = ok
        if all_ok:
            break
        if bufs:
            for _, ptr in bufs:
                _abi.lib().ldpc535_host_free(ptr)
            bufs = None
        e2e_cw //= 2
        if e2e_cw < 100_000:
            raise SystemExit("cannot pin host memory for the end-to-end leg")
    (h_sym, p1), (h_bytes, p2), (h_synd, p3), (h_iters, p4) = bufs
    _abi.check(_abi.lib().ldpc535_memcpy_d2h(code.handle, p1, d_sym.data_ptr(), e2e_cw * 512), "d2h")
    h_sym_c = h_sym.view(np.complex64)

    def e2e_pass():
        code.decode(h_sym_c, method=L.METHOD_SUMPRODUCT, max_iters=MAX_ITERS, early_stop=False,
                    n_win=e2e_cw, out=(h_bytes, h_synd, h_iters))

    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_pass()
    hs0 = code.host_stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = e2e_cw * world * e2e_steps * K_INFO / (e2e_ms * 1e-3) / 1e9
    e2e_same = bool(np.array_equal(h_bytes, d_bytes[:e2e_cw].cpu().numpy()))
    hp = code.host_path()
    hs1 = code.host_stats()
    n_packed, n_raw = hs1["chunks_packed"] - hs0["chunks_packed"], hs1["chunks_raw"] - hs0["chunks_raw"]
    host_path = {0: "copy engine reads the caller's pinned buffer directly",
                 1: "%d host threads pack the real parts into pinned staging (halves PCIe bytes)" % hp["pack_threads"],
                 2: "measured: %d host threads pack the real parts of a 128 MiB chunk into pinned staging while they do a byte "
                    "faster than the copy engine, otherwise it reads the caller's pinned buffer directly (rank 0 in the "
                    "timed passes: %d chunks packed, %d raw)" % (hp["pack_threads"], n_packed, n_raw)}[hp["pack_pinned"]]
    pcie_bytes = (hs1["h2d_bytes"] - hs0["h2d_bytes"]) // e2e_steps * world      # rank 0's mix, every rank

    # ---- the box's host->device ceiling with every rank copying at once: plain pinned copies of the
    # same buffer in the pipeline's chunk size, no kernel, no packing ----
    chunk = 128 << 20
    h2d_total = min(e2e_cw * 512, 2 << 30)
    d_sink = torch.empty(chunk, dtype=torch.uint8, device="cuda")

    def h2d_sweep():
        for off in range(0, h2d_total, chunk):
            _abi.lib().ldpc535_memcpy_h2d(code.handle, d_sink.data_ptr(), C.c_void_p(p1.value + off),
                                          min(chunk, h2d_total - off))
    h2d_sweep()
    barrier()
    t0 = time.perf_counter()
    h2d_sweep()
    h2d_s = max_over_ranks(time.perf_counter() - t0)
    h2d_ceiling = h2d_total * world / h2d_s / 1e9          # GB/s, all ranks together
    del d_sink

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
        f_max = float((clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        mufu_peak = code.probe_pipe_peak(_abi.PIPE_MUFU)          # measured now, on this GPU
        fp64_peak = code.probe_pipe_peak(_abi.PIPE_FP64)
        t_k = kern_ms * 1e-3
        hbm_achieved = n_cw * BYTES_PER_CW / t_k / 1e9
        edge_it = n_cw * E_EDGES * MAX_ITERS / t_k
        mufu = (MUFU_PER_EDGE_ITER * n_cw * E_EDGES * MAX_ITERS + n_cw * N_SYM) / t_k
        xu_nominal = SM_COUNT * XU_LANES * f_max
        kname = code.kernel_name(L.METHOD_SUMPRODUCT)
        per_cw, tsrc = traffic_from_profiles({"c4-thread": "decode_c4_thread_kernel"}.get(kname, kname))
        roofline = {
            # the roofline that governs a fixed-50-iteration sum-product: the SM special-function
            # pipe; the HBM figure sits beside it
            "bound": "sm_xu", "achieved": mufu / 1e12, "peak": mufu_peak / 1e12, "unit": "T MUFU/s",
            "frac": mufu / mufu_peak,
            "peak_source": "ldpc535_probe_pipe_peak(MUFU) on this GPU in this run: 1 ex2 : 2 lg2 mix, best of 16/32/64 "
                           "warps per SM = %.2f lanes/clk/SM at the %.0f MHz max clock" % (mufu_peak / SM_COUNT / f_max, f_max / 1e6),
            "def": "executed MUFU per launch = 3 per edge-iteration (1 ex2 + 2 lg2, spa_math.cuh) x E=168 x 50 "
                   "iterations + 64 start-up ex2, per codeword, / the kernel's mean launch time",
            "edge_iterations_per_s": edge_it,
            "frac_algorithmic_4xu_nominal": XU_PER_EDGE_ITER * edge_it / xu_nominal,
            "kernel": kname, "kernel_ms": kern_ms,
            "traffic": None if per_cw is None else n_cw * per_cw, "traffic_unit": "bytes per launch",
            "traffic_source": tsrc, "algorithmic_bytes": n_cw * BYTES_PER_CW,
            "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_achieved / hbm_peak, "peak_source": peak_src}}
        line = {"metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config(n_cw, world),
                "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": e2e_cw * world * 512,
                        "d2h_bytes_per_step": e2e_cw * world * 6, "steps": e2e_steps,
                        "codewords_per_gpu": e2e_cw,
                        "ms_per_step": e2e_ms / e2e_steps,
                        "api": "ldpc535_decode_batch (pinned host buffers, 3-slot stage/H2D/kernel/D2H pipeline)",
                        "host_path": host_path, "host_cores_per_rank": share, "numa": numa_note,
                        "pcie_h2d_bytes_per_step": pcie_bytes,
                        "h2d_ceiling_gbs": h2d_ceiling,
                        "h2d_ceiling_def": "all %d ranks copying their pinned shard at once, 128 MiB cudaMemcpyAsync "
                                           "chunks, no kernel" % world,
                        "frac_of_h2d_ceiling": pcie_bytes / (e2e_ms / e2e_steps * 1e-3) / 1e9 / h2d_ceiling,
                        "matches_device_path": e2e_same},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "check": {"frames_recovered": frame_ok, "frames_not_at_max_iters": bad_iters,
                          "inputs": "Philox4x32-10, seed %d, counter = global codeword index (shard of rank r "
                                    "starts at r x codewords_per_gpu)" % SEED,
                          "rank0_shard_digest": shard_digest}}
    for p in (p1, p2, p3, p4):
        _abi.lib().ldpc535_host_free(p)
    if rank == 0 and world == 1 and not args.no_extras:
        del d_bytes, d_synd, d_iters
        line["configs"] = extra_configs(L, torch, np, code, stream, sp, d_data, d_sym, n_cw, first_frame,
                                        mufu_peak, fp64_peak, hbm_peak, f_max)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.cpu_per_core)
        print(json.dumps(line))
    code.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--codewords", type=int, default=10_000_000, help="codewords per GPU per step")
    ap.add_argument("--cpu-per-core", type=int, default=20_000, help="cpu_baseline sample per host core (BASELINE.md 5.3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the `configs` measurements (other BASELINE configs)")
    ap.add_argument("--kernel", default=None, help="force a decoder kernel family (warp/block/c4-thread)")
    ap.add_argument("--pack-pinned", type=int, default=-1, help="e2e host path: 1 pack, 0 raw DMA, -1 from the core share")
    ap.add_argument("--pack-threads", type=int, default=0, help="e2e packing team size (0: cores / ranks on this host)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
