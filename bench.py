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
  roofline  HBM bytes of the decode kernel against the measured copy bandwidth (the schema's
            bound) plus `governing`: the SM special-function / issue roofline that actually
            bounds a 50-iteration sum-product (DESIGN.md "Rooflines").
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
# XU-pipe ceiling measured on this pool's B200 with tools/microbench/mufu_peak.cu (independent
# ex2/lg2 streams, 1 ex2 : 2 lg2 mix): 15.9 MUFU lanes/clk/SM at 1965 MHz (profiles/r1_microbench.txt)
XU_MEASURED_PEAK = 4.63e12


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
            "codewords_per_s": n / dt, "port_same_work": port}


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
def pinned_array(L, nbytes, dtype, shape):
    import numpy as np
    from ldpc_ece535a import _abi
    p = C.c_void_p()
    _abi.check(_abi.lib().ldpc535_host_alloc(nbytes, C.byref(p)), "host_alloc")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def run_gpu(args):
    import numpy as np
    import torch
    import ldpc_ece535a as L
    from ldpc_ece535a import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_cw = args.codewords
    code = L.Code(None, device=local)
    if args.kernel:
        code.set_kernel(args.kernel)
    stream = torch.cuda.Stream()            # a real stream: NULL would mean "the handle's own"
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert sp.value, "expected a non-default stream"

    # ---- synthetic shard, generated on the device (seed = 535 + global shard index) ----
    gen = torch.Generator(device="cuda")
    gen.manual_seed(535 + rank)
    d_data = torch.randint(0, 256, (n_cw, 4), dtype=torch.uint8, device="cuda", generator=gen)
    d_sym = torch.empty((n_cw, N_SYM, 2), dtype=torch.float32, device="cuda")
    code.encode_dev(d_data.data_ptr(), n_cw, d_sym.data_ptr(), stream=sp)
    sigma = float(np.sqrt(10.0 ** (-EBN0_DB / 10.0)))
    step = 1 << 20
    for a in range(0, n_cw, step):
        b = min(n_cw, a + step)
        d_sym[a:b, :, 0] += sigma * torch.randn((b - a, N_SYM), device="cuda", generator=gen)
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
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = e2e_cw * world * e2e_steps * K_INFO / (e2e_ms * 1e-3) / 1e9
    e2e_same = bool(np.array_equal(h_bytes, d_bytes[:e2e_cw].cpu().numpy()))
    hp = code.host_path()
    host_path = ("%d host threads pack the real parts into pinned staging (halves PCIe bytes)" % hp["pack_threads"]
                 if hp["pack_pinned"] else "copy engine reads the caller's pinned buffer directly")
    pcie_bytes = e2e_cw * world * (256 if hp["pack_pinned"] else 512)

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
        t_k = kern_ms * 1e-3
        hbm_achieved = n_cw * BYTES_PER_CW / t_k / 1e9
        edge_it = n_cw * E_EDGES * MAX_ITERS / t_k
        xu_peak_ops = SM_COUNT * XU_LANES * f_max
        roofline = {
            "bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": hbm_achieved / hbm_peak,
            "frac_governing": MUFU_PER_EDGE_ITER * edge_it / XU_MEASURED_PEAK,   # see `governing` below
            # ncu --set full, dram__bytes_read+write of this kernel: 523.2 B per codeword
            # (profiles/r1_kernels_ncu_full.txt) against 518 B algorithmic
            "traffic": n_cw * 523.2 if code.kernel_name(L.METHOD_SUMPRODUCT) == "c4-thread" else None,
            "traffic_unit": "bytes per launch", "algorithmic_bytes": n_cw * BYTES_PER_CW,
            "peak_source": peak_src,
            "kernel": code.kernel_name(L.METHOD_SUMPRODUCT), "kernel_ms": kern_ms,
            "note": "a fixed-50-iteration sum-product is bound by the SM special-function/issue "
                    "pipes, not HBM (SURVEY 8d); `governing` is the roofline that applies",
            "governing": {"bound": "sm_xu", "achieved": MUFU_PER_EDGE_ITER * edge_it / 1e12,
                          "peak": XU_MEASURED_PEAK / 1e12, "unit": "T MUFU/s",
                          "frac": MUFU_PER_EDGE_ITER * edge_it / XU_MEASURED_PEAK,
                          "edge_iterations_per_s": edge_it,
                          "frac_algorithmic_4xu_nominal": XU_PER_EDGE_ITER * edge_it / xu_peak_ops,
                          "def": "executed MUFU (3 per edge-iteration: 1 ex2 + 2 lg2) x E=168 x 50 iterations per "
                                 "codeword against the XU-pipe peak measured with tools/microbench/mufu_peak.cu; "
                                 "frac_algorithmic_4xu_nominal uses SURVEY 8d's 4 XU ops per edge-iteration "
                                 "against 148 SMs x 16 lanes x max SM clock"}}
        line = {"metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config(n_cw, world),
                "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": e2e_cw * world * 512,
                        "d2h_bytes_per_step": e2e_cw * world * 6, "steps": e2e_steps,
                        "codewords_per_gpu": e2e_cw,
                        "ms_per_step": e2e_ms / e2e_steps,
                        "api": "ldpc535_decode_batch (pinned host buffers, 3-slot stage/H2D/kernel/D2H pipeline)",
                        "host_path": host_path, "pcie_h2d_bytes_per_step": pcie_bytes,
                        "matches_device_path": e2e_same},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "check": {"frames_recovered": frame_ok, "frames_not_at_max_iters": bad_iters}}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.cpu_per_core)
        print(json.dumps(line))
    for p in (p1, p2, p3, p4):
        _abi.lib().ldpc535_host_free(p)
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
    ap.add_argument("--cpu-per-core", type=int, default=4000, help="cpu_baseline sample per host core")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel", default=None, help="force a decoder kernel family (warp/block/c4-thread)")
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
