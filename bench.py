#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json: acquisition-step candidates/sec.

A "step" is one pass of the hot path over one batch of synthetic candidates: b200bo_acquire on M candidate
columns against a resident N-observation GP factor (k*, both triangular solves, mu, sigma^2, score, arg-max; the
one-off fit is outside the step).  Default workload = the metric text of BASELINE.json (N=2048, D=8, SEArd, EI,
M=65536 LHS candidates per GPU); side blocks: configs[1] (cfg2), configs[2] with gradient (cfg3), configs[4] shard (cfg5), configs[3] (MAP sweep).

  python bench.py --gpus N --steps K --warmup W        (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                 (CPU restatement of the reference path on the host cores)

`value`  : device-resident inputs, per-step CUDA-event pairs on the launching stream, max over ranks.
`e2e`    : the public host-pointer API (pinned host candidates -> H2D -> fused launch -> D2H best -> rank exchange).
`roofline`: dominant kernel (acq_i8_gemm_kernel, tcgen05 kind::i8) in int8 TOP/s against the self-measured UTCIMMA peak.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kernel, D, N, M per GPU, acquisition, gradient)
    "cfg2": dict(kernel="Mat52Ard", D=6, N=2048, M=65536, acq="UCB", grad=False,
                 desc="BASELINE configs[1]: Hartmann-6 D=6, N=2048, Mat52Ard, UCB(BrochuBetaScaling), M=65536 LHS candidates"),
    "cfg3": dict(kernel="SEArd", D=32, N=4096, M=262144, acq="EI", grad=True,
                 desc="BASELINE configs[2]: synthetic D=32, N=4096, SEArd, EI + gradient, M=262144 LHS candidates"),
    "cfg5": dict(kernel="SEArd", D=16, N=8192, M=131072, acq="TS", grad=False,
                 desc="BASELINE configs[4] per-GPU shard: synthetic D=16, N=8192, SEArd, ThompsonSamplingSimple, M=1048576/8"),
    "metric": dict(kernel="SEArd", D=8, N=2048, M=65536, acq="EI", grad=False,
                   desc="metric text: GP posterior + EI at N=2048, D=8, M=65536 candidates"),
    "target": dict(kernel="SEArd", D=8, N=4096, M=262144, acq="EI", grad=False,
                   desc="north_star target: GP posterior + EI at N=4096, D=8, M=262144 candidates"),
}
METRIC = "acquisition-step candidates/sec (GP posterior + acquisition score + arg-max over an M-candidate sweep)"


def workload_config(w, world):
    """the `config` object both arms print (same keys, so the driver's same_config check compares like with like)"""
    return {"workload": w["desc"], "candidates_per_gpu": w["M"], "N": w["N"], "D": w["D"], "kernel": w["kernel"], "acquisition": w["acq"],
            "gradient": w["grad"], "parallelism": f"candidate-sharded x{world}, replicated factor, one 16 B/rank all-gather",
            "l2": "256 MiB memset between timed steps (outside the per-step event pairs) flushes the 126 MB L2"}


def synth(w, seed=2):
    """synthetic inputs: X ~ U[0,1]^{D x N}; cfg2: y = -hartmann6; else smooth bumps + noise."""
    from oracle import gp_oracle as orc   # data generators only (hartmann6 / LHS); not on the timed path
    rng = np.random.default_rng(seed)
    D, N = w["D"], w["N"]
    X = rng.random((D, N))
    if D == 6 and w["kernel"] == "Mat52Ard":
        y = -orc.hartmann6(X); ll = np.zeros(D)
    else:
        c = rng.random((D, 8))
        y = sum(np.exp(-0.5 * np.sum((X - c[:, k:k + 1]) ** 2, axis=0) / 0.15) for k in range(8)) + np.exp(-2.0) * rng.standard_normal(N)
        ll = np.full(D, np.log(np.sqrt(D) * 0.25))
    return np.asfortranarray(X), y, ll


def acq_params(w, y):
    from oracle.gp_oracle import brochu_beta
    return {"UCB": (brochu_beta(w["D"], w["N"]),), "EI": (float(np.max(y)),), "PI": (float(np.max(y)),), "TS": (),
            "MI": (1.0, 0.1), "MaxMean": ()}[w["acq"]]


def candidates(w, rank, seed=20):
    from oracle.gp_oracle import latin_hypercube_sampling
    return np.asfortranarray(latin_hypercube_sampling(np.zeros(w["D"]), np.ones(w["D"]), w["M"], np.random.default_rng(seed + rank)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, w):
    """--impl reference: the reference's CPU path restated (oracle/oracle.c, 'port'): per-candidate predict_f loop +
    functor, candidates split across all host threads.  Each step = a bounded sample of the workload's candidates."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gp_oracle as orc
    from oracle.c_oracle import COracle
    X, y, ll = synth(w)
    gp = orc.GPOracle(w["D"], w["kernel"], "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    co = COracle(gp)
    par = acq_params(w, y)
    Xs = candidates(w, 0)
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: ask for all host threads explicitly
    nthr = os.cpu_count() or 1
    t0 = time.perf_counter(); r = co.acquire(w["acq"], par, Xs[:, :256], want_grad=w["grad"], nthreads=nthr); pilot = (time.perf_counter() - t0) / 256
    budget = min(2.0, 150.0 / max(args.steps + args.warmup, 1))
    sample = int(max(256, min(w["M"], budget / pilot)))
    for _ in range(args.warmup):
        co.acquire(w["acq"], par, Xs[:, :sample], want_grad=w["grad"], nthreads=nthr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = co.acquire(w["acq"], par, Xs[:, :sample], want_grad=w["grad"], nthreads=nthr)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    cb = {"value": val, "unit": "candidates/s", "cores": int(r["threads"]), "kind": "port",
          "sample": f"{sample} of {w['M']} candidates per step, per-candidate predict_f loop (dtrsv-shaped) + functor, "
                    f"oracle/oracle.c with OpenMP over candidates; host has {os.cpu_count()} logical CPUs"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f64", "data": "synthetic", "config": workload_config(w, args.gpus),
                      "cpu_baseline": cb, "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def roofline_i8(w, M, gemm_ms, step_ms, peak_i8, peak_fp64, peaks, traffic):
    """dominant kernel = acq_i8_gemm_kernel (tcgen05.mma kind::i8).  Algorithmic work per candidate: the value path is
    v = W k* over the lower triangle = N^2/2 multiply-adds per slice product x 28 products (p + q <= 6) = 28 N^2 int8 ops (2 per MAC); a
    gradient launch adds the second triangular product w = W^T v, another 28 N^2."""
    N = float(w["N"])
    ops = 28.0 * N * N * (2.0 if w["grad"] else 1.0)
    fp64_flops = N * N * (2.0 if w["grad"] else 1.0)
    ach = ops * M / (gemm_ms * 1e-3) * 1e-12
    bf16 = peaks.get("bf16_tflops")
    Np = float((int(N) + 127) // 128 * 128)
    ch = min(float(int(64 * 2 ** 20 / (7 * Np))), float(M))
    tr = traffic or {}
    return {"kernel": "acq_i8_gemm_kernel (tcgen05.mma.cta_group::1.kind::i8, 128x64x32, TMEM accumulators)", "bound": "tensor",
            "achieved": ach, "peak": peak_i8, "unit": "TFLOP/s", "frac": ach / peak_i8, "traffic": tr.get("acq_i8_gemm_kernel_per_launch"),
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE acq_i8_gemm_kernel launch (one chunk of candidates), ncu capture of this "
                            "workload in profiles/ncu_traffic.json; algorithmic bytes of such a launch = the int8 slices of the triangular factor "
                            "inverse (7 N^2 / 2) + the chunk's k* slices (7 CH N)",
            "algorithmic_bytes_per_launch": 3.5 * Np * Np + 7.0 * ch * Np, "traffic_per_step_all_kernels": tr.get("bytes_per_step"),
            "ops": "int8 multiply-adds of the 28 error-free slice products, 2 ops each (achieved and peak are int8 TOP/s)",
            "peak_source": "self-measured tcgen05.mma kind::i8 rate of this GPU (b200bo_i8_peak_tops: back-to-back 128x256x32 MMAs on "
                           "resident operands, all SMs); MEASURED_PEAKS.json has no int8 figure" +
                           (f" -- its bf16 burst figure x 2 would be {2 * bf16:.0f}" if bf16 else ""),
            "frac_of_nominal": ach / 4500.0,
            "peak_note": "the self-measured peak moves with the power state of the run (3.86-4.58 POP/s seen under sw_power_cap); frac_of_nominal "
                         "divides by the nominal dense int8 rate of 4500 TOP/s instead",
            "algorithmic_int8_ops_per_candidate": ops, "launch_ms_sum_per_step": gemm_ms,
            "launch_timing": "CUDA-event pairs around every acq_i8_gemm_kernel launch of one step on the launching stream (one chunk lane)",
            "step_view": {"achieved": ops * M / (step_ms * 1e-3) * 1e-12, "frac": ops * M / (step_ms * 1e-3) * 1e-12 / peak_i8,
                          "note": "same ops over the WHOLE step (kernel evaluation, scores and arg-max included)"},
            "fp64_equivalent": {"flops_per_candidate": fp64_flops, "achieved_tflops": fp64_flops * M / (step_ms * 1e-3) * 1e-12,
                                "dmma_peak_tflops": peak_fp64,
                                "note": "the FP64 flops the replaced triangular solves would need, over the whole step; the round-1 DMMA kernel "
                                        "ran at 0.85 of the self-measured DMMA peak"}}


def run_gpu_workload(ctx, w, steps, warmup, want_fit_side=True, m_local=None):
    """one workload on this rank's GPU: device-resident arm, end-to-end arm, slice-product timing pass.  m_local overrides the
    candidates per rank (strong scaling: a fixed total split over the ranks)."""
    if m_local is not None:
        w = dict(w, M=int(m_local))
    import torch
    import b200bo
    from b200bo import _lib
    dist, dev, stream, world, rank, flush = ctx["dist"], ctx["dev"], ctx["stream"], ctx["world"], ctx["rank"], ctx["flush"]
    D, N, M = w["D"], w["N"], w["M"]
    X, y, ll = synth(w)
    model = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.gp._Kernel(w["kernel"], ll, 0.0), logNoise=-2.0, capacity=N,
                           device=ctx["local_rank"])
    model.fit(X, y)                                  # one-off fit (identical, replicated on every rank)
    model.fit(X, y)                                  # second fit: warm timings for the side metrics
    fit_ms = {k: model.timing_ms(v) for k, v in dict(kmat=_lib.T_KMAT, chol=_lib.T_CHOL, syrk=_lib.T_SYRK, alpha=_lib.T_ALPHA).items()}
    par = np.array(acq_params(w, y), float)
    kind = _lib.ACQ_KINDS[w["acq"]]
    _lib.check(_lib.lib.b200bo_set_stream(model._h, C.c_void_p(stream.cuda_stream)), model._h)
    if world > 1:
        from b200bo.dist import comm_attach
        comm_attach(model, device=dev)               # the library's own communicator: the all-gather + merge run inside libb200bo

    # ---- device-resident arm -----------------------------------------------------------------------------------
    Xs_host = torch.from_numpy(candidates(w, rank).T.copy()).pin_memory()          # [M][D] == D x M column-major, pinned
    dXs = Xs_host.to(dev)
    dbest = torch.zeros(2, dtype=torch.float64, device=dev)                        # b200bo_best_t {f64 value; i64 index}: the GLOBAL best
    dgrad = torch.empty((M, D), dtype=torch.float64, device=dev) if w["grad"] else None
    pp = par.ctypes.data_as(C.POINTER(C.c_double)) if par.size else None
    offset = rank * M

    def step_dev():
        # fused launch(es) + (N > 1) the ONE exchange step: a 272 B/rank ncclAllGather and the merge, all enqueued by the library
        _lib.check(_lib.lib.b200bo_acquire_dev(model._h, kind, pp, par.size, C.c_void_p(dXs.data_ptr()), M, 50, offset, None,
                                               C.c_void_p(dgrad.data_ptr()) if dgrad is not None else None, None, None,
                                               C.c_void_p(dbest.data_ptr())), model._h)

    for _ in range(warmup):
        step_dev(); flush.zero_()
    torch.cuda.synchronize(dev)
    launches0 = model.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    for e0, e1 in evs:
        e0.record(stream); step_dev(); e1.record(stream)
        flush.zero_()                                                              # L2 flush between timed steps
        torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches = (model.launch_count - launches0) // steps
    t_dev = sum(e0.elapsed_time(e1) for e0, e1 in evs) * 1e-3
    gb = dbest.cpu().numpy()
    best_v, best_i = float(gb[0]), int(gb.view(np.int64)[1])

    # ---- slice-product timing pass (roofline numerator): event pairs around every acq_i8_gemm_kernel launch, one chunk lane ----
    model.set_knob("acq_gemm_timing", 1)
    gemm_ms = []
    for _ in range(3):
        step_dev(); torch.cuda.synchronize(dev)
        gemm_ms.append(model.timing_ms(_lib.T_ACQ_GEMM))
        flush.zero_(); torch.cuda.synchronize(dev)
    model.set_knob("acq_gemm_timing", 0)
    gemm_ms = float(np.median(gemm_ms))

    # ---- end-to-end arm: public host API, pinned host candidates, H2D + D2H inside the timed region -------------
    Xs_np = Xs_host.numpy().T                                                      # D x M view of the pinned buffer (F-order)
    grad_np = torch.empty((M, D), dtype=torch.float64).pin_memory().numpy().T if w["grad"] else None   # pinned result buffer (D x M, F-order)
    def step_e2e():
        r = model.acquire(w["acq"], par, Xs_np, seed=50, idx_offset=offset, want_values=False, want_grad=w["grad"], grad_out=grad_np)
        return r["best_value"], r["best_index"]                                    # global on every rank (library-side exchange)
    for _ in range(warmup):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        bv_e2e, bi_e2e = step_e2e()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_e2e = e0.elapsed_time(e1) * 1e-3

    tt = torch.tensor([t_dev, t_e2e, gemm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, gemm_ms = (float(v) for v in tt.cpu())
    out = dict(model=model, X=X, y=y, ll=ll, par=par, t_dev=t_dev, t_e2e=t_e2e, gemm_ms=gemm_ms, launches=int(launches), fit_ms=fit_ms,
               best=(best_v, best_i, float(bv_e2e), int(bi_e2e)))
    if world > 1:
        model.comm_destroy()
    return out


def line_for(w, r, steps, world, peaks, traffic_key):
    """value / e2e / roofline of one measured workload"""
    model, M, D = r["model"], w["M"], w["D"]
    total = M * world
    step_ms = r["t_dev"] / steps * 1e3
    if "peak_i8" not in peaks:
        peaks["peak_i8"] = model.i8_peak_tops(); peaks["peak_fp64"] = model.fp64_peak_tflops()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(traffic_key)
    except Exception:
        pass
    d2h = 16 + 8 * D + (8 * D * M if w["grad"] else 0)
    return {"value": total * steps / r["t_dev"], "ms_per_step": step_ms,
            "e2e": {"value": total * steps / r["t_e2e"], "unit": "candidates/s", "h2d_bytes_per_step": int(8 * D * M * world),
                    "d2h_bytes_per_step": int(d2h * world), "ms_per_step": r["t_e2e"] / steps * 1e3},
            "gpu_launches": r["launches"],
            "roofline": roofline_i8(w, M, r["gemm_ms"], step_ms, peaks["peak_i8"], peaks["peak_fp64"], peaks, traffic),
            "best": {"value": r["best"][0], "index": r["best"][1], "e2e_value": r["best"][2], "e2e_index": r["best"][3]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the cfg3 / cfg5 side blocks and the N=4096 fit metrics")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libb200bo has no CPU fallback (use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import b200bo
    from b200bo import _lib

    stream = torch.cuda.Stream(dev)                 # a real (non-default) stream shared by torch and the library (NCCL included)
    torch.cuda.set_stream(stream)
    ctx = dict(dist=dist, dev=dev, stream=stream, world=world, rank=rank, local_rank=local_rank,
               flush=torch.empty(256 << 20, dtype=torch.uint8, device=dev))        # > 126 MB L2
    saved_out = None
    if world > 1:                                   # NCCL banner of the library's own communicator: keep stdout clean
        sys.stdout.flush(); saved_out = os.dup(1); os.dup2(2, 1)
    clk = ClockSampler(local_rank)
    clk.__enter__()                                  # sampled from the warm-up to the end of the e2e arm (all under load)
    r = run_gpu_workload(ctx, w, args.steps, args.warmup)
    clk.__exit__(None, None, None)
    if saved_out is not None:
        sys.stdout.flush(); os.dup2(saved_out, 1); os.close(saved_out)
    D, N, M = w["D"], w["N"], w["M"]
    model, fit_ms = r["model"], r["fit_ms"]
    # the north_star's target shape (N=4096, D=8, M=262144) at this GPU count: every rank takes part (library-side exchange).  Weak: 262144
    # candidates per GPU, like the headline; strong: 262144 candidates in total, split over the ranks.
    tgt = {}
    if not args.no_side and args.workload != "target":
        try:
            wt = WORKLOADS["target"]
            tgt["weak"] = (wt, run_gpu_workload(ctx, wt, 3, 3))
            if world > 1:
                ws = dict(wt, M=wt["M"] // world)
                tgt["strong"] = (ws, run_gpu_workload(ctx, ws, 3, 3))
        except Exception as exc:
            tgt = {"error": str(exc)}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ln = line_for(w, r, args.steps, world, peaks, args.workload)
        peak_fp64 = peaks["peak_fp64"]
        nblk, far_flops = (N + 127) // 128, 0.0          # flops of the K=512 trailing updates timed by T_SYRK (csrc/chol.cu schedule)
        for p0 in range(0, nblk, 4):
            p1, p2 = min(p0 + 4, nblk), min(p0 + 8, nblk)
            if p1 >= nblk:
                break
            tiles = sum(max(0, (2 * bi + 2) - 2 * p2) for bi in range(p1, nblk))
            far_flops += tiles * 128 * 64 * (p1 - p0) * 128 * 2.0
        kbytes = 8.0 * N * N + 8.0 * N * D
        side = {"kmat_assembly": {"ms": fit_ms["kmat"], "achieved_gbs": kbytes / (fit_ms["kmat"] * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                                  "frac": kbytes / (fit_ms["kmat"] * 1e-3) * 1e-9 / hbm_peak, "algorithmic_bytes": kbytes},
                "cholesky": {"ms": fit_ms["chol"], "tflops_fp64": (N ** 3 / 3.0) / (fit_ms["chol"] * 1e-3) * 1e-12,
                             "syrk_k512_ms": fit_ms["syrk"], "syrk_k512_flops": far_flops,
                             "syrk_k512_tflops_fp64": far_flops / max(fit_ms["syrk"], 1e-9) * 1e-9, "peak_tflops_fp64": peak_fp64,
                             "syrk_k512_engine": "tcgen05.mma kind::i8, 7-slice error-free product, TMEM accumulators (csrc/syrk_i8.cu); "
                                                 "timed on the second stream while the next outer panel runs"},
                "alpha_ms": fit_ms["alpha"]}
        if "error" in tgt:
            side["target_n4096_d8_m262144"] = tgt
        elif tgt:
            blk = {}
            for kind, (wk, rk) in tgt.items():
                b = line_for(wk, rk, 3, world, peaks, "target")
                b["config"] = workload_config(wk, world)
                b["scaling"] = kind
                b["total_candidates"] = wk["M"] * world
                blk[kind] = b
            side["target_n4096_d8_m262144"] = blk
        if world == 1 and not args.no_side:
            # the north_star's fit targets are quoted at N=4096, D=8 (SEArd): measure that fit beside the workload's own (outside every
            # timed region; best of three warm refits, library CUDA-event timers around K1 / the factorisation / the K=512 updates)
            try:
                rng4 = np.random.default_rng(4)
                X4 = rng4.random((8, 4096)); y4 = np.sin(3 * X4.sum(0)) + 0.1 * rng4.standard_normal(4096)
                m4 = b200bo.B200GPE(8, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(8, np.log(np.sqrt(8) * 0.25)), 0.0), logNoise=-2.0,
                                    capacity=4096, device=local_rank)
                m4.set_knob("chol_graph", 2)       # capture the factorisation graph at the second fit (default: a shape's sixth consecutive one)
                t4 = []
                for _ in range(4):
                    m4.fit(X4, y4)
                    t4.append([m4.timing_ms(v) for v in (_lib.T_KMAT, _lib.T_CHOL, _lib.T_SYRK, _lib.T_ALPHA)])
                k4, c4, s4, a4 = (min(t[i] for t in t4[1:]) for i in range(4))
                kb4 = 8.0 * 4096 * 4096 + 8.0 * 4096 * 8
                far4 = 0.0
                for p0 in range(0, 32, 4):
                    p1, p2 = min(p0 + 4, 32), min(p0 + 8, 32)
                    if p1 >= 32:
                        break
                    far4 += sum(max(0, (2 * bi + 2) - 2 * p2) for bi in range(p1, 32)) * 128 * 64 * (p1 - p0) * 128 * 2.0
                side["fit_n4096_d8"] = {"kmat_ms": k4, "kmat_gbs": kb4 / (k4 * 1e-3) * 1e-9, "kmat_frac_of_hbm": kb4 / (k4 * 1e-3) * 1e-9 / hbm_peak,
                                        "kmat_algorithmic_bytes": kb4, "cholesky_ms": c4, "syrk_k512_ms": s4,
                                        "syrk_k512_tflops_fp64": far4 / max(s4, 1e-9) * 1e-9,
                                        "syrk_k512_int8_tops": far4 * 28.0 / max(s4, 1e-9) * 1e-9, "int8_peak_tops": peaks["peak_i8"],
                                        "alpha_ms": a4, "fit_total_ms": k4 + c4 + a4}
                # cfg4 of BASELINE.json: the MAP sweep (64 settings, mll + gradient) on the same N=4096, D=8 model
                grid = [(lnz, l) for lnz in np.linspace(-3.0, 0.0, 8) for l in np.linspace(-1.5, 0.5, 8)]
                Theta = np.stack([np.concatenate([[lnz, 0.0], np.full(8, l), [0.0]]) for lnz, l in grid], axis=1)
                m4.mll_sweep(Theta[:, :8])
                t0 = time.perf_counter(); m4.mll_sweep(Theta); t_sw = time.perf_counter() - t0
                t0 = time.perf_counter(); m4.mll_sweep(Theta, want_grad=False); t_sv = time.perf_counter() - t0
                side["cfg4_map_sweep"] = {"workload": "BASELINE configs[3]: N=4096, D=8, 64 (logNoise, length-scale) settings, mll + dmll",
                                          "seconds_with_gradient": t_sw, "seconds_values_only": t_sv, "settings": 64}
                # the look-ahead schedule + CUDA graph against the in-order schedule of round 1 (same kernels), N=4096
                sched = {}
                for name, (sc, gr) in dict(in_order=(0, 0), look_ahead_eager=(1, 0), look_ahead_graph=(1, 2)).items():
                    m4.set_knob("chol_sched", sc); m4.set_knob("chol_graph", gr)
                    ts = []
                    for _ in range(4):
                        m4.fit(X4, y4); ts.append(m4.timing_ms(_lib.T_CHOL))
                    sched[name + "_ms"] = min(ts[1:])
                m4.set_knob("chol_sched", -1); m4.set_knob("chol_graph", -1)
                side["fit_n4096_d8"]["cholesky_schedules"] = sched
                # the reference's derivative-free search for ThompsonSamplingSimple (GN_DIRECT_L, maxeval = 2000) and the joint
                # posterior sample of myrand(model, X::Matrix), both inside the library
                m4.fit(X4, y4)
                lb4, ub4 = np.zeros(8), np.ones(8)
                m4.acquire_direct("TS", (), lb4, ub4, maxeval=2000)
                t0 = time.perf_counter(); rd = m4.acquire_direct("TS", (), lb4, ub4, maxeval=2000, seed=1); t_d = time.perf_counter() - t0
                Xj = rng4.random((8, 1024))
                m4.rand_joint(Xj)
                t0 = time.perf_counter(); rj = m4.rand_joint(Xj, seed=1); t_j = time.perf_counter() - t0
                side["direct_l_search"] = {"workload": "N=4096, D=8, ThompsonSamplingSimple, maxeval=2000 (src/acquisition.jl:7-9)", "seconds": t_d,
                                           "evaluations": rd["evals"], "device_launch_batches": rd["batches"]}
                side["joint_posterior_sample"] = {"workload": "N=4096, D=8, one joint draw over M=1024 points (src/models/gp.jl:7)", "seconds": t_j,
                                                  "make_posdef_tries": rj["tries"]}
                del m4
            except Exception as exc:                                   # a side metric must never take the bench line down
                side["fit_n4096_d8"] = {"error": str(exc)}
            # the other single-GPU configurations of BASELINE.json, each with its own roofline (3 timed steps)
            for name in ("cfg3", "cfg5", "cfg2"):
                if name == args.workload:
                    continue
                try:
                    w2 = WORKLOADS[name]
                    r2 = run_gpu_workload(ctx, w2, 3, 3)
                    blk = line_for(w2, r2, 3, 1, peaks, name)
                    blk["config"] = workload_config(w2, 1)
                    side["workload_" + name] = blk
                    del r2
                except Exception as exc:
                    side["workload_" + name] = {"error": str(exc)}
        out = {"metric": METRIC, "value": ln["value"], "unit": "candidates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ln["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": workload_config(w, world), "e2e": ln["e2e"], "gpu_launches": ln["gpu_launches"],
               "roofline": ln["roofline"], "side_metrics": side, "clocks": clk.summary(), "best": ln["best"]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(w, r["X"], r["y"], r["ll"], r["par"])
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(w, X, y, ll, par):
    """the CPU restatement timed on this box's host cores on bounded samples: C1 reference-shaped per-candidate loop
    (all threads and one thread), C2 best-effort batched dtrsm, C3 fit, C4 one MAP objective + gradient evaluation."""
    import scipy.linalg as sl
    from oracle import gp_oracle as orc
    from oracle.c_oracle import COracle
    t0 = time.perf_counter()
    gp = orc.GPOracle(w["D"], w["kernel"], "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    t_fit = time.perf_counter() - t0
    co = COracle(gp)
    Xs = candidates(w, 0)
    nthr = os.cpu_count() or 1
    t0 = time.perf_counter(); r = co.acquire(w["acq"], tuple(par), Xs[:, :256], want_grad=w["grad"], nthreads=nthr); pilot = (time.perf_counter() - t0) / 256
    sample = int(max(256, min(w["M"], 10.0 / pilot)))
    t0 = time.perf_counter(); r = co.acquire(w["acq"], tuple(par), Xs[:, :sample], want_grad=w["grad"], nthreads=nthr); dt = time.perf_counter() - t0
    s1 = int(max(64, min(sample, 3.0 / (pilot * nthr))))
    t0 = time.perf_counter(); co.acquire(w["acq"], tuple(par), Xs[:, :s1], want_grad=w["grad"], nthreads=1); dt1 = time.perf_counter() - t0
    # C2: the same math batched -- k* block, ONE dtrsm (OpenBLAS, all threads), vectorised epilogue
    B = 4096
    L = np.ascontiguousarray(gp.U.T)
    t0 = time.perf_counter(); nb = 0
    while time.perf_counter() - t0 < 4.0 and (nb + 1) * B <= w["M"]:
        Ks = gp.cov(gp.X, Xs[:, nb * B:(nb + 1) * B])
        v = sl.solve_triangular(L, Ks, lower=True, check_finite=False)
        mu = gp.beta + Ks.T @ gp.alpha; s2 = np.maximum(gp.sf2 - np.einsum("ij,ij->j", v, v), 0.0)
        orc.acq_value(w["acq"], tuple(par), mu, s2, eps=np.zeros(B) if w["acq"] == "TS" else None)
        nb += 1
    dt2 = time.perf_counter() - t0
    # C4: one evaluation of the MAP objective and its gradient at the workload's own size
    t0 = time.perf_counter(); gp.mll_dmll(gp.get_params()); t_map = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "candidates/s", "cores": int(r["threads"]), "kind": "port",
            "sample": f"C1: first {sample} of {w['M']} LHS candidates, per-candidate predict_f loop (dtrsv-shaped) + functor, oracle/oracle.c, "
                      f"OpenMP over candidates; host has {os.cpu_count()} logical CPUs; {dt:.1f} s",
            "c1_one_thread": {"value": s1 / dt1, "cores": 1, "sample": f"{s1} candidates, {dt1:.1f} s"},
            "c2_batched_dtrsm": {"value": nb * B / dt2, "cores": nthr, "blas": "NumPy/SciPy OpenBLAS", "gradient": False,
                                 "sample": f"{nb} blocks of {B} candidates: k* block, one dtrsm, vectorised functor; {dt2:.1f} s"},
            "c3_fit_seconds": {"value": t_fit, "what": f"assembly + dpotrf + alpha + mll at N={w['N']} (oracle/gp_oracle.py, LAPACK)"},
            "c4_map_eval_seconds": {"value": t_map, "what": f"mll + dmll for one theta at N={w['N']}, D={w['D']} (dpotrf, inverse, D+2 traces)"}}


if __name__ == "__main__":
    main()
