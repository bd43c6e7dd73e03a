#!/usr/bin/env python
"""bench.py -- KL-NMF iterations/sec on the BASELINE.json workload (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode tf32r|tf32|tf32x3|fp64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one reference iteration (objective + ratio + coefficient update + dictionary
update, nmf.py:212-222) over the whole workload.  Default workload = BASELINE.json configs[4]
(cfg5: dense fit, n=4,000,000 x f=8192, k=512), STRONG scaling: the same 4M samples are
row-sharded over the N ranks; the only exchange is the all-reduce of the k x f numerator.

Prints ONE JSON line on rank 0.  `value` = iterations/s with X resident in HBM (device time,
CUDA events on the engine's stream, max over ranks); `e2e` = the same metric through the
public API (KLdivNMF.fit_transform on a pinned HOST array: H2D of X and D2H of W inside the
timed region); `roofline` = the dominant contraction kernel against the measured TF32 peak;
`cpu_baseline` = the reference itself (baseline/_ref, installed by baseline/install_reference.sh; the float64 oracle
port when that directory is absent) on this box's host cores on a bounded row-subsample; `workloads` = the other two
large configs (cfg3 transform, cfg4 sparse fit) measured the same way in the same run.  The arithmetic mode is the
library's default (`_native.DEFAULT_MODE`); `alt_modes` carries the others.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_total, f, k, kind, description)
    "cfg5": (4_000_000, 8192, 512, "dense_fit", "cfg5 dense fit n=4000000 f=8192 k=512 (BASELINE.json configs[4])"),
    "cfg3": (1_000_000, 4096, 256, "dense_transform", "cfg3 transform n=1000000 f=4096 k=256 (BASELINE.json configs[2])"),
    "cfg4": (2_000_000, 50_000, 256, "sparse_fit", "cfg4 sparse fit n=2000000 f=50000 density=0.005 k=256 (BASELINE.json configs[3])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("KLNMF_BENCH_MODE", ""), help="default: the library's default mode")
    ap.add_argument("--workload", default=os.environ.get("KLNMF_BENCH_WORKLOAD", "cfg5"), choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the total sample count (development only)")
    ap.add_argument("--k", type=int, default=0, help="override the number of components (development only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--alt-mode", default=os.environ.get("KLNMF_BENCH_ALT", "tf32,tf32x3"),
                    help="further arithmetic modes reported under alt_modes, comma separated ('' to skip)")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg3 / cfg4 / mixed-stack entries under `workloads`")
    ap.add_argument("--balance", action="store_true",
                    help="several ranks: size the shards by each GPU's measured samples/s instead of the equal split "
                         "(experimental: under a power cap a GPU's speed depends on how long it idles in the all-reduce, "
                         "and the calibration overshoots -- profiles/r2_run9.log)")
    ap.add_argument("--e2e-iters", type=int, default=50,
                    help="second end-to-end call at the reference's iteration count (experiment.py:20-25); 0 to skip")
    args = ap.parse_args()
    if not args.mode:
        from multimodal_b200 import _native
        args.mode = _native.DEFAULT_MODE
    return args


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the float64 numpy oracle on host cores, bounded row-subsample
# ------------------------------------------------------------------------------------------------
def load_reference():
    """The UNMODIFIED reference estimator from baseline/_ref (baseline/install_reference.sh), or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "multimodal")):
        return None
    np.Inf = np.inf                    # numpy >= 2: nmf.py:206 says np.Inf
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        from multimodal.lib.nmf import KLdivNMF
        return KLdivNMF
    except Exception:
        return None


def cpu_sample(workload, n_total, f, k, kind, steps, warmup):
    from oracle import klnmf_oracle as O
    import scipy.sparse as sp
    # torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm is meant to use every host core it can
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count(), user_api="blas")
    except Exception:
        pass
    rs = np.random.RandomState(0)
    if kind == "sparse_fit":
        n_cpu = 4096
        X = sp.random(n_cpu, f, density=0.005, random_state=rs, format="csr")
        X.data = 1.0 - X.data
        sample = "sparse n_cpu=%d of n=%d rows, f=%d, density 0.005, k=%d" % (n_cpu, n_total, f, k)
    else:
        n_cpu = 16384 if k >= 512 else 32768      # about 3 s of float64 numpy work per iteration on 16 cores
        X = rs.random_sample((n_cpu, f))
        sample = "dense n_cpu=%d of n=%d rows, f=%d, k=%d" % (n_cpu, n_total, f, k)
    np.random.seed(0)
    H = O.init_dictionary(k, f)
    fit = kind != "dense_transform"
    Ref = load_reference()
    if Ref is not None:
        # the reference's own public API and stock code path: KLdivNMF(...).fit_transform / transform, tol = 0 as the
        # learner passes it (learner.py:12,39); one call of `warmup` iterations, then one timed call of `steps`
        def call(iters):
            est = Ref(n_components=k, max_iter=iters, tol=0)
            est._init_dictionary = H
            est.components_ = H
            t0 = time.perf_counter()
            W, errs = est.fit_transform(X, _fit=fit, return_errors=True)
            dt = time.perf_counter() - t0
            assert len(errs) == iters, "the reference stopped early (%d of %d iterations)" % (len(errs), iters)
            return dt
        if warmup > 0:
            call(warmup)
        t_step = call(steps) / steps
        times = [t_step] * steps
        kind_name = "reference"
    else:
        W = np.asarray(X.dot(H.T))
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.error(X, W, H)                      # the reference computes the objective every iteration (nmf.py:214)
            W, H = O.update(X, W, H, fit=fit)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        t_step = float(np.mean(times))
        kind_name = "port"
    # cost is exactly linear in n at fixed f, k: extrapolate to the full workload
    value = (1.0 / t_step) * (n_cpu / float(n_total))
    cores = os.cpu_count()
    try:
        from threadpoolctl import threadpool_info
        nth = [i.get("num_threads") for i in threadpool_info() if i.get("user_api") == "blas"]
        if nth:
            cores = max(nth)
    except Exception:
        pass
    return {"value": value, "unit": "iterations/s", "cores": cores, "kind": kind_name,
            "sample": sample + ", %d timed iterations, %.2f s each, extrapolated linearly in n" % (len(times), t_step) +
            ("; the unmodified reference KLdivNMF from baseline/_ref, one fit_transform call (W0 = X.H0^T included)"
             if kind_name == "reference" else "; float64 numpy port (oracle/klnmf_oracle.py): baseline/_ref is absent")}, t_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total, f, k, kind, desc = WORKLOADS[args.workload]
    if args.n:
        n_total = args.n
    steps, warmup = max(1, min(args.steps, 8)), max(0, min(args.warmup, 2))
    base, t_step = cpu_sample(args.workload, n_total, f, k, kind, steps, warmup)
    line = {"impl": "reference", "metric": "KL-NMF iterations/sec", "value": base["value"], "unit": "iterations/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / base["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n": n_total, "f": f, "k": k, "note":
                       ("the unmodified reference (baseline/_ref: multimodal.lib.nmf.KLdivNMF, float64 numpy/scipy)"
                        if base["kind"] == "reference" else "float64 numpy restatement of the reference "
                        "(oracle/klnmf_oracle.py)") + " on host cores; bounded row-subsample extrapolated in n"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def allreduce_max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize(local)


def timed_run(eng, steps, warmup, fit, dist, local):
    """W untimed iterations, then EXACTLY `steps` timed ones, barrier + sync on both sides."""
    ninf = -float("inf")
    if warmup > 0:
        eng.run(warmup, ninf, fit)
    barrier(dist, local)
    c0 = eng.counters()
    errs, _ = eng.run(steps, ninf, fit)
    barrier(dist, local)
    c1 = eng.counters()
    ms, cnt = eng.last_run_profile()
    assert len(errs) == steps, "the timed region did not run %d iterations" % steps
    total_ms = allreduce_max(dist, local, ms["total"])
    return total_ms, ms, cnt, c1["launches"] - c0["launches"], errs


def make_engine(_native, n_local, f, k, kind, mode, local, rank, world, H0, scratch):
    eng = _native.Engine(n_local, f, k, mode=mode, device=local, scratch_limit=scratch)
    if kind == "sparse_fit":
        eng.fill_csr_synthetic(int(round(f * 0.005)), 1234 + rank)
    else:
        eng.fill_dense_synthetic(1234 + rank)
    if world > 1:
        from multimodal_b200 import distributed as D
        eng.comm_attach(D.rank_comm(local))
    eng.set_dictionary(H0)
    eng.init_coefficients()
    return eng


def ncu_traffic(kernel_key, rows_per_launch, f, k, mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the committed
    `ncu --set full` summary (profiles/ncu_traffic.json, written by tools/ncu_summarise.py). Only a capture of
    the SAME launch shape counts; anything else is reported as null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        best = None
        for e in json.load(open(p)):
            if e["kernel"] == kernel_key and e["f"] == f and e["k"] == k and e["mode"] == mode:
                if best is None or abs(e["rows"] - rows_per_launch) < abs(best["rows"] - rows_per_launch):
                    best = e
        if best is not None:
            # every byte of these kernels scales with the rows of the launch (X, Q, W panels; the dictionary is
            # 16.8 MB): a capture at the panel size is scaled to the average launch of this run
            return best["dram_bytes"] * rows_per_launch / float(best["rows"])
    except Exception:
        pass
    return None


def phase_roofline(ms, cnt, n_local, f, k, kind, steps, peaks, mode):
    if kind == "sparse_fit":
        # compulsory bytes of one sparse iteration (CSR once, W r+w, H r + numerator w)
        nnz = n_local * int(round(f * 0.005))
        alg = 8.0 * nnz + 4.0 * (n_local + 1) + 8.0 * n_local * k + 8.0 * k * f
        t = (ms["ratio"] + ms["numerator"]) / max(steps, 1) * 1e-3
        ach = alg / t / 1e9
        # what really bounds the path: every stored entry gathers one dictionary column (rows pass) and one coefficient
        # row (numerator pass) of 4 k bytes from the L2 -- against the L2 -> SM read bandwidth measured on this GPU
        gather = 2.0 * 4.0 * k * nnz
        l2 = None
        if peaks.get("l2_gbs"):
            l2 = {"bound": "l2", "achieved": gather / t / 1e9, "peak": peaks["l2_gbs"], "unit": "GB/s",
                  "frac": gather / t / 1e9 / peaks["l2_gbs"], "gather_bytes": gather,
                  "peak_source": "klnmf_l2_read_bench on this GPU: best of a 24 MB and a 48 MB buffer read 200 times by every SM, eight 512-byte pieces in flight per warp (ld.global.cg.v4)",
                  "peak_by_buffer_mb": peaks.get("l2_gbs_by_mb"),
                  "rows_pass_gbs": 0.5 * gather / (ms["ratio"] / max(steps, 1) * 1e-3) / 1e9,
                  "numerator_pass_gbs": 0.5 * gather / (ms["numerator"] / max(steps, 1) * 1e-3) / 1e9}
        return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "l2_roofline": l2,
                "traffic": ncu_traffic("sparse", n_local, f, k, "tf32" if mode != "fp64" else "fp64"),
                "algorithmic_bytes": alg, "kernel": "sparse_rows_full_kernel + sparse_numerator_bcsc_kernel",
                "peak_source": peaks["source"],
                "phase_ms_per_step": {p: ms[p] / max(steps, 1) for p in ("ratio", "numerator", "dictionary", "allreduce")},
                "note": "one launch = one iteration's rows pass + numerator pass; bound in practice by the L2 gather "
                        "traffic of 2 x nnz x 4k bytes (DESIGN.md 4.4), not by HBM"}
    names = {"ratio": "tc_gemm_kernel<ratio: S=W.H, Q=(X+eps)/(S+eps), KL>",
             "coefficient": "tc_gemm_kernel<coefficient: W'=W(.)(Q.H^T)>",
             "numerator": "tc_gemm_kernel<numerator: N+=W'^T.Q>"}
    if mode == "fp64":
        names = {k_: v.replace("tc_gemm_kernel", "generic_gemm_kernel<double>") for k_, v in names.items()}
    fused = kind != "sparse_fit" and mode in ("tf32", "tf32r") and os.environ.get("KLNMF_FUSED", "1") != "0" and \
        (k <= 128 or (k <= 256 and os.environ.get("KLNMF_FUSED256", "1") != "0"))
    per_launch = {p_: 2.0 for p_ in names}                             # each contraction is 2 n k f per iteration
    if fused:
        # one kernel per row panel does both contractions of the coefficient half-step (dense_fused.cu, k <= 128; on
        # clusters of two CTAs for 128 < k <= 256, dense_fused256.cu): 4 n k f, timed under the "ratio" phase
        names["ratio"] = ("fused_coef_kernel" if k <= 128 else "fused_coef256_kernel") + \
            "<S=W.H -> Q=(X+eps)/(S+eps), KL -> G+=Q.H^T -> W'=W(.)G>"
        per_launch["ratio"] = 4.0
    dom = max(names, key=lambda p_: ms[p_])
    launches = max(cnt[dom], 1)
    flops_per_launch = per_launch[dom] * n_local * k * f * steps / launches
    avg_s = ms[dom] / launches * 1e-3
    ach = flops_per_launch / avg_s / 1e12
    peak = peaks["bf16_sustained"] / 2.0                                 # TF32 dense = half the BF16 rate
    if mode == "fp64":
        peak = 40.0
    rows_per_launch = int(round(n_local * steps / float(launches)))
    r = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
         "traffic": ncu_traffic(("fused" if k <= 128 else "fused256") if (fused and dom == "ratio") else dom, rows_per_launch, f, k, mode),
         "rows_per_launch": rows_per_launch,
         "kernel": names[dom], "avg_launch_ms": ms[dom] / launches, "launches": launches,
         "peak_source": ("TF32 = 1/2 x sustained cuBLAS bf16, " + peaks["source"]) if mode != "fp64" else "nominal B200 FP64",
         "phase_ms_per_step": {p: ms[p] / max(steps, 1) for p in ("ratio", "coefficient", "numerator", "dictionary", "allreduce")}}
    if mode == "tf32x3":
        r["issued_frac"] = 3.0 * ach / peak      # three TF32 MMAs are issued per algorithmic product
    # the whole iteration against the same peak: SURVEY 8d's algorithmic work, 6 n k f (fit) or 4 n k f (transform),
    # over the time of its contraction phases on this rank
    t_step = sum(ms[p_] for p_ in ("ratio", "coefficient", "numerator")) / max(steps, 1) * 1e-3
    if t_step > 0:
        alg = (4.0 if kind == "dense_transform" else 6.0) * n_local * k * f
        r["whole_step"] = {"achieved": alg / t_step / 1e12, "frac": alg / t_step / 1e12 / peak, "unit": "TFLOP/s",
                           "algorithmic_flop": alg}
    return r


def modality_dims(f):
    """Widths of the three concatenated modalities of cfg5 (sound + image + motion): 4096 + 3072 + 1024 at f = 8192."""
    a, b = f // 2, (3 * f) // 8
    return [a, b, f - a - b]


def run_e2e(args, n_local, f, k, kind, mode, local, rank, world, dist, X_host, H0, steps, warm=True):
    """The public API on HOST data: H2D of X, `steps` iterations, D2H of the result, all timed.
    Dense fit on one GPU goes through MultimodalLearner.train on three modality blocks (column ranges of the pinned
    array, one coefficient each: the scaled concatenation is formed on the device) and reads the dictionary back;
    transforms and the sparse fit go through KLdivNMF and read the coefficients back."""
    from multimodal_b200.lib.nmf import KLdivNMF
    from multimodal_b200.learner import MultimodalLearner
    learner_path = world == 1 and kind == "dense_fit" and not hasattr(X_host, "nnz") and f >= 8
    dims = modality_dims(f)
    offs = [0, dims[0], dims[0] + dims[1], f]
    mods, coefs = ['sound', 'image', 'motion'], [1.0, 0.5, 2.0]

    def call(X, iters):
        if learner_path:
            lr = MultimodalLearner(mods, dims, coefs, k, mode=mode, device=local)
            np.random.seed(0)
            lr.train([X[:, offs[i]:offs[i + 1]] for i in range(3)], iters)
            return lr.dico
        est = KLdivNMF(n_components=k, max_iter=iters, tol=0, mode=mode, device=local)
        est._init_dictionary = H0
        if kind == "dense_transform":
            est.components_ = H0
            return est.transform(X)
        return est.fit_transform(X)

    def call_sharded(X, iters):
        from multimodal_b200 import distributed as D
        sh = D.ShardedNMF(k, max_iter=iters, tol=0, mode=mode, device=local)
        sh.components_ = H0
        if kind == "dense_fit":
            # the same contract as the one-GPU leg (MultimodalLearner.train -> KLdivNMF.fit): the dictionary comes back,
            # the coefficients of the shard stay on the device (nmf.py:259-273 throws them away)
            return sh.fit(X, X.shape[0] * world, H0=H0).components_
        return sh.fit_transform(X, X.shape[0] * world, H0=H0, fit=(kind != "dense_transform"))

    # one untimed warm-up call on a small slice: the process-wide pinned staging buffers, kernel attributes, the
    # allocator and (several ranks) the NCCL communicator are set up once per process, not once per call
    if warm:
        small = X_host[:min(X_host.shape[0], 65536)]
        if world == 1:
            call(small, 2)
        else:
            call_sharded(small, 2)
    barrier(dist, local)
    t0 = time.perf_counter()
    if world > 1:
        out = call_sharded(X_host, steps)
    else:
        out = call(X_host, steps)
    barrier(dist, local)
    dt = time.perf_counter() - t0
    dt = allreduce_max(dist, local, dt)
    h2d = X_host.nbytes if not hasattr(X_host, "nnz") else (X_host.data.nbytes + X_host.indices.nbytes + X_host.indptr.nbytes)
    if learner_path or (world > 1 and kind == "dense_fit"):
        d2h = out.nbytes                      # the trained dictionary; the fit keeps the coefficients on the device
    else:
        d2h = out.nbytes + (H0.nbytes if kind != "dense_transform" else 0)
    return dt, h2d, d2h, learner_path


def gather_phases(dist, local, ms, steps):
    """Per-phase ms per step as {phase: [min over ranks, max over ranks]} -- rank skew (compute) and wire time
    (all-reduce) separate in the record: a rank that waits in the all-reduce shows a short compute and a long wait."""
    names = ["ratio", "coefficient", "numerator", "dictionary", "allreduce", "total"]
    vals = [ms[n] / max(steps, 1) for n in names]
    if dist is None:
        return {n: [v, v] for n, v in zip(names, vals)}
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device="cuda:%d" % local)
    allv = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, t)
    m = torch.stack(allv).cpu().numpy()
    return {n: [float(m[:, i].min()), float(m[:, i].max())] for i, n in enumerate(names)}


def measure_workload(args, name, mode, ctx, with_clocks, keep_engine=False, steps=None, warmup=None):
    """One workload, one arithmetic mode: W warm-up iterations, K timed ones, device time, max over ranks."""
    _native, D, rank, world, local, dist, peaks = ctx
    n_total, f, k, kind, desc = WORKLOADS[name]
    if name == args.workload:
        if args.n:
            n_total = args.n
            desc += " [n overridden to %d]" % n_total
        if args.k:
            k = args.k
            desc += " [k overridden to %d]" % k
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    bounds = D.shard_bounds(n_total, world)
    n_local = bounds[rank + 1] - bounds[rank]
    fit = kind != "dense_transform"
    np.random.seed(0)
    H0 = np.abs(np.random.random((k, f))) + .01
    H0 = H0 / (1.e-16 + H0.sum(axis=1, keepdims=True))
    if dist is not None:
        H0 = D.broadcast_object(H0, 0)

    def build(rows):
        x_bytes = rows * f * (8 if mode == "fp64" else 4)
        # leave room for X + the W ping-pong (+ lo parts) on a 192 GB part
        scratch = (8 << 30) if x_bytes > (100 << 30) else (16 << 30)
        return make_engine(_native, rows, f, k, kind, mode, local, rank, world, H0, scratch)

    eng = build(n_local)
    balance = None
    if dist is not None and fit and args.balance:
        # Every fit iteration ends in an all-reduce that waits for the slowest shard, and the GPUs of one box differ by
        # several per cent under their power caps.  Calibrate on the equal split (untimed), then size the shards in
        # proportion to the measured samples per second of each GPU (distributed.shard_bounds(weights=...)).
        import torch
        balance = {"rounds": [], "note": "shards sized by each GPU's measured samples/s on untimed calibration runs "
                   "(4 discarded + 8 measured iterations per round, steady state under the power cap)"}
        for rnd in range(2):
            eng.run(4, -float("inf"), fit)
            eng.run(8, -float("inf"), fit)
            ms_cal, _ = eng.last_run_profile()
            t_local = sum(ms_cal[p_] for p_ in ("ratio", "coefficient", "numerator", "dictionary"))
            t = torch.tensor([n_local / max(t_local, 1e-6), t_local], dtype=torch.float64, device="cuda:%d" % local)
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            speeds = [float(v[0].item()) for v in allv]
            times = [float(v[1].item()) for v in allv]
            spread = (max(times) - min(times)) / max(times)
            balance["rounds"].append({"rows_per_rank": [bounds[r + 1] - bounds[r] for r in range(world)],
                                      "compute_ms_per_step": [x / 8.0 for x in times], "spread": spread})
            if spread < 0.01:
                break
            bounds = D.shard_bounds(n_total, world, weights=speeds)
            eng.close()
            n_local = bounds[rank + 1] - bounds[rank]
            eng = build(n_local)
        balance["rows_per_rank"] = [bounds[r + 1] - bounds[r] for r in range(world)]
    sampler = ClockSampler(local) if with_clocks else None
    if sampler:
        sampler.start()
    total_ms, ms, cnt, launches, errs = timed_run(eng, steps, warmup, fit, dist, local)
    clocks = sampler.stop() if sampler else None
    roof = phase_roofline(ms, cnt, n_local, f, k, kind, steps, peaks, mode)
    roof["phase_ms_per_step_min_max_over_ranks"] = gather_phases(dist, local, ms, steps)
    roof["shard_balance"] = balance
    out = {"value": steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
           "launches": int(launches), "roofline": roof, "clocks": clocks, "desc": desc, "n_total": n_total, "f": f, "k": k,
           "kind": kind, "n_local": n_local, "H0": H0, "errs": errs}
    if keep_engine:
        out["engine"] = eng
    else:
        eng.close()
    return out


def measure_mixed_stack(_native, device, steps, warmup, n=65536, fd=4096, fs=50000, per_row=250, k=256):
    """a wide dense modality next to a sparse one (learner.py:53-56).  The hybrid stack (dense block through
    the contraction engine, CSR block through the sparse passes) and the all-CSR stack the reference would
    build (array_utils.py:5-9), the same fit on both, device-timed like the other workloads."""
    import scipy.sparse as sp
    rs = np.random.RandomState(0)
    dense = rs.random_sample((n, fd)).astype(np.float32)
    dense[dense < 0.1] = 0.0                                    # 10 % structural zeros in the dense modality
    lo = (np.arange(per_row) * fs) // per_row
    width = ((np.arange(per_row) + 1) * fs) // per_row - lo
    idx = (lo[None, :] + (rs.random_sample((n, per_row)) * width[None, :]).astype(np.int64)).astype(np.int32)
    vals = (1.0 - rs.random_sample((n, per_row))).astype(np.float32)
    sparse = sp.csr_matrix((vals.ravel(), idx.ravel(), np.arange(n + 1, dtype=np.int64) * per_row), shape=(n, fs))
    f = fd + fs
    np.random.seed(1)
    H0 = np.abs(np.random.random((k, f))) + .01
    H0 = H0 / (1.e-16 + H0.sum(axis=1, keepdims=True))
    out = {"workload": "mixed stack fit: n=%d, dense modality %d columns + CSR modality %d columns (%d stored entries per row), k=%d"
                       % (n, fd, fs, per_row, k), "unit": "iterations/s", "steps": steps, "warmup": warmup}
    old = os.environ.get("KLNMF_HYBRID")
    try:
        for name, flag in (("all_csr_stack", "0"), ("hybrid_stack", "1")):
            os.environ["KLNMF_HYBRID"] = flag
            with _native.Engine(n, f, k, mode=_native.DEFAULT_MODE, device=device) as e:
                e.set_stacked_blocks([dense, sparse], [1.0, 1.0])
                assert e.is_hybrid() == (flag == "1")
                e.set_dictionary(H0)
                e.init_coefficients()
                e.run(warmup, 0.0, True)
                errs, _ = e.run(steps, 0.0, True)
                ms, _ = e.last_run_profile()
                out[name] = {"ms_per_step": ms["total"] / steps, "value": steps / (ms["total"] * 1e-3),
                             "objective_last": float(errs[-1])}
    finally:
        if old is None:
            os.environ.pop("KLNMF_HYBRID", None)
        else:
            os.environ["KLNMF_HYBRID"] = old
    out["hybrid_speedup"] = out["all_csr_stack"]["ms_per_step"] / out["hybrid_stack"]["ms_per_step"]
    return out


def run_ours(args):
    rank, world, local, dist = dist_setup(args)
    os.environ.setdefault("KLNMF_PROFILE", "1")
    from multimodal_b200 import _native
    from multimodal_b200 import distributed as D
    peaks = measured_peaks()
    try:
        # the roofline's denominator: the best of a 24 MB buffer (the W' block of the numerator pass) and a 48 MB one
        # (the 51 MB dictionary of the rows pass), read the way the passes gather (eight 512-byte pieces in flight per warp)
        peaks["l2_gbs_by_mb"] = {mb: _native.l2_read_bandwidth(bytes=mb << 20, device=local) for mb in (24, 48)}
        peaks["l2_gbs"] = max(peaks["l2_gbs_by_mb"].values())
    except Exception:
        peaks["l2_gbs"] = None
    ctx = (_native, D, rank, world, local, dist, peaks)

    main = measure_workload(args, args.workload, args.mode, ctx, True, keep_engine=True)
    eng = main.pop("engine")
    n_total, f, k, kind, n_local, H0, errs = (main[x] for x in ("n_total", "f", "k", "kind", "n_local", "H0", "errs"))
    fit = kind != "dense_transform"

    # ---- end-to-end leg through the public API on host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        x_bytes = n_local * f * 4
        avail = psutil.virtual_memory().available
        try:
            lim = open("/sys/fs/cgroup/memory.max").read().strip()
            if lim.isdigit():
                avail = min(avail, int(lim))
        except Exception:
            pass
        avail = avail / max(1, world)       # every rank of the box pins its own shard
        rows = n_local
        if kind != "sparse_fit" and x_bytes * 2.5 > avail:
            rows = max(1024, int(avail / 2.5 / (f * 4)) // 1024 * 1024)
        rows = min(rows, n_local)
        iters_list = [args.steps] + ([args.e2e_iters] if args.e2e_iters and args.e2e_iters != args.steps else [])
        if kind == "sparse_fit":
            # host scipy CSR (same stratified pattern as the device generator), pinned arrays; bounded row count,
            # scaled linearly in n (cost is exactly linear in the rows of a shard)
            import torch
            import scipy.sparse as sp
            eng.close()
            m = int(round(f * 0.005))
            rows = min(n_local, 262144)
            rs = np.random.RandomState(77 + rank)
            ind = torch.empty((rows * m,), dtype=torch.int32, pin_memory=True).numpy()
            val = torch.empty((rows * m,), dtype=torch.float32, pin_memory=True).numpy()
            ptr = torch.empty((rows + 1,), dtype=torch.int64, pin_memory=True).numpy()
            lo = (np.arange(m, dtype=np.int64) * f) // m
            width = ((np.arange(m, dtype=np.int64) + 1) * f) // m - lo
            blk = 16384
            for r0 in range(0, rows, blk):
                r1 = min(rows, r0 + blk)
                u = rs.random_sample((r1 - r0, m))
                ind[r0 * m:r1 * m] = (lo[None, :] + (u * width[None, :]).astype(np.int64)).astype(np.int32).ravel()
                val[r0 * m:r1 * m] = (1.0 - rs.random_sample((r1 - r0) * m)).astype(np.float32)
            ptr[:] = np.arange(rows + 1, dtype=np.int64) * m
            Xh = sp.csr_matrix((val, ind, ptr), shape=(rows, f), copy=False)
            what = "KLdivNMF.fit_transform on a host scipy CSR matrix (pinned arrays)"
        else:
            import torch
            Xh = torch.empty((rows, f), dtype=torch.float32, pin_memory=True).numpy()
            if rows == n_local:
                eng.get_dense(Xh)
                eng.close()
            else:
                eng.close()
                with _native.Engine(rows, f, k, mode=args.mode, device=local) as e2:
                    e2.fill_dense_synthetic(1234 + rank)
                    e2.get_dense(Xh)
            what = None
        scale = rows / float(n_local)
        legs = []
        for it_n in iters_list:
            dt, h2d, d2h, via_learner = run_e2e(args, rows, f, k, kind, args.mode, local, rank, world, dist, Xh, H0, it_n,
                                                warm=(it_n == iters_list[0]))
            legs.append({"value": it_n / dt * scale, "unit": "iterations/s", "iterations": it_n,
                         "h2d_bytes_per_step": int(h2d / it_n), "d2h_bytes_per_step": int(d2h / it_n), "seconds": dt,
                         "rows_per_rank": rows})
        if what is None:
            what = ("MultimodalLearner.train (three modality blocks of %s columns scaled and concatenated on the "
                    "device; the dictionary is read back)" % "+".join(str(d) for d in modality_dims(f))) if via_learner \
                else ("KLdivNMF.transform" if not fit else "KLdivNMF.fit_transform") + " on a pinned host float32 array"
        e2e = legs[0]
        e2e["note"] = ("one %s call of %d iterations (after an untimed warm-up call on a 65536-row slice); X crosses PCIe "
                       "once per call, so per-step bytes are the call's bytes / steps" % (what, args.steps)) + \
            ("; %d ranks, each through distributed.ShardedNMF%s on its own pinned shard" % (world, ".fit (the dictionary is read back)" if kind == "dense_fit" else "") if world > 1 else "") + \
            ("" if rows == n_local else "; host RAM too small for the full shard: measured on %d rows and scaled "
             "linearly in n" % rows)
        if len(legs) > 1:
            e2e["at_reference_iterations"] = legs[1]     # the reference's 50 iterations (experiment.py:20-25)
        del Xh
    try:
        eng.close()
    except Exception:
        pass

    alt = {}
    for am in [m for m in (args.alt_mode or "").split(",") if m and m != args.mode]:
        if kind == "sparse_fit" and am != "fp64":
            continue                      # the sparse path computes in FP32 FMA in every TF32 mode
        r = measure_workload(args, args.workload, am, ctx, False)
        alt[am] = {"value": r["value"], "ms_per_step": r["ms_per_step"], "roofline_frac": r["roofline"]["frac"],
                   "roofline_achieved": r["roofline"]["achieved"], "kernel": r["roofline"]["kernel"],
                   "phase_ms_per_step": r["roofline"].get("phase_ms_per_step")}

    # ---- the other two large configs, driver-timed in the same run ----------------------------------------------
    extra = {}
    if not args.no_extra and args.workload == "cfg5" and not args.n and not args.k:
        for name in ("cfg3", "cfg4"):
            r = measure_workload(args, name, args.mode, ctx, True, steps=max(args.steps, 5), warmup=max(args.warmup, 3))
            extra[name] = {"workload": r["desc"], "value": r["value"], "unit": "iterations/s", "ms_per_step": r["ms_per_step"],
                           "steps": r["steps"], "warmup": r["warmup"], "rows_per_rank": r["n_local"], "mode": args.mode,
                           "gpu_launches": r["launches"], "roofline": r["roofline"], "clocks": r["clocks"]}

        if world == 1:
            try:
                extra["mixed_stack"] = measure_mixed_stack(_native, local, max(args.steps, 5), max(args.warmup, 3))
            except Exception as ex:                       # an extra: never takes the headline line down with it
                extra["mixed_stack"] = {"error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_sample(args.workload, n_total, f, k, kind, 4, 1)

    if rank == 0:
        line = {"metric": "KL-NMF iterations/sec", "value": main["value"], "unit": "iterations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.mode,
                "data": "synthetic",
                "config": {"workload": main["desc"], "n": n_total, "f": f, "k": k, "mode": args.mode,
                           "mode_is_library_default": args.mode == _native.DEFAULT_MODE,
                           "rows_per_rank": n_local, "l2": "inputs (%.1f GB per rank) are larger than the 126 MB L2"
                           % (n_local * f * 4 / 1e9), "objective_first_last": [float(errs[0]), float(errs[-1])]},
                "clocks": main["clocks"], "e2e": e2e, "gpu_launches": main["launches"], "roofline": main["roofline"],
                "cpu_baseline": cpu, "alt_modes": alt, "workloads": extra}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
