#!/usr/bin/env python
"""bench.py -- candidates scored per second on the acquisition hot path.

Workload (BASELINE.json configs[4], SURVEY.md section 8d "C5"): n_train = 1024, d = 10,
2 objectives (ZDT1 values), injected hyper-parameters (sigma_f2 = (1, 2), ell = 0.7 / 0.8),
full GP posterior (mean + std for BOTH objectives) + 2-objective EHVI + arg-max, over a
counter-generated uniform candidate pool of m candidates per GPU per step (weak scaling).

  value  : candidates/s, whole job, inputs resident on the device (the pool is generated
           on-chip from a counter RNG: 0 input bytes), CUDA-event time per step, max over ranks.
  e2e    : same metric through the host entry of the C ABI (ombo_score_host): the candidate
           pool lives in PINNED HOST memory, is copied host->device inside the timed region
           (chunked, overlapped with scoring) and the 16-byte result is read back.
  roofline: the dominant kernel (posterior: K1+K2) timed with CUDA events on its stream
           inside the timed region (C-ABI profile hooks), algorithmic FLOPs/candidate from
           SURVEY section 8d.
  cpu_baseline: the numpy/LAPACK oracle port (oracle/oracle.py) on the host cores on a
           bounded sample of the same workload.
  --impl reference : the same oracle port as the measured arm (the reference is pure Python,
           cannot travel to the GPU box, and scores one candidate per call; the batched port
           is its best case).  Rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM, N_OBJ = 1024, 10, 2


def zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.column_stack([f1, g * (1 - np.sqrt(f1 / g))])


def workload(n=N_TRAIN, d=DIM):
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    Y = zdt1(X)
    ells = [0.7 * np.ones(d), 0.8 * np.ones(d)]
    sf2 = [1.0, 2.0]
    return X, Y, ells, sf2


def algorithmic_flops_per_candidate(n, d, G, G_var, P):
    """F = G_var n^2 + G (2n + n(3d+10)) + A, A(EHVI-2D) ~ 30 P."""
    return G_var * n * n + G * (2 * n + n * (3 * d + 10)) + 30 * P


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_ev = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def summary(self):
        self._stop_ev.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                              ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "reasons": sorted(reasons)}


def cpu_oracle_rate(budget_s, chunk=4096, semantics="exact"):
    """Batched FP64 posterior + vectorised EHVI of the oracle port on the host cores."""
    from oracle import oracle as O
    X, Y, ells, sf2 = workload()
    st = [O.gp_fit_state(X, Y[:, i], ells[i], sf2[i]) for i in range(N_OBJ)]
    PF, r = O.calc_pf(Y), Y.max(0)
    from scipy.stats import norm, qmc
    cache = norm.ppf(qmc.Sobol(d=2, scramble=True, seed=0).random_base2(m=5))
    lo, hi = np.zeros(DIM), np.ones(DIM)
    done, t0 = 0, time.perf_counter()
    best = (-np.inf, -1)
    while True:
        Xc = O.candidates_from_counter(1, done, chunk, lo, hi)
        post = [O.gp_posterior(s, Xc) for s in st]
        a = O.ehvi_batched(post[0][0], post[1][0], post[0][1], post[1][1], PF, r, cache, semantics)
        i = O.argmax_lowest_index(a)
        if a[i] > best[0]:
            best = (float(a[i]), done + i)
        done += chunk
        el = time.perf_counter() - t0
        if el >= budget_s and done >= 2 * chunk:
            break
    return done / el, done, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    t_budget = 6.0
    times, total = [], 0
    for it in range(args.warmup + args.steps):
        rate, done, el = cpu_oracle_rate(t_budget if it >= args.warmup else 2.0)
        if it >= args.warmup:
            times.append(el)
            total += done
    value = total / sum(times)
    line = {
        "impl": "reference", "metric": "candidates scored/sec (GP mean+std+EHVI)", "value": value,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5: n_train=1024 d=10 k=2 EHVI-2D (exact semantics), counter-generated pool",
                   "sample_per_step": total // max(1, len(times))},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{total} candidates in {len(times)} steps, batched numpy/LAPACK oracle "
                                   "(4096-row chunks), all BLAS threads"},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default=os.environ.get("OMBO_BENCH_PRECISION", "auto"))
    ap.add_argument("--log2m", type=int, default=0, help="candidates per GPU per step = 2**log2m")
    ap.add_argument("--semantics", default="exact")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="skip the (untimed, separately reported) device hyper-parameter fit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import optimobo_b200 as ob
    from optimobo_b200 import _cabi
    from optimobo_b200.distributed import score_sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    precision = args.precision
    if precision == "auto":
        precision = "fast" if _cabi.fast_path_available() else "fp64"
    log2m = args.log2m or (24 if precision == "fast" else 20)
    m_per_gpu = 1 << log2m
    m_total = m_per_gpu * world

    X, Y, ells, sf2 = workload()
    t0 = time.perf_counter()
    models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=dev) for i in range(N_OBJ)]
    torch.cuda.synchronize()
    refresh_first_ms = 1e3 * (time.perf_counter() - t0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for mdl in models:
        mdl.refresh()
    ev1.record()
    torch.cuda.synchronize()
    refresh_ms = ev0.elapsed_time(ev1)

    # hyper-parameter fit on the device (section 8f-1): reported separately, not part of the metric
    fit_ms = None
    if rank == 0 and not args.no_fit:
        from optimobo_b200.fit import fit_hyperparameters_device
        t0 = time.perf_counter()
        fit_hyperparameters_device(X, Y[:, 0], max_f_eval=40, device=dev)
        fit_ms = 1e3 * (time.perf_counter() - t0)

    # first front + hypervolume of the evaluated sample (section 8f-2), device vs host; reported, not in the metric
    prep_ms = None
    if rank == 0:
        ob.device_prep.calc_pf(Y, dev)                                   # warm-up
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(10):
            pf_d = ob.device_prep.calc_pf(Y, dev)
            hv_d = ob.device_prep.hypervolume(Y, Y.max(0), dev)
        t1 = time.perf_counter()
        pf_h, hv_h = ob.host_prep.calc_pf(Y), ob.host_prep.hypervolume(Y, Y.max(0))
        t2 = time.perf_counter()
        assert np.array_equal(pf_d, pf_h) and hv_d == hv_h
        prep_ms = {"device": 1e2 * (t1 - t0), "host": 1e3 * (t2 - t1)}

    cache = ob.host_prep.cached_samples(2, 5, seed=0)
    PF, r = ob.host_prep.calc_pf(Y), Y.max(0)
    spec = ob.spec_ehvi(r, PF, cache, args.semantics)
    pool = ob.CandidatePool.counter(m_total, np.zeros(DIM), np.ones(DIM), seed=1)
    ctx = _cabi.Context.get(local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        return score_sharded(models, spec, pool, precision=precision, rescoring=(world > 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.launch_count(reset=True)
    ctx.profile(True)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    result = None
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the events)
        starts[i].record()
        result = step()
        ends[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    n_prof, prof_ms = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count(reset=True)
    clocks = sampler.summary() if sampler else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    t = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = m_total * args.steps / (total_ms / 1e3)

    # ---- end-to-end: host candidates through the C-ABI host entry ------------------------------
    e2e = None
    if not args.no_e2e:
        host_dtype = torch.float32 if precision == "fast" else torch.float64
        gen = torch.Generator().manual_seed(1234 + rank)
        Xh = torch.rand((m_per_gpu, DIM), dtype=host_dtype, generator=gen).pin_memory()
        for _ in range(2):
            ob.propose_host(models, spec, Xh, precision=precision, index_base=rank * m_per_gpu)
        barrier()
        k_e2e = max(2, min(args.steps, 5))
        w0 = time.perf_counter()
        es = torch.cuda.Event(enable_timing=True); ee = torch.cuda.Event(enable_timing=True)
        es.record()
        for _ in range(k_e2e):
            bv, bi = ob.propose_host(models, spec, Xh, precision=precision, index_base=rank * m_per_gpu)
        ee.record()
        barrier()
        te = torch.tensor([es.elapsed_time(ee)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": m_total * k_e2e / (float(te.item()) / 1e3), "unit": "candidates/s",
               "h2d_bytes_per_step": int(Xh.numel() * Xh.element_size()), "d2h_bytes_per_step": 16,
               "steps": k_e2e, "host_dtype": str(host_dtype).replace("torch.", ""),
               "wall_s": time.perf_counter() - w0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------
    n, d, P = N_TRAIN, DIM, len(PF)
    fl_cand = algorithmic_flops_per_candidate(n, d, N_OBJ, N_OBJ, P)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if precision == "fast":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")
    else:
        # no FP64 peak in MEASURED_PEAKS.json: measure cuBLAS DGEMM here
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        best = 1e9
        for _ in range(6):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(); torch.matmul(a, b); e_.record(); torch.cuda.synchronize()
            best = min(best, s_.elapsed_time(e_))
        peak = 2 * 4096 ** 3 / (best / 1e3) / 1e12
        peak_src = "cuBLAS DGEMM 4096^3 measured in this run (best of 6)"
    launch_ms = prof_ms / max(1, n_prof)
    # one posterior launch covers ONE GP for one chunk of <= 2^20 candidates: its algorithmic share is
    # (F - A) / G per candidate (the acquisition term A belongs to k_acquire)
    cand_per_launch = m_per_gpu * args.steps * N_OBJ / max(1, n_prof)
    fl_launch = (fl_cand - 30 * P) / N_OBJ * cand_per_launch
    achieved = fl_launch / (launch_ms / 1e3) / 1e12 if n_prof else None
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the same size (ncu --set full capture)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[precision]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "k_posterior_" + ("fast" if precision == "fast" else "fp64"),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src, "launches": n_prof, "avg_launch_ms": launch_ms,
                "candidates_per_launch": cand_per_launch,
                "kernel_share_of_step": prof_ms / total_ms if total_ms else None,
                "flops_per_candidate": fl_cand}
    if precision == "fast" and achieved:
        # what the tensor cores actually execute: 3 16-bit products per split-precision MAC, and the 128-wide
        # diagonal blocks of the triangular factor are multiplied densely (68 of 64 ideal units at n = 1024)
        nch = N_TRAIN // 128
        units = sum(2 * c + 2 for c in range(nch)) - 0.5 * nch
        issued = 3 * units * (128 * 64) * 2              # FLOP per candidate per GP: units of 128 columns x 64 K
        issued_tflops = issued * cand_per_launch / (launch_ms / 1e3) / 1e12
        roofline["issued"] = {"tflops": issued_tflops, "frac": issued_tflops / peak,
                              "note": "bf16x3 split (fp16x3 planes for ill-conditioned GPs): 3 tensor-core products per algorithmic MAC, "
                                      "diagonal blocks dense; this is the tensor-pipe work the kernel sustains"}

    cpu = None
    if not args.no_cpu_baseline:
        rate, done, el = cpu_oracle_rate(args.cpu_seconds, semantics=args.semantics)
        cpu = {"value": rate, "unit": "candidates/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{done} candidates of the same pool in {el:.1f} s: batched numpy/LAPACK oracle "
                         "(FP64 posterior x2 + vectorised EHVI, 4096-row chunks), all BLAS threads"}

    line = {
        "metric": "candidates scored/sec (GP mean+std+EHVI)", "value": value, "unit": "candidates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3" if precision == "fast" else "f64", "data": "synthetic",
        "config": {"workload": f"C5: n_train={n} d={d} k=2, full posterior + EHVI-2D ({args.semantics} semantics) + arg-max",
                   "candidates_per_gpu_per_step": m_per_gpu, "precision": precision,
                   "l2": "256 MB flush between timed iterations",
                   "pool": "counter-generated on device (value) / pinned host buffer (e2e)",
                   "parallelism": f"dp{world} (pool sharded, GP state replicated)"},
        "ms_per_bo_iter": {"gp_refresh_x2": refresh_ms, "score_and_reduce": total_ms / args.steps,
                           "first_refresh_incl_init": refresh_first_ms,
                           "hyperparameter_fit_40_evals_one_gp": fit_ms,
                           "first_front_and_hypervolume_n1024": prep_ms},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "best": {"value": result[0], "index": result[1]}, "wall_s_timed_region": wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
