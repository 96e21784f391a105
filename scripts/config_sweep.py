"""Candidates/s of the BASELINE.json configurations C2-C5 at their full pool sizes, both precision modes where
they apply, on one GPU (device-resident counter pool, CUDA events, 3 warm-up + 5 timed passes, 256 MB L2 flush
between passes).  Prints one markdown table row per case."""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import optimobo_b200 as ob
import test_gpu_fullsize as T

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("| config | precision | candidates | ms per pass | candidates/s |\n|---|---|---|---|---|")
for name, make in T.CONFIGS.items():
    X, ys, ells, sf2, spec, lo, hi, m, precision, _ = make()
    models = [ob.GPModel(X, y, e, s, device=dev) for y, e, s in zip(ys, ells, sf2)]
    pool = ob.CandidatePool.counter(m, lo, hi, seed=1)
    n = len(X)
    for prec in (("fp64", "fast") if n >= 256 else ("fp64",)):
        mm = m if prec == "fast" or n <= 256 else min(m, 1 << 20)
        if prec == "fast" and n <= 256:
            mm = 1 << 24                      # C2 in the fast mode: the pool size the round-1 table quotes
        p = ob.CandidatePool.counter(mm, lo, hi, seed=1)
        for _ in range(3):
            ob.score(models, spec, p, precision=prec, sync=False)
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ob.score(models, spec, p, precision=prec, sync=False); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        print(f"| {name.split('_')[0]} (n={n}, d={X.shape[1]}, {len(models)} GPs) | {prec} | 2^{int(np.log2(mm))} | {ms:.2f} | {mm / ms * 1e3:.3e} |", flush=True)
