"""One FP64 scoring pass of the C2 configuration (n = 256, d = 10, 2 GPs, 2^20 candidates) -- for ncu captures."""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import optimobo_b200 as ob
import test_gpu_fullsize as T
X, ys, ells, sf2, spec, lo, hi, m, precision, _ = T.CONFIGS["C2_zdt1_n256_2e20_fp64"]()
models = [ob.GPModel(X, y, e, s, device='cuda:0') for y, e, s in zip(ys, ells, sf2)]
pool = ob.CandidatePool.counter(m, lo, hi, seed=1)
for _ in range(3):
    r = ob.score(models, spec, pool, precision="fp64")
torch.cuda.synchronize()
print(r.best_value, r.best_index)
