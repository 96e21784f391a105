import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import optimobo_b200 as ob
import test_gpu_fullsize as T
from oracle import oracle as O
X, ys, ells, sf2, spec, lo, hi, m, precision, oracle_acq = T.CONFIGS["C2_zdt1_n256_2e20_fp64"]()
models = [ob.GPModel(X, y, e, s, device='cuda:0') for y, e, s in zip(ys, ells, sf2)]
print("kappa", [mm.conditioning for mm in models])
pool = ob.CandidatePool.counter(1 << 24, lo, hi, seed=1)
for prec in ("fast", "fp64"):
    p = pool if prec == "fast" else ob.CandidatePool.counter(1 << 20, lo, hi, seed=1)
    for _ in range(2): ob.score(models, spec, p, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = ob.score(models, spec, p, precision=prec, sync=False); e1.record(); torch.cuda.synchronize()
    print(prec, p.m, e0.elapsed_time(e1), "ms", p.m / e0.elapsed_time(e1) * 1e3, "cand/s", r.best_value, r.best_index)
small = ob.CandidatePool.counter(1 << 14, lo, hi, seed=1)
a_f = ob.score(models, spec, small, precision="fast", want_acq=True).acq.cpu().numpy()
a_d = ob.score(models, spec, small, precision="fp64", want_acq=True).acq.cpu().numpy()
print("acq max abs diff / max", np.abs(a_f - a_d).max() / np.abs(a_d).max(), "argmax equal", a_f.argmax() == a_d.argmax())
