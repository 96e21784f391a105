"""f8c fast kernel (fp16 + 2 x e4m3 corrections, CTA pair) against the FP64 CUDA path: ragged shapes + timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimobo_b200 as ob
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_parity import make_problem

DEV = "cuda:0"
cases = [(100, 4, 129), (128, 10, 1000), (256, 10, 5000), (300, 7, 2049), (512, 12, 3000), (700, 7, 5001),
         (1024, 10, 148 * 128 * 2 + 77), (1536, 10, 2048), (2048, 10, 4096)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
worst = 0.0
for n, d, m in cases:
    X, Y, ells, sf2 = make_problem(n, d)
    for i in range(2):
        gp = ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV)
        Xc = np.random.default_rng(n + d).random((m, d))
        mu64, var64 = ob.posterior([gp], Xc, precision="fp64")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mu, var = ob.posterior([gp], Xc, precision="fast")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        s64, s = var64[0].sqrt().cpu().numpy(), var[0].sqrt().cpu().numpy()
        m64, mf = mu64[0].cpu().numpy(), mu[0].cpu().numpy()
        es = np.abs(s - s64)
        rel = (es / s64).max()
        worst = max(worst, rel)
        print(f"n={n} d={d} m={m} gp{i} fmt={gp.plane_format} kappa={gp.conditioning:.1f} "
              f"max|dsig|/sf={es.max()/np.sqrt(sf2[i]):.2e} max rel dsig={rel:.2e} "
              f"max|dmu|/max|mu|={np.abs(mf-m64).max()/np.abs(m64).max():.2e} t={dt*1e3:.1f} ms", flush=True)
        mu2, var2 = ob.posterior([gp], Xc, precision="fast")
        assert torch.equal(mu2, mu) and torch.equal(var2, var), "not reproducible"
print("worst rel sigma error", worst)
