"""One (n, d, m) case of the fast mode against the FP64 CUDA path (debug helper): python scripts/f8c_one.py n d m"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import optimobo_b200 as ob
from test_gpu_parity import make_problem
n, d, m = (int(a) for a in sys.argv[1:4])
X, Y, ells, sf2 = make_problem(n, d)
gp = ob.GPModel(X, Y[:, 0], ells[0], sf2[0], device="cuda:0")
Xc = np.random.default_rng(n + d).random((m, d))
mu64, var64 = ob.posterior([gp], Xc, precision="fp64")
torch.cuda.synchronize()
mu, var = ob.posterior([gp], Xc, precision="fast")
torch.cuda.synchronize()
es = (var[0].sqrt() - var64[0].sqrt()).abs()
print(f"n={n} d={d} m={m} fmt={gp.plane_format} max|dsig|/sf={float(es.max())/np.sqrt(sf2[0]):.2e} "
      f"max|dmu|={float((mu-mu64).abs().max()):.2e}", flush=True)
