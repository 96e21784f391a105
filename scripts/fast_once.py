import sys, numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
n, d = 1024, 10
m = int(sys.argv[1])
rng = np.random.default_rng(0); X = rng.random((n, d)); y = np.sin(X.sum(1))
gp = ob.GPModel(X, y, 0.7 * np.ones(d), 1.0, device='cuda:0')
Xc = rng.random((m, d))
mu, var = ob.posterior([gp], Xc, precision='fast')
torch.cuda.synchronize()
print(m, "ok", float(var[0].mean()))
