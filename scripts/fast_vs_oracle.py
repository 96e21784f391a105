import sys, numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
from oracle import oracle as O
n, d = 1024, 10
rng = np.random.default_rng(0); X = rng.random((n, d)); y = np.sin(X.sum(1))
gp = ob.GPModel(X, y, 0.7 * np.ones(d), 1.0, device='cuda:0')
st = O.gp_fit_state(X, y, 0.7 * np.ones(d), 1.0)
for m in (128 * 148, 128 * 149, 128 * 300, 1 << 17):
    Xc = rng.random((m, d))
    mu, var = ob.posterior([gp], Xc, precision='fast')
    torch.cuda.synchronize()
    sel = rng.integers(0, m, 2000)
    mu_o, var_o = O.gp_posterior(st, Xc[sel])
    sd, sdo = np.sqrt(var[0].cpu().numpy()[sel]), np.sqrt(var_o)
    print(m, "ok: sd rel err", float((np.abs(sd - sdo) / sdo).max()), "mu err", float(np.abs(mu[0].cpu().numpy()[sel] - mu_o).max()), flush=True)
