"""Robustness / timing at training-set sizes beyond C5 (n = 2048, 4096): parity of both precision modes against
the oracle on a sample, and candidates/s of one GP."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
from oracle import oracle as O
d = 10
for n in (2048, 4096):
    rng = np.random.default_rng(n); X = rng.random((n, d)); y = np.sin(X.sum(1))
    ell = 0.7 * np.ones(d)
    t0 = time.perf_counter(); gp = ob.GPModel(X, y, ell, 1.0, device='cuda:0'); torch.cuda.synchronize(); t_ref = time.perf_counter() - t0
    st = O.gp_fit_state(X, y, ell, 1.0)
    Xc = rng.random((1 << 17, d)); sel = rng.integers(0, len(Xc), 1500)
    mu_o, var_o = O.gp_posterior(st, Xc[sel])
    for prec in ("fp64", "fast"):
        ob.posterior([gp], Xc[:4096], precision=prec); torch.cuda.synchronize()
        t0 = time.perf_counter(); mu, var = ob.posterior([gp], Xc, precision=prec); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        sd, sdo = np.sqrt(var[0].cpu().numpy()[sel]), np.sqrt(var_o)
        print(f"n={n} {prec}: refresh {t_ref*1e3:.1f} ms  sd rel err {np.abs(sd - sdo).max() / sdo.max():.2e} (max rel {(np.abs(sd-sdo)/sdo).max():.2e})  "
              f"mu err {np.abs(mu[0].cpu().numpy()[sel] - mu_o).max():.2e}  {len(Xc)/dt:.3e} cand/s", flush=True)
