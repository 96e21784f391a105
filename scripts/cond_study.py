"""Fast-mode absolute sigma error (in units of sigma_f) against a cheap conditioning proxy from the Cholesky
factor: kappa = (max L_ii / min L_ii)^2.  Used to calibrate PoolOptimiserBase._precision_for."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
rng = np.random.default_rng(0)
print("n d ell kappa_proxy err_sd/sigma_f err_mu/scale")
for n, d, ell in ((1024, 10, 0.7), (1024, 10, 1.5), (1024, 5, 0.9), (1024, 5, 0.5), (1024, 5, 0.3), (1024, 3, 0.5),
                  (1024, 3, 0.2), (512, 2, 0.3), (512, 2, 0.1), (2048, 10, 0.7), (2048, 6, 0.6), (300, 10, 2.0),
                  (1024, 12, 1.0), (1024, 20, 2.0), (1024, 2, 0.05)):
    X = rng.random((n, d)); y = np.sin(3 * X.sum(1))
    gp = ob.GPModel(X, y, ell * np.ones(d), 1.5, device='cuda:0')
    dg = torch.diagonal(gp.L)[:n]
    kappa = float((dg.max() / dg.min()) ** 2)
    Xc = rng.random((20000, d))
    mu_f, var_f = ob.posterior([gp], Xc, precision='fast')
    mu_d, var_d = ob.posterior([gp], Xc, precision='fp64')
    e_sd = float((var_f[0].sqrt() - var_d[0].sqrt()).abs().max() / np.sqrt(1.5))
    e_mu = float((mu_f[0] - mu_d[0]).abs().max() / max(1.0, float(mu_d[0].abs().max())))
    print(n, d, ell, f"{kappa:.2e} {e_sd:.2e} {e_mu:.2e}", flush=True)
