"""A few likelihood + gradient evaluations at n = 1024, d = 10 (for an ncu launch list of the fit path)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from optimobo_b200.fit import fit_hyperparameters_device
rng = np.random.default_rng(0)
X = rng.random((1024, 10)); y = np.sin(3 * X[:, 0]) + X[:, 1:].sum(1)
info = {}
fit_hyperparameters_device(X, y, max_f_eval=4, device="cuda:0", info=info)
torch.cuda.synchronize()
print("ok", info)
