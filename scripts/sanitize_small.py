"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): refresh, FP64 + fast posterior,
every acquisition kind, arg-max, host entry."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import optimobo_b200 as ob
from optimobo_b200 import scalarisations as S

rng = np.random.default_rng(0)
n, d, m = 150, 5, 700
X = rng.random((n, d))
Y = np.column_stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0]), X[:, 1] ** 2 + X[:, 2]])
models = [ob.GPModel(X, Y[:, i], 0.7 * np.ones(d), 1.0 + i, device="cuda:0") for i in range(3)]
pool = ob.CandidatePool.counter(m, np.zeros(d), np.ones(d), seed=1)
cache2, cache3 = ob.host_prep.cached_samples(2, 4, seed=0), ob.host_prep.cached_samples(3, 4, seed=0)
PF = ob.host_prep.calc_pf(Y[:, :2]); r = Y[:, :2].max(0)
specs = [ob.spec_ehvi(r, PF, cache2), ob.spec_ehvi(r, PF, cache2, "exact"),
         ob.spec_ehvi3d(Y.max(0) + 1, ob.host_prep.calc_pf(Y), cache3),
         ob.spec_expected_decomposition([.5, .5], S.PBI(Y[:, :2].min(0), r), 0.3, cache2),
         ob.spec_ei(0.2), ob.spec_constrained_ei(0.2, 2), ob.spec_pareto_ei(0.2),
         ob.spec_hv_poi(ob.host_prep.decompose_into_cells(PF, Y[:, :2].min(0), r))]
for prec in ("fp64", "fast"):
    for sp in specs:
        res = ob.score(models[:max(sp.n_models, 1)], sp, pool, precision=prec, want_acq=True, want_posterior=True)
        assert np.isfinite(res.best_value) or res.best_value == -np.inf
Xh = torch.rand((m, d), dtype=torch.float64).pin_memory()
print(ob.propose_host(models[:2], specs[1], Xh, precision="fast"))
torch.cuda.synchronize()
print("sanitize run ok")
