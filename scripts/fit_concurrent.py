"""Two objectives' device hyper-parameter fits: one after the other against concurrently (threads + streams)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import optimobo_b200 as ob
from optimobo_b200.algorithms.base import PoolOptimiserBase
from optimobo_b200.fit import fit_hyperparameters_device
rng = np.random.default_rng(0)
n, d = 1024, 10
X = rng.random((n, d))
Y = np.column_stack([np.sin(3 * X[:, 0]) + X[:, 1:].sum(1), np.cos(2 * X[:, 1]) * X[:, 2:].sum(1)])
class Opt(PoolOptimiserBase):
    pass
o = Opt.__new__(Opt)
o.hyperparameters, o.fit_on_device, o.max_f_eval, o.device = None, True, 40, "cuda:0"
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    serial = [o._fit_model(X, Y[:, i]) for i in range(2)]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    conc = o._fit_models(X, [Y[:, 0], Y[:, 1]])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"serial {1e3 * (t1 - t0):.1f} ms, concurrent {1e3 * (t2 - t1):.1f} ms")
for a, b in zip(serial, conc):
    assert np.array_equal(a.lengthscale, b.lengthscale) and a.variance == b.variance, (a.lengthscale, b.lengthscale)
    assert torch.equal(a.state, b.state)
print("same hyper-parameters and state: ok")
