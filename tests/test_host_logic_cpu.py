"""CPU: host-side logic around the hot path -- hyper-parameter fit gradients, the Problem shim, the
constrained-ParEGO bookkeeping, the packed arg-max key and its all-reduce over gloo (world_size 2)."""
import os

import numpy as np
import pytest

from optimobo_b200.fit import _nlml_and_grad, fit_hyperparameters
from optimobo_b200.problem import ElementwiseProblem, Problem
from optimobo_b200.distributed import pack_key_host, unpack_key


def test_fit_gradient_matches_finite_differences():
    rng = np.random.default_rng(0)
    X = rng.random((25, 3)); y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2
    for kernel in ("matern52", "rbf"):
        th = np.array([0.3, -0.2, 0.1, 0.4])
        f0, g = _nlml_and_grad(th, X, y, kernel, 1e-8)
        for i in range(len(th)):
            e = np.zeros_like(th); e[i] = 1e-6
            fd = (_nlml_and_grad(th + e, X, y, kernel, 1e-8)[0] - _nlml_and_grad(th - e, X, y, kernel, 1e-8)[0]) / 2e-6
            assert abs(fd - g[i]) <= 1e-4 * max(1.0, abs(fd)), (kernel, i, fd, g[i])


def test_fit_improves_likelihood_and_matches_sklearn_optimum():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.default_rng(1)
    X = rng.random((40, 2)); y = np.cos(4 * X[:, 0]) * X[:, 1]
    ell, sf2 = fit_hyperparameters(X, y)
    th = np.log(np.concatenate(([sf2], ell)))
    assert _nlml_and_grad(th, X, y, "matern52", 1e-8)[0] < _nlml_and_grad(np.zeros(3), X, y, "matern52", 1e-8)[0]
    gpr = GaussianProcessRegressor(ConstantKernel(1.0) * Matern(length_scale=np.ones(2), nu=2.5), alpha=1e-8,
                                   n_restarts_optimizer=0).fit(X, y)
    assert -gpr.log_marginal_likelihood_value_ >= _nlml_and_grad(th, X, y, "matern52", 1e-8)[0] - 1e-3 * abs(
        gpr.log_marginal_likelihood_value_) - 1e-2
    assert fit_hyperparameters(X, np.ones(40))[1] == 1.0          # constant targets: start point returned


def test_problem_shims():
    class MyProblem(ElementwiseProblem):          # README example 1
        def __init__(self):
            super().__init__(n_var=2, n_obj=2, xl=np.array([-2, -2]), xu=np.array([2, 2]))

        def _evaluate(self, x, out, *args, **kwargs):
            out["F"] = [100 * (x[0] ** 2 + x[1] ** 2), (x[0] - 1) ** 2 + x[1] ** 2]

    class BNH(Problem):                          # README example 2
        def __init__(self):
            super().__init__(n_var=2, n_obj=2, n_ieq_constr=2, vtype=float)
            self.xl = np.zeros(self.n_var); self.xu = np.array([5.0, 3.0])

        def _evaluate(self, x, out, *args, **kwargs):
            out["F"] = [4 * x[:, 0] ** 2 + 4 * x[:, 1] ** 2, (x[:, 0] - 5) ** 2 + (x[:, 1] - 5) ** 2]

        def _evaluate_constraints(self, x, out, *args, **kwargs):
            out["G"] = [(1 / 25) * ((x[:, 0] - 5) ** 2 + x[:, 1] ** 2 - 25),
                        -1 / 7.7 * ((x[:, 0] - 8) ** 2 + (x[:, 1] + 3) ** 2 - 7.7)]

    p = MyProblem()
    np.testing.assert_allclose(p.evaluate(np.array([1.0, 0.5])), [125.0, 0.25])
    assert p.evaluate(np.zeros((3, 2))).shape == (3, 2)
    b = BNH()
    np.testing.assert_allclose(b.evaluate(np.array([1.0, 1.0])), [8.0, 32.0])
    g = b.evaluate_constraints(np.array([[1.0, 1.0], [5.0, 3.0]]))
    assert g.shape == (2, 2) and g[0, 0] < 0
    with pytest.raises(AssertionError):
        p.evaluate(np.zeros(3))


def test_constrained_parego_bookkeeping():
    from optimobo_b200.algorithms.cparego import _ConstrainedParEGO
    rng = np.random.default_rng(3)
    n = 60
    agg = rng.random(n); Y = rng.random((n, 2)); G = rng.normal(size=(n, 2))
    inf = np.any(G > 0, axis=1)
    pen = _ConstrainedParEGO._penalise(agg, G, inf, True)
    assert np.array_equal(pen[~inf], agg[~inf]) and np.all(pen[inf] >= agg[~inf].min())

    class Dummy(_ConstrainedParEGO):
        def __init__(self):
            self.n_vars, self.n_obj, self.n_ieq_constr, self.n_eq_constr = 2, 2, 2, 0
    o = Dummy()
    fi, ii = np.flatnonzero(~inf), np.flatnonzero(inf)
    for N_max in (10, 30, 200):
        sel = o.select_subset(fi, ii, pen, Y, G, np.array([0.5, 0.5]), N_max)
        assert len(sel) == min(N_max, n) and len(set(sel.tolist())) == len(sel)
    sel = o.select_subset(fi, np.array([], dtype=int), pen, Y, G, np.array([0.5, 0.5]), 10)
    assert set(sel).issubset(set(fi)) and len(sel) == 10
    assert o.select_current_best(fi, ii, pen, G) == pen[fi].max()


def test_packed_key_orders_like_value_then_lowest_index():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(size=200), [0.0, -0.0, np.inf, -np.inf, 1e-30, -1e-30, 3.5, 3.5]])
    idxs = rng.integers(0, 2 ** 32 - 1, size=len(vals))
    keys = [pack_key_host(v, int(i)) for v, i in zip(vals, idxs)]
    for a in range(0, len(vals), 7):
        for b in range(0, len(vals), 5):
            va, vb = np.float32(vals[a]), np.float32(vals[b])
            if va > vb:
                assert keys[a] > keys[b]
            elif va == vb and not (va == 0 and np.signbit(va) != np.signbit(vb)) and idxs[a] < idxs[b]:
                assert keys[a] > keys[b]
    v, i = unpack_key(pack_key_host(1.7174754575557043, 774849))
    assert i == 774849 and v == float(np.float32(1.7174754575557043))
    assert unpack_key(pack_key_host(-np.inf, -1))[0] == -np.inf
    assert -(2 ** 63) <= min(keys) and max(keys) < 2 ** 63


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from optimobo_b200.distributed import allreduce_best_key
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank owns a shard's local best (value, global index); ties across ranks -> lowest index
    local = [(0.25, 100), (0.75, 1000 + rank), (0.75, 5 - rank)][rank % 3] if world > 2 else [(0.75, 9), (0.75, 4)][rank]
    key = torch.tensor([pack_key_host(*local)], dtype=torch.int64)
    allreduce_best_key(key)
    out[rank] = unpack_key(int(key.item()))
    dist.destroy_process_group()


def test_allreduce_best_key_gloo_world2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1] == (0.75, 4)


def _gloo_exact_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from optimobo_b200.distributed import allgather_best_exact
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    # (a) denormal-range values (late-BO EI, constrained EI x PoF): float32 would flush both to 0 and pick index 3
    # (b) values that agree to float32 precision but not in FP64   (c) equal values: lowest index   (d) NaN never wins
    cases = [[(3e-50, 3), (7e-50, 900)], [(1.0 + 1e-12, 50), (1.0, 7)], [(0.75, 9), (0.75, 4)], [(float("nan"), 1), (-2.0, 8)]]
    for case in cases:
        v, i = case[rank]
        pair = torch.tensor([0, i], dtype=torch.int64)
        pair[:1].view(torch.float64)[0] = v
        res.append(allgather_best_exact(pair))
    out[rank] = res
    dist.destroy_process_group()


def test_allgather_best_exact_gloo_world2():
    """The default multi-rank reduce: one all-gather of the 16-byte (FP64 value, index) pairs, exact for every
    value range (the packed float32 key is not: ADVICE r1)."""
    import torch.multiprocessing as mp
    from optimobo_b200.distributed import pick_best
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gloo_exact_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1] == [(7e-50, 900), (1.0 + 1e-12, 50), (0.75, 4), (-2.0, 8)]
    # the packed key really does lose case (a): both float32 images are 0 -> lowest index
    assert unpack_key(max(pack_key_host(3e-50, 3), pack_key_host(7e-50, 900)))[1] == 3
    assert pick_best([np.nan, np.nan], [5, 2]) == (-np.inf, 2)


# ------------------------------------------------------------------------------------------
# TuRBO bookkeeping (turbo.py:119-154, :340-380): trust-region length, batch selection
# ------------------------------------------------------------------------------------------
def _turbo_problem():
    from optimobo_b200.problem import Problem

    class P(Problem):
        def __init__(self):
            super().__init__(n_var=3, n_obj=2, xl=np.zeros(3), xu=np.ones(3))

        def _evaluate(self, x, out, *args, **kwargs):
            out["F"] = [x[:, 0], 1 - x[:, 0] + x[:, 1:].sum(1)]
    return P()


def test_turbo_1_trust_region_and_selection():
    from optimobo_b200.algorithms import TuRBO_1
    t = TuRBO_1(_turbo_problem(), batch_size=2, ideal_point=[0, 0], max_point=[1, 3], seed=0)
    assert t.n_cand == 300 and t.failtol == 2 and t.succtol == 3 and t.length == 0.4
    # select_candidates: the minimiser of each sample, never the same candidate twice
    X_cand = np.arange(15, dtype=float).reshape(5, 3) / 15
    y_cand = np.array([[3.0, 3.0], [1.0, 1.0], [2.0, 0.5], [5.0, 5.0], [4.0, 4.0]]).reshape(5, 1, 2)
    picked = t.select_candidates(X_cand, y_cand.copy())
    np.testing.assert_array_equal(picked, X_cand[[1, 2]])
    y_same = np.array([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0], [4.0, 4.0], [5.0, 5.0]]).reshape(5, 1, 2)
    np.testing.assert_array_equal(t.select_candidates(X_cand, y_same.copy()), X_cand[[0, 1]])
    # _adjust_length: three successes double the region (capped), `failtol` failures halve it
    t._restart()
    t._aggregated_samples = np.array([[1.0]])
    for _ in range(3):
        t._adjust_length(np.array([0.5]))
    assert t.length == 0.8 and t.succcount == 0
    t._adjust_length(np.array([2.0])); t._adjust_length(np.array([2.0]))
    assert t.length == 0.4 and t.failcount == 0
    t._adjust_length(np.array([0.9995]))                 # within 1e-3 relative of the best: a failure
    assert t.failcount == 1
    np.testing.assert_allclose(t.denormalise(t.normalise(np.array([[0.2, 0.4, 0.9]]))), [[0.2, 0.4, 0.9]])


def test_turbo_m_selection_across_regions():
    from optimobo_b200.algorithms import TuRBO_M
    t = TuRBO_M(_turbo_problem(), [0, 0], [1, 3], batch_size=2, n_trust_regions=2, seed=0)
    assert t.failtol == 5 and t.length.shape == (2,)
    t.n_cand = 3
    X_cand = np.arange(18, dtype=float).reshape(2, 3, 3) / 18
    y_cand = np.array([[[5.0, 0.1], [4.0, 9.0], [3.0, 9.0]], [[0.2, 9.0], [6.0, 9.0], [7.0, 0.05]]])
    X_next, idx = t._select_candidates(X_cand, y_cand.copy())
    np.testing.assert_array_equal(X_next, np.stack([X_cand[1, 0], X_cand[1, 2]]))
    np.testing.assert_array_equal(idx[:, 0], [1, 1])
    # a region's length reacts to the samples IT proposed (objective 0 of its own history, turbo.py:348)
    t.ysample = np.array([[1.0, 0.0], [2.0, 0.0], [0.3, 0.0]])
    t._idx = np.array([[0], [0], [1]])
    t._adjust_length(np.array([0.5]), 0)
    assert t.succcount[0] == 1 and t.failcount[0] == 0
    t._adjust_length(np.array([0.31, 0.4]), 1)
    assert t.succcount[1] == 0 and t.failcount[1] == 2
