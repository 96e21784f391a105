"""GPU: the re-hosted optimiser surface (`.solve()` of the six classes the reference exposes) runs
end to end on the pool-scoring path, and the sharded multi-GPU entry agrees with the single-GPU one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import optimobo_b200 as ob  # noqa: E402
import optimobo_b200.scalarisations as sc  # noqa: E402
from optimobo_b200.algorithms import (EMO, KEEP, MonoSurrogateOptimiser, MultiSurrogateOptimiser, ParEGO,  # noqa: E402
                                      ParEGO_C1, ParEGO_C2)
from optimobo_b200.problem import ElementwiseProblem, Problem  # noqa: E402
from oracle import oracle as O  # noqa: E402


class MyProblem(ElementwiseProblem):     # README example 1 (BASELINE config 1)
    def __init__(self):
        super().__init__(n_var=2, n_obj=2, xl=np.array([-2, -2]), xu=np.array([2, 2]))

    def _evaluate(self, x, out, *args, **kwargs):
        out["F"] = [100 * (x[0] ** 2 + x[1] ** 2), (x[0] - 1) ** 2 + x[1] ** 2]


class BNH(Problem):                      # README example 2 (BASELINE config 3)
    def __init__(self):
        super().__init__(n_var=2, n_obj=2, n_ieq_constr=2, vtype=float)
        self.xl = np.zeros(self.n_var)
        self.xu = np.array([5.0, 3.0])

    def _evaluate(self, x, out, *args, **kwargs):
        out["F"] = [4 * x[:, 0] ** 2 + 4 * x[:, 1] ** 2, (x[:, 0] - 5) ** 2 + (x[:, 1] - 5) ** 2]

    def _evaluate_constraints(self, x, out, *args, **kwargs):
        out["G"] = [(1 / 25) * ((x[:, 0] - 5) ** 2 + x[:, 1] ** 2 - 25),
                    -1 / 7.7 * ((x[:, 0] - 8) ** 2 + (x[:, 1] + 3) ** 2 - 7.7)]


class DTLZ2(Problem):
    def __init__(self, n_var=6, n_obj=3):
        super().__init__(n_var=n_var, n_obj=n_obj, xl=np.zeros(n_var), xu=np.ones(n_var))

    def _evaluate(self, x, out, *args, **kwargs):
        g = ((x[:, 2:] - 0.5) ** 2).sum(1)
        a, b = 0.5 * np.pi * x[:, 0], 0.5 * np.pi * x[:, 1]
        out["F"] = [(1 + g) * np.cos(a) * np.cos(b), (1 + g) * np.cos(a) * np.sin(b), (1 + g) * np.sin(a)]


KW = dict(n_candidates=1 << 14, seed=0, device="cuda:0", max_f_eval=200)


def _check(res, n_init, budget, n_obj=2):
    assert res.ysample.shape == (n_init + budget, n_obj) and res.Xsample.shape[0] == n_init + budget
    assert len(res.hypervolume_convergence) >= 1 and len(res.pf_approx) >= 1
    assert res.pf_inputs.shape[0] == res.pf_approx.shape[0]
    assert np.all(np.isfinite(res.ysample))
    assert len(res.timings) >= 1 and res.timings[0]["score_s"] > 0


def test_multisurrogate_readme_tchebicheff_improves_hypervolume():
    opt = MultiSurrogateOptimiser(MyProblem(), [0, 0], [700, 12], **KW)
    res = opt.solve(budget=12, n_init_samples=20, sample_exponent=3, acquisition_func=sc.Tchebicheff([0, 0], [700, 12]))
    _check(res, 20, 12)
    hv = res.hypervolume_convergence
    assert hv[-1] >= hv[0] and all(b >= a - 1e-9 for a, b in zip(hv, hv[1:]))
    final = ob.host_prep.hypervolume(res.ysample, [700, 12])
    assert final > hv[0]


def test_multisurrogate_default_ehvi_and_unknown_bounds():
    res = MultiSurrogateOptimiser(MyProblem(), **KW).solve(budget=4, n_init_samples=10)
    _check(res, 10, 4)
    res = MultiSurrogateOptimiser(MyProblem(), [0, 0], [700, 12], semantics="exact", **KW).solve(budget=4, n_init_samples=10)
    _check(res, 10, 4)


def test_multisurrogate_three_objectives_ehvi3d():
    res = MultiSurrogateOptimiser(DTLZ2(), [0, 0, 0], [2.5, 2.5, 2.5], **KW).solve(budget=3, n_init_samples=12)
    _check(res, 12, 3, n_obj=3)


def test_monosurrogate_parego_keep_emo():
    agg = sc.AugmentedTchebicheff([0, 0], [700, 12])
    _check(MonoSurrogateOptimiser(MyProblem(), [0, 0], [700, 12], **KW).solve(agg, budget=4, n_init_samples=8), 8, 4)
    _check(ParEGO(MyProblem(), [0, 0], [700, 12], **KW).solve(sc.Tchebicheff([0, 0], [700, 12]), budget=4, n_init_samples=8), 8, 4)
    _check(KEEP(MyProblem(), [0, 0], [700, 12], **KW).solve(sc.PBI([0, 0], [700, 12]), budget=3, n_init_samples=8), 8, 3)
    _check(EMO(MyProblem(), [0, 0], [700, 12], **KW).solve(budget=4, n_init_samples=8), 8, 4)


def test_constrained_parego_bnh():
    for cls in (ParEGO_C1, ParEGO_C2):
        res = cls(BNH(), [0, 0], [150, 60], **KW).solve(sc.Tchebicheff([0, 0], [150, 60]), budget=11, n_init_samples=12)
        assert res.ysample.shape == (12 + 11, 2) and res.gsample.shape == (23, 2)
        assert len(res.y_feasible) + len(res.y_infeasible) == 23
        assert np.all(np.any(res.gsample > 0, axis=1) == np.r_[[False] * 0, np.any(res.gsample > 0, axis=1)])


def test_reference_acquisition_callables_run_unmodified_on_gpmodel(golden):
    """The GPy-shaped surface: the ORACLE's per-candidate restatement of the reference call pattern
    (predict one x at a time) evaluated on GPModel equals the batched GPU acquisition."""
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    X = rng.random((60, 3))
    Y = np.column_stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0])])
    models = [ob.GPModel(X, Y[:, i], 0.6 * np.ones(3), 1.0 + i, device="cuda:0") for i in range(2)]
    cache = golden["acq_in_cache2"]
    PF, r = ob.host_prep.calc_pf(Y), Y.max(0)
    Xc = rng.random((40, 3))
    batched = ob.EHVI(Xc, models, r, PF, cache)
    one_at_a_time = []
    for x in Xc:                                   # util_functions.py:154-167 call pattern
        preds = [m.predict(np.asarray([x])) for m in models]
        mu = np.array([[p[0][0][0] for p in preds]]); var = np.array([[p[1][0][0] for p in preds]])
        one_at_a_time.append(O.ehvi_batched(mu[:, 0], mu[:, 1], var[:, 0], var[:, 1], PF, r, cache)[0])
    np.testing.assert_allclose(batched, one_at_a_time, rtol=1e-6, atol=1e-9 * np.abs(batched).max())


def test_unmodified_reference_ehvi_runs_on_gpmodel(golden):
    """The boundary claim of INTEGRATION.md section B: the reference's OWN `util_functions.EHVI` (util_functions.py:136-167,
    loaded unmodified through oracle/ref_loader from baseline/_ref, /root/reference or $OPTIMOBO_REF) takes
    `optimobo_b200.GPModel` objects where it expects GPy models -- one x per call, `model.predict(np.asarray([x]))` --
    and returns what the batched GPU acquisition returns in reference semantics."""
    from oracle import ref_loader as R
    if not R.reference_available():
        pytest.skip("reference package not reachable (baseline/_ref, /root/reference, $OPTIMOBO_REF)")
    uf = R.load_reference().util_functions
    rng = np.random.default_rng(0)
    X = rng.random((60, 3))
    Y = np.column_stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0])])
    models = [ob.GPModel(X, Y[:, i], 0.6 * np.ones(3), 1.0 + i, device="cuda:0") for i in range(2)]
    cache = golden["acq_in_cache2"]
    PF, r = uf.calc_pf(Y), Y.max(0)                       # the reference's own front
    np.testing.assert_array_equal(PF, ob.host_prep.calc_pf(Y))
    Xc = rng.random((40, 3))
    ref_vals = np.array([float(np.asarray(uf.EHVI(x, models, r, PF, cache)).reshape(-1)[0]) for x in Xc])
    batched = ob.EHVI(Xc, models, r, PF, cache)           # FP64 mode, reference semantics
    # np.cov of the affine-mapped samples (util_functions.py:163) rounds differently from var0 * C: 1e-7, as in the goldens
    np.testing.assert_allclose(batched, ref_vals, rtol=1e-7, atol=1e-9 * np.abs(ref_vals).max())
    assert int(np.argmax(batched)) == int(np.argmax(ref_vals))


def test_score_sharded_single_process_matches_score():
    from optimobo_b200.distributed import score_sharded
    rng = np.random.default_rng(0)
    X = rng.random((128, 5)); y = np.sin(X.sum(1))
    gp = ob.GPModel(X, y, 0.8 * np.ones(5), 1.0, device="cuda:0")
    pool = ob.CandidatePool.counter(50000, np.zeros(5), np.ones(5), seed=3)
    v, i = score_sharded([gp], ob.spec_ei(y.min(), 0.0), pool)
    ref = ob.score([gp], ob.spec_ei(y.min(), 0.0), pool)
    assert (v, i) == (ref.best_value, ref.best_index)
    # emulate 4 ranks by hand: shard bests combined through the packed key == global best
    from optimobo_b200.distributed import pack_key_host, unpack_key
    keys = []
    for rnk in range(4):
        r = ob.score([gp], ob.spec_ei(y.min(), 0.0), pool.shard(rnk, 4))
        keys.append(pack_key_host(r.best_value, r.best_index))
    assert unpack_key(max(keys))[1] == ref.best_index


def test_refinement_never_worse_and_usually_better():
    rng = np.random.default_rng(1)
    X = rng.random((200, 6)); y = ((X - 0.37) ** 2).sum(1)
    gp = ob.GPModel(X, y, 0.8 * np.ones(6), 1.0, device="cuda:0")
    pool = ob.CandidatePool.counter(1 << 14, np.zeros(6), np.ones(6), seed=2)
    spec = ob.spec_ei(y.min(), 0.0)
    x0, f0, i0 = ob.propose([gp], spec, pool)
    x1, f1, i1 = ob.propose([gp], spec, pool, refine_rounds=3)
    assert f1 <= f0 and np.all(x1 >= 0) and np.all(x1 <= 1)
    assert (i1 == i0 and np.array_equal(x0, x1)) or (i1 == -1 and f1 < f0)
    np.testing.assert_allclose(ob.expected_improvement(x1, gp, y.min())[0], -f1, rtol=1e-9)


def test_every_acquisition_kind_in_both_precisions():
    import runpy, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cwd = os.getcwd()
    os.chdir(root)
    try:
        runpy.run_path(os.path.join(root, "scripts", "sanitize_small.py"), run_name="__main__")
    finally:
        os.chdir(cwd)


def test_nccl_two_ranks_agree_with_single_gpu():
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(root, "scripts", "check_multigpu.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "multi-GPU check ok" in out.stdout


# ------------------------------------------------------------------------------------------
# C1: the whole outer loop against an oracle twin.  With injected
# hyper-parameters, FP64 scoring, no refinement and the counter-generated pool (bit-identical on
# host and device) the GPU optimiser and a numpy loop built from oracle/ functions must propose
# the SAME points, so evaluated samples and hypervolume trajectories coincide.
# ------------------------------------------------------------------------------------------
class ReadmeProblem(Problem):
    def __init__(self):
        super().__init__(n_var=2, n_obj=2, xl=np.array([-2.0, -2.0]), xu=np.array([2.0, 2.0]))

    def _evaluate(self, x, out, *args, **kwargs):
        out["F"] = [100 * (x[:, 0] ** 2 + x[:, 1] ** 2), (x[:, 0] - 1) ** 2 + x[:, 1] ** 2]


def _oracle_twin_ehvi(problem, seed, budget, n_init, sample_exponent, m, hyper, max_point):
    rng = np.random.default_rng(seed)
    ranges = list(zip(problem.xl, problem.xu))
    X = ob.host_prep.latin_hypercube(n_init, ranges, rng)
    Y = np.asarray([problem.evaluate(x) for x in X])
    cache = ob.host_prep.cached_samples(2, sample_exponent, seed=int(rng.integers(0, 2 ** 31)))
    n_dirs = len(O.das_dennis(100, 2))
    hv = []
    for _ in range(budget):
        hv.append(O.hypervolume(Y, max_point))
        rng.integers(0, n_dirs)                                     # the reference draws a direction every iteration
        states = [O.gp_fit_state(X, Y[:, i], hyper[0], hyper[1]) for i in range(2)]
        pool_seed = int(rng.integers(1, 2 ** 62))
        Xc = O.candidates_from_counter(pool_seed, 0, m, problem.xl, problem.xu)
        (mu0, v0), (mu1, v1) = [O.gp_posterior(s, Xc) for s in states]
        acq = O.ehvi_batched(mu0, mu1, v0, v1, O.calc_pf(Y), np.asarray(max_point, float), cache, "reference")
        x = Xc[O.argmax_lowest_index(acq)]
        X, Y = np.vstack((X, x)), np.vstack((Y, problem.evaluate(x)))
    return X, Y, hv


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_c1_solve_trajectory_matches_oracle_twin(seed):
    problem, max_point = ReadmeProblem(), [700.0, 12.0]
    hyper = (np.array([1.5, 1.5]), 4000.0)
    budget, n_init, m = 5, 12, 1 << 12
    opt = MultiSurrogateOptimiser(problem, [0, 0], max_point, n_candidates=m, precision="fp64",
                                                semantics="reference", seed=seed, hyperparameters=hyper,
                                                refine_rounds=0)
    res = opt.solve(budget=budget, n_init_samples=n_init, sample_exponent=3)
    Xo, Yo, hvo = _oracle_twin_ehvi(problem, seed, budget, n_init, 3, m, hyper, max_point)
    np.testing.assert_allclose(res.Xsample, Xo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(res.ysample, Yo, rtol=1e-12)
    np.testing.assert_allclose(res.hypervolume_convergence, hvo, rtol=1e-12)


def test_device_prep_inside_solve_changes_nothing():
    """prep_on_device=True (first front, hypervolume, cells on the GPU) gives the same run as the
    host versions: the device kernels are bit-identical to host_prep."""
    hyper = (np.array([1.5, 1.5]), 4000.0)
    runs = []
    for on_dev in (False, True):
        opt = MultiSurrogateOptimiser(ReadmeProblem(), [0, 0], [700.0, 12.0], n_candidates=1 << 12,
                                      precision="fp64", seed=3, hyperparameters=hyper, refine_rounds=0,
                                      prep_on_device=on_dev)
        runs.append(opt.solve(budget=4, n_init_samples=10, sample_exponent=3))
    np.testing.assert_array_equal(runs[0].Xsample, runs[1].Xsample)
    np.testing.assert_array_equal(runs[0].hypervolume_convergence, runs[1].hypervolume_convergence)
    emo = []
    for on_dev in (False, True):
        opt = EMO(ReadmeProblem(), [0, 0], [700.0, 12.0], n_candidates=1 << 12, precision="fp64", seed=5,
                  hyperparameters=hyper, refine_rounds=0, prep_on_device=on_dev)
        emo.append(opt.solve(budget=3, n_init_samples=10))
    np.testing.assert_array_equal(emo[0].Xsample, emo[1].Xsample)


def test_turbo_1_runs_and_improves():
    """TuRBO-1 (turbo.py) on the device sampling path: batches of Thompson samples over the trust region."""
    from optimobo_b200.algorithms import TuRBO_1
    opt = TuRBO_1(ReadmeProblem(), batch_size=4, ideal_point=[0, 0], max_point=[700.0, 12.0], seed=0, max_f_eval=100)
    assert opt.n_cand == 200 and opt.failtol == 1
    res = opt.solve(sc.Tchebicheff([0, 0], [700.0, 12.0]), budget=45, n_init_samples=8)
    assert res.ysample.shape[0] >= 45 and res.ysample.shape[1] == 2 and res.Xsample.shape[0] == res.ysample.shape[0]
    assert np.all(res.Xsample >= -2 - 1e-12) and np.all(res.Xsample <= 2 + 1e-12)
    agg = np.maximum(0.5 * res.ysample[:, 0] / 700.0, 0.5 * res.ysample[:, 1] / 12.0)
    assert agg[8:].min() <= agg[:8].min()                 # the model-guided batches find a better point than the design
    assert len(res.pf_approx) >= 1 and len(res.hypervolume_convergence) >= 1


def test_turbo_m_runs():
    from optimobo_b200.algorithms import TuRBO_M
    opt = TuRBO_M(ReadmeProblem(), [0, 0], [700.0, 12.0], batch_size=3, n_trust_regions=2, seed=1, max_f_eval=60)
    res = opt.solve(sc.Tchebicheff([0, 0], [700.0, 12.0]), budget=30, n_init_samples=6)
    assert res.ysample.shape[0] >= 30 and res.Xsample.shape == (res.ysample.shape[0], 2)
    assert set(np.unique(opt._idx)) <= {-1, 0, 1} and len(opt._idx) == len(res.ysample)
    assert np.all(np.isfinite(res.ysample)) and len(res.pf_approx) >= 1
