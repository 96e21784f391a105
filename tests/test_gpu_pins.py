"""Round-2 parity pins (VERDICT r1 "Next round" item 1 and the ADVICE findings), all through the C ABI:

* the CUDA posterior against sklearn's GaussianProcessRegressor.predict(return_std=True) DIRECTLY -- the live
  call at /root/reference/optimobo/util_functions.py:265 -- so row a1 is pinned against an implementation the
  builder did not write (sklearn 1.9 is in the image on the GPU box too);
* the device scalarisation switch against the reference-minted k = 2 AND k = 3 fixtures, all 12 functions;
* `GPModel.from_gpy` on a GPy-shaped stub, two devices in one process, large-n refresh + posterior;
* the fast mode at north_star's 1e-3 bound on the acquisition values too.
"""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import optimobo_b200 as ob  # noqa: E402
from optimobo_b200 import _cabi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_gpu_parity import make_problem  # noqa: E402

DEV = "cuda:0"


# ------------------------------------------------------------------------------------------
# a1 against sklearn, directly (util_functions.py:265: model.predict(X, return_std=True))
# ------------------------------------------------------------------------------------------
def _sklearn_gp(X, y, ell, sf2, kernel="matern52"):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    base = Matern(length_scale=ell, nu=2.5) if kernel == "matern52" else RBF(length_scale=ell)
    gp = GaussianProcessRegressor(kernel=ConstantKernel(sf2) * base, alpha=1e-8, optimizer=None, normalize_y=False)
    return gp.fit(X, y)


@pytest.mark.parametrize("n,d,m", [(256, 10, 4096), (512, 12, 4096), (1024, 10, 4096)])     # C2, C4, C5 shapes
def test_posterior_vs_sklearn_direct(n, d, m):
    X, Y, ells, sf2 = make_problem(n, d)
    Xc = np.random.default_rng(17).random((m, d))
    for i in range(2):
        sk = _sklearn_gp(X, Y[:, i], ells[i], sf2[i])
        mu_s, sd_s = sk.predict(Xc, return_std=True)
        gp = ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV)
        # FP64 mode, the sklearn-shaped surface of GPModel: rtol 1e-6 (north_star)
        mu, sd = gp.predict(Xc, return_std=True)
        np.testing.assert_allclose(mu, mu_s, rtol=1e-6, atol=1e-7 * np.abs(mu_s).max())
        np.testing.assert_allclose(sd, sd_s, rtol=1e-6, atol=1e-7 * np.sqrt(sf2[i]))
        # fast mode: rtol 1e-3 (north_star)
        if _cabi.fast_path_available():
            mu_f, sd_f = gp.predict(Xc, return_std=True, precision="fast")
            np.testing.assert_allclose(mu_f, mu_s, rtol=1e-3, atol=1e-3 * np.abs(mu_s).max())
            np.testing.assert_allclose(sd_f, sd_s, rtol=1e-3, atol=1e-4 * np.sqrt(sf2[i]))


def test_posterior_vs_sklearn_rbf():
    n, d, m = 200, 5, 2000
    X, Y, ells, sf2 = make_problem(n, d)
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel
    # RBF Gram matrices are numerically singular without noise: sklearn's alpha plays GPy's noise + jitter
    sk = GaussianProcessRegressor(kernel=ConstantKernel(sf2[0]) * RBF(length_scale=ells[0]), alpha=1e-3 + 1e-8,
                                  optimizer=None).fit(X, Y[:, 0])
    Xc = np.random.default_rng(3).random((m, d))
    mu_s, sd_s = sk.predict(Xc, return_std=True)
    gp = ob.GPModel(X, Y[:, 0], ells[0], sf2[0], noise=1e-3, kernel="rbf", device=DEV)
    mu, var = gp.predict(Xc)                              # GPy surface: variance INCLUDING the noise term
    np.testing.assert_allclose(mu[:, 0], mu_s, rtol=1e-6, atol=1e-6 * np.abs(mu_s).max())
    np.testing.assert_allclose(np.sqrt(var[:, 0] - 1e-3), sd_s, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------
# a7: all 12 scalarisations on the device against the reference-minted fixtures, k = 2 and k = 3
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("name", O.SCALARISATIONS)
def test_device_scalarisations_golden(golden, name, k):
    from optimobo_b200 import scalarisations as S
    kw = dict(p=8) if name == "ExponentialWeightedCriterion" else {}
    obj = getattr(S, name)(golden[f"sc{k}_in_ideal"], golden[f"sc{k}_in_max"], **kw)
    F, w = golden[f"sc{k}_in_F"], golden[f"sc{k}_in_w"]
    got = ob.scalarise_on_device(obj, F, w, DEV)
    np.testing.assert_allclose(got, golden[f"sc{k}_out_{name}_batch"], rtol=1e-9)
    np.testing.assert_allclose(got, golden[f"sc{k}_out_{name}_single"], rtol=1e-9)
    # and the host plugin object agrees with its device twin
    np.testing.assert_allclose(got, np.asarray(obj(F.copy(), w)).reshape(-1), rtol=1e-9)


def test_scalarise_entry_is_loud():
    from optimobo_b200 import scalarisations as S
    obj = S.Tchebicheff([0, 0, 0, 0, 0], [1, 1, 1, 1, 1])
    with pytest.raises(ValueError, match="objectives"):
        ob.scalarise_on_device(obj, np.zeros((3, 5)), np.ones(5) / 5, DEV)

    class Mine(S.Scalarisation):
        pass
    with pytest.raises((TypeError, NotImplementedError, AttributeError)):
        ob.scalarise_on_device(Mine([0, 0], [1, 1]), np.zeros((3, 2)), np.ones(2) / 2, DEV)


# ------------------------------------------------------------------------------------------
# boundary: GPModel.from_gpy on a GPy-shaped object (GPy itself is not installable offline)
# ------------------------------------------------------------------------------------------
def test_from_gpy_adopts_a_gpy_shaped_model():
    n, d = 60, 3
    X, Y, ells, sf2 = make_problem(n, d)
    stub = types.SimpleNamespace(
        X=X, Y=Y[:, :1], kern=types.SimpleNamespace(lengthscale=ells[0], variance=np.array([sf2[0] * 1.7])),
        Gaussian_noise=types.SimpleNamespace(variance=np.array([0.0])))
    gp = ob.GPModel.from_gpy(stub, device=DEV)
    assert gp.n == n and gp.d == d and gp.noise == 0.0 and gp.variance == pytest.approx(1.7 * sf2[0])
    np.testing.assert_allclose(gp.kern.lengthscale, ells[0])
    st = O.gp_fit_state(X, Y[:, 0], ells[0], 1.7 * sf2[0])
    Xc = np.random.default_rng(0).random((500, d))
    mu_o, var_o = O.gp_posterior(st, Xc)
    mu, var = gp.predict(Xc)                              # GPy surface
    assert mu.shape == (500, 1) and var.shape == (500, 1)
    np.testing.assert_allclose(mu[:, 0], mu_o, rtol=1e-6, atol=1e-7 * np.abs(mu_o).max())
    np.testing.assert_allclose(np.sqrt(var[:, 0]), np.sqrt(var_o), rtol=1e-6, atol=2e-6 * np.sqrt(1.7 * sf2[0]))


# ------------------------------------------------------------------------------------------
# several devices driven by ONE process (per-device kernel attributes, per-device contexts)
# ------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    n, d, m = 300, 6, 3000
    X, Y, ells, sf2 = make_problem(n, d)
    Xc = np.random.default_rng(1).random((m, d))
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=dev) for i in range(2)]
        spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), ob.host_prep.cached_samples(2, 5, seed=0), "exact")
        for precision in ("fp64", "fast"):
            r = ob.score(models, spec, ob.CandidatePool.explicit(Xc, device=dev), precision=precision, want_acq=True)
            outs.append((dev, precision, r.acq.cpu().numpy(), r.best_index))
        pool = ob.CandidatePool.counter(1000, np.zeros(d), np.ones(d), seed=2)
        assert pool.rows(5, 3, dev).device == torch.device(dev)
    for a, b in ((0, 2), (1, 3)):
        assert np.array_equal(outs[a][2], outs[b][2]) and outs[a][3] == outs[b][3]      # deterministic kernels


# ------------------------------------------------------------------------------------------
# K3 + posterior at the large end of the header's range (n_pad = 2048, 4096)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(2048, 10), (4000, 8)])
def test_large_n_refresh_and_posterior(n, d):
    X, Y, ells, sf2 = make_problem(n, d)
    Xc = np.random.default_rng(2).random((1024, d))
    gp = ob.GPModel(X, Y[:, 0], ells[0], sf2[0], device=DEV)
    st = O.gp_fit_state(X, Y[:, 0], ells[0], sf2[0], form="direct")
    L = gp.L.cpu().numpy()
    np.testing.assert_allclose(L, st["L"], rtol=1e-7, atol=1e-9)
    mu_o, var_o = O.gp_posterior(st, Xc)
    mu, var = ob.posterior([gp], Xc)
    np.testing.assert_allclose(mu[0].cpu().numpy(), mu_o, rtol=1e-6, atol=1e-6 * np.abs(mu_o).max())
    np.testing.assert_allclose(np.sqrt(var[0].cpu().numpy()), np.sqrt(var_o), rtol=1e-6, atol=2e-6 * np.sqrt(sf2[0]))
    if _cabi.fast_path_available() and gp.conditioning <= ob.GPModel.FAST_MODE_CONDITIONING_LIMIT:
        mu_f, var_f = ob.posterior([gp], Xc, precision="fast")
        np.testing.assert_allclose(mu_f[0].cpu().numpy(), mu_o, rtol=1e-3, atol=1e-3 * np.abs(mu_o).max())
        np.testing.assert_allclose(np.sqrt(var_f[0].cpu().numpy()), np.sqrt(var_o), rtol=1e-3, atol=1e-3 * np.sqrt(sf2[0]))


# ------------------------------------------------------------------------------------------
# fast mode at north_star's bound: sigma AND acquisition values within 1e-3, same selection
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(256, 10), (512, 12), (1024, 10)])
def test_fast_mode_meets_1e3_on_sigma_and_ehvi(n, d):
    """north_star: std and acquisition values within 1e-3 in the fast mode.  sigma = sqrt(sigma_f2 - |V|^2) loses
    relative accuracy as sigma -> 0 (cancellation), so the bound carries the absolute floor the
    formats are measured at (<= 4e-4 sigma_f up to kappa = 100 for f8c); the acquisition value
    is held to 1e-3 of its own value plus 1e-3 of the pool maximum for the far tails of Phi, and the selected
    candidate must be the FP64 one whenever the top-two gap exceeds the tolerance."""
    if not _cabi.fast_path_available():
        pytest.skip("fast path not built")
    m = 1 << 14
    X, Y, ells, sf2 = make_problem(n, d)
    models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV) for i in range(2)]
    assert max(g.conditioning for g in models) <= 2e3            # the regime the fast mode is specified for
    pool = ob.CandidatePool.counter(m, np.zeros(d), np.ones(d), seed=1)
    spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), ob.host_prep.cached_samples(2, 5, seed=0), "exact")
    fast = ob.score(models, spec, pool, precision="fast", want_acq=True, want_posterior=True)
    ref = ob.score(models, spec, pool, precision="fp64", want_acq=True, want_posterior=True)
    for i in range(2):
        s_f, s_r = fast.var[i].sqrt().cpu().numpy(), ref.var[i].sqrt().cpu().numpy()
        np.testing.assert_allclose(s_f, s_r, rtol=1e-3, atol=4e-4 * np.sqrt(sf2[i]))
        np.testing.assert_allclose(fast.mu[i].cpu().numpy(), ref.mu[i].cpu().numpy(), rtol=1e-3,
                                   atol=1e-3 * ref.var[i].sqrt().max().item())        # |d mu| <= 1e-3 sigma scale
    a_f, a_r = fast.acq.cpu().numpy(), ref.acq.cpu().numpy()
    np.testing.assert_allclose(a_f, a_r, rtol=1e-3, atol=1e-3 * np.abs(a_r).max())
    order = np.argsort(-a_r)
    if a_r[order[0]] - a_r[order[1]] > 1e-3 * a_r[order[0]]:
        assert fast.best_index == ref.best_index
    assert a_r[fast.best_index] >= a_r[order[0]] * (1 - 1e-3)


def test_precision_auto_everywhere():
    """'auto' is accepted by every public entry point, not only by the optimisers."""
    n, d = 300, 6
    X, Y, ells, sf2 = make_problem(n, d)
    models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV) for i in range(2)]
    spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), ob.host_prep.cached_samples(2, 5, seed=0))
    pool = ob.CandidatePool.counter(2000, np.zeros(d), np.ones(d), seed=3)
    want = ob.resolve_precision(models, "auto")
    assert want == ("fast" if _cabi.fast_path_available() else "fp64")
    a = ob.score(models, spec, pool, precision="auto")
    b = ob.score(models, spec, pool, precision=want)
    assert (a.best_index, a.best_value) == (b.best_index, b.best_value)
    Xh = ob.CandidatePool.counter(2000, np.zeros(d), np.ones(d), seed=3).rows(0, 2000, DEV).cpu().numpy()
    v, i = ob.propose_host(models, spec, Xh, precision="auto")
    assert i == ob.propose_host(models, spec, Xh, precision=want)[1]
    x, neg, idx = ob.propose(models, spec, pool, precision="auto")
    assert idx == a.best_index
    with pytest.raises(ValueError, match="precision"):
        ob.score(models, spec, pool, precision="fp32")


def test_refresh_models_concurrent_matches_serial():
    """gp.refresh_models (one stream + host thread per model) leaves the same state blobs as serial refreshes."""
    X, Y, ells, sf2 = make_problem(700, 7)
    a = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV) for i in range(2)]
    b = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV, refresh=False) for i in range(2)]
    ob.refresh_models(b)
    torch.cuda.synchronize()
    for ma, mb in zip(a, b):
        assert mb.refreshed and mb.plane_format == ma.plane_format
        for f in ("L", "Linv", "alpha"):             # (the blob also holds scratch fields a refresh never initialises)
            assert torch.equal(getattr(ma, f), getattr(mb, f)), f
    pool = ob.CandidatePool.counter(4096, np.zeros(7), np.ones(7), seed=2)
    for prec in ("fp64", "fast"):
        ra = ob.score(a, None, pool, precision=prec, want_posterior=True)
        rb = ob.score(b, None, pool, precision=prec, want_posterior=True)
        assert torch.equal(ra.mu, rb.mu) and torch.equal(ra.var, rb.var)
