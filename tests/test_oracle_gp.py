"""CPU: the GP posterior restatement (GPy semantics) against sklearn's
GaussianProcessRegressor -- the only independent implementation on this box.  Row a1 is
"parity unpinned" by the reference (GPy absent); this is the cross-check that exists."""
import numpy as np
import pytest

from oracle import oracle as O


def zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.column_stack([f1, g * (1 - np.sqrt(f1 / g))])


@pytest.mark.parametrize("n,d,ell,sf2", [(64, 2, 0.6, 1.0), (256, 10, 0.8, 2.0)])
def test_posterior_vs_sklearn(n, d, ell, sf2):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.default_rng(0)
    X = rng.random((n, d)); y = zdt1(X)[:, 1] if d > 2 else np.sin(3 * X[:, 0]) + X[:, 1]
    Xs = rng.random((500, d))
    ells = ell * np.ones(d)
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(sf2) * Matern(length_scale=ells, nu=2.5),
                                   alpha=1e-8, optimizer=None, normalize_y=False).fit(X, y)
    mu_s, sd_s = gpr.predict(Xs, return_std=True)
    for form in ("gpy", "direct"):
        st = O.gp_fit_state(X, y, ells, sf2, 0.0, 1e-8, form=form)
        mu, var = O.gp_posterior(st, Xs)
        np.testing.assert_allclose(mu, mu_s, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(np.sqrt(var), sd_s, rtol=1e-5, atol=1e-6 * np.sqrt(sf2))
        mu2, sd2 = O.gp_posterior_std(st, Xs)
        np.testing.assert_allclose(sd2, sd_s, rtol=1e-5, atol=1e-6 * np.sqrt(sf2))


def test_interpolation_and_floor():
    rng = np.random.default_rng(1)
    X = rng.random((40, 3)); y = X.sum(1)
    st = O.gp_fit_state(X, y, 0.7 * np.ones(3), 1.5)
    mu, var = O.gp_posterior(st, X)
    np.testing.assert_allclose(mu, y, atol=1e-5)
    assert var.min() >= 1e-15 and var.max() < 1e-6


def test_rbf_kernel():
    rng = np.random.default_rng(2)
    X = rng.random((30, 4)); y = np.cos(X.sum(1))
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, RBF
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(1.3) * RBF(length_scale=0.9 * np.ones(4)),
                                   alpha=1e-8, optimizer=None).fit(X, y)
    Xs = rng.random((100, 4))
    mu_s, sd_s = gpr.predict(Xs, return_std=True)
    st = O.gp_fit_state(X, y, 0.9 * np.ones(4), 1.3, kernel=O.KERNEL_RBF)
    mu, var = O.gp_posterior(st, Xs)
    np.testing.assert_allclose(mu, mu_s, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.sqrt(var), sd_s, rtol=1e-3, atol=1e-5)
