"""Hyper-parameter fit of the GP surrogate on the host (outside the hot path).

The reference fits every model with GPy: `GPRegression(X, y, Matern52(ARD)); noise fixed to 0;
model.optimize(max_f_eval=1000)` (optimisers.py:226-231 and the sites in SURVEY section 0.1), i.e.
L-BFGS on the exact marginal likelihood starting from variance = 1, lengthscale = 1.  This is the
same objective in log-parameters with analytic gradients (float64 numpy / LAPACK); the refreshed
state (L, L^-1, alpha) is then rebuilt ON THE DEVICE by `GPModel.refresh()` (K3).
"""
from __future__ import annotations

import numpy as np
from scipy import linalg as sla
from scipy.optimize import minimize

_S5 = np.sqrt(5.0)


def _nlml_and_grad(theta, X, y, kernel, jitter):
    n, d = X.shape
    sf2 = np.exp(theta[0])
    ell = np.exp(theta[1:])
    Z = X / ell
    diff2 = (Z[:, None, :] - Z[None, :, :]) ** 2              # (n, n, d)
    r2 = diff2.sum(-1)
    if kernel == "matern52":
        r = np.sqrt(r2)
        e = np.exp(-_S5 * r)
        K0 = (1.0 + _S5 * r + (5.0 / 3.0) * r2) * e
        dK_dr2_scaled = (5.0 / 3.0) * (1.0 + _S5 * r) * e       # = -d k0 / d(log ell_j) / diff2_j
    else:
        K0 = np.exp(-0.5 * r2)
        dK_dr2_scaled = K0
    K = sf2 * K0 + jitter * np.eye(n)
    try:
        L = sla.cholesky(K, lower=True)
    except sla.LinAlgError:
        return 1e25, np.zeros_like(theta)
    alpha = sla.cho_solve((L, True), y)
    nlml = 0.5 * y @ alpha + np.log(np.diag(L)).sum() + 0.5 * n * np.log(2 * np.pi)
    Kinv = sla.cho_solve((L, True), np.eye(n))
    W = Kinv - np.outer(alpha, alpha)                           # dNLML/dK = 0.5 W
    g = np.empty_like(theta)
    g[0] = 0.5 * np.sum(W * (sf2 * K0))
    WK = W * (sf2 * dK_dr2_scaled)
    g[1:] = 0.5 * np.einsum("ij,ijk->k", WK, diff2)
    return nlml, g


def fit_hyperparameters(X, y, kernel="matern52", max_f_eval=1000, jitter=1e-8, init=None):
    """Returns (lengthscale (d,), variance).  Start point variance = 1, lengthscale = 1 like GPy."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    d = X.shape[1]
    theta0 = np.zeros(d + 1) if init is None else np.log(np.concatenate(([init[1]], np.asarray(init[0], float))))
    if np.ptp(y) == 0.0:                                        # constant targets: nothing to fit
        return np.exp(theta0[1:]), float(np.exp(theta0[0]))
    res = minimize(_nlml_and_grad, theta0, args=(X, y, kernel, jitter), jac=True, method="L-BFGS-B",
                   bounds=[(-12.0, 14.0)] + [(-7.0, 9.0)] * d, options=dict(maxfun=max_f_eval))
    theta = res.x if np.isfinite(res.fun) else theta0
    return np.exp(theta[1:]), float(np.exp(theta[0]))


def fit_hyperparameters_device(X, y, kernel="matern52", max_f_eval=1000, jitter=1e-8, init=None, device="cuda:0",
                               info=None):
    """Same optimisation with every likelihood / gradient evaluation on the GPU
    (`ombo_gp_nlml_grad`: K3 refresh + K^-1 + gradient reduction, FP64); scipy's L-BFGS-B drives it from
    the host.  Returns (lengthscale (d,), variance); `info` (a dict) receives scipy's evaluation count `nfev`."""
    import ctypes as C

    import torch

    from . import _cabi
    from .gp import _KERNELS, _as_device, current_stream_ptr
    dev = _as_device(device)
    Xd = torch.as_tensor(np.asarray(X, dtype=np.float64)).to(dev).contiguous()
    yd = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1)).to(dev).contiguous()
    n, d = Xd.shape
    theta0 = np.zeros(d + 1) if init is None else np.log(np.concatenate(([init[1]], np.asarray(init[0], float))))
    if float(yd.max() - yd.min()) == 0.0:
        return np.exp(theta0[1:]), float(np.exp(theta0[0]))
    state = torch.empty(_cabi.state_bytes(n, d), dtype=torch.uint8, device=dev)
    ctx = _cabi.Context.get(dev.index)
    out = (C.c_double * (d + 2))()

    def fun(theta):
        ell = (C.c_double * d)(*np.exp(theta[1:]).tolist())
        spec = _cabi.GpSpec(n=n, d=d, kernel=_KERNELS[kernel], reserved=0, sigma_f2=float(np.exp(theta[0])),
                            sigma_n2=0.0, jitter=jitter, X=Xd.data_ptr(), y=yd.data_ptr(), ell=ell)
        with torch.cuda.device(dev):
            rc = _cabi.lib().ombo_gp_nlml_grad(ctx.handle, C.byref(spec), C.c_void_p(state.data_ptr()), out,
                                               current_stream_ptr(dev))
        if rc == _cabi.ERR_NOT_PD:
            return 1e25, np.zeros_like(theta)
        _cabi.check(rc)
        v = np.frombuffer(out, dtype=np.float64, count=d + 2).copy()
        return float(v[0]), v[1:]

    res = minimize(fun, theta0, jac=True, method="L-BFGS-B", bounds=[(-12.0, 14.0)] + [(-7.0, 9.0)] * d,
                   options=dict(maxfun=max_f_eval))
    if info is not None:
        info["nfev"] = int(res.nfev)
    theta = res.x if np.isfinite(res.fun) and res.fun < 1e24 else theta0
    return np.exp(theta[1:]), float(np.exp(theta[0]))
