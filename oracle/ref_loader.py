"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *reference's own* acquisition arithmetic (aje220/OptiMOBO,
`optimobo/util_functions.py`, `optimobo/scalarisations.py` and the
`optimobo/algorithms/*.py` methods) inside this container so that golden
vectors can be minted from it and the numpy restatement in `oracle/` can be
pinned against it.

The reference imports GPy, pygmo, pymoo and matplotlib at module import time
(util_functions.py:3-4, scalarisations.py:2, optimisers.py:6-8,13); none of
them is installed here and there is no network, so they are stubbed in
`sys.modules`.  Nothing from the stubs is ever *used* by the functions we
call, except:
  * `pygmo.fast_non_dominated_sorting` (util_functions.py:76) -> restated
    O(n^2) dominance filter returning the first front, and
  * `pygmo.hypervolume(points).compute(r)` (util_functions.py:198-206) ->
    restated exact hypervolume (2-D sweep / 3-D slicing).
numpy >= 2 removed `np.product` (util_functions.py:410, emo.py:220): aliased.

The reference tree does not exist on the GPU box; callers must handle
`reference_available() == False` (tests skip, fixtures are committed).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

# Lookup order: $OPTIMOBO_REF, the read-only tree of the build container, and `baseline/_ref/` -- the
# unmodified reference package installed by `python -m pip install --no-index --no-build-isolation --no-deps
# --target baseline/_ref <copy of /root/reference>` (recorded in DESIGN.md; git-ignored, NOT gpurun-ignored, so it
# travels to the GPU box and the reference's own call pattern can be timed there).
_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_CANDIDATES = [os.environ.get("OPTIMOBO_REF", ""), "/root/reference", os.path.join(_REPO_ROOT, "baseline", "_ref")]


def reference_root():
    for p in _REF_CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "optimobo", "util_functions.py")):
            return p
    return None


def reference_available() -> bool:
    return reference_root() is not None


# ----------------------------------------------------------------------------
# restated third-party pieces the reference calls (pygmo is absent)
# ----------------------------------------------------------------------------
def _first_front_indices(Y):
    """Indices of the first non-dominated front (minimisation), pygmo order
    semantics: ascending original index."""
    Y = np.asarray(Y, dtype=float)
    n = len(Y)
    keep = []
    for i in range(n):
        dominated = False
        for j in range(n):
            if j == i:
                continue
            if np.all(Y[j] <= Y[i]) and np.any(Y[j] < Y[i]):
                dominated = True
                break
        if not dominated:
            keep.append(i)
    return keep


def _fast_non_dominated_sorting(points):
    idx = _first_front_indices(points)
    return [np.asarray(idx)], None, None, None


def _hv_exact(points, ref):
    """Exact hypervolume (minimisation) for 1-3 objectives by slicing."""
    P = np.asarray(points, dtype=float)
    ref = np.asarray(ref, dtype=float)
    if P.ndim == 1:
        P = P[None, :]
    if np.any(P > ref):
        # pygmo raises ValueError for points outside the reference point
        raise ValueError("A reference point is invalid: a point is outside it")
    k = P.shape[1]
    if k == 1:
        return float(ref[0] - P[:, 0].min())
    if k == 2:
        Q = P[np.argsort(P[:, 0])]
        hv, best = 0.0, ref[1]
        for a, b in Q:
            if b < best:
                hv += (ref[0] - a) * (best - b)
                best = b
        return float(hv)
    # k >= 3: slice along the last objective
    order = np.argsort(P[:, -1])
    Q = P[order]
    hv = 0.0
    for i in range(len(Q)):
        z_lo = Q[i, -1]
        z_hi = Q[i + 1, -1] if i + 1 < len(Q) else ref[-1]
        if z_hi > z_lo:
            hv += _hv_exact(Q[: i + 1, :-1], ref[:-1]) * (z_hi - z_lo)
    return float(hv)


class _Hypervolume:
    def __init__(self, points):
        self.points = np.asarray(points, dtype=float)

    def compute(self, ref):
        return _hv_exact(self.points, ref)


def _install_stubs():
    if "optimobo_ref_stubs_installed" in sys.modules:
        return
    if not hasattr(np, "product"):
        np.product = np.prod  # removed in numpy 2

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = mod("matplotlib")
            mpl.pyplot = mod("matplotlib.pyplot")
    mod("pygmo", fast_non_dominated_sorting=_fast_non_dominated_sorting,
        hypervolume=_Hypervolume)
    gpy = mod("GPy")
    gpy.plotting = mod("GPy.plotting", change_plotting_library=lambda *a, **k: None)
    gpy.models = mod("GPy.models")
    gpy.kern = mod("GPy.kern")
    pymoo = mod("pymoo")
    pymoo.util = mod("pymoo.util")
    pymoo.util.ref_dirs = mod("pymoo.util.ref_dirs", get_reference_directions=None)
    pymoo.indicators = mod("pymoo.indicators")
    pymoo.indicators.hv = mod("pymoo.indicators.hv", HV=None)
    pymoo.gradient = mod("pymoo.gradient")
    pymoo.gradient.toolbox = mod("pymoo.gradient.toolbox")
    pymoo.util.cache = mod("pymoo.util.cache", Cache=lambda f: f)
    pymoo.util.misc = mod("pymoo.util.misc", at_least_2d_array=None)
    sys.modules["optimobo_ref_stubs_installed"] = types.ModuleType("x")


_loaded = {}


def load_reference():
    """Returns a namespace with the reference modules:
    .util_functions, .scalarisations, .optimisers, .parego, .cparego, .keep, .emo
    """
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected $OPTIMOBO_REF, /root/reference or baseline/_ref)")
    _install_stubs()
    # Import under the reference's own package name.  The product package is
    # called `optimobo_b200`, so there is no clash.
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    _loaded["util_functions"] = importlib.import_module("optimobo.util_functions")
    _loaded["scalarisations"] = importlib.import_module("optimobo.scalarisations")
    for name in ("optimisers", "parego", "cparego", "keep", "emo"):
        try:
            _loaded[name] = importlib.import_module("optimobo.algorithms." + name)
        except Exception as e:  # pragma: no cover - diagnostic only
            _loaded[name] = None
            _loaded[name + "_error"] = repr(e)
    return types.SimpleNamespace(**_loaded)


class ShimGP:
    """GPy-shaped model: predict(X (m,d)) -> (mean (m,1), var (m,1)), backed by a
    python callable posterior(X) -> (mu (m,), var (m,)).  Used to drive the
    reference's unmodified functions (they call `model.predict(np.asarray([X]))`)."""

    def __init__(self, posterior):
        self._post = posterior

    def predict(self, X, return_std=False):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        mu, var = self._post(X)
        mu = np.asarray(mu, dtype=float).reshape(-1)
        var = np.asarray(var, dtype=float).reshape(-1)
        if return_std:  # sklearn surface (util_functions.py:265)
            return mu, np.sqrt(np.maximum(var, 0.0))
        return mu.reshape(-1, 1), var.reshape(-1, 1)
