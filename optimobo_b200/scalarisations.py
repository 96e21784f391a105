"""The twelve scalarisation objects of `optimobo.scalarisations` as the plugin surface of the
B200 path (same class names, constructor arguments, `obj(F, weights)`, `set_bounds`).

Host side: each class evaluates itself in float64 numpy (needed by the BO loops for
`min_y g(y, w)`, optimisers.py:250, and for aggregating new samples).  Device side: `device_spec()`
maps the *live* object (bounds mutate every iteration through `set_bounds`,
scalarisations.py:29-34) to the enum + parameter block the CUDA acquisition kernel switches on
(csrc/acquisition.cu::scalarise).  Unknown subclasses raise -- there is no CPU fallback for the
pool scoring.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Scalarisation", "WeightedSum", "Tchebicheff", "AugmentedTchebicheff", "ModifiedTchebicheff",
           "ExponentialWeightedCriterion", "WeightedNorm", "WeightedPower", "WeightedProduct", "PBI",
           "IPBI", "QPBI", "APD"]


class Scalarisation:
    """Base: normalises F' = (F - ideal)/(max - ideal) and defers to `_g(Fp, w)` on (S,k) rows.
    `obj(F, w)` returns a flat array: shape (1,) for one objective vector, (S,) for (S,k) input
    (scalarisations.py:17-27)."""

    sc_id = -1

    def __init__(self, ideal_point=None, max_point=None):
        self.ideal_point = ideal_point
        self.max_point = max_point

    def set_bounds(self, new_lower, new_upper):
        self.ideal_point = new_lower
        self.max_point = new_upper

    def _normalise(self, F):
        F = np.atleast_2d(np.asarray(F, dtype=np.float64))
        lo = np.asarray(self.ideal_point, dtype=np.float64)
        hi = np.asarray(self.max_point, dtype=np.float64)
        return (F - lo) / (hi - lo)

    def _g(self, Fp, w):  # pragma: no cover - abstract
        raise NotImplementedError

    def _do(self, F, weights):
        with np.errstate(all="ignore"):
            return self._g(self._normalise(F), np.asarray(weights, dtype=np.float64).reshape(-1))

    def do(self, F, weights, **kw):
        return np.asarray(self._do(F, weights)).reshape(-1)

    __call__ = do

    def params(self):
        return (0.0, 0.0, 0.0, 0.0)

    def device_spec(self):
        if self.sc_id < 0:
            raise TypeError(f"{type(self).__name__} has no CUDA implementation (and there is no CPU fallback)")
        if self.ideal_point is None or self.max_point is None:
            raise ValueError("scalarisation bounds are not set (ideal_point / max_point)")
        return self.sc_id, tuple(float(v) for v in self.params())


class WeightedSum(Scalarisation):
    sc_id = 0

    def _g(self, Fp, w):
        return (Fp * w).sum(1)


class Tchebicheff(Scalarisation):
    sc_id = 1

    def _g(self, Fp, w):
        return (w * Fp).max(1)


class AugmentedTchebicheff(Scalarisation):
    sc_id = 2

    def __init__(self, ideal_point=None, max_point=None, alpha=0.0001):
        super().__init__(ideal_point, max_point)
        self.alpha = alpha

    def params(self):
        return (self.alpha, 0, 0, 0)

    def _g(self, Fp, w):
        a = np.abs(Fp)
        return (a * w).max(1) + self.alpha * a.sum(1)


class ModifiedTchebicheff(Scalarisation):
    sc_id = 3

    def __init__(self, ideal_point=None, max_point=None, alpha=1):
        super().__init__(ideal_point, max_point)
        self.alpha = alpha

    def params(self):
        return (self.alpha, 0, 0, 0)

    def _g(self, Fp, w):
        a = np.abs(Fp)
        return ((a + self.alpha * a.sum(1, keepdims=True)) * w).max(1)


class ExponentialWeightedCriterion(Scalarisation):
    sc_id = 4

    def __init__(self, ideal_point=None, max_point=None, p=100, **kwargs):
        super().__init__(ideal_point, max_point)
        self.p = p

    def params(self):
        return (self.p, 0, 0, 0)

    def _g(self, Fp, w):
        return (np.exp(self.p * w - 1) * np.exp(self.p * Fp)).sum(1)


class WeightedNorm(Scalarisation):
    sc_id = 5

    def __init__(self, ideal_point=None, max_point=None, p=3):
        super().__init__(ideal_point, max_point)
        self.p = p

    def params(self):
        return (self.p, 0, 0, 0)

    def _g(self, Fp, w):
        return np.power((np.power(np.abs(Fp), self.p) * w).sum(1), 1 / self.p)


class WeightedPower(Scalarisation):
    sc_id = 6

    def __init__(self, ideal_point=None, max_point=None, p=3):
        super().__init__(ideal_point, max_point)
        self.p = p

    def params(self):
        return (self.p, 0, 0, 0)

    def _g(self, Fp, w):
        return ((Fp ** self.p) * w).sum(1)


class WeightedProduct(Scalarisation):
    sc_id = 7

    def _g(self, Fp, w):
        return np.prod((Fp + 100000) ** w, axis=1)


class _PenaltyBoundary(Scalarisation):
    def __init__(self, ideal_point=None, max_point=None, theta=5):
        super().__init__(ideal_point, max_point)
        self.theta = theta

    def params(self):
        return (self.theta, 0, 0, 0)

    @staticmethod
    def _d1_d2(Fp, w):
        wn = w / np.linalg.norm(w)
        d1 = (Fp * wn).sum(1)
        d2 = np.linalg.norm(Fp - d1[:, None] * wn, axis=1)
        return d1, d2


class PBI(_PenaltyBoundary):
    sc_id = 8

    def _g(self, Fp, w):
        d1, d2 = self._d1_d2(Fp, w)
        return d1 + self.theta * d2


class IPBI(_PenaltyBoundary):
    sc_id = 9

    def _g(self, Fp, w):
        d1, d2 = self._d1_d2(Fp, w)
        return self.theta * d2 - d1


class QPBI(_PenaltyBoundary):
    sc_id = 10

    def __init__(self, ideal_point=None, max_point=None, theta=5, alpha=5.0, H=5.0):
        super().__init__(ideal_point, max_point, theta)
        self.alpha = alpha
        self.H = H

    def params(self):
        return (self.theta, self.alpha, self.H, 0)

    def _g(self, Fp, w):
        d1, d2 = self._d1_d2(Fp, w)
        k = Fp.shape[1]
        span = np.sum(np.asarray(self.max_point, float) - np.asarray(self.ideal_point, float))
        d_star = self.alpha * ((1.0 / float(self.H)) * (1.0 / float(k)) * span)
        return d1 + self.theta * d2 * (d2 / d_star)


class APD(Scalarisation):
    sc_id = 11

    def __init__(self, ideal_point=None, max_point=None, FE=1, FE_max=10, gamma=0.010304664101210016):
        super().__init__(ideal_point, max_point)
        self.FE = FE
        self.FE_max = FE_max
        self.gamma = gamma

    def params(self):
        return (self.FE, self.FE_max, self.gamma, 0)

    def _g(self, Fp, w):
        k = Fp.shape[1]
        norm_f = np.linalg.norm(Fp, axis=1)
        f = np.where(np.all(Fp == 0, axis=1, keepdims=True), 1e-5, Fp)
        wv = w if np.any(w != 0) else np.full(k, 1e-5)
        cosang = ((f / np.linalg.norm(f, axis=1, keepdims=True)) * (wv / np.linalg.norm(wv))).sum(1)
        theta = np.arccos(np.clip(cosang, -1.0, 1.0))
        return (1 + k * (self.FE / self.FE_max) * theta / self.gamma) * norm_f
