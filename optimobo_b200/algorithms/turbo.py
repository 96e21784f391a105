"""TuRBO-1 with the reference's surface (turbo.py:17-309) on the device sampling path.

The only GP-consuming step of TuRBO is a joint posterior draw over the trust-region candidates,
`GP.posterior_samples(X_cand, size=batch_size)` (turbo.py:116): up to 5000 candidates, i.e. a
5000 x 5000 posterior covariance and its factorisation per batch.  Here that draw is
`GPModel.posterior_samples` (covariance, Cholesky and product on the device in FP64,
`ombo_posterior_joint_samples`); the trust-region bookkeeping is
host orchestration re-hosted as is.  Additive knobs as for the other optimisers: device, seed,
hyperparameters (skip the fit), max_f_eval, fit_on_device.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.stats import qmc

from .. import host_prep, result
from .base import PoolOptimiserBase


class TuRBO_1(PoolOptimiserBase):
    """https://doi.org/10.48550/arXiv.1910.01739"""

    def __init__(self, test_problem, batch_size, ideal_point=None, max_point=None, **kw):
        super().__init__(test_problem, ideal_point, max_point, **kw)
        self.batch_size = int(batch_size)
        self.n_evals = 0
        self.Xsample = np.zeros((0, self.n_vars))
        self.ysample = np.zeros((0, self.n_obj))
        self.n_cand = min(100 * self.n_vars, 5000)                       # turbo.py:36
        self.length_min, self.length_max, self.length_init = 0.5 ** 7, 1.6, 0.4
        self.length = self.length_init
        self.ref_dirs = host_prep.get_reference_directions("das-dennis", self.n_obj, n_partitions=10)
        self.failtol = np.ceil(np.max([4.0 / self.batch_size, self.n_vars / self.batch_size]))
        self.succtol = 3

    def normalise(self, X):
        return (np.asarray(X) - np.asarray(self.lower)) / (np.asarray(self.upper) - np.asarray(self.lower))

    def denormalise(self, X):
        return X * (self.upper - self.lower) + self.lower

    def create_candidates(self, Xsample, ysample, GP, length):
        """Candidates in the trust region around the incumbent and `batch_size` joint posterior draws over
        them (turbo.py:75-117).  X is scaled to [0,1]^d."""
        assert Xsample.min() >= 0.0 and Xsample.max() <= 1.0
        x_center = Xsample[int(np.argmin(ysample)), :][None, :]
        weights = np.asarray(GP.lengthscale, dtype=float)
        weights = weights / weights.mean()
        weights = weights / np.prod(np.power(weights, 1.0 / len(weights)))
        lb = np.clip(x_center - weights * length / 2.0, 0.0, 1.0)
        ub = np.clip(x_center + weights * length / 2.0, 0.0, 1.0)
        sample = qmc.Sobol(d=self.n_vars, scramble=False).random(n=self.n_cand)
        sample = qmc.scale(sample, lb[0], np.maximum(ub[0], lb[0] + 1e-12))
        prob_perturb = min(20.0 / self.n_vars, 1.0)
        mask = self.rng.random((self.n_cand, self.n_vars)) <= prob_perturb
        ind = np.where(np.sum(mask, axis=1) == 0)[0]
        if len(ind):
            mask[ind, self.rng.integers(0, max(self.n_vars - 1, 1), size=len(ind))] = True
        X_cand = x_center.copy() * np.ones((self.n_cand, self.n_vars))
        X_cand[mask] = sample[mask]
        y_cand = GP.posterior_samples(X_cand, size=self.batch_size, rng=self.rng)
        return X_cand, y_cand

    def _restart(self):
        self._Xsample, self._ysample = [], []
        self.failcount = self.succcount = 0
        self.length = self.length_init

    def _adjust_length(self, fX_next):
        best = np.min(self._aggregated_samples)
        if np.min(fX_next) < best - 1e-3 * math.fabs(best):
            self.succcount += 1
            self.failcount = 0
        else:
            self.succcount = 0
            self.failcount += 1
        if self.succcount == self.succtol:
            self.length = min(2.0 * self.length, self.length_max)
            self.succcount = 0
        elif self.failcount == self.failtol:
            self.length /= 2.0
            self.failcount = 0

    def select_candidates(self, X_cand, y_cand):
        """The minimiser of each sampled function, never the same candidate twice (turbo.py:142-154)."""
        X_next = np.ones((self.batch_size, self.n_vars))
        for i in range(self.batch_size):
            indbest = int(np.argmin(y_cand[:, 0, i]))
            X_next[i, :] = X_cand[indbest, :]
            y_cand[indbest, :] = np.inf
        return X_next

    def get_random_weight(self):
        return self.ref_dirs[self.rng.integers(0, len(self.ref_dirs))]

    def solve(self, aggregation_func, budget=100, n_init_samples=5):
        self.budget = budget
        problem = self.test_problem
        hypervolume_convergence = []
        uniform = np.full(self.n_obj, 1.0 / self.n_obj)                 # the reference hard-wires [0.5, 0.5]
        while self.n_evals < self.budget:
            if len(self.ysample):
                self._update_bounds(self.ysample, aggregation_func)
                hypervolume_convergence.append(self._hypervolume(self.ysample))
            else:
                hypervolume_convergence.append(0.0)
            self._restart()
            Xsample, ysample = self._initial_design(n_init_samples)
            if not len(self.ysample):
                self._update_bounds(ysample, aggregation_func)
            aggregated = np.asarray([aggregation_func(y, uniform) for y in ysample]).flatten()
            self.n_evals += n_init_samples
            self._Xsample, self._ysample = Xsample.copy(), ysample.copy()
            self._aggregated_samples = aggregated.reshape(-1, 1)
            self.Xsample = np.vstack((self.Xsample, Xsample))
            self.ysample = np.vstack((self.ysample, ysample))
            while self.n_evals < self.budget and self.length >= self.length_min:
                Xn = self.normalise(self._Xsample)
                GP = self._fit_model(Xn, self._aggregated_samples[:, 0])
                X_cand, y_cand = self.create_candidates(Xn, self._aggregated_samples, GP, self.length)
                X_next = self.denormalise(self.select_candidates(X_cand, y_cand))
                ref_dir = self.get_random_weight()
                y_next = np.asarray([self._objective_function(problem, x) for x in X_next])
                aggregated_next = np.asarray([aggregation_func(y, ref_dir) for y in y_next]).reshape(-1, 1)
                self._adjust_length(aggregated_next)
                self.n_evals += self.batch_size
                self._Xsample = np.vstack((self._Xsample, X_next))
                self._ysample = np.vstack((self._ysample, y_next))
                self._aggregated_samples = np.vstack((self._aggregated_samples, aggregated_next))
                self.Xsample = np.vstack((self.Xsample, X_next))
                self.ysample = np.vstack((self.ysample, y_next))
        mask = self._pareto_members(self.ysample)
        return result.Res(self.ysample[mask], self.Xsample[mask], self.ysample, self.Xsample, hypervolume_convergence,
                          self.n_obj, n_init_samples)


class TuRBO_M(TuRBO_1):
    """TuRBO-m (turbo.py:309-573): `n_trust_regions` local models, every batch takes the best Thompson
    samples across all regions.  Same device path per region as TuRBO_1."""

    def __init__(self, test_problem, ideal_point, max_point, batch_size, n_trust_regions, **kw):
        self.n_trust_regions = int(n_trust_regions)
        super().__init__(test_problem, batch_size, ideal_point, max_point, **kw)
        self.succtol = 3
        self.failtol = max(5, self.n_vars)
        self.aggregated_samples = np.zeros((0, 1))
        self._restart()

    def _restart(self):
        self._idx = np.zeros((0, 1), dtype=int)          # which trust region proposed which sample
        self.failcount = np.zeros(self.n_trust_regions, dtype=int)
        self.succcount = np.zeros(self.n_trust_regions, dtype=int)
        self.length = self.length_init * np.ones(self.n_trust_regions)

    def _adjust_length(self, fX_next, i):
        fX_min = self.ysample[self._idx[:, 0] == i, 0].min()      # the reference's target value (turbo.py:348)
        if fX_next.min() < fX_min - 1e-3 * math.fabs(fX_min):
            self.succcount[i] += 1
            self.failcount[i] = 0
        else:
            self.succcount[i] = 0
            self.failcount[i] += len(fX_next)
        if self.succcount[i] == self.succtol:
            self.length[i] = min(2.0 * self.length[i], self.length_max)
            self.succcount[i] = 0
        elif self.failcount[i] >= self.failtol:
            self.length[i] /= 2.0
            self.failcount[i] = 0

    def _select_candidates(self, X_cand, y_cand):
        X_next = np.zeros((self.batch_size, self.n_vars))
        idx_next = np.zeros((self.batch_size, 1), dtype=int)
        for k in range(self.batch_size):
            i, j = np.unravel_index(np.argmin(y_cand[:, :, k]), (self.n_trust_regions, self.n_cand))
            X_next[k, :] = X_cand[i, j, :]
            idx_next[k, 0] = i
            y_cand[i, j, :] = np.inf
        return X_next, idx_next

    def _init_region(self, i, aggregation_func, weights, n_init_samples):
        Xsample, ysample = self._initial_design(n_init_samples)
        aggregated = np.asarray([aggregation_func(y, weights) for y in ysample]).reshape(-1, 1)
        self.n_evals += n_init_samples
        self._idx = np.vstack((self._idx, i * np.ones((n_init_samples, 1), dtype=int)))
        self.Xsample = np.vstack((self.Xsample, Xsample))
        self.ysample = np.vstack((self.ysample, ysample))
        self.aggregated_samples = np.vstack((self.aggregated_samples, aggregated))

    def solve(self, aggregation_func, budget, n_init_samples):
        self.budget = budget
        assert self.n_trust_regions > 1 and isinstance(budget, int)
        assert budget > self.n_trust_regions * n_init_samples, "Not enough trust regions to do initial evaluations"
        assert budget > self.batch_size, "Not enough evaluations to do a single batch"
        problem = self.test_problem
        hypervolume_convergence = []
        uniform = np.full(self.n_obj, 1.0 / self.n_obj)
        for i in range(self.n_trust_regions):
            self._init_region(i, aggregation_func, self.get_random_weight(), n_init_samples)
        self._update_bounds(self.ysample, aggregation_func)
        while self.n_evals < self.budget:
            self._update_bounds(self.ysample, aggregation_func)
            hypervolume_convergence.append(self._hypervolume(self.ysample))
            X_cand = np.zeros((self.n_trust_regions, self.n_cand, self.n_vars))
            y_cand = np.inf * np.ones((self.n_trust_regions, self.n_cand, self.batch_size))
            for i in range(self.n_trust_regions):
                idx = np.where(self._idx == i)[0]
                Xn = self.normalise(self.Xsample[idx, :])
                aggre = self.aggregated_samples[idx, :]
                GP = self._fit_model(Xn, aggre[:, 0])
                X_cand[i], y_i = self.create_candidates(Xn, aggre, GP, length=self.length[i])
                y_cand[i] = y_i.reshape(self.n_cand, self.batch_size)
            X_next, idx_next = self._select_candidates(X_cand, y_cand)
            X_next = self.denormalise(X_next)
            ref_dir = self.get_random_weight()
            y_next = np.asarray([self._objective_function(problem, x) for x in X_next])
            aggregated_next = np.asarray([aggregation_func(y, ref_dir) for y in y_next]).reshape(-1)
            for i in range(self.n_trust_regions):
                idx_i = np.where(idx_next == i)[0]
                if len(idx_i) > 0:
                    self._adjust_length(aggregated_next[idx_i], i)
            self.n_evals += self.batch_size
            self.Xsample = np.vstack((self.Xsample, X_next))
            self.ysample = np.vstack((self.ysample, y_next))
            self.aggregated_samples = np.vstack((self.aggregated_samples, aggregated_next.reshape(-1, 1)))
            self._idx = np.vstack((self._idx, idx_next))
            for i in range(self.n_trust_regions):
                if self.length[i] < self.length_min:              # converged: restart this region
                    self._idx[self._idx[:, 0] == i, 0] = -1
                    self.length[i] = self.length_init
                    self.succcount[i] = self.failcount[i] = 0
                    self._init_region(i, aggregation_func, uniform, n_init_samples)
        mask = self._pareto_members(self.ysample)
        return result.Res(self.ysample[mask], self.Xsample[mask], self.ysample, self.Xsample, hypervolume_convergence,
                          self.n_obj, n_init_samples)
