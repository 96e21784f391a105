"""KEEP (keep.py:153-320): ParEGO plus a second GP on the 0/1 Pareto-membership labels;
fitness = pareto_mean(x) * EI_scalar(x) (keep.py:142-150), EI with sqrt(var + 1e-6)."""
from __future__ import annotations

import numpy as np

from .. import host_prep, result
from ..acquisition import spec_pareto_ei
from .base import PoolOptimiserBase


class KEEP(PoolOptimiserBase):
    def solve(self, aggregation_func, budget=100, n_init_samples=5):
        problem = self.test_problem
        Xsample, ysample = self._initial_design(n_init_samples)
        ref_dirs = host_prep.get_reference_directions("das-dennis", problem.n_obj, n_partitions=100)
        hypervolume_convergence = []
        for _ in range(budget):
            self._update_bounds(ysample, aggregation_func)
            hypervolume_convergence.append(self._hypervolume(ysample))
            ref_dir = ref_dirs[self.rng.integers(0, len(ref_dirs))]
            aggregated_samples = np.asarray([aggregation_func(y, ref_dir) for y in ysample]).flatten()
            scalar_model = self._fit_model(Xsample, aggregated_samples)
            probs = self._pareto_members(ysample).astype(float)          # keep.py:226-233
            pareto_model = self._fit_model(Xsample, probs)
            current_best = aggregated_samples[np.argmin(aggregated_samples)]
            next_X, _ = self._propose([pareto_model, scalar_model], spec_pareto_ei(current_best))
            next_y = self._objective_function(problem, next_X)
            ysample = np.vstack((ysample, next_y))
            Xsample = np.vstack((Xsample, next_X))
        mask = self._pareto_members(ysample)
        res = result.Res(ysample[mask], Xsample[mask], ysample, Xsample, hypervolume_convergence, problem.n_obj,
                         n_init_samples)
        res.timings = self.timings
        return res
