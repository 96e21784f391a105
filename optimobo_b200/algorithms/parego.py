"""ParEGO (parego.py:148-298): GP on a randomly weighted scalarisation, EI with the reference's
sqrt(var + 1e-6) (parego.py:139).  The reference maximises EI with a 20-member EA over 1000
generations (26,000 single-x predicts per iteration); here the pool arg-max does it in one pass."""
from __future__ import annotations

import numpy as np

from .. import host_prep, result
from ..acquisition import spec_ei
from .base import PoolOptimiserBase


class ParEGO(PoolOptimiserBase):
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
            model = self._fit_model(Xsample, aggregated_samples)
            current_best = aggregated_samples[np.argmin(aggregated_samples)]
            next_X, _ = self._propose([model], spec_ei(current_best, 1e-6))
            next_y = self._objective_function(problem, next_X)
            ysample = np.vstack((ysample, next_y))
            Xsample = np.vstack((Xsample, next_X))
        mask = self._pareto_members(ysample)
        res = result.Res(ysample[mask], Xsample[mask], ysample, Xsample, hypervolume_convergence, problem.n_obj,
                         n_init_samples)
        res.timings = self.timings
        return res
