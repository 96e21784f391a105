"""Constrained ParEGO: ParEGO_C1 (penalised scalarisation + EI, cparego.py:12-405) and ParEGO_C2
(EI x prod probability-of-feasibility over constraint GPs, cparego.py:408-874).

Only the acquisition seam runs on the GPU (`_expected_improvement` / `consraint_ei` over a candidate
pool).  The penalty / subset-selection bookkeeping is tiny, data-dependent host logic (SURVEY row 5:
out of scope for kernels); it is restated on index arrays instead of the reference's hstack-ed rows,
keeping its formulas -- including the operator precedence of `xi_bar` (cparego.py:761-765) and
`select_current_best` returning the LARGEST feasible scalarised value (cparego.py:498-512).
Deviations, both deliberate: infeasibility scores reuse the stored constraint values instead of
re-evaluating the expensive problem (cparego.py:553-560 notes it should), and the final
feasible/infeasible split covers every evaluated sample.
"""
from __future__ import annotations

import numpy as np

from .. import host_prep, result
from ..acquisition import spec_constrained_ei, spec_ei
from .base import PoolOptimiserBase


class _ConstrainedParEGO(PoolOptimiserBase):
    def __init__(self, test_problem, ideal_point=None, max_point=None, **kw):
        super().__init__(test_problem, ideal_point, max_point, **kw)
        self.n_ieq_constr = test_problem.n_ieq_constr
        self.n_eq_constr = test_problem.n_eq_constr

    # ---- penalisation of infeasible scalarised values (cparego.py:738-801) ------------------
    @staticmethod
    def _penalise(agg, gsample, infeasible, feasible_any):
        agg = agg.copy()
        if not infeasible.any():
            return agg
        v_max = gsample.max(axis=0)
        safe = np.where(v_max > 0, v_max, 1.0)                  # a never-violated constraint contributes 0
        xi = (np.maximum(gsample, 0.0) / safe).sum(axis=1) / gsample.shape[1]      # xi_single
        scores = xi[infeasible]
        lo, hi = agg.min(), agg.max()
        if feasible_any:
            feas_idx = np.flatnonzero(~infeasible)
            s_star = agg[feas_idx[np.argmin(agg[feas_idx])]]
        else:
            inf_idx = np.flatnonzero(infeasible)
            s_star = agg[inf_idx[np.argmin(scores)]]
        s_bar = (agg[infeasible] - lo) / (hi - lo)
        if len(scores) == 1:
            xi_bar = (scores - lo) / (hi - lo)
        else:   # reference precedence: xi - (min / (max - min))
            xi_bar = scores - scores.min() / (scores.max() - scores.min())
        s_dot = np.maximum(agg[infeasible], s_star)
        agg[infeasible] = s_dot + np.exp(2 * (s_bar + xi_bar) - 1) / (np.exp(2) - 1)
        return agg

    # ---- subset selection (cparego.py:548-644) -------------------------------------------------
    def _best_performing(self, idx, N, agg, ysample, ref_dir):
        order = idx[np.argsort(agg[idx], kind="stable")]
        head, rest = order[: N // 2], order[N // 2:]
        if len(rest) == 0:
            return head
        dist = np.linalg.norm(ysample[rest] - ref_dir, axis=1)
        return np.concatenate([head, rest[np.argsort(dist, kind="stable")][: N - N // 2]])

    def select_subset(self, feasible_idx, infeasible_idx, agg, ysample, gsample, ref_dir, N_max):
        H = N_max // 2
        xi = np.maximum(gsample, 0.0).sum(axis=1)
        inf_sorted = infeasible_idx[np.argsort(xi[infeasible_idx], kind="stable")]
        nf, ni = len(feasible_idx), len(infeasible_idx)
        if nf + ni < N_max:
            return np.concatenate([feasible_idx, infeasible_idx])
        if ni == 0:
            return self._best_performing(feasible_idx, N_max, agg, ysample, ref_dir)
        if nf == 0:
            first = inf_sorted[:H]
            rest = np.setdiff1d(infeasible_idx, first)
            return np.concatenate([first, self._best_performing(rest, N_max - len(first), agg, ysample, ref_dir)])
        if ni >= H and nf >= H:
            first = self._best_performing(feasible_idx, H, agg, ysample, ref_dir)
            return np.concatenate([first, inf_sorted[: N_max - len(first)]])
        if ni < H <= nf:
            return np.concatenate([infeasible_idx,
                                   self._best_performing(feasible_idx, N_max - ni, agg, ysample, ref_dir)])
        return np.concatenate([feasible_idx, inf_sorted[: N_max - nf]])

    def select_current_best(self, feasible_idx, infeasible_idx, agg, gsample):
        if len(feasible_idx) == 0:
            xi = np.maximum(gsample[infeasible_idx], 0.0).sum(axis=1)
            return agg[infeasible_idx[np.argmax(xi)]]
        return agg[feasible_idx[np.argmax(agg[feasible_idx])]]

    # ---- shared loop -----------------------------------------------------------------------------
    def _solve(self, aggregation_func, budget, n_init_samples, N_max, constrained_acquisition):
        self.aggregation_func = aggregation_func
        problem = self.test_problem
        Xsample, ysample = self._initial_design(n_init_samples)
        gsample = np.asarray([self._constraint_function(problem, x) for x in Xsample])
        ref_dirs = host_prep.get_reference_directions("das-dennis", problem.n_obj, n_partitions=10)
        assert budget >= len(ref_dirs), \
            "For " + str(self.n_obj) + " dimensions, the budget must be above " + str(len(ref_dirs))
        hypervolume_convergence = []
        for _ in range(budget // len(ref_dirs)):
            hypervolume_convergence.append(self._hypervolume(ysample, ysample.max(axis=0)))
            ref_dirs = ref_dirs[self.rng.permutation(len(ref_dirs))]
            for ref_dir in ref_dirs:
                aggregation_func.set_bounds(ysample.min(axis=0), ysample.max(axis=0))
                agg = np.asarray([aggregation_func(y, ref_dir) for y in ysample]).flatten()
                unpenalised_best = agg.min()
                infeasible = np.any(gsample > 0, axis=1)
                agg = self._penalise(agg, gsample, infeasible, bool((~infeasible).any()))
                feas_idx, inf_idx = np.flatnonzero(~infeasible), np.flatnonzero(infeasible)
                sel = self.select_subset(feas_idx, inf_idx, agg, ysample, gsample, ref_dir, N_max)
                model_input = Xsample[sel]
                agg_model = self._fit_model(model_input, agg[sel])
                if constrained_acquisition:
                    constraint_models = self._fit_models(model_input, [gsample[sel, c] for c in range(gsample.shape[1])])
                    current_best = self.select_current_best(feas_idx, inf_idx, agg, gsample)
                    next_X, _ = self._propose([agg_model] + constraint_models,
                                              spec_constrained_ei(current_best, len(constraint_models)))
                else:
                    next_X, _ = self._propose([agg_model], spec_ei(unpenalised_best, 0.0))
                ysample = np.vstack((ysample, self._objective_function(problem, next_X)))
                Xsample = np.vstack((Xsample, next_X))
                gsample = np.vstack((gsample, self._constraint_function(problem, next_X)))
        infeasible = np.any(gsample > 0, axis=1)
        y_feasible, y_infeasible = ysample[~infeasible], ysample[infeasible]
        X_feasible, X_infeasible = Xsample[~infeasible], Xsample[infeasible]
        if len(y_feasible):
            mask = self._pareto_members(y_feasible)
            pf_approx, pf_inputs = y_feasible[mask], X_feasible[mask]
        else:
            pf_approx, pf_inputs = y_feasible, X_feasible
        res = result.Constrained_Res(y_infeasible, y_feasible, X_infeasible, X_feasible, pf_approx, pf_inputs, ysample,
                                     Xsample, hypervolume_convergence, problem.n_obj, n_init_samples)
        res.gsample = gsample
        res.timings = self.timings
        return res


class ParEGO_C1(_ConstrainedParEGO):
    def solve(self, aggregation_func, budget=50, n_init_samples=5, N_max=100):
        return self._solve(aggregation_func, budget, n_init_samples, N_max, constrained_acquisition=False)


class ParEGO_C2(_ConstrainedParEGO):
    def solve(self, aggregation_func, budget=10, n_init_samples=5, N_max=100):
        return self._solve(aggregation_func, budget, n_init_samples, N_max, constrained_acquisition=True)
