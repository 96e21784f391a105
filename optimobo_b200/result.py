"""Result holders with the reference's attribute names (result.py:4,55; README.md:92-103)."""
from __future__ import annotations


class Res:
    def __init__(self, pf_approx, pf_inputs, ysample, Xsample, hypervolume_convergence, n_obj, n_init_samples):
        self.pf_approx = pf_approx
        self.pf_inputs = pf_inputs
        self.ysample = ysample
        self.Xsample = Xsample
        self.xsample = Xsample          # README spelling
        self.hypervolume_convergence = hypervolume_convergence
        self.n_obj = n_obj
        self.n_init_samples = n_init_samples
        self.timings = []               # per-iteration dict(fit_s, refresh_s, score_s) added by the B200 path

    def plot_pareto_front(self):        # plotting needs matplotlib, which is optional here
        import matplotlib.pyplot as plt
        n0 = self.n_init_samples
        if self.n_obj == 2:
            plt.scatter(self.ysample[n0:, 0], self.ysample[n0:, 1], color="red", label="Samples.")
            plt.scatter(self.pf_approx[:, 0], self.pf_approx[:, 1], color="green", label="PF approximation.")
            plt.scatter(self.ysample[:n0, 0], self.ysample[:n0, 1], color="blue", label="Initial samples.")
            plt.xlabel(r"$f_1(x)$"); plt.ylabel(r"$f_2(x)$"); plt.legend()
        else:
            ax = plt.figure().add_subplot(projection="3d")
            ax.scatter(*self.ysample[n0:, :3].T, color="red", label="Samples.")
            ax.scatter(*self.pf_approx[:, :3].T, color="green", label="PF approximation.")
            ax.legend()

    def plot_hv_convergence(self):
        import matplotlib.pyplot as plt
        plt.plot(self.hypervolume_convergence)


class Constrained_Res(Res):
    def __init__(self, y_infeasible, y_feasible, X_infeasible, X_feasible, pf_approx, pf_inputs, ysample, Xsample,
                 hypervolume_convergence, n_obj, n_init_samples):
        super().__init__(pf_approx, pf_inputs, ysample, Xsample, hypervolume_convergence, n_obj, n_init_samples)
        self.X_infeasible, self.X_feasible = X_infeasible, X_feasible
        self.y_feasible, self.y_infeasible = y_feasible, y_infeasible
