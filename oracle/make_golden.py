"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN
CODE (oracle/ref_loader.py) on seeded inputs.  Run in the build container only
(`python -m oracle.make_golden`); /root/reference does not exist on the GPU box, the
committed fixtures travel instead.

Every "out_*" array in the fixtures is the return value of an unmodified reference
function; the "in_*" arrays are the inputs it was called with.  The reference evaluates
one candidate per call, so posterior values are injected through
a table-lookup shim model: candidate x = [row_index, 0, ...].
"""
from __future__ import annotations

import os

import numpy as np
from scipy.stats import norm, qmc

from oracle.ref_loader import ShimGP, load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sobol_cache(k, exponent, seed):
    """optimisers.py:121-141 with a seed (the reference is unseeded)."""
    s = qmc.Sobol(d=k, scramble=True, seed=seed).random_base2(m=exponent)
    return np.asarray([norm.ppf(s[:, i]) for i in range(k)]).T.copy()


def table_models(mu, var):
    """k shim models whose predict(np.asarray([x])) returns row int(x[0]) of (mu, var)."""
    models = []
    for i in range(mu.shape[1]):
        def post(X, i=i):
            r = np.asarray(X)[:, 0].astype(int)
            return mu[r, i], var[r, i]
        models.append(ShimGP(post))
    return models


def main():
    R = load_reference()
    uf, sc = R.util_functions, R.scalarisations
    rng = np.random.default_rng(20261018)
    os.makedirs(OUT, exist_ok=True)
    G = {}

    # ---------------- anchors quoted in SURVEY section 8c ---------------------------
    PF3 = np.array([[.1, .9], [.4, .5], [.8, .2]])
    G["anchor_ehvi2d_aux"] = np.asarray(uf.EHVI_2D_aux(PF3, [1, 1], [.3, .6], [.2, .3])).reshape(-1)
    G["anchor_wfg"] = np.asarray([uf.wfg(PF3.tolist(), [1, 1])])
    G["anchor_cells"] = uf.decompose_into_cells(PF3, [0, 0], [1, 1])

    # ---------------- a7: scalarisations ----------------------------------------
    names = ["WeightedSum", "Tchebicheff", "AugmentedTchebicheff", "ModifiedTchebicheff",
             "ExponentialWeightedCriterion", "WeightedNorm", "WeightedPower", "WeightedProduct",
             "PBI", "IPBI", "QPBI", "APD"]
    for k in (2, 3):
        F = rng.uniform(-0.2, 1.4, size=(48, k)) * np.array([700, 12, 3.0][:k])
        F[0] = np.array([0.0, 0.0, 0.0][:k])             # APD zero-vector guard (:390)
        ideal = np.array([0.0, 0.0, 0.0][:k])
        maxp = np.array([700.0, 12.0, 3.0][:k])
        w = rng.dirichlet(np.ones(k))
        G[f"sc{k}_in_F"], G[f"sc{k}_in_w"] = F, w
        G[f"sc{k}_in_ideal"], G[f"sc{k}_in_max"] = ideal, maxp
        for nme in names:
            kw = dict(p=8) if nme == "ExponentialWeightedCriterion" else {}
            obj = getattr(sc, nme)(ideal, maxp, **kw)
            G[f"sc{k}_out_{nme}_batch"] = np.asarray(obj(F.copy(), w)).reshape(-1)
            G[f"sc{k}_out_{nme}_single"] = np.asarray([obj(f.copy(), w)[0] for f in F])

    # ---------------- posterior tables used by the acquisition goldens ----------
    m = 96
    mu2 = np.column_stack([rng.uniform(0.0, 1.0, m), rng.uniform(0.0, 6.0, m)])
    var2 = np.column_stack([rng.uniform(1e-4, 0.08, m), rng.uniform(1e-3, 0.9, m)])
    var2[:4, 0] = 1e-15                                   # GPy variance floor
    X2 = np.column_stack([np.arange(m), np.zeros(m)])
    cache2 = sobol_cache(2, 5, seed=0)
    Ytrain = np.column_stack([rng.uniform(0, 1, 40), rng.uniform(0, 6, 40)])
    Ytrain[:, 1] = 6.0 * (1 - np.sqrt(Ytrain[:, 0])) + rng.uniform(0, 1.5, 40)
    PF = uf.calc_pf(Ytrain)
    r = Ytrain.max(0)
    G.update(acq_in_mu2=mu2, acq_in_var2=var2, acq_in_cache2=cache2, acq_in_Y=Ytrain,
             acq_out_calc_pf=PF, acq_in_ref=r)
    models2 = table_models(mu2, var2)

    # a3: EHVI (reference, 2-D)
    G["acq_out_EHVI"] = np.asarray([uf.EHVI(x, models2, r, PF, cache2)[0] for x in X2])
    # a4: EHVI_2D_aux with true stds (for the 'exact' mode minus its extra stripe)
    G["acq_out_EHVI_2D_aux_truestd"] = np.asarray(
        [uf.EHVI_2D_aux(PF, r, mu2[i], np.sqrt(var2[i]))[0] for i in range(m)])

    # a6: expected_decomposition for the 12 scalarisations (S=32 and S=8)
    ideal2, max2 = np.array([0.0, 0.0]), np.array([1.2, 7.5])
    w2 = np.array([0.37, 0.63])
    G.update(ed_in_ideal=ideal2, ed_in_max=max2, ed_in_w=w2)
    cache8 = sobol_cache(2, 3, seed=1)
    G["ed_in_cache8"] = cache8
    for nme in names:
        kw = dict(p=8) if nme == "ExponentialWeightedCriterion" else {}
        obj = getattr(sc, nme)(ideal2, max2, **kw)
        gmin = float(np.min([obj(y, w2) for y in Ytrain]))            # optimisers.py:250
        G[f"ed_in_gmin_{nme}"] = np.asarray([gmin])
        G[f"ed_out_{nme}"] = np.asarray(
            [uf.expected_decomposition(x, models2, w2, obj, gmin, cache2) for x in X2])
        G[f"ed8_out_{nme}"] = np.asarray(
            [uf.expected_decomposition(x, models2, w2, obj, gmin, cache8) for x in X2])

    # a5: EHVI_3D (reference box chosen so every sample is inside; pygmo raises otherwise)
    mu3 = rng.uniform(0.1, 0.9, size=(m, 3))
    var3 = rng.uniform(1e-4, 0.02, size=(m, 3))
    cache3 = sobol_cache(3, 5, seed=2)
    Y3 = rng.uniform(0.2, 1.0, size=(30, 3))
    PF3d = uf.calc_pf(Y3)
    r3 = np.array([2.5, 2.5, 2.5])
    models3 = table_models(mu3, var3)
    X3 = np.column_stack([np.arange(m), np.zeros(m), np.zeros(m)])
    from oracle.ref_loader import _hv_exact
    G.update(e3_in_mu=mu3, e3_in_var=var3, e3_in_cache=cache3, e3_in_pf=PF3d, e3_in_ref=r3,
             e3_in_sminus=np.asarray([_hv_exact(PF3d, r3)]))
    G["e3_out_EHVI_3D"] = np.asarray([uf.EHVI_3D(x, models3, r3, PF3d, cache3) for x in X3])

    # a8: EI, the five copies (two distinct epsilons)
    class P:  # minimal problem stub for the optimiser constructors
        n_var, n_obj, xl, xu = 2, 2, np.zeros(2), np.ones(2)
        n_ieq_constr, n_eq_constr = 2, 0
    best = 0.45
    G["ei_in_best"] = np.asarray([best])
    mono = R.optimisers.MonoSurrogateOptimiser(P(), [0, 0], [1, 1])
    G["ei_out_mono"] = np.asarray([mono._expected_improvement(x, models2[0], best)[0] for x in X2])
    par = R.parego.ParEGO(P(), [0, 0], [1, 1])
    G["ei_out_parego"] = np.asarray([par._expected_improvement(x, models2[0], best)[0] for x in X2])
    keep = R.keep.KEEP(P(), [0, 0], [1, 1])
    G["ei_out_keep"] = np.asarray([keep._expected_improvement(x, models2[0], best)[0] for x in X2])
    # a10: KEEP pareto EI: pareto model = models2[1] (mean only), scalar model = models2[0]
    G["pei_out_keep"] = np.asarray(
        [np.asarray(keep.pareto_expected_improvement(x, models2[1], models2[0], best)).reshape(-1)[0]
         for x in X2])
    # a9: ParEGO_C2 EI x prod PoF: agg model + 2 constraint models
    muc = np.column_stack([mu2[:, 0], rng.uniform(-1.0, 1.0, m), rng.uniform(-0.5, 0.5, m)])
    varc = np.column_stack([var2[:, 0], rng.uniform(1e-6, 0.3, m), rng.uniform(1e-6, 0.3, m)])
    modelsc = table_models(muc, varc)
    c2 = R.cparego.ParEGO_C2(P(), [0, 0], [1, 1])
    G.update(cei_in_mu=muc, cei_in_var=varc)
    G["cei_out_c2"] = np.asarray([c2.consraint_ei(x, modelsc[0], modelsc[1:], best)[0] for x in X2])
    G["pof_out_c2"] = np.asarray(
        [np.asarray(c2.probability_of_feasibility(x, modelsc[1])).reshape(-1)[0] for x in X2])

    # a11: EMO hypervolume-based PoI, cells from the reference's own decomposition
    emo = R.emo.EMO(P(), np.array([0.0, 0.0]), np.array([1.1, 7.5]))
    cells = emo.decompose_into_cells(PF)
    G["emo_in_ideal"], G["emo_in_max"] = emo.ideal_point, emo.max_point
    G["emo_out_cells"] = cells
    G["emo_out_poi"] = np.asarray(
        [emo.hypervolume_based_PoI(x, models2, Ytrain, cells) for x in X2])
    # module-level twin (util_functions.py:414) on a few random fronts
    for t in range(3):
        Yt = rng.uniform(0, 1, size=(12 + 5 * t, 2))
        pf_t = uf.calc_pf(Yt)
        G[f"cells{t}_in_pf"] = pf_t
        G[f"cells{t}_out"] = uf.decompose_into_cells(pf_t, [0.0, 0.0], [1.0, 1.0])
        G[f"cells{t}_wfg"] = np.asarray([uf.wfg(sorted(pf_t.tolist()), [1.0, 1.0])])

    np.savez_compressed(os.path.join(OUT, "acq_golden.npz"), **G)
    print("wrote", os.path.join(OUT, "acq_golden.npz"), len(G), "arrays")


if __name__ == "__main__":
    main()
