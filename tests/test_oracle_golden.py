"""CPU: the numpy restatement (oracle/oracle.py) against fixtures minted by executing the
reference's own code (oracle/make_golden.py -> tests/golden/acq_golden.npz), plus the
hand-checkable anchors quoted in SURVEY.md section 8c."""
import numpy as np
import pytest

from oracle import oracle as O

RTOL = 1e-9   # restatement vs reference: same float64 formulas, different evaluation order


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_anchors(golden):
    PF = np.array([[.1, .9], [.4, .5], [.8, .2]])
    ref = O.ehvi2d_aux_batched(PF, [1, 1], np.array([.3]), np.array([.6]), np.array([.2]), np.array([.3]))
    close(ref, golden["anchor_ehvi2d_aux"])
    close(ref, [0.0700010532552525], rtol=1e-12)
    exact = O.ehvi2d_aux_batched(PF, [1, 1], np.array([.3]), np.array([.6]), np.array([.2]),
                                 np.array([.3]), exact=True)
    close(exact, [0.0768782210398855], rtol=1e-12)     # textbook n+1 stripes
    close(O.hypervolume(PF, [1, 1]), 0.39, rtol=1e-12)
    close(O.hypervolume(PF, [1, 1]), golden["anchor_wfg"][0], rtol=1e-12)
    close(O.decompose_into_cells_2d(PF, [0, 0], [1, 1]), golden["anchor_cells"], rtol=0, atol=0)


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("name", O.SCALARISATIONS)
def test_scalarisations(golden, k, name):
    F, w = golden[f"sc{k}_in_F"], golden[f"sc{k}_in_w"]
    kw = dict(p=8) if name == "ExponentialWeightedCriterion" else {}
    got = O.scalarise(name, F, w, golden[f"sc{k}_in_ideal"], golden[f"sc{k}_in_max"], **kw)
    close(got, golden[f"sc{k}_out_{name}_batch"])
    close(got, golden[f"sc{k}_out_{name}_single"])


def test_scalarisation_survey_values():
    F, w = np.array([0.2, 0.7]), np.array([0.3, 0.7])
    want = dict(WeightedSum=0.55, Tchebicheff=0.49, AugmentedTchebicheff=0.49009, ModifiedTchebicheff=1.12,
                WeightedNorm=0.62359685, WeightedPower=0.2425, WeightedProduct=100000.54999974,
                PBI=1.1817579, IPBI=-0.26261287, QPBI=0.76442676, APD=2.5167245,
                ExponentialWeightedCriterion=2.32773204e+60)
    for n, v in want.items():
        close(O.scalarise(n, F, w, [0, 0], [1, 1]), v, rtol=2e-8)


def test_calc_pf(golden):
    close(O.calc_pf(golden["acq_in_Y"]), golden["acq_out_calc_pf"], rtol=0)


def test_ehvi_reference(golden):
    mu, var = golden["acq_in_mu2"], golden["acq_in_var2"]
    got = O.ehvi_batched(mu[:, 0], mu[:, 1], var[:, 0], var[:, 1], golden["acq_out_calc_pf"],
                         golden["acq_in_ref"], golden["acq_in_cache2"], "reference")
    want = golden["acq_out_EHVI"]
    ok = np.isfinite(want)
    assert ok.sum() > 80
    # np.cov of the affine-mapped samples vs var0*C: rounding differs -> 1e-7
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-7, atol=1e-9 * np.abs(want[ok]).max())
    assert np.array_equal(np.isnan(got), np.isnan(want))


def test_ehvi_truestd(golden):
    mu, var = golden["acq_in_mu2"], golden["acq_in_var2"]
    got = O.ehvi2d_aux_batched(golden["acq_out_calc_pf"], golden["acq_in_ref"], mu[:, 0], mu[:, 1],
                               np.sqrt(var[:, 0]), np.sqrt(var[:, 1]))
    close(got, golden["acq_out_EHVI_2D_aux_truestd"], rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("name", O.SCALARISATIONS)
def test_expected_decomposition(golden, name):
    mu, var = golden["acq_in_mu2"], golden["acq_in_var2"]
    kw = dict(p=8) if name == "ExponentialWeightedCriterion" else {}
    for key, cache in (("ed_out_", golden["acq_in_cache2"]), ("ed8_out_", golden["ed_in_cache8"])):
        got = O.expected_decomposition_batched(mu, var, golden["ed_in_w"], name, golden["ed_in_ideal"],
                                               golden["ed_in_max"], golden[f"ed_in_gmin_{name}"][0], cache,
                                               "reference", **kw)
        close(got, golden[key + name], rtol=1e-9, atol=1e-14)


def test_ehvi3d(golden):
    got = O.ehvi3d_batched(golden["e3_in_mu"], golden["e3_in_var"], golden["e3_in_ref"],
                           golden["e3_in_sminus"][0], golden["e3_in_cache"])
    close(got, golden["e3_out_EHVI_3D"], rtol=1e-10)
    close(O.hypervolume(golden["e3_in_pf"], golden["e3_in_ref"]), golden["e3_in_sminus"][0], rtol=1e-12)


def test_ei_family(golden):
    mu, var = golden["acq_in_mu2"], golden["acq_in_var2"]
    best = golden["ei_in_best"][0]
    close(O.expected_improvement(mu[:, 0], var[:, 0], best, 0.0), golden["ei_out_mono"], rtol=1e-10, atol=1e-300)
    close(O.expected_improvement(mu[:, 0], var[:, 0], best, 1e-6), golden["ei_out_parego"], rtol=1e-10)
    close(O.expected_improvement(mu[:, 0], var[:, 0], best, 1e-6), golden["ei_out_keep"], rtol=1e-10)
    close(O.pareto_ei(mu[:, 1], mu[:, 0], var[:, 0], best), golden["pei_out_keep"], rtol=1e-10)
    close(O.constrained_ei(golden["cei_in_mu"], golden["cei_in_var"], best), golden["cei_out_c2"],
          rtol=1e-10, atol=1e-300)
    close(O.probability_of_feasibility(golden["cei_in_mu"][:, 1], golden["cei_in_var"][:, 1]),
          golden["pof_out_c2"], rtol=1e-12)


def test_survey_constant_model_values():
    close(O.expected_improvement(np.array([.5]), np.array([.04]), .6), [0.13955931], rtol=1e-7)
    close(O.expected_improvement(np.array([.5]), np.array([.04]), .6, 1e-6), [0.13956019], rtol=1e-7)
    mu = np.array([[.5, -.1, .2]]); var = np.array([[.04, .01, .02]])
    close(O.constrained_ei(mu, var, .6), [0.0092396], rtol=1e-5)
    close(O.pareto_ei(np.array([.7]), np.array([.5]), np.array([.04]), .6), [0.09769213], rtol=1e-7)
    PF = np.array([[.1, .9], [.4, .5], [.8, .2]])
    cells = O.decompose_into_cells_2d(PF, [0, 0], [1, 1])
    close(O.hv_poi_batched(np.array([[.3, .6]]), np.array([[.04, .09]]), cells), [0.018681436421498922], rtol=1e-9)


def test_emo(golden):
    pf = golden["acq_out_calc_pf"]
    cells = O.decompose_into_cells_2d(pf, golden["emo_in_ideal"], golden["emo_in_max"])
    close(cells, golden["emo_out_cells"], rtol=0)
    got = O.hv_poi_batched(golden["acq_in_mu2"], golden["acq_in_var2"], cells)
    close(got, golden["emo_out_poi"], rtol=1e-10, atol=1e-300)
    for t in range(3):
        close(O.decompose_into_cells_2d(golden[f"cells{t}_in_pf"], [0., 0.], [1., 1.]), golden[f"cells{t}_out"], rtol=0)
        close(O.hypervolume(golden[f"cells{t}_in_pf"], [1., 1.]), golden[f"cells{t}_wfg"][0], rtol=1e-12)


def test_das_dennis():
    W = O.das_dennis(100, 2)
    assert W.shape == (101, 2) and np.allclose(W.sum(1), 1)
    assert O.das_dennis(10, 3).shape == (66, 3)


def test_counter_generator_properties():
    u = O.counter_uniform(1, 0, 4096, 10)
    assert u.shape == (4096, 10) and u.min() >= 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 0.01
    # any shard of the global index range reproduces the same values
    close(O.counter_uniform(1, 1000, 96, 10), u[1000:1096], rtol=0)
    # known-answer: pins the generator for the CUDA twin
    z = O.counter_uniform(7, 5, 1, 3)[0]
    assert np.all(z == O.counter_uniform(7, 0, 6, 3)[5])
