"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls
without a GPU), the ctypes structs match the C layout, the product fails loudly without a CUDA
device, and the host-side logic (plugin scalarisations, host prep) matches the
reference-minted fixtures."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import optimobo_b200 as ob
from optimobo_b200 import _cabi, host_prep, scalarisations as S
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "optimobo_b200.h")).read()
    declared = set(re.findall(r"\b(ombo_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.EXPORTS)
    assert lib.ombo_abi_version() == 1


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "optimobo_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(ombo_gp_spec),sizeof(ombo_gp),sizeof(ombo_pool),sizeof(ombo_acq),sizeof(ombo_best));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_cabi.GpSpec), C.sizeof(_cabi.Gp), C.sizeof(_cabi.Pool), C.sizeof(_cabi.Acq),
                     C.sizeof(_cabi.Best)]
    assert C.sizeof(_cabi.Best) == 16


def test_state_layout_queries():
    assert _cabi.lib().ombo_n_pad(1) == 128 and _cabi.lib().ombo_n_pad(128) == 128 and _cabi.lib().ombo_n_pad(129) == 256
    total = _cabi.state_bytes(1024, 10)
    seen = []
    for f in range(9):
        off, cnt = _cabi.state_field(1024, 10, f)
        assert off % 256 == 0 and off < total
        seen.append(off)
    assert len(set(seen)) == 9
    with pytest.raises(_cabi.OmboError):
        _cabi.state_bytes(0, 3)
    with pytest.raises(_cabi.OmboError):
        _cabi.state_bytes(10, 33)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.OmboError, match="no CUDA device"):
        _cabi.Context.get(0)
    with pytest.raises(RuntimeError):
        ob.GPModel(np.zeros((3, 2)), np.zeros(3), [1.0, 1.0], device="cpu")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "optimobo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f


# ---- plugin scalarisations (host side) vs fixtures minted from the reference -----------------
@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("name", O.SCALARISATIONS)
def test_plugin_scalarisations(golden, k, name):
    kw = dict(p=8) if name == "ExponentialWeightedCriterion" else {}
    obj = getattr(S, name)(golden[f"sc{k}_in_ideal"], golden[f"sc{k}_in_max"], **kw)
    F, w = golden[f"sc{k}_in_F"], golden[f"sc{k}_in_w"]
    out = obj(F, w)
    assert out.shape == (len(F),)
    np.testing.assert_allclose(out, golden[f"sc{k}_out_{name}_batch"], rtol=1e-9)
    one = obj(F[3], w)
    assert one.shape == (1,)
    np.testing.assert_allclose(one[0], golden[f"sc{k}_out_{name}_single"][3], rtol=1e-9)
    sc_id, params = obj.device_spec()
    assert sc_id == O.SCALARISATIONS.index(name) and len(params) == 4
    obj.set_bounds([1, 1, 1][:k], [2, 2, 2][:k])
    assert list(obj.ideal_point) == [1, 1, 1][:k]


def test_host_prep_matches_oracle_and_golden(golden):
    Y = golden["acq_in_Y"]
    np.testing.assert_array_equal(host_prep.calc_pf(Y), golden["acq_out_calc_pf"])
    pf = golden["acq_out_calc_pf"]
    np.testing.assert_array_equal(host_prep.decompose_into_cells(pf, golden["emo_in_ideal"], golden["emo_in_max"]),
                                  golden["emo_out_cells"])
    for t in range(3):
        np.testing.assert_array_equal(host_prep.decompose_into_cells(golden[f"cells{t}_in_pf"], [0., 0.], [1., 1.]),
                                      golden[f"cells{t}_out"])
        np.testing.assert_allclose(host_prep.hypervolume(golden[f"cells{t}_in_pf"], [1., 1.]),
                                   golden[f"cells{t}_wfg"][0], rtol=1e-12)
    np.testing.assert_allclose(host_prep.hypervolume(golden["e3_in_pf"], golden["e3_in_ref"]),
                               golden["e3_in_sminus"][0], rtol=1e-12)
    y = host_prep.ehvi_stripes(pf, golden["acq_in_ref"])
    y1, y2 = O.ehvi_stripes(pf, golden["acq_in_ref"])
    np.testing.assert_array_equal(y, np.stack([y1, y2]))
    assert host_prep.cache_covariance(golden["acq_in_cache2"]) == O.cache_cov(golden["acq_in_cache2"])
    W = host_prep.get_reference_directions("das-dennis", 2, n_partitions=100)
    assert W.shape == (101, 2)
    np.testing.assert_allclose(np.sort(W, 0), np.sort(O.das_dennis(100, 2), 0))
    assert host_prep.get_reference_directions("das-dennis", 3, n_partitions=10).shape == (66, 3)
    lhs = host_prep.latin_hypercube(20, [(-2, 2), (0, 5)], np.random.default_rng(0))
    assert lhs.shape == (20, 2)
    assert sorted(np.floor((lhs[:, 0] + 2) / 4 * 20).astype(int)) == list(range(20))   # one per stratum
    c = host_prep.cached_samples(2, 5, seed=0)
    assert c.shape == (32, 2)
    np.testing.assert_allclose(c, golden["acq_in_cache2"])


def test_pool_sharding_covers_range():
    pool = ob.CandidatePool.counter(1000003, np.zeros(4), np.ones(4), seed=1)
    shards = [pool.shard(r, 8) for r in range(8)]
    assert sum(s.m for s in shards) == pool.m
    assert shards[0].index_base == 0
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.index_base + a.m == b.index_base
