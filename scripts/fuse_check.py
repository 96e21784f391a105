"""Fused K4 + K5 epilogue of the f8c kernel against the unfused launch sequence (same FP32 EHVI function): the
per-candidate values and the selected candidate must agree; ragged pool sizes, explicit and counter pools."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimobo_b200 as ob
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_parity import make_problem

DEV = "cuda:0"
cases = [(100, 4, 1), (128, 10, 63), (256, 10, 257), (300, 7, 5001), (512, 12, 3000), (700, 7, 148 * 128 + 5),
         (1024, 10, (1 << 20) + 3), (1536, 10, 4096)]
for sem in ("exact", "reference"):
    for n, d, m in cases:
        X, Y, ells, sf2 = make_problem(n, d)
        models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=DEV) for i in range(2)]
        PF, r = ob.host_prep.calc_pf(Y), Y.max(0)
        cache = ob.host_prep.cached_samples(2, 5, seed=0)
        spec = ob.spec_ehvi(r, PF, cache, sem)
        for pool in (ob.CandidatePool.counter(m, np.zeros(d), np.ones(d), seed=3),
                     ob.CandidatePool.explicit(np.random.default_rng(1).random((m, d)), device=DEV)):
            fused = ob.score(models, spec, pool, precision="fast", want_acq=True)                      # fused (no posterior asked)
            plain = ob.score(models, spec, pool, precision="fast", want_acq=True, want_posterior=True)  # K4 launch
            only = ob.score(models, spec, pool, precision="fast")                                       # arg-max only
            a, b = fused.acq.cpu().numpy(), plain.acq.cpu().numpy()
            same = np.array_equal(a, b, equal_nan=True)
            err = np.nanmax(np.abs(a - b)) / max(np.nanmax(np.abs(b)), 1e-300)
            print(f"{sem} n={n} d={d} m={m} P={len(PF)} fmt={[g.plane_format for g in models]} identical={same} maxerr/max={err:.2e} "
                  f"best fused=({fused.best_value:.6g},{fused.best_index}) plain=({plain.best_value:.6g},{plain.best_index}) only=({only.best_value:.6g},{only.best_index})", flush=True)
            assert err < 1e-5, err
            assert fused.best_index == only.best_index and fused.best_value == only.best_value
            if sem == "exact":     # (reference semantics: model 1 runs the mean-only kernel unless the posterior is asked for)
                assert same and fused.best_index == plain.best_index and fused.best_value == plain.best_value
print("fuse check ok")
