"""Timing of the joint posterior draw behind TuRBO (section 8f rank 4): device (ombo_posterior_joint_samples) vs the
numpy/LAPACK oracle on the host, same Z.  Prints markdown rows."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
from oracle import oracle as O
print("| n_train | d | candidates | samples | device ms | host (oracle) ms | max abs diff / scale |\n|---|---|---|---|---|---|---|")
for n, d, m, S in ((100, 10, 1000, 4), (100, 32, 3200, 4), (256, 10, 5000, 4), (1024, 10, 5000, 8)):
    rng = np.random.default_rng(n + m)
    X = rng.random((n, d)); y = np.sin(X.sum(1))
    ell = 0.5 * np.sqrt(d) * np.ones(d)
    gp = ob.GPModel(X, y, ell, 1.0, device='cuda:0')
    Xc = rng.random((m, d)); Z = rng.standard_normal((m, S))
    gp.posterior_samples_from(Xc, Z, jitter=1e-6); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        got = gp.posterior_samples_from(Xc, Z, jitter=1e-6)
    dev_ms = (time.perf_counter() - t0) / 3 * 1e3
    st = O.gp_fit_state(X, y, ell, 1.0, form="direct")
    t0 = time.perf_counter(); want = O.posterior_samples(st, Xc, Z, 1e-6); host_ms = (time.perf_counter() - t0) * 1e3
    print(f"| {n} | {d} | {m} | {S} | {dev_ms:.1f} | {host_ms:.0f} | {np.abs(got[:, 0, :] - want).max() / np.abs(want).max():.1e} |", flush=True)
