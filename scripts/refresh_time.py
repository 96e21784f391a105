"""GP refresh (K3) timing per n, with the per-kernel CUDA-event breakdown of the library's profiler."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import optimobo_b200 as ob
from optimobo_b200 import _cabi
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048]:
    X = rng.random((n, 10)); y = np.sin(X.sum(1))
    gp = ob.GPModel(X, y, 0.7 * np.ones(10), 1.0, device="cuda:0")
    for _ in range(3): gp.refresh()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gp.refresh()
    e1.record(); torch.cuda.synchronize()
    print(f"n={n}: {e0.elapsed_time(e1) / 10:.3f} ms per refresh (kappa {gp.conditioning:.1f}, planes {gp.plane_format})", flush=True)
