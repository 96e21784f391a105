"""Refresh results (L, L^-1, alpha) of the library in the tree against a second build (argv[1]): bit comparison at
several n and both kernels, and the refresh time of each."""
import ctypes, os, subprocess, sys
import numpy as np
if len(sys.argv) > 2:      # child: dump states with the library given
    import torch
    sys.path.insert(0, ".")
    import optimobo_b200 as ob
    out = {}
    rng = np.random.default_rng(0)
    for n, d, kern in [(100, 3, "rbf"), (130, 3, "rbf"), (256, 10, "matern52"), (700, 7, "matern52"), (1024, 10, "matern52"), (1536, 10, "rbf")]:
        X = rng.random((n, d)); y = np.sin(X.sum(1))
        gp = ob.GPModel(X, y, 0.7 * np.ones(d), 1.3, kernel=kern, device="cuda:0")
        for f in ("L", "Linv", "alpha"):
            out[f"{n}_{kern}_{f}"] = getattr(gp, f).cpu().numpy()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gp.refresh()
        e1.record(); torch.cuda.synchronize()
        print(sys.argv[2], n, kern, f"{e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
other = sys.argv[1]
so = "optimobo_b200/liboptimobo_b200.so"
subprocess.run([sys.executable, __file__, "child", "/tmp/bits_new.npz"], check=True)
os.rename(so, so + ".keep"); subprocess.run(["cp", other, so], check=True)
try:
    subprocess.run([sys.executable, __file__, "child", "/tmp/bits_old.npz"], check=True)
finally:
    os.rename(so + ".keep", so)
a, b = np.load("/tmp/bits_new.npz"), np.load("/tmp/bits_old.npz")
bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
print("bit-identical" if not bad else f"DIFFERENT: {bad}")
