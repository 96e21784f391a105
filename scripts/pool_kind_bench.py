"""C5 scoring pass with the counter-generated pool against an explicit device pool (f32 / f64): how much of a step is the
candidate-coordinate preparation of warp 2 (counter hash in 64-bit integer + FP64 arithmetic)."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import optimobo_b200 as ob
import bench as B
X, Y, ells, sf2 = B.workload()
dev = "cuda:0"
models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device=dev) for i in range(2)]
spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), ob.host_prep.cached_samples(2, 5, seed=0), "exact")
m, d = 1 << 24, X.shape[1]
pools = {"counter": ob.CandidatePool.counter(m, np.zeros(d), np.ones(d), seed=1)}
Xf = torch.rand((m, d), dtype=torch.float32, device=dev)
pools["explicit f32"] = ob.CandidatePool.explicit(Xf, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, p in pools.items():
    for _ in range(2): ob.score(models, spec, p, precision="fast", sync=False)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ob.score(models, spec, p, precision="fast", sync=False); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{name:14s} {np.median(ts):.2f} ms per 2^24 (min {min(ts):.2f})", flush=True)
