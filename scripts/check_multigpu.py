"""torchrun --nproc-per-node N scripts/check_multigpu.py : the sharded path over NCCL must return the
same (value, index) on every rank as a single-GPU pass over the whole pool."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimobo_b200 as ob  # noqa: E402
from optimobo_b200.distributed import propose_sharded, score_sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(0)
n, d, m = 300, 6, 1 << 18
X = rng.random((n, d))
Y = np.column_stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0])])
models = [ob.GPModel(X, Y[:, i], 0.7 * np.ones(d), 1.0 + i, device=dev) for i in range(2)]
cache = ob.host_prep.cached_samples(2, 5, seed=0)
spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), cache, "exact")
pool = ob.CandidatePool.counter(m, np.zeros(d), np.ones(d), seed=5)
for precision in ("fp64", "fast"):
    v, i = score_sharded(models, spec, pool, precision=precision)
    whole = ob.score(models, spec, pool, precision=precision)
    x, neg, idx = propose_sharded(models, spec, pool, precision=precision)
    got = torch.tensor([v, float(i)], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(got) for _ in range(world)]
    dist.all_gather(allv, got)
    assert all(torch.equal(a, allv[0]) for a in allv), "ranks disagree"
    assert i == whole.best_index == idx, (precision, i, whole.best_index)
    assert abs(v - whole.best_value) <= 1e-12 * abs(whole.best_value), (v, whole.best_value)
    assert np.array_equal(x, pool.rows(i, 1, dev)[0].cpu().numpy()) and neg == -v
if rank == 0:
    print(f"multi-GPU check ok: world={world} best={v} @ {i}")
dist.destroy_process_group()
