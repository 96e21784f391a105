import sys, numpy as np, torch
sys.path.insert(0, ".")
import optimobo_b200 as ob
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
X = rng.random((n, 10)); y = np.sin(X.sum(1))
gp = ob.GPModel(X, y, 0.7 * np.ones(10), 1.0, device="cuda:0")
gp.refresh(); gp.refresh()
torch.cuda.synchronize()
print("ok")
