"""Joules per C5 scoring step from NVML's total-energy counter (the GPU runs at its power cap: time follows energy).
Run once per mode: OMBO_FAST_DBG unset (full), 1 (K1 only: no MMAs), 2 (MMA + pipeline only: no K1 arithmetic)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import pynvml
import optimobo_b200 as ob
import bench as B
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
X, Y, ells, sf2 = B.workload()
models = [ob.GPModel(X, Y[:, i], ells[i], sf2[i], device="cuda:0") for i in range(2)]
spec = ob.spec_ehvi(Y.max(0), ob.host_prep.calc_pf(Y), ob.host_prep.cached_samples(2, 5, seed=0), "exact")
pool = ob.CandidatePool.counter(1 << 24, np.zeros(10), np.ones(10), seed=1)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for _ in range(5): ob.score(models, spec, pool, precision="fast", sync=False)
torch.cuda.synchronize()
e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t0 = time.perf_counter()
for _ in range(steps): ob.score(models, spec, pool, precision="fast", sync=False)
torch.cuda.synchronize()
t1 = time.perf_counter(); e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
print(f"dbg={os.environ.get('OMBO_FAST_DBG', '0'):>3s}  {1e3 * (t1 - t0) / steps:7.2f} ms/step  {(e1 - e0) / 1e3 / steps:6.2f} J/step  "
      f"avg {(e1 - e0) / 1e3 / (t1 - t0):6.0f} W  (SM clock after {clk} MHz, limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3:.0f} W)", flush=True)
