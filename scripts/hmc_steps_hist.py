"""C4 shape after its warm-up: how ragged are the chains' trajectory lengths?  (the ensemble step runs
max(L) + 1 leap-frog stages; a 64-row tile of the stage GEMM is useful only while one of its chains runs)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import smcmc_b200
from smcmc_b200 import binding as b
from hmc_bench import precision
n, E = 500, int(os.environ.get("HMC_CHAINS", "16384"))
eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
eng.set_error_matrix(precision(n))
eng.set_dummy_mode(b.DUMMY_TENSOR)
eng.hmc_set(b.HMC_USER_GRADIENT, 1)
eng.hmc_start(np.ones(n))
for warm in (40, 80, 200):
    eng.hmc_step(warm if warm == 40 else warm - prev); eng.sync(); prev = warm
    s = eng.hmc_scalars()
    L = np.abs(s["leapfrog"]).astype(int)
    print("after %d steps: mean L %.2f, min %d, max %d, mean eps %.4g" % (warm, L.mean(), L.min(), L.max(), np.abs(s["mean_epsilon"]).mean() if "mean_epsilon" in s else float("nan")))
    print("  histogram:", {int(k): int(v) for k, v in zip(*np.unique(L, return_counts=True))})
    stages = L.max() + 1
    tiles = L.reshape(-1, 64).max(axis=1)
    useful_rows = (L + 1).sum()
    print("  stages run %d; useful stage-rows / all = %.3f; tile-stages with an active row / all = %.3f; sorted by L: %.3f"
          % (stages, useful_rows / (stages * E), (tiles + 1).sum() / (stages * len(tiles)),
             (np.sort(L)[::-1].reshape(-1, 64).max(axis=1) + 1).sum() / (stages * len(tiles))))
print("keys", list(s.keys()))
