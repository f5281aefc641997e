"""Time the pair kernel of the library named by SMCMC_B200_LIB (tuning experiments)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
from smcmc_b200 import synth
E, N = 4096, 1000000
events = synth.make_mc_sample(N // 3 + 1, N - N // 3 - 1, 2)
data = synth.make_data_histograms(33334, 33334, 2)
eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
eng.set_fake_events(events)
eng.set_fake_data(data, 0.1)
eng.start(np.random.default_rng(0).uniform(-1, 1, (E, 9)))
eng.enable_kernel_timing(True)
eng.step(3); eng.sync(); eng.pair_kernel_stats(reset=True)
eng.step(20); eng.sync()
ms, n = eng.pair_kernel_stats()
print("%-40s pair kernel %.4f ms/launch" % (os.environ.get("SMCMC_B200_LIB", "default"), ms / n), flush=True)
