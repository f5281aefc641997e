"""One chain over 16.8 M events (kFakeStream) for ncu."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
N = 16777216
events = smcmc_b200.synth.make_mc_sample(N // 3 + 1, N - N // 3 - 1, 2)
data = smcmc_b200.synth.make_data_histograms(33334, 33334, 2)
eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 1, seed=3)
eng.set_fake_events(events)
eng.set_fake_data(data, 0.006)
eng.start(np.random.default_rng(0).uniform(-1, 1, (1, 9)))
eng.step(4); eng.sync()
print("done")
