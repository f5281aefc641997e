"""Short run of the unbinned likelihood at C2's size (4096 chains x 1M events) for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
from smcmc_b200 import synth
E, N = 4096, 1000000
events = synth.make_mc_sample(N // 3 + 1, N - N // 3 - 1, 2)
eng = smcmc_b200.Engine(smcmc_b200.LLH_UNBINNED, 9, E, seed=3)
eng.set_unbinned_events(events)
eng.start(np.random.default_rng(0).uniform(-1, 1, (E, 9)))
eng.step(3); eng.sync()
print("done", eng.get("acceptance").mean())
