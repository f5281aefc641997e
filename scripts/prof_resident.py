"""One wave of resident chains (888 chains x 50 dims, THorrific) for ncu: kStepsResident runs
RES_STEPS whole Metropolis steps per launch (csrc/proposal_resident.cuh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
E = int(os.environ.get("RES_CHAINS", "888"))
K = int(os.environ.get("RES_STEPS", "200"))
eng = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, 50, E, seed=4)
eng.start(np.zeros(50))
eng.step(50); eng.sync()
import time
t = time.perf_counter(); eng.step(K); eng.sync(); dt = time.perf_counter() - t
print("resident: %d chains x %d steps in %.3f ms = %.2f us per step, %.3e chain-steps/s, %d launches"
      % (E, K, dt * 1e3, dt / K * 1e6, E * K / dt, eng.launch_count()))
