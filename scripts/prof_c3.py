"""Short run of C3 (65536 chains x 50 dims, per-chain adaptation) for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
eng = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, 50, 65536, seed=4)
if os.environ.get("C3_POOLED"):
    from smcmc_b200 import binding as b
    eng.prop_set(b.PROP_POOLED_EVERY, int(os.environ["C3_POOLED"]))
eng.start(np.zeros(50))
eng.step(int(os.environ.get("C3_STEPS", "40"))); eng.sync()
print("done", eng.get("acceptance").mean())
