"""Short pooled-adaptation run at n = 500 x 16384 chains (tensor-core proposal) for the ncu launch list."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
from smcmc_b200 import binding as b
n, E = 500, 16384
eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=4)
eng.prop_set(b.PROP_POOLED_EVERY, 16)
eng.prop_set(b.PROP_POOLED_TENSOR, 1)
eng.start(np.zeros(n))
eng.step(int(os.environ.get("POOLED_STEPS", "8"))); eng.sync()
print("done")
