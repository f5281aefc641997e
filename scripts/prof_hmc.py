"""Short C4 run (TENSOR mode, 16384 chains x 500 dims) for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import smcmc_b200
from smcmc_b200 import binding as b
from hmc_bench import precision
n, E = 500, 16384
eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
eng.set_error_matrix(precision(n))
eng.set_dummy_mode(b.DUMMY_TENSOR)
eng.hmc_set(b.HMC_USER_GRADIENT, 1)
eng.hmc_start(np.ones(n))
eng.hmc_step(int(os.environ.get("HMC_STEPS", "2"))); eng.sync()
print("done")
