import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo/root-simple-mcmc_b200"); sys.path.insert(0, "/root/repo")
import smcmc_b200
from smcmc_b200 import binding as b
n, E = 500, 16384
eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=4)
eng.prop_set(b.PROP_POOLED_EVERY, 16)
eng.prop_set(b.PROP_POOLED_TENSOR, 1)
eng.start(np.zeros(n))
eng.step(20); eng.sync()
