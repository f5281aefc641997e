"""A/B timing of the C4 leap-frog stage inside ONE process (same box, same clocks): 16384 chains x
500 dims, TENSOR mode, trajectory length fixed at 20 (21 gradients per step), fused stage
(kHmcLeapDmma) against gradient kernel + kHmcKickDrift (SMCMC_HMC_NO_FUSE=1), alternating."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import smcmc_b200
from smcmc_b200 import binding as b
from hmc_bench import precision
n, E, L, steps = 500, int(os.environ.get("AB_CHAINS", "16384")), 20, 10
eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
eng.set_error_matrix(precision(n))
eng.set_dummy_mode(b.DUMMY_TENSOR)
eng.hmc_set(b.HMC_USER_GRADIENT, 1)
eng.hmc_set(b.HMC_LEAPFROG, L)
eng.hmc_start(np.ones(n))
eng.hmc_step(3); eng.sync()
for rep in range(3):
    for nofuse in (0, 1):
        if nofuse:
            os.environ["SMCMC_HMC_NO_FUSE"] = "1"
        else:
            os.environ.pop("SMCMC_HMC_NO_FUSE", None)
        eng.hmc_step(1); eng.sync()
        t = time.perf_counter(); eng.hmc_step(steps); eng.sync(); dt = time.perf_counter() - t
        evals = E * steps * (L + 2)
        print("rep %d %s: %.3f ms/step, %.3f ms per leap-frog stage, %.2f TFLOP/s over the step"
              % (rep, "unfused" if nofuse else "fused  ", 1e3 * dt / steps, 1e3 * dt / steps / (L + 1),
                 evals * 2.0 * n * n / dt / 1e12), flush=True)
