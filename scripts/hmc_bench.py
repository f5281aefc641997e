"""C4 (BASELINE.json configs[3]): TSimpleHMC on a 500-dimensional Gaussian with
the analytic gradient, E chains.  Prints steps/s, gradients/s and the FP64 rate
of the gradient contraction (2 n^2 flop per gradient)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200"))
sys.path.insert(0, ROOT)
import smcmc_b200
from smcmc_b200 import binding as b


def precision(n, seed=5):
    """Dense random SPD precision matrix ('a dense random-SPD variant')."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n))
    m = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))
    return 0.5 * (m + m.T)


def run(n, E, steps, burn, tensor=0):
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
    eng.set_error_matrix(precision(n))
    eng.set_dummy_mode(b.DUMMY_TENSOR if tensor else b.DUMMY_EXACT)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_start(np.ones(n))                       # SimpleHMC.C:45
    eng.hmc_step(burn)
    eng.sync()
    s0 = eng.hmc_scalars()
    t = time.perf_counter()
    eng.hmc_step(steps)
    eng.sync()
    dt = time.perf_counter() - t
    s1 = eng.hmc_scalars()
    grads = float((s1["gradient_count"] - s0["gradient_count"]).sum())
    pots = float((s1["potential_count"] - s0["potential_count"]).sum())
    print("%s n=%d E=%d: %.2f ms/step, %.3e chain-steps/s, %.3e gradients/s, %.3e potentials/s, "
          "gradient+potential contraction %.2f TFLOP/s (2n^2 each), leapfrog %.1f, eps %.4f, acceptance %.3f"
          % ("tensor" if tensor else "exact ", n, E, 1e3 * dt / steps, E * steps / dt, grads / dt, pots / dt, (grads + pots) * 2 * n * n / dt / 1e12,
             np.abs(s1["leapfrog"]).mean(), s1["mean_epsilon"].mean(), s1["acceptance"].mean()), flush=True)
    eng.close()


if __name__ == "__main__":
    cfgs = [(500, 1, 20, 5, 0), (500, 1024, 20, 5, 0), (500, 1024, 20, 5, 1), (500, 16384, 10, 3, 0),
            (500, 16384, 10, 3, 1), (100, 4096, 30, 5, 0), (100, 4096, 30, 5, 1)]
    if len(sys.argv) > 1:
        cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        run(*c)
