"""Steps/s of the other BASELINE.json configs (analytic likelihoods)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
def run(name, kind, dim, E, steps, err=None):
    eng = smcmc_b200.Engine(kind, dim, E, seed=4)
    if err is not None: eng.set_error_matrix(err)
    eng.start(np.zeros((E, dim)))
    eng.step(20); eng.sync()
    t = time.time(); eng.step(steps); eng.sync(); dt = time.time() - t
    tri = dim * (dim + 1) // 2
    bytes_per = (2 * tri + dim * (dim + 1) // 2 + 6 * dim) * 8
    print("%-28s E=%6d n=%3d: %.3f ms/step, %.3e chain-steps/s, ~%.0f GB/s algorithmic, acc %.3f" % (
        name, E, dim, 1e3 * dt / steps, E * steps / dt, E * steps / dt * bytes_per / 1e9, eng.get("acceptance").mean()), flush=True)
run("C1 unit gauss 1 chain", 0, 5, 1, 2000)
run("C1 dummy100 1 chain", 1, 100, 1, 500, np.eye(100))
run("C3 horrific", 2, 50, 65536, 100)
run("C3 asym", 3, 50, 65536, 100)
run("unit9 4096", 0, 9, 4096, 500)
run("dummy100 4096", 1, 100, 4096, 50, np.eye(100))
