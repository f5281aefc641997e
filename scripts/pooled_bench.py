"""Pooled adaptation at large dimension: per-warp z.U against the DMMA GEMM."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
from smcmc_b200 import binding as b
for n, E in ((128, 16384), (256, 16384), (500, 16384)):
    for tensor in (0, 1):
        eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=4)
        eng.prop_set(b.PROP_POOLED_EVERY, 16)
        eng.prop_set(b.PROP_POOLED_TENSOR, tensor)
        eng.start(np.zeros(n))
        eng.step(20); eng.sync()
        t = time.perf_counter(); eng.step(50); eng.sync(); dt = time.perf_counter() - t
        print("pooled n=%d E=%d %s: %.3f ms/step, %.3e chain-steps/s, proposal contraction %.2f TFLOP/s (2n^2 per chain-step)"
              % (n, E, "tensor" if tensor else "warp  ", 1e3 * dt / 50, E * 50 / dt, E * 50 * 2.0 * n * n / dt / 1e12), flush=True)
