import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo/root-simple-mcmc_b200"); sys.path.insert(0, "/root/repo")
import smcmc_b200
from smcmc_b200 import synth
N = 4000000
ev = synth.make_mc_sample(N // 3, N - N // 3, seed=2)
data = synth.make_data_histograms(33334, 33334, seed=2)
for E in (1, 8, 16, 17, 64, 256):
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
    eng.set_fake_events(ev); eng.set_fake_data(data, 0.02)
    x0 = np.random.default_rng(1).uniform(-1, 1, (E, 9))
    eng.start(x0); eng.step(3); eng.sync()
    t = time.perf_counter(); eng.step(20); eng.sync(); dt = (time.perf_counter() - t) / 20
    print("E %4d  %.3f ms/step  %.3e pairs/s" % (E, dt * 1e3, E * N / dt), flush=True)
    os.environ["SMCMC_GRAPH"] = "1"
    eng.step(20); eng.sync()
    t = time.perf_counter(); eng.step(200); eng.sync(); dt = (time.perf_counter() - t) / 200
    del os.environ["SMCMC_GRAPH"]
    print("        %.3f ms/step replayed from a CUDA graph (SMCMC_GRAPH=1, 200 steps, graph cached)" % (dt * 1e3), flush=True)
    t = time.perf_counter(); eng.step(200); eng.sync(); dt = (time.perf_counter() - t) / 200
    print("        %.3f ms/step plain loop, 200 steps" % (dt * 1e3), flush=True)
