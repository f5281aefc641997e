"""Quick timing of the headline config (no CPU baseline)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
E, N = 4096, 1000000
events = smcmc_b200.synth.make_mc_sample(N // 3 + 1, N - N // 3 - 1, 2)
data = smcmc_b200.synth.make_data_histograms(33334, 33334, 2)
eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
eng.set_fake_events(events)
expo = smcmc_b200.synth.exposure_ratio(eng, data); eng.set_fake_data(data, expo)
eng.start(np.random.default_rng(0).uniform(-1, 1, (E, 9)))
eng.enable_kernel_timing(True)
eng.step(5); eng.sync(); eng.pair_kernel_stats(reset=True)
t = time.time(); eng.step(20); eng.sync(); dt = time.time() - t
ms, n = eng.pair_kernel_stats()
print("20 steps wall %.3f s -> %.1f MH steps/s ; pair kernel %.3f ms/launch ; pairs/s %.3e" % (dt, E * 20 / dt, ms / n, E * N * n / (ms * 1e-3)))
print("filter check (256 pts):", eng.fake_filter_check(np.random.default_rng(1).uniform(-1, 1, (256, 9))))
