"""Short run of the headline config for ncu (few launches of the pair kernel)."""
import sys, os
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
eng.step(4); eng.sync()
print("done", eng.get("acceptance").mean())
if os.environ.get("PAIRS_FILTER_CHECK"):
    pairs, unsure, bad = eng.fake_filter_check(eng.get("proposed"))
    print("filter check: pairs %d, undecided %d (%.4f %%), disagreements %d" % (pairs, unsure, 100.0 * unsure / pairs, bad))
