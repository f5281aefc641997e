"""Small invocations of the kernels added late in round 1, for compute-sanitizer
(memcheck / racecheck): kStepsResident, kAcceptLocal, kPoolAccumulateDmma (both
templates), kProposePooledTile, kFakeFinish with split bins, kHmcExxtFlush, kVaatPropose."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200 as sm
from smcmc_b200 import binding as b, synth

# resident kernel: odd and even n, a uniform dimension, UpdateProposal inside the launch
for n, E, kind in ((7, 5, sm.LLH_ASYM), (20, 3, sm.LLH_HORRIFIC), (6, 2, sm.LLH_HARD)):
    eng = sm.Engine(kind, n, E, seed=3)
    eng.prop_set(b.PROP_ACCEPTANCE_WINDOW, 12.0)
    if n == 7:
        eng.set_uniform(2, -1.0, 1.0)
    eng.start(np.full(n, 0.01))
    eng.step(150); eng.sync()
    assert np.all(eng.get("total_steps") == 150)
    eng.close()
# three-launch step with the fused accept, 33 chains (a warp tile with one chain)
os.environ["SMCMC_NO_RESIDENT"] = "1"
eng = sm.Engine(sm.LLH_HORRIFIC, 9, 33, seed=4); eng.start(np.zeros(9)); eng.step(40); eng.sync(); eng.close()
eng = sm.Engine(sm.LLH_UNIT_GAUSS, 7, 40, seed=4, proposal=sm.PROPOSAL_VAAT); eng.start(np.zeros(7)); eng.step(60); eng.sync(); eng.close()
# pooled: tile proposal + tensor-core accumulation, diagonal block only and with off-diagonal blocks
for n, E in ((9, 70), (60, 37)):
    eng = sm.Engine(sm.LLH_UNIT_GAUSS, n, E, seed=5)
    eng.prop_set(b.PROP_POOLED_EVERY, 4)
    eng.start(np.random.default_rng(1).normal(0, 1, (E, n)))
    eng.step(9); eng.sync()
    assert eng.get("pooled_count")[0] == E * 8
    eng.close()
del os.environ["SMCMC_NO_RESIDENT"]
# event likelihood with few chains: kFakeFinish with the bins split over CTAs
events, data = synth.fake_inputs(300, 300, 10, seed=6)
eng = sm.Engine(sm.LLH_FAKE, 9, 3, seed=8)
eng.set_fake_events(events); eng.set_fake_data(data, 0.1)
eng.start(np.random.default_rng(1).uniform(-1, 1, (3, 9))); eng.step(5); eng.sync(); eng.close()
# HMC with the deferred fEXXT update (ring of 3), covariance read at the end
os.environ["SMCMC_HMC_DEFER"] = "3"
n, E = 6, 5
eng = sm.Engine(sm.LLH_UNIT_GAUSS, n, E, seed=9)
eng.hmc_start(np.full(n, 0.5)); eng.hmc_step(40, 3); eng.sync()
cov = eng.hmc_get("covariance")
assert np.all(np.isfinite(cov))
eng.close()
print("sanitize_small ok")
