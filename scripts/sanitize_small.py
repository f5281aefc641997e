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
del os.environ["SMCMC_HMC_DEFER"]
# round 2 kernels: fused leap-frog stage + 3-stage DMMA pipeline (even and odd n, ragged tiles), pooled HMC
# covariance (kHmcPooledFold / Trigger / Spectrum / Apply past the 32-step threshold)
def spd(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n))
    m = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))
    return 0.5 * (m + m.T)
for n, E in ((70, 66), (37, 9)):
    eng = sm.Engine(sm.LLH_DUMMY, n, E, seed=6)
    eng.set_error_matrix(spd(n, n))
    eng.set_dummy_mode(b.DUMMY_TENSOR)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_set(b.HMC_POOLED_COVARIANCE, 1)
    eng.hmc_start(np.full(n, 0.5)); eng.hmc_step(36, 0); eng.sync()
    assert eng.hmc_get("pooled_scalars")[7] >= 1
    eng.close()
# pooled Metropolis at n >= 64: tiled Gram statistics (kPoolGramDmma), CTA-wide factorisation, DMMA proposal
eng = sm.Engine(sm.LLH_UNIT_GAUSS, 130, 70, seed=5)
eng.prop_set(b.PROP_POOLED_EVERY, 4); eng.prop_set(b.PROP_POOLED_TENSOR, 1)
eng.start(np.random.default_rng(1).normal(0, 1, (70, 130))); eng.step(9); eng.sync(); eng.close()
# debugging modes of the proposal, event-sharded layout of the count table is covered by the 2-GPU test
eng = sm.Engine(sm.LLH_UNIT_GAUSS, 5, 37, seed=2)
eng.start(np.zeros(5)); eng.force_step(np.full(5, 0.1)); eng.step(3, 2); eng.set_scan(2); eng.step(4); eng.set_scan(-1); eng.step(3); eng.sync(); eng.close()
# streaming event likelihood with irregular records folded into kFakeStream
ev2 = np.concatenate([events[:2000], events[:3]])
ev2["Type"][-3:] = -1
eng = sm.Engine(sm.LLH_FAKE, 9, 2, seed=8)
eng.set_fake_events(ev2); eng.set_fake_data(data, 0.1)
eng.start(np.random.default_rng(1).uniform(-1, 1, (2, 9))); eng.step(4); eng.sync(); eng.close()
print("sanitize_small ok")
