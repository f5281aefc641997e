"""sMCMC::TSimpleMCMC<L, TProposeVAATStep> on the device (csrc/vaat.cuh) against
the golden chains of the reference build (TProposeVAATStep.H:22-307 through
TSimpleMCMC::Step, tests/golden/make_golden.py vaat) and against the port on an
ensemble.

Required: identical accept / reject sequence, identical order of the proposed
coordinates (the shuffled index queue, :176-190), identical trial counts; points,
per-dimension step sizes and acceptances to 1e-12 (the only difference is the last
ulp of pow() between CUDA and glibc in the step-size update, :245-251).
"""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-12):
    scale = np.maximum(np.abs(b), 1.0)
    return np.all(np.abs(a - b) <= rtol * scale)


@pytest.mark.parametrize("name", ["vaat_unit5", "vaat_unit9_hints", "vaat_horrific75"])
def test_golden_vaat_chain(name):
    import smcmc_b200
    from smcmc_b200 import binding
    from golden.make_golden import VAAT_CHAINS
    assert torch.cuda.is_available()
    kind, dim, seed, chain, nsteps, configure = VAAT_CHAINS[name]
    g = golden("vaat.npz")
    lo = max(0, chain - 1)
    eng = smcmc_b200.Engine(kind, dim, 3, seed=seed, chain_offset=lo, proposal=smcmc_b200.PROPOSAL_VAAT)
    if configure:                                     # the same calls the golden run made
        eng.set_uniform(2, -1.5, 2.0)
        eng.set_gaussian(4, 0.5)
        eng.prop_set(binding.PROP_ACCEPTANCE_RIGIDITY, 1.5)
    eng.prop_set(binding.PROP_ACCEPTANCE_WINDOW, 7.0)  # overridden by Start, as in the reference (:208)
    ok = eng.start(np.zeros(dim))
    c = chain - lo
    assert ok[c] == int(g[name + "/ok"][0])
    assert eng.get("acceptance_window")[0] == 100
    first = eng.step_trace(nsteps - 300)
    eng.prop_set(binding.PROP_ACCEPTANCE_WINDOW, 37.0)
    second = eng.step_trace(300)
    tr = {k: np.concatenate([first[k][:, c], second[k][:, c]]) for k in first}
    assert np.array_equal(tr["accepted"], g[name + "/accepted"])
    assert close(tr["points"], g[name + "/x"])
    assert close(tr["llh_proposed"], g[name + "/llh_proposed"], 1e-11)
    assert close(tr["sigma"], g[name + "/sigma"])
    # which coordinate moved at every step is the shuffle order: identical
    moved_dev = np.argmax(np.abs(np.diff(np.vstack([np.zeros(dim), tr["points"]]), axis=0)) > 0, axis=1)
    moved_ref = np.argmax(np.abs(np.diff(np.vstack([np.zeros(dim), g[name + "/x"]]), axis=0)) > 0, axis=1)
    assert np.array_equal(moved_dev, moved_ref)
    assert close(eng.get("vaat_sigma")[c], g[name + "/final_sigma"])
    assert close(eng.get("vaat_acceptance")[c], g[name + "/final_acceptance"])
    assert np.array_equal(eng.get("vaat_acceptance_trials")[c], g[name + "/final_acceptance_trials"])
    misc = g[name + "/final_misc"]
    assert eng.get("trials")[c] == misc[0] and eng.get("successes")[c] == misc[1]
    assert eng.get("vaat_last_index")[c] == misc[2] and eng.get("vaat_queue")[c] == misc[3]
    assert close(eng.get("step_rms")[c], g[name + "/final_step_rms"][0])
    assert close(eng.get("sigma")[c], g[name + "/final_sigma"].mean(), 1e-11)
    assert eng.get("total_steps")[c] == nsteps and eng.get("llh_calls")[c] == nsteps + 1


def test_ensemble_against_the_port(checkers):
    """300 chains of the event likelihood with the VAAT proposal: every chain equals the port's."""
    import smcmc_b200
    events, data = smcmc_b200.synth.fake_inputs(60, 60, 10, seed=5)
    chains, steps = 300, 120
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, chains, seed=31, proposal=smcmc_b200.PROPOSAL_VAAT)
    eng.set_fake_events(events)
    eng.set_fake_data(data, 0.1)
    rng = np.random.default_rng(1)
    x0 = rng.uniform(-1, 1, (chains, 9))
    assert eng.start(x0).all()
    tr = eng.step_trace(steps)
    for c in (0, 1, 150, 299):
        o = checkers.CpuChain("orc", checkers.LLH_FAKE, 9, 31, c, vaat=True)
        o.set_fake(events, data, 0.1)
        o.start(x0[c])
        want = o.step(steps)
        assert np.array_equal(tr["accepted"][:, c], want["accepted"]), c
        assert close(tr["points"][:, c], want["x"])
    assert 0.2 < tr["accepted"].mean() < 0.9


def test_vaat_rejects_what_the_reference_does_not_have():
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 4, 8, seed=3, proposal=smcmc_b200.PROPOSAL_VAAT)
    eng.start(np.zeros((8, 4)))
    eng.step(10)
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.update_proposal()
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.save_state()
    ada = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 4, 8, seed=3)
    ada.start(np.zeros((8, 4)))
    with pytest.raises(smcmc_b200.SmcmcError):
        ada.get("vaat_sigma")
