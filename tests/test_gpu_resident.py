"""kStepsResident (csrc/proposal_resident.cuh): smcmc_step(nsteps) for a chain-local
likelihood runs all the steps in ONE launch with the chain's adaptive state resident
in shared memory.  Required:
  * the chains are BIT-IDENTICAL to the ones the three-launch step gives
    (SMCMC_NO_RESIDENT=1), whatever the split of the steps over calls;
  * the end state equals the golden runs of the reference build (tests/golden/chains.npz)
    to the tolerance test_gpu_chains.py states for the traced path.
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CHAINS, bind_likelihood_inputs, configure_golden, golden, golden_chain
from test_gpu_chains import _set_field, close

pytestmark = pytest.mark.gpu

FIELDS = ("accepted", "proposed", "accepted_llh", "proposed_llh", "sigma", "trials", "successes", "next_update",
          "total_steps", "llh_calls", "step_rms", "center", "covariance", "decomposition", "status", "acceptance",
          "acceptance_trials", "acceptance_rigidity", "covariance_trials", "center_trials", "sigma_trace")


def _fields(eng):
    return {f: eng.get(f) for f in FIELDS}


def _run(monkeypatch, resident, make, steps):
    import smcmc_b200
    if resident:
        monkeypatch.delenv("SMCMC_NO_RESIDENT", raising=False)
    else:
        monkeypatch.setenv("SMCMC_NO_RESIDENT", "1")
    eng = make(smcmc_b200)
    for k in steps:
        eng.step(k)
    eng.sync()
    out = _fields(eng)
    out["launches"] = eng.launch_count()
    out["tail"] = eng.step_trace(5, want=("accepted", "points"))["points"]   # traced steps: the three-launch path
    return out


def _horrific(sm):
    eng = sm.Engine(sm.LLH_HORRIFIC, 20, 257, seed=9)
    eng.start(np.zeros(20))
    return eng


def _asym50(sm):
    eng = sm.Engine(sm.LLH_ASYM, 50, 300, seed=4)
    eng.start(np.full(50, 0.01))
    return eng


def _unit_hints(sm):
    eng = sm.Engine(sm.LLH_UNIT_GAUSS, 5, 33, seed=1)
    eng.set_gaussian(3, 2.0)
    eng.set_uniform(4, -5, 5)
    eng.set_correlation(3, 4, 0.3)
    eng.start(np.zeros(5))
    return eng


def _hard(sm):
    eng = sm.Engine(sm.LLH_HARD, 6, 40, seed=13)
    eng.start(np.full(6, 0.5))
    return eng


def _dummy(sm):
    n = 12
    rng = np.random.default_rng(2)
    m = rng.normal(size=(n, n))
    err = m @ m.T / n + np.eye(n)
    eng = sm.Engine(sm.LLH_DUMMY, n, 70, seed=6)
    eng.set_error_matrix(err)
    eng.start(np.zeros(n))
    return eng


def _frozen(sm):
    from smcmc_b200 import binding
    eng = sm.Engine(sm.LLH_ASYM, 24, 50, seed=11)
    eng.prop_set(binding.PROP_ACCEPTANCE_WINDOW, 12.0)       # UpdateProposal after 12 accepted steps
    eng.prop_set(binding.PROP_COVARIANCE_FROZEN, 1.0)
    eng.start(np.random.default_rng(5).uniform(-0.01, 0.01, (50, 24)))
    return eng


def _deweighted(sm):
    from smcmc_b200 import binding
    eng = sm.Engine(sm.LLH_HORRIFIC, 24, 50, seed=11)
    eng.prop_set(binding.PROP_ACCEPTANCE_WINDOW, 12.0)
    eng.prop_set(binding.PROP_COVARIANCE_DEWEIGHT, 0.37)     # fractional trial counts
    eng.prop_set(binding.PROP_COVARIANCE_WINDOW, 150.0)
    eng.start(np.random.default_rng(5).uniform(-0.01, 0.01, (50, 24)))
    return eng


def _asym100(sm):
    eng = sm.Engine(sm.LLH_ASYM, 100, 6, seed=4, chain_offset=10)   # one chain per CTA, two CTAs per SM
    eng.start(np.full(100, 0.01))
    return eng


@pytest.mark.parametrize("make,steps", [
    (_frozen, (300, 60)),
    (_deweighted, (500,)),
    (_asym100, (900, 1)),
    (_horrific, (1500, 3, 1, 200)),       # several UpdateProposal passes per chain
    (_asym50, (400, 100)),
    (_unit_hints, (2500, 700)),           # a uniform dimension: the generic proposal loop
    (_hard, (1800,)),
    (_dummy, (1500, 2)),
])
def test_resident_equals_the_three_launch_step(monkeypatch, make, steps):
    assert torch.cuda.is_available()
    a = _run(monkeypatch, True, make, steps)
    b = _run(monkeypatch, False, make, steps)
    assert set(a) == set(b) and "accepted" in a and "covariance" in a
    for k in a:
        if k == "launches":
            continue
        assert np.array_equal(a[k], b[k]), k
    assert np.all(a["total_steps"] == sum(steps) + 0)
    assert a["launches"] < b["launches"] / 50     # one launch per call instead of three per step


# (a user functor -- kind 8 -- has no code inside the resident kernel: it steps through its own launches)
@pytest.mark.parametrize("name", sorted(k for k in GOLDEN_CHAINS if GOLDEN_CHAINS[k][0] != 8))
def test_resident_reaches_the_golden_end_state(name):
    import smcmc_b200
    from oracle.cpu_checkers import STATE_FIELDS
    kind, dim, seed, chain, nsteps, start = GOLDEN_CHAINS[name]
    g = golden("chains.npz")
    want = golden_chain(g, name)
    lo = max(0, chain - 2)
    eng = smcmc_b200.Engine(kind, dim, 4, seed=seed, chain_offset=lo)
    bind_likelihood_inputs(eng, kind, g)
    configure_golden(name, eng, _set_field)
    x0 = np.zeros(dim) if start is None else np.full(dim, start)
    eng.start(x0)
    before = eng.launch_count()
    eng.step(nsteps)
    eng.sync()
    assert eng.launch_count() == before + 1 or dim >= 100     # n = 100: the state of a chain does not fit beside three others
    c = chain - lo
    exact = name in ("unit9_frozen_sigma",)
    n = dim
    packed = np.array([want["final_cov"][i, j] for i in range(n) for j in range(i + 1)])
    if exact:
        assert np.array_equal(eng.get("accepted")[c], want["x"][-1])
        assert np.array_equal(eng.get("covariance")[c], packed)
        assert np.array_equal(eng.get("decomposition")[c], want["final_decomp"])
        assert np.array_equal(eng.get("center")[c], want["final_center"])
    else:
        assert close(eng.get("accepted")[c], want["x"][-1])
        assert close(eng.get("covariance")[c], packed, 1e-10)
        assert np.allclose(eng.get("decomposition")[c], want["final_decomp"], rtol=1e-9, atol=1e-12)
    scal = dict(zip(STATE_FIELDS, want["final_scalars"]))
    assert eng.get("trials")[c] == scal["trials"]
    assert eng.get("successes")[c] == scal["successes"]
    assert eng.get("next_update")[c] == scal["next_update"]
    assert eng.get("total_steps")[c] == scal["total_steps"]
    assert eng.get("llh_calls")[c] == scal["llh_calls"]
