"""kAcceptLocal (csrc/accept_local.cuh): for a chain-local likelihood and a step without a
trace the likelihood of the proposed point and the Metropolis rule run in ONE launch, the
proposed rows staged coalesced through shared memory.  The chains must be bit-identical to
the ones kSimpleLikelihood + kAccept give (SMCMC_NO_ACCEPT_LOCAL=1)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIELDS = ("accepted", "proposed", "accepted_llh", "proposed_llh", "sigma", "trials", "successes", "next_update",
          "total_steps", "llh_calls", "step_rms", "center", "covariance", "status", "acceptance")


def _run(monkeypatch, fused, make, steps):
    import smcmc_b200
    monkeypatch.setenv("SMCMC_NO_RESIDENT", "1")          # the step-by-step path is the one under test
    if fused:
        monkeypatch.delenv("SMCMC_NO_ACCEPT_LOCAL", raising=False)
    else:
        monkeypatch.setenv("SMCMC_NO_ACCEPT_LOCAL", "1")
    eng = make(smcmc_b200)
    for k in steps:
        eng.step(k)
    eng.sync()
    out = {f: eng.get(f) for f in FIELDS}
    out["launches"] = eng.launch_count()
    out["tail"] = eng.step_trace(4, want=("accepted", "points"))["points"]
    return out


def _horrific(sm):
    eng = sm.Engine(sm.LLH_HORRIFIC, 20, 257, seed=9)          # 257 chains: a warp tile with one chain
    eng.start(np.zeros(20))
    return eng


def _asym_odd(sm):
    eng = sm.Engine(sm.LLH_ASYM, 7, 1000, seed=3)
    eng.start(np.full(7, 0.01))
    return eng


def _hard(sm):
    eng = sm.Engine(sm.LLH_HARD, 6, 40, seed=13)
    eng.start(np.full(6, 0.5))
    return eng


def _vaat(sm):
    eng = sm.Engine(sm.LLH_UNIT_GAUSS, 7, 64, seed=4, proposal=sm.PROPOSAL_VAAT)
    eng.start(np.zeros(7))
    return eng


def _metropolis_box(sm):
    eng = sm.Engine(sm.LLH_HORRIFIC, 50, 96, seed=21)           # starts at the edge of the box: -1e30 proposals
    eng.start(np.full(50, 0.999))
    return eng


@pytest.mark.parametrize("make", [_horrific, _asym_odd, _hard, _vaat, _metropolis_box])
def test_fused_accept_equals_likelihood_kernel_plus_accept(monkeypatch, make):
    assert torch.cuda.is_available()
    steps = (300, 1, 40)
    a = _run(monkeypatch, True, make, steps)
    b = _run(monkeypatch, False, make, steps)
    for k in a:
        if k == "launches":
            continue
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert a["launches"] < b["launches"]
    assert np.all(a["llh_calls"] >= sum(steps))
