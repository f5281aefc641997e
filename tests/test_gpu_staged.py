"""kProposeStaged (csrc/proposal_staged.cuh: covariance row and Cholesky factor
moved by TMA, covariance update in shared memory, shared-divisor division) must
leave every chain in EXACTLY the state kPropose (csrc/proposal.cuh, per-lane
global loads, __ddiv_rn) leaves it in: same bits in every accepted point, in the
covariance (TSimpleMCMC.H:1795-1820), its factor (:1103-1120), the central point
and every scalar, through several UpdateProposal steps (:1824-1826).  The golden
chains of test_gpu_chains.py pin the staged kernel against the reference build;
this file pins it against the global-memory kernel on ensembles and dimensions
the golden set does not hold.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIELDS = ("accepted", "center", "covariance", "decomposition", "sigma", "acceptance", "acceptance_trials",
          "acceptance_rigidity", "trials", "successes", "next_update", "covariance_trials", "center_trials",
          "step_rms", "accepted_llh", "sigma_trace", "total_steps")


def _run(kind, dim, chains, steps, generic, configure=None, seed=11):
    import smcmc_b200
    from smcmc_b200 import binding
    if generic:
        os.environ["SMCMC_PROPOSE_GENERIC"] = "1"
    else:
        os.environ.pop("SMCMC_PROPOSE_GENERIC", None)
    try:
        eng = smcmc_b200.Engine(kind, dim, chains, seed=seed)
    finally:
        os.environ.pop("SMCMC_PROPOSE_GENERIC", None)
    # a short acceptance window: the first UpdateProposal comes after 12 accepted steps
    eng.prop_set(binding.PROP_ACCEPTANCE_WINDOW, 12.0)
    if configure:
        configure(eng, binding)
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-0.01, 0.01, (chains, dim))
    ok = eng.start(x0)
    assert ok.all()
    tr = eng.step_trace(steps, want=("accepted", "points", "llh_proposed"))
    out = {f: eng.get(f) for f in FIELDS}
    out["trace_accepted"] = tr["accepted"]
    out["trace_points"] = tr["points"]
    out["trace_llh"] = tr["llh_proposed"]
    out["launches"] = eng.launch_count()
    return out


def _same(a, b):
    for k in a:
        if k == "launches":
            continue
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def _uniform_dims(eng, binding):
    eng.set_uniform(3, -0.5, 0.75)
    eng.set_uniform(20, -1.0, 1.0)
    eng.set_gaussian(7, 0.3)
    eng.set_correlation(1, 2, 0.4)


def _frozen(eng, binding):
    eng.prop_set(binding.PROP_COVARIANCE_FROZEN, 1.0)


def _deweighted(eng, binding):
    # fractional trial counts: the divisor of the covariance update is not an integer
    eng.prop_set(binding.PROP_COVARIANCE_DEWEIGHT, 0.37)
    eng.prop_set(binding.PROP_COVARIANCE_WINDOW, 150.0)


@pytest.mark.parametrize("kind_name,dim,configure", [
    ("LLH_HORRIFIC", 50, None),
    ("LLH_ASYM", 50, None),
    ("LLH_UNIT_GAUSS", 9, None),
    ("LLH_UNIT_GAUSS", 33, _uniform_dims),
    ("LLH_UNIT_GAUSS", 2, None),
    ("LLH_UNIT_GAUSS", 64, _deweighted),
    ("LLH_UNIT_GAUSS", 21, _frozen),
])
def test_staged_equals_global_memory_kernel(kind_name, dim, configure):
    import smcmc_b200
    assert torch.cuda.is_available()
    kind = getattr(smcmc_b200, kind_name)
    chains, steps = 333, 400
    staged = _run(kind, dim, chains, steps, False, configure)
    plain = _run(kind, dim, chains, steps, True, configure)
    _same(staged, plain)
    # the chains moved and the proposal was refactored along the way
    assert staged["trace_accepted"].sum() > chains
    assert (staged["successes"] > 12).mean() > 0.5     # most chains went through UpdateProposal


def test_large_dimension_keeps_the_global_memory_kernel():
    """n = 200: the rows do not fit four warps per SM; the engine must still run."""
    import smcmc_b200
    a = _run(smcmc_b200.LLH_UNIT_GAUSS, 200, 8, 30, False)
    b = _run(smcmc_b200.LLH_UNIT_GAUSS, 200, 8, 30, True)
    _same(a, b)


def test_shared_divisor_division_is_correctly_rounded():
    """2^32 random (numerator, divisor) cases, including the ends of the
    significand range, tiny, huge, zero, infinite and NaN numerators."""
    from smcmc_b200 import binding
    assert binding.selftest_division(1 << 32, seed=7) == 0
    assert binding.selftest_division(1 << 28, seed=123456789) == 0
