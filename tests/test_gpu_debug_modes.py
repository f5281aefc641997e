"""The proposal's debugging modes and the remaining setters of the reference API
on the device: ForceStep (TSimpleMCMC.H:811-818), SetScanDimension (:820-830, the
short-circuits of operator() :671-704), SetEstimatedCenter (:733-739); a repeated
Start() (InitializeState runs once, :1680-1681); trace rows of chains that are not
running."""
import numpy as np
import pytest

from helpers import configure_debug_modes, golden, golden_chain, run_debug_modes

pytestmark = pytest.mark.gpu


class _AsChain:
    """The Engine behind the method names of oracle.cpu_checkers.CpuChain, for one
    traced chain of the ensemble."""

    def __init__(self, eng, c):
        self.eng, self.c = eng, c

    def step(self, n, metropolis=0):
        tr = self.eng.step_trace(n, metropolis)
        c = self.c
        return {"accepted": tr["accepted"][:, c], "llh_accepted": tr["llh_accepted"][:, c],
                "llh_proposed": tr["llh_proposed"][:, c], "x": tr["points"][:, c], "sigma": tr["sigma"][:, c]}

    def force_step(self, x):
        self.eng.force_step(x)

    def set_scan(self, d):
        self.eng.set_scan(d)

    def set_center(self, v):
        self.eng.set_center(v)


def test_debug_modes_match_the_reference_build():
    """Chain 5 of seed 61 inside a 6-chain ensemble: the golden run "debug9" of the
    reference build -- accept sequence identical, points to 1e-12 (the adaptive step
    size goes through pow)."""
    import smcmc_b200
    want = golden_chain(golden("chains.npz"), "debug9")
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 9, 6, seed=61, chain_offset=2)
    configure_debug_modes(eng)
    assert eng.start(np.full(9, 0.2)).all()
    got = run_debug_modes(_AsChain(eng, 3))
    assert np.array_equal(got["accepted"], want["accepted"])
    assert np.allclose(got["x"], want["x"], rtol=1e-12, atol=1e-14)
    assert np.allclose(got["llh_proposed"], want["llh_proposed"], rtol=1e-11, atol=1e-13)
    assert np.allclose(got["sigma"], want["sigma"], rtol=1e-12)
    assert np.array_equal(got["x"][60], np.linspace(-0.4, 0.4, 9))          # the forced point, taken
    from oracle.cpu_checkers import STATE_FIELDS
    scal = dict(zip(STATE_FIELDS, want["final_scalars"]))
    assert eng.get("trials")[3] == scal["trials"]                            # forced / scan steps are not trials
    assert eng.get("total_steps")[3] == scal["total_steps"] == 187
    assert eng.get("llh_calls")[3] == scal["llh_calls"]
    assert np.allclose(eng.get("center")[3], want["final_center"], rtol=1e-12, atol=1e-14)
    assert abs(eng.get("step_rms")[3] - scal["step_rms"]) <= 1e-12 * scal["step_rms"]


def test_force_step_per_chain_and_many_steps():
    """One forced point per chain; smcmc_step(n) uses it for its first step only."""
    import smcmc_b200
    E, n = 300, 4
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=3)
    eng.start(np.zeros(n))
    pts = np.random.default_rng(1).normal(0, 1, (E, n))
    eng.force_step(pts)
    eng.step(1, 2)
    assert np.array_equal(eng.get("accepted"), pts)
    assert np.all(eng.get("trials") == 0) and np.all(eng.get("total_steps") == 1)
    eng.force_step(np.full(n, 0.25))
    eng.step(10)                                   # first step forced, nine regular steps
    assert np.all(eng.get("total_steps") == 11) and np.all(eng.get("trials") == 9)
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.force_step(np.zeros(n + 1))


def test_second_start_keeps_the_adapted_proposal(checkers):
    """Start() again: the point moves, trials / acceptance / covariance / next update stay
    (reference :246-276 with InitializeState's fStateInitialized guard :1680-1681) --
    bit for bit against the port, frozen step size."""
    import smcmc_b200
    from smcmc_b200 import binding
    cc = checkers
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 5, 3, seed=8)
    eng.prop_set(binding.PROP_ACCEPTANCE_RIGIDITY, -1.0)
    eng.prop_set(binding.PROP_SIGMA, 0.6)
    eng.start(np.zeros(5))
    eng.step(200)
    before = {k: eng.get(k).copy() for k in ("trials", "successes", "covariance", "next_update", "acceptance")}
    assert eng.start(np.full(5, 0.3)).all()
    for k, v in before.items():
        assert np.array_equal(eng.get(k), v), k
    assert np.array_equal(eng.get("accepted"), np.full((3, 5), 0.3))
    tr = eng.step_trace(100)
    for c in range(3):
        o = cc.CpuChain("orc", cc.LLH_UNIT_GAUSS, 5, 8, c)
        o.set(cc.SET_ACCEPTANCE_RIGIDITY, -1.0)
        o.set(cc.SET_SIGMA, 0.6)
        o.start(np.zeros(5))
        o.step(200)
        o.start(np.full(5, 0.3))
        w = o.step(100)
        assert np.array_equal(tr["accepted"][:, c], w["accepted"])
        assert np.array_equal(tr["points"][:, c], w["x"])


def test_trace_rows_of_chains_that_are_not_running():
    """A chain whose Start() failed does not step; its trace rows repeat its standing
    state instead of holding whatever the buffer contained."""
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 6, 4, seed=2)
    x0 = np.zeros((4, 6))
    x0[2] = 1e200                                  # likelihood -inf: Start fails for this chain (:265-268)
    ok = eng.start(x0)
    assert list(ok) == [1, 1, 0, 1]
    tr = eng.step_trace(12)
    assert np.all(tr["accepted"][:, 2] == 0)
    assert np.all(tr["points"][:, 2] == 1e200)
    assert np.all(tr["llh_proposed"][:, 2] == -np.inf)
    assert np.all(np.isfinite(tr["step_rms"])) and np.all(np.isfinite(tr["sigma"]))
    assert tr["accepted"][:, [0, 1, 3]].sum() > 0
