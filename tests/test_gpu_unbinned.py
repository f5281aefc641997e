"""The unbinned mixture likelihood (SMCMC_LLH_UNBINNED, BASELINE.json
configs[4]) on the device against its CPU checker.  The functor is defined in
this repository (the reference's likelihood is binned), so the checker is the
oracle port, pinned in tests/test_oracle.py against a numpy statement of the
definition.  Requirement: 1e-12 relative (the device adds the per-event terms
chunk by chunk, the checker sequentially; exp/log are CUDA's and glibc's)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def setup(n_sig, n_bkg, chains, seed=9):
    import smcmc_b200
    from smcmc_b200 import synth
    events = synth.make_mc_sample(n_sig, n_bkg, seed=seed)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNBINNED, 9, chains, seed=12)
    eng.set_unbinned_events(events)
    return eng, events


def oracle_llh(events, pts):
    from oracle import cpu_checkers as cc
    c = cc.CpuChain("orc", cc.LLH_UNBINNED, 9, 12, 0)
    c.set_fake(events, np.zeros(150), 1.0)
    return np.array([c.llh(p) for p in pts])


def test_likelihood_matches_the_checker():
    eng, events = setup(7000, 13000, 64)
    rng = np.random.default_rng(1)
    pts = np.concatenate([np.zeros((1, 9)), rng.uniform(-1, 1, (150, 9)), rng.normal(0, 4, (49, 9))])
    dev = eng.eval(pts)
    ref = oracle_llh(events, pts)
    assert np.all(np.isfinite(dev))
    assert np.max(np.abs(dev / ref - 1.0)) < 1e-12


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 4097])
def test_ragged_sample_sizes(n):
    import smcmc_b200
    from smcmc_b200 import synth
    events = synth.make_mc_sample(n // 3, n - n // 3, seed=3)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNBINNED, 9, 5, seed=1)
    eng.set_unbinned_events(events)
    pts = np.random.default_rng(n).uniform(-1, 1, (37, 9))
    dev = eng.eval(pts)
    ref = oracle_llh(events, pts) if n else np.zeros(37)
    assert np.allclose(dev, ref, rtol=1e-12, atol=0)


def test_non_positive_mass_makes_the_likelihood_nan():
    """log of a non-positive mass: NaN in the checker, NaN on the device, and
    the sampler treats the point like the reference treats any non-finite
    likelihood (TSimpleMCMC.H:432-436): rejected."""
    import smcmc_b200
    eng, events = setup(50, 50, 4)
    bad = events.copy()
    bad["Mass"][7] = -1.0
    eng.set_unbinned_events(bad)
    assert np.all(np.isnan(eng.eval(np.zeros((3, 9)))))
    assert np.all(np.isnan(oracle_llh(bad, np.zeros((1, 9)))))


def test_metropolis_chains_follow_the_checker():
    from oracle import cpu_checkers as cc
    eng, events = setup(1500, 2500, 6)
    x0 = np.random.default_rng(8).uniform(-1, 1, (6, 9))
    assert eng.start(x0).all()
    tr = eng.step_trace(150)
    for c in (0, 5):
        o = cc.CpuChain("orc", cc.LLH_UNBINNED, 9, 12, c)
        o.set_fake(events, np.zeros(150), 1.0)
        o.start(x0[c])
        want = o.step(150)
        assert np.array_equal(tr["accepted"][:, c], want["accepted"])
        assert np.allclose(tr["llh_accepted"][:, c], want["llh_accepted"], rtol=1e-11)
        assert np.allclose(tr["points"][:, c], want["x"], rtol=1e-9, atol=1e-12)
