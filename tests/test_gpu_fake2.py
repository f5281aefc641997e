"""Parity of the device likelihood SMCMC_LLH_FAKE2 (libsmcmc_b200.so, through
the C ABI) with the oracle for example2/FakeLikelihood.H:58-118, 222-289: the
event corrections and cuts of example/ filled into separate signal and
background histograms, each renormalised to the event counts x[0], x[1] by its
integral, plus the penalty terms.

The (chain, event) pair kernel and its integer count table are the ones of
SMCMC_LLH_FAKE; what is new on the device is the weight set of
example2/SystematicCorrection.H:75-117 and kFake2Finish.  Tolerance: 1e-12
relative on bin contents and on the log-likelihood (last-ulp differences of
atan/log between CUDA's and the host's libm); the Metropolis accept sequence
must be identical.
"""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def hist_close(a, b):
    return np.all((np.abs(a - b) <= RTOL * np.abs(b)) | (np.isnan(a) & np.isnan(b)))


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def make_engine(events, data, chains=32, seed=1, chain_offset=0):
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE2, 9, chains, seed=seed, chain_offset=chain_offset)
    eng.set_fake_events(events)
    eng.set_fake_data(data, 1.0)
    return eng


@pytest.mark.parametrize("tag", ["", "_irregular"])
def test_golden_grid(tag):
    """Values produced by the reference build (tests/golden/make_golden.py fake2)."""
    assert torch.cuda.is_available()
    g = golden("fake2_likelihood.npz")
    eng = make_engine(g["events" + tag], g["data"])
    llh = eng.eval(g["points"])
    hist = eng.fake_histograms(g["points"])
    assert hist_close(hist, g["hist" + tag])
    assert rel(llh, g["llh" + tag]) < RTOL


def test_against_oracle_on_seeded_inputs(checkers):
    import smcmc_b200
    events, data = smcmc_b200.synth.fake2_inputs(2000, 2000, 10, seed=41)     # ~92 000 events
    eng = make_engine(events, data, chains=300)
    orc = checkers.CpuChain("orc", checkers.LLH_FAKE2, 9, 1, 0)
    orc.set_fake(events, data, 1.0)
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.normal(0, 1, (200, 9)), rng.normal(0, 5, (100, 9))])
    pts[:, 0] = rng.uniform(-200, 5000, 300)
    pts[:, 1] = rng.uniform(-200, 5000, 300)
    pts[7, 2] = 2000.0          # every event cut: x / 0 normalisation, NaN likelihood
    llh = eng.eval(pts)
    hist = eng.fake_histograms(pts[:40])
    counts = eng.fake_counts(pts[:40])
    for i in range(40):
        assert np.array_equal(counts[i], orc.fake_counts(pts[i])), i
        assert hist_close(hist[i], orc.fake_hist(pts[i])), i
    want = np.array([orc.llh(p) for p in pts])
    assert np.isnan(want[7]) and np.isnan(llh[7])
    keep = ~np.isnan(want)
    assert rel(llh[keep], want[keep]) < RTOL
    assert (pts[keep, 0] < 0).any() and (pts[keep, 1] < 0).any()      # penalty branches exercised


def test_golden_chains():
    """example2/FakeMCMC.C's schedule in miniature: burn-in, ResetProposal,
    burn-in, UpdateProposal, run -- the reference's accept / reject sequence."""
    g = golden("fake2_likelihood.npz")
    for chain in (0, 5):
        eng = make_engine(g["events"], g["data"], chains=2, seed=777, chain_offset=chain)
        eng.set_gaussian(0, 15.0)
        eng.set_gaussian(1, 15.0)
        ok = eng.start(np.tile(g["chain%d_x0" % chain], (2, 1)))
        assert ok.all()
        parts = [eng.step_trace(80)]
        eng.reset_proposal()
        parts.append(eng.step_trace(80))
        eng.update_proposal()
        parts.append(eng.step_trace(140))
        acc = np.concatenate([p["accepted"][:, 0] for p in parts])
        pts = np.concatenate([p["points"][:, 0] for p in parts])
        llh = np.concatenate([p["llh_proposed"][:, 0] for p in parts])
        assert np.array_equal(acc, g["chain%d_accepted" % chain])
        assert acc.sum() > 15
        scale = np.maximum(np.abs(g["chain%d_x" % chain]), 1.0)
        assert np.all(np.abs(pts - g["chain%d_x" % chain]) <= 1e-11 * scale)
        assert rel(llh, g["chain%d_llh_proposed" % chain]) < 1e-10


def test_event_order_and_tiling_do_not_matter():
    """Counts are integers: 4096 chains (16 point tiles) give the same values as 3."""
    g = golden("fake2_likelihood.npz")
    pts = np.tile(g["points"][:3], (1366, 1))[:4096]
    big = make_engine(g["events"], g["data"], chains=4096).eval(pts)
    small = make_engine(g["events"], g["data"], chains=3).eval(g["points"][:3])
    assert np.array_equal(big[:3], small)
    assert np.array_equal(big[3:6], small)
