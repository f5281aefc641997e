"""Parity of the device event likelihood (libsmcmc_b200.so, called through the
C ABI) with the oracle: example/FakeLikelihood.H:47-81,188-216.

Tolerances: the integer event counts per (weight class, histogram, bin) must
be IDENTICAL to the oracle's; bin contents (count x weight, summed in the
reference's order) and the log-likelihood must agree to 1e-12 relative, the
bound BASELINE.json states for FP64 -- the only difference left is the last
ulp of exp/atan/log between CUDA's and the host's libm.
"""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def hist_close(a, b):
    return np.all(np.abs(a - b) <= RTOL * np.abs(b))


def make_engine(events, data, exposure, chains=32, seed=1):
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, chains, seed=seed)
    eng.set_fake_events(events)
    eng.set_fake_data(data, exposure)
    return eng


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("tag", ["", "_irregular"])
def test_golden_grid(tag):
    """The 19-point parameter grid of example/TestLikelihood.C plus random
    points, values produced by the reference build."""
    assert torch.cuda.is_available()
    g = golden("fake_likelihood.npz")
    eng = make_engine(g["events" + tag], g["data"], float(g["exposure"]))
    llh = eng.eval(g["points"])
    hist = eng.fake_histograms(g["points"])
    assert hist_close(hist, g["hist" + tag])
    assert rel(llh, g["llh" + tag]) < RTOL


def test_against_oracle_on_seeded_inputs(checkers):
    import smcmc_b200
    events, data = smcmc_b200.synth.fake_inputs(3000, 3000, 10, seed=31)      # 90 000 events
    eng = make_engine(events, data, 1.0, chains=300)
    expo = smcmc_b200.synth.exposure_ratio(eng, data)
    eng.set_fake_data(data, expo)
    orc = checkers.CpuChain("orc", checkers.LLH_FAKE, 9, 1, 0)
    orc.set_fake(events, data, 1.0)
    sim0 = orc.fake_hist(np.zeros(9))
    total = 0.0
    for h in range(3):
        part = 0.0
        for v in sim0[h * 50:(h + 1) * 50]:
            part += v
        total += part
    assert expo == float(sum(data[:50]) + sum(data[50:100]) + sum(data[100:])) / total
    orc.set_fake(events, data, expo)
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-1, 1, (200, 9)), rng.normal(0, 6, (100, 9))])
    llh = eng.eval(pts)                      # 300 points: not a multiple of the 256-chain tile
    hist = eng.fake_histograms(pts[:40])
    counts = eng.fake_counts(pts[:40])
    for i in range(40):
        assert np.array_equal(counts[i], orc.fake_counts(pts[i])), i
        assert hist_close(hist[i], orc.fake_hist(pts[i])), i
    want = np.array([orc.llh(p) for p in pts])
    assert rel(llh, want) < RTOL


def test_generic_path_equals_fast_path(monkeypatch):
    """Every event through the straight per-pair transcription (the path the
    irregular events take) gives the same integer counts as the fast path."""
    import smcmc_b200
    g = golden("fake_likelihood.npz")
    fast = make_engine(g["events"], g["data"], float(g["exposure"]))
    monkeypatch.setenv("SMCMC_FAKE_FORCE_GENERIC", "1")
    slow = make_engine(g["events"], g["data"], float(g["exposure"]))
    monkeypatch.delenv("SMCMC_FAKE_FORCE_GENERIC")
    pts = g["points"]
    assert np.array_equal(fast.fake_counts(pts), slow.fake_counts(pts))
    assert np.array_equal(fast.fake_histograms(pts), slow.fake_histograms(pts))
    assert np.array_equal(fast.eval(pts), slow.eval(pts))


def test_exact_mode_equals_filter_mode(monkeypatch):
    """SMCMC_FAKE_EXACT=1 sends every pair through the FP64 arithmetic; the
    default FP32-filter + FP64-fallback evaluation must count identically."""
    g = golden("fake_likelihood.npz")
    filt = make_engine(g["events"], g["data"], float(g["exposure"]))
    monkeypatch.setenv("SMCMC_FAKE_EXACT", "1")
    exact = make_engine(g["events"], g["data"], float(g["exposure"]))
    monkeypatch.delenv("SMCMC_FAKE_EXACT")
    pts = np.concatenate([g["points"], np.random.default_rng(4).normal(0, 12, (60, 9))])
    assert np.array_equal(filt.fake_counts(pts), exact.fake_counts(pts))
    assert np.array_equal(filt.eval(pts), exact.eval(pts))


def test_filter_decisions_never_differ_from_fp64():
    """Every (point, event) pair evaluated both ways: a decision taken by the
    FP32 interval filter must equal the FP64 decision; undecided pairs go to
    FP64 anyway.  Typical points leave well under 1% of the pairs undecided."""
    import smcmc_b200
    events, data = smcmc_b200.synth.fake_inputs(4000, 4000, 10, seed=77)     # 120 000 events
    eng = make_engine(events, data, 0.1, chains=8)
    rng = np.random.default_rng(12)
    typical = rng.uniform(-1, 1, (256, 9))
    pairs, unsure, bad = eng.fake_filter_check(typical)
    assert pairs == 256 * len(events) and bad == 0
    assert unsure < 0.01 * pairs
    wild = np.concatenate([rng.normal(0, 10, (200, 9)), rng.normal(0, 60, (56, 9))])
    wild[5, 3] = -900.0
    wild[6, 4] = 1e5
    wild[7, 2] = np.nan
    wild[8, 6] = np.inf
    pairs, unsure, bad = eng.fake_filter_check(wild)
    assert pairs == 256 * len(events) and bad == 0


def test_event_order_does_not_matter():
    """Integer counting: any permutation inside the signal block and inside
    the background block leaves every bit of the result unchanged."""
    g = golden("fake_likelihood.npz")
    ev = g["events"]
    rng = np.random.default_rng(0)
    sig = ev[ev["Type"] == 0]
    bkg = ev[ev["Type"] != 0]
    shuffled = np.concatenate([rng.permutation(sig), rng.permutation(bkg)])
    a = make_engine(ev, g["data"], float(g["exposure"]))
    b = make_engine(shuffled, g["data"], float(g["exposure"]))
    assert np.array_equal(a.eval(g["points"]), b.eval(g["points"]))
    assert np.array_equal(a.fake_histograms(g["points"]), b.fake_histograms(g["points"]))


@pytest.mark.parametrize("nev", [0, 1, 2, 127, 128, 129, 8191, 8192, 8193, 20001])
def test_ragged_event_counts(checkers, nev):
    """Empty sample, one event, tile (128) and chunk (8192) boundaries."""
    import smcmc_b200
    events = smcmc_b200.synth.make_mc_sample(nev // 3, nev - nev // 3, seed=nev + 1)
    data = smcmc_b200.synth.make_data_histograms(300, 300, seed=2)
    eng = make_engine(events, data, 0.37, chains=5)
    orc = checkers.CpuChain("orc", checkers.LLH_FAKE, 9, 1, 0)
    orc.set_fake(events, data, 0.37)
    pts = np.random.default_rng(nev).uniform(-2, 2, (5, 9))
    hist = eng.fake_histograms(pts)
    counts = eng.fake_counts(pts)
    llh = eng.eval(pts)
    for i in range(5):
        assert np.array_equal(counts[i], orc.fake_counts(pts[i]))
        assert hist_close(hist[i], orc.fake_hist(pts[i]))
    assert rel(llh, np.array([orc.llh(p) for p in pts])) < RTOL


@pytest.mark.parametrize("kernel", ["stream", "pairs"])
def test_extreme_parameters(checkers, kernel):
    """Parameter points that push masses out of range, collapse the width, or
    are not finite: the cut / overflow handling must follow the reference
    (FakeLikelihood.H:203-205 and TH1's under/overflow bins).  Twelve points go through the
    streaming kernel (<= 16 chains); "pairs" evaluates them three times over in one call so that
    kFakePairs takes them -- including the chains whose constants leave the range the FP32 filter's
    error analysis covers (|c2| > 24, |w2| > 2^24, non-finite: every pair in FP64) and chains
    just inside it, whose bound is so wide that most pairs are undecided."""
    g = golden("fake_likelihood.npz")
    eng = make_engine(g["events"], g["data"], float(g["exposure"]), chains=64)
    orc = checkers.CpuChain("orc", checkers.LLH_FAKE, 9, 1, 0)
    orc.set_fake(g["events"], g["data"], float(g["exposure"]))
    pts = np.zeros((18, 9))
    pts[12, 2] = 200.0        # c2 = 25.5: outside the filter's range
    pts[13, 2] = -250.0       # c2 = -39.4
    pts[14, 3] = 200.0        # width e^20: w2 > 2^24
    pts[15, 3] = 160.0        # w2 = 1.3e7 < 2^24: filter on, bound useless
    pts[16, 2] = 160.0        # c2 = 19.8: filter on, K = 2^19.8
    pts[17, 2], pts[17, 3] = -140.0, 100.0
    pts[0, 2] = 40.0          # mass scale e^4: everything above 500
    pts[1, 2] = -60.0         # everything in the first bin
    pts[2, 3] = -400.0        # width -> 0: all events at the nominal mass
    pts[3, 3] = 60.0          # width e^6
    pts[4, 4] = 500.0         # skew saturates at 0.3
    pts[5, 4] = -500.0
    pts[6, 5] = 7000.0        # separation scale overflows to inf
    pts[7, 6] = -8000.0       # separation scale underflows to 0
    pts[8, 7] = 1e6
    pts[9, 8] = -1e6
    pts[10, 2] = np.nan
    pts[11, 3] = np.inf
    base = len(pts)
    if kernel == "pairs":
        pts = np.concatenate([pts, pts, pts])
    hist = eng.fake_histograms(pts)
    counts = eng.fake_counts(pts)
    llh = eng.eval(pts)
    for i in range(len(pts)):
        if i >= base:                                  # the repeats: identical to the first copy
            assert np.array_equal(counts[i], counts[i - base]) and np.array_equal(llh[i], llh[i - base], equal_nan=True), i
            continue
        assert np.array_equal(counts[i], orc.fake_counts(pts[i])), i
        want_hist = orc.fake_hist(pts[i])
        ok = np.isfinite(want_hist)
        assert hist_close(hist[i][ok], want_hist[ok]), i
        assert np.array_equal(np.isfinite(hist[i]), ok), i
        want = orc.llh(pts[i])
        if np.isfinite(want):
            assert abs(llh[i] - want) <= RTOL * abs(want), i
        else:
            assert np.isnan(llh[i]) == np.isnan(want) and (np.isnan(want) or llh[i] == want), i


def test_missing_inputs_fail_loudly():
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 4)
    with pytest.raises(smcmc_b200.SmcmcError) as err:
        eng.eval(np.zeros((1, 9)))
    assert err.value.status == -2
    with pytest.raises(smcmc_b200.SmcmcError) as err:
        eng.step(1)
    assert err.value.status == -1          # Step before Start: std::invalid_argument (:371-374)


@pytest.mark.parametrize("chains", [1, 3, 8, 16])
def test_streaming_kernel_counts_equal_the_pair_kernel(chains, monkeypatch):
    """Up to 16 chains the events are streamed once with the chains looped per event (kFakeStream);
    more chains use the chain-per-thread pair kernel.  Same filter, same FP64 fallback: the integer
    counts, and therefore the histograms and likelihoods, are identical."""
    import smcmc_b200
    events, data = smcmc_b200.synth.fake_inputs(700, 900, 10, seed=13)        # 25 000 events, ragged tiles
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(-1, 1, (4, 9)), rng.normal(0, 6, (4, 9)), rng.normal(0, 2, (24, 9))])[:chains]
    stream = make_engine(events, data, 0.1, chains=chains)
    c_stream, l_stream = stream.fake_counts(pts), stream.eval(pts)
    monkeypatch.setenv("SMCMC_FAKE_NO_STREAM", "1")
    pair = make_engine(events, data, 0.1, chains=chains)
    c_pair, l_pair = pair.fake_counts(pts), pair.eval(pts)
    monkeypatch.delenv("SMCMC_FAKE_NO_STREAM")
    assert np.array_equal(c_stream, c_pair)
    assert np.array_equal(l_stream, l_pair)
    assert c_stream.sum() > 1000 * chains
    # and the exact (all-FP64) evaluation agrees as well
    monkeypatch.setenv("SMCMC_FAKE_EXACT", "1")
    exact = make_engine(events, data, 0.1, chains=chains)
    assert np.array_equal(exact.fake_counts(pts), c_stream)


def test_pair_kernel_is_deterministic_under_repetition():
    """compute-sanitizer is closed on this pool, so the barrier-free tile hand-over, the shared-memory
    RED counters, the provisional count / take-back of undecided pairs and the per-warp FP64 queues of
    kFakePairs are stressed the only way left: the same 700 points (a partly filled last CTA), 40
    evaluations, on an event set with ragged class sizes -- every integer of every count table must come
    out the same, and equal to the all-FP64 evaluation."""
    import smcmc_b200
    import os
    events, data = smcmc_b200.synth.fake_inputs(9000, 7000, 6, seed=5)
    eng = make_engine(events, data, 0.15, chains=700)
    pts = np.random.default_rng(8).normal(0, 3, (700, 9))
    first = eng.fake_counts(pts)
    for _ in range(40):
        assert np.array_equal(eng.fake_counts(pts), first)
    os.environ["SMCMC_FAKE_EXACT"] = "1"
    try:
        exact = make_engine(events, data, 0.15, chains=700)
    finally:
        del os.environ["SMCMC_FAKE_EXACT"]
    assert np.array_equal(exact.fake_counts(pts), first)
