"""BASELINE.json's full size (4096 chains x 1 000 000 events): the oracle
cannot run this in seconds, so the checks are size-independent properties plus
a sampled comparison with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHAINS, EVENTS = 4096, 1_000_000


@pytest.fixture(scope="module")
def setup():
    import smcmc_b200
    from smcmc_b200 import synth
    signal = EVENTS // 3 + 1
    events = synth.make_mc_sample(signal, EVENTS - signal, seed=2)
    data = synth.make_data_histograms(33334, 33334, seed=2)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, CHAINS, seed=3)
    eng.set_fake_events(events)
    expo = synth.exposure_ratio(eng, data)
    eng.set_fake_data(data, expo)
    x = np.random.default_rng(5).uniform(-1, 1, (CHAINS, 9))
    return eng, events, data, expo, x


def test_filter_against_fp64_on_every_pair(setup):
    eng, events, data, expo, x = setup
    pairs, unsure, bad = eng.fake_filter_check(x)
    # a handful of events (|logSigma| > 36, i.e. reconstructed masses of a few
    # keV) are kept out of the FP32 path at upload and always evaluated in FP64
    assert CHAINS * (EVENTS - 20) <= pairs <= CHAINS * EVENTS
    assert bad == 0
    assert unsure < 0.005 * pairs


def test_counts_are_conserved_and_additive(setup):
    """Every event lands in at most one bin: per chain, the counts sum to the
    number of events minus the cut ones, class by class; and the counts of the
    full sample equal the sum of the counts of its two halves (a checksum of
    checksums over 4 x 10^9 pairs)."""
    import smcmc_b200
    eng, events, data, expo, x = setup
    full = eng.fake_counts(x[:512]).astype(np.int64)
    n_sig0 = int(((events["Type"] == 0) & (events["MuDk"] == 0)).sum())
    n_sig1 = int(((events["Type"] == 0) & (events["MuDk"] > 0)).sum())
    n_bkg0 = int(((events["Type"] > 0) & (events["MuDk"] == 0)).sum())
    n_bkg1 = int(((events["Type"] > 0) & (events["MuDk"] > 0)).sum())
    assert np.all(full[:, 0:100].sum(1) <= n_sig0) and np.all(full[:, 100:150].sum(1) <= n_sig1)
    assert np.all(full[:, 150:250].sum(1) <= n_bkg0) and np.all(full[:, 250:300].sum(1) <= n_bkg1)
    assert np.all(full[:, 300:] == 0)
    assert np.all(full[:, 0:100].sum(1) > 0.95 * n_sig0)          # signal sits inside [0,500)
    halves = []
    for part in (events[0::2], events[1::2]):
        e = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 512, seed=3)
        e.set_fake_events(part)
        e.set_fake_data(data, expo)
        halves.append(e.fake_counts(x[:512]).astype(np.int64))
    assert np.array_equal(full, halves[0] + halves[1])


def test_sampled_chains_against_oracle(setup, checkers):
    eng, events, data, expo, x = setup
    llh = eng.eval(x)
    orc = checkers.CpuChain("orc", checkers.LLH_FAKE, 9, 1, 0)
    orc.set_fake(events, data, expo)
    for c in (0, 1777, 4095):
        want = orc.llh(x[c])
        assert abs(llh[c] - want) <= 1e-12 * abs(want)
        assert np.array_equal(eng.fake_counts(x[c:c + 1])[0], orc.fake_counts(x[c]))
    assert np.all(np.isfinite(llh))


def test_ensemble_runs_and_adapts(setup):
    eng, events, data, expo, x = setup
    assert eng.start(x).all()
    eng.step(30)
    acc = eng.get("acceptance")
    assert np.all(np.isfinite(acc)) and 0.02 < acc.mean() < 0.9
    assert np.all(eng.get("total_steps") == 30)
    assert np.all(eng.get("llh_calls") == 31)
    assert np.all(eng.get("status") == 0)
