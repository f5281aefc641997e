"""Synthetic inputs of the shapes BASELINE.json names (there is no network
for real data): the toy generators of the reference restated with numpy.

  * MC sample   : Simulated::MakeSample            example/Simulated.H:17-53
  * toy data    : FakeData::MakeSample / FillData  example/FakeData.H:32-117

The distributions follow the reference; the random stream is numpy's
(the reference's own runs are not reproducible either: FakeMCMC.C:19 seeds
from the clock).  tests/test_synth.py checks the moments against the
reference generators run through the oracle build.
"""
import numpy as np

from .binding import EVENT_DTYPE


def _positive_normal(rng, mean, sigma):
    """Gaus(mean, sigma), redrawn while negative (Simulated.H:33-35,45-47)."""
    out = rng.normal(mean, sigma)
    bad = out < 0
    while bad.any():
        out[bad] = rng.normal(np.broadcast_to(mean, out.shape)[bad],
                              np.broadcast_to(sigma, out.shape)[bad])
        bad = out < 0
    return out


def make_mc_sample(signal, background, seed, variant=1):
    """Simulated::MakeSample(signal, background): signal events first.
    variant=2: example2/Simulated.H:17-63 (at least 1000 signal and as many
    background events, the background counted only while Mass < 500; wider
    separation and mass-resolution distributions)."""
    if variant == 2:
        return _make_mc_sample2(signal, background, seed)
    rng = np.random.default_rng(seed)
    n = signal + background
    ev = np.zeros(n, EVENT_DTYPE)
    s, b = slice(0, signal), slice(signal, n)
    ev["TrueMass"][s] = 135.0
    ev["TrueMass"][b] = rng.uniform(0.0, 1000.0, background)
    ev["TrueMassSigma"] = 0.3 * ev["TrueMass"]
    ev["Mass"] = _positive_normal(rng, ev["TrueMass"].copy(), ev["TrueMassSigma"].copy())
    ev["Type"][s] = 0
    ev["Type"][b] = 1
    ev["Separation"][s] = np.abs(rng.exponential(100.0, signal))
    ev["Separation"][b] = np.abs(rng.normal(0.0, 50.0, background))
    ev["MuDk"][s] = rng.uniform(size=signal) < 0.05
    ev["MuDk"][b] = rng.uniform(size=background) < 0.5
    return ev


def _make_mc_sample2(signal, background, seed):
    rng = np.random.default_rng(seed)
    signal = max(signal, 1000)                                   # example2/Simulated.H:19
    background = max(background, signal)                         # :21
    # background events are generated until `background` of them have Mass < 500 (:33-37)
    rows = []
    need = background
    while need > 0:
        k = int(need * 2.2) + 16
        tm = rng.uniform(0.0, 1000.0, k)
        sg = 0.4 * tm                                            # :54
        m = _positive_normal(rng, tm.copy(), sg.copy())
        ok = np.cumsum(m < 500.0)
        stop = int(np.searchsorted(ok, need)) + 1 if ok[-1] >= need else k
        rows.append((tm[:stop], sg[:stop], m[:stop]))
        need -= int(ok[stop - 1])
    tm = np.concatenate([r[0] for r in rows])
    sg = np.concatenate([r[1] for r in rows])
    mb = np.concatenate([r[2] for r in rows])
    nb = len(tm)
    ev = np.zeros(signal + nb, EVENT_DTYPE)
    s, b = slice(0, signal), slice(signal, signal + nb)
    ev["TrueMass"][s] = 135.0
    ev["TrueMassSigma"][s] = 0.3 * 135.0                         # :44
    ev["Mass"][s] = _positive_normal(rng, ev["TrueMass"][s].copy(), ev["TrueMassSigma"][s].copy())
    ev["TrueMass"][b] = tm
    ev["TrueMassSigma"][b] = sg
    ev["Mass"][b] = mb
    ev["Type"][s] = 0
    ev["Type"][b] = 1
    ev["Separation"][s] = np.abs(rng.exponential(150.0, signal))          # :47
    ev["Separation"][b] = np.abs(rng.normal(0.0, 70.0, nb))               # :59
    ev["MuDk"][s] = rng.uniform(size=signal) < 0.05
    ev["MuDk"][b] = rng.uniform(size=nb) < 0.5
    return ev


def find_bin(x, nbins=50, lo=0.0, hi=500.0):
    """TAxis::FindBin: 0 underflow, nbins+1 overflow."""
    x = np.asarray(x, dtype=np.float64)
    b = 1 + (nbins * (x - lo) / (hi - lo)).astype(np.int64)
    b = np.where(x < lo, 0, b)
    b = np.where(~(x < hi), nbins + 1, b)
    return b


def make_data_histograms(signal, background, seed, variant=1):
    """FakeData::FillData(signal, background) -> data150 (Close, Separated,
    DecayTag; 50 bins each on [0,500)).  variant=2: the truth distributions of
    example2/FakeData.H:37-43."""
    rng = np.random.default_rng(seed)

    def redraw(mean, sigma, floor):
        v = rng.normal(mean, sigma)
        while v < floor:
            v = rng.normal(mean, sigma)
        return v

    if variant == 2:
        scale = redraw(1.0, 0.01, 0.80)
        resolution = redraw(0.4, 0.05, 0.15)
        sig_sep = rng.normal(150.0, 10.0)
        bkg_sep = rng.normal(70.0, 10.0)
        fake_mudk, mudk_frac = redraw(0.05, 0.01, 0.03), redraw(0.50, 0.01, 0.3)
    else:
        scale = redraw(1.0, 0.15, 0.80)
        resolution = redraw(0.4, 0.05, 0.15)
        sig_sep = rng.normal(150.0, 20.0)
        bkg_sep = rng.normal(70.0, 10.0)
        fake_mudk, mudk_frac = 0.05, 0.50
    # signal
    width = np.log(1.0 + resolution)
    m = scale * 135.0 * np.exp(rng.normal(0.0, width, signal))
    bad = (m < 0.0) | (m > 500.0)
    while bad.any():
        m[bad] = scale * 135.0 * np.exp(rng.normal(0.0, width, int(bad.sum())))
        bad = (m < 0.0) | (m > 500.0)
    sep_s = np.abs(rng.exponential(abs(sig_sep), signal))
    tag_s = rng.uniform(size=signal) < fake_mudk
    # background
    mb = rng.uniform(0.0, 500.0, background)
    sep_b = np.abs(rng.normal(0.0, abs(bkg_sep), background))
    tag_b = rng.uniform(size=background) < mudk_frac
    if variant != 2:                                   # example/FakeData.H:87 (not in example2)
        tag_b = tag_b | (rng.uniform(size=background) < fake_mudk)
    mass = np.concatenate([m, mb])
    sep = np.concatenate([sep_s, sep_b])
    tag = np.concatenate([tag_s, tag_b])
    hist = np.where(tag, 2, np.where(sep < 100.0, 0, 1))
    bins = find_bin(mass)
    data = np.zeros((3, 52))
    np.add.at(data, (hist, bins), 1.0)
    return data[:, 1:51].reshape(150).copy()


def fake_inputs(data_signal, data_background, oversample, seed):
    """FakeLikelihood::Init(dataSignal, dataBackground, mcOversample)
    (example/FakeLikelihood.H:86-100): events + data histograms.  The exposure
    ratio (:107-138) needs a histogram fill and is computed by the caller
    through the engine (see `exposure_ratio`)."""
    data = make_data_histograms(data_signal, data_background, seed)
    events = make_mc_sample(int(oversample * data_signal),
                            int(2 * oversample * data_background), seed + 1)
    return events, data


def fake2_inputs(data_signal, data_background, oversample, seed):
    """example2's FakeLikelihood::Init(dataSignal, dataBackground, mcOversample)
    (example2/FakeLikelihood.H:123-137): events + data histograms.  There is no
    exposure ratio in example2: the histograms are renormalised to the event
    counts x[0], x[1] at every evaluation."""
    data = make_data_histograms(data_signal, data_background, seed, variant=2)
    events = make_mc_sample(int(oversample * data_signal),
                            int(2 * oversample * data_background), seed + 1, variant=2)
    return events, data


def exposure_ratio(engine, data150):
    """Corrections.ExposureRatio = data/mc at the nominal point with the ratio
    set to 1 (FakeLikelihood.H:107-138)."""
    engine.set_fake_data(data150, 1.0)
    sim = engine.fake_histograms(np.zeros((1, 9)))[0]

    def integral(h):            # TH1::Integral: bins 1..50 in order
        total = 0.0
        for v in h:
            total += float(v)
        return total

    data150 = np.asarray(data150, dtype=np.float64)
    data = integral(data150[0:50]) + integral(data150[50:100]) + integral(data150[100:150])
    mc = integral(sim[0:50]) + integral(sim[50:100]) + integral(sim[100:150])
    return data / mc
