"""The product's synthetic-input builder follows the distributions of the
reference's toy generators (example/Simulated.H:31-53, FakeData.H:32-117)."""
import numpy as np
import pytest


def test_mc_sample_shape_and_classes():
    from smcmc_b200 import synth
    ev = synth.make_mc_sample(20000, 40000, seed=3)
    assert len(ev) == 60000 and ev.dtype.itemsize == 48
    s, b = ev[:20000], ev[20000:]
    assert np.all(s["Type"] == 0) and np.all(b["Type"] == 1)
    assert np.all(s["TrueMass"] == 135.0) and np.all(s["TrueMassSigma"] == 0.3 * 135.0)
    assert np.all(ev["Mass"] >= 0) and np.all(ev["Separation"] >= 0)
    assert abs(s["MuDk"].mean() - 0.05) < 0.006 and abs(b["MuDk"].mean() - 0.5) < 0.01
    assert abs(s["Separation"].mean() - 100.0) < 3.0
    assert abs(b["Separation"].mean() - 50.0 * np.sqrt(2 / np.pi)) < 1.5
    assert abs(b["TrueMass"].mean() - 500.0) < 8.0


def test_data_histograms():
    from smcmc_b200 import synth
    d = synth.make_data_histograms(5000, 5000, seed=4)
    assert d.shape == (150,) and d.sum() == 10000 and np.all(d >= 0)
    assert np.all(d == np.round(d))


def test_against_reference_generators(checkers, have_ref):
    if not have_ref:
        pytest.skip("reference-backed checker not built here")
    from smcmc_b200 import synth
    ev_ref, data_ref, expo_ref = checkers.ref_generate(11, 3000, 3000, 10.0)
    ev = synth.make_mc_sample(30000, 60000, seed=5)
    assert len(ev_ref) == len(ev) == 90000
    for sl in (slice(0, 30000), slice(30000, None)):
        for f in ("Mass", "Separation", "TrueMass", "MuDk"):
            a, b = ev_ref[f][sl].astype(float), ev[f][sl].astype(float)
            scale = max(abs(a.mean()), a.std(), 1e-9)
            assert abs(a.mean() - b.mean()) < 0.03 * scale, f
            assert abs(a.std() - b.std()) < 0.05 * scale + 1e-12, f
    assert data_ref.sum() == 6000
    assert 0.03 < expo_ref < 0.3


def test_example2_generators_against_reference(checkers, have_ref):
    """example2/Simulated.H:17-63 and example2/FakeData.H:32-120."""
    if not have_ref:
        pytest.skip("reference-backed checker not built here")
    from smcmc_b200 import synth
    ev_ref, data_ref = checkers.ref2_generate(13, 2000, 2000, 10.0)
    ev, data = synth.fake2_inputs(2000, 2000, 10.0, seed=8)
    ns_ref, ns = int((ev_ref["Type"] == 0).sum()), int((ev["Type"] == 0).sum())
    assert ns_ref == ns == 20000
    # background events are drawn until 40000 of them lie below 500
    for sample in (ev_ref, ev):
        b = sample[sample["Type"] == 1]
        assert int((b["Mass"] < 500.0).sum()) == 40000
        assert np.all(sample["Type"][:20000] == 0) and np.all(sample["Type"][20000:] == 1)
    assert abs(len(ev_ref) - len(ev)) < 0.03 * len(ev)
    for typ in (0, 1):
        a, b = ev_ref[ev_ref["Type"] == typ], ev[ev["Type"] == typ]
        for f in ("Mass", "Separation", "TrueMass", "TrueMassSigma"):
            x, y = a[f].astype(float), b[f].astype(float)
            scale = max(abs(x.mean()), x.std(), 1e-9)
            assert abs(x.mean() - y.mean()) < 0.03 * scale, (typ, f)
            assert abs(x.std() - y.std()) < 0.05 * scale + 1e-12, (typ, f)
        want = 0.05 if typ == 0 else 0.5                  # Bernoulli tag probabilities (:48, :60)
        assert abs(a["MuDk"].mean() - want) < 0.012 and abs(b["MuDk"].mean() - want) < 0.012
    assert data_ref.sum() == data.sum() == 4000
    for h in range(3):       # the split over Close / Separated / DecayTag
        assert abs(data_ref[h * 50:(h + 1) * 50].sum() - data[h * 50:(h + 1) * 50].sum()) < 200
