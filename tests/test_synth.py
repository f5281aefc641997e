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
