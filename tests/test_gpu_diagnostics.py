"""On-device ensemble diagnostics (smcmc_diag_*, csrc/diagnostics.cuh) against
the reference's offline formulas evaluated with numpy on the traced points:

  MakeCovariance.C:63-89        mean_i = sum x_i / N ; cov_ij = sum x_i x_j / N - mean_i mean_j
  MakeAutocorrelation.C:96-148  a_i(lag) = (<x_i(t) x_i(t-lag)> - mean_i^2) / var_i

with all chains of the ensemble pooled, plus Gelman-Rubin's R-hat across the
chains.  The device sums run in a different order than numpy's: 1e-9 relative.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_diagnostics(pts, lags):
    """pts: [steps, chains, n] accepted points after every step."""
    T, E, n = pts.shape
    flat = pts.reshape(T * E, n)
    mean = flat.sum(0) / (T * E)
    cov = flat.T @ flat / (T * E) - np.outer(mean, mean)
    var = np.diag(cov)
    rho = np.zeros((len(lags), n))
    for k, lag in enumerate(lags):
        prod = (pts[lag:] * pts[:T - lag]).sum((0, 1)) / ((T - lag) * E)
        rho[k] = (prod - mean * mean) / var
    cm = pts.mean(0)                              # chain means [E, n]
    W = pts.var(0, ddof=1).mean(0)
    BoverN = cm.var(0, ddof=1)
    rhat = np.sqrt(((T - 1) / T * W + BoverN) / W)
    tau = np.zeros(n)
    for i in range(n):
        S, pl, pr = 0.0, 0.0, 1.0
        for k, lag in enumerate(lags):
            r = rho[k, i]
            w = lag - pl
            m = (r - pr) / w
            if r > 0:
                S += w * pr + m * w * (w + 1) / 2
                pl, pr = float(lag), r
            else:
                kk = np.floor(pr / -m)
                S += kk * pr + m * kk * (kk + 1) / 2
                break
        tau[i] = 1 + 2 * S
    return mean, cov, rho, rhat, tau


@pytest.mark.parametrize("chains,dim,steps,max_lag", [(64, 5, 400, 48), (300, 9, 150, 16), (7, 3, 90, 0)])
def test_diagnostics_match_the_offline_formulas(chains, dim, steps, max_lag):
    import smcmc_b200
    assert torch.cuda.is_available()
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, dim, chains, seed=21)
    rng = np.random.default_rng(2)
    eng.start(rng.normal(0, 1, (chains, dim)))
    eng.step(50)                                  # not accumulated
    eng.diag_enable(max_lag)
    tr = eng.step_trace(steps, want=("points",))
    d = eng.diag_get()
    assert d["samples"] == chains * steps and d["steps"] == steps
    lags = list(d["lags"])
    if max_lag:
        assert lags[:6] == [1, 2, 3, 4, 6, 8] and max(lags) <= max_lag
    else:
        assert lags == []
    mean, cov, rho, rhat, tau = reference_diagnostics(tr["points"], lags)
    assert np.allclose(d["mean"], mean, rtol=1e-9, atol=1e-12)
    assert np.allclose(d["covariance"], cov, rtol=1e-9, atol=1e-12)
    assert np.allclose(d["rhat"], rhat, rtol=1e-9)
    if max_lag:
        assert np.allclose(d["autocorrelation"], rho, rtol=1e-8, atol=1e-10)
        assert np.allclose(d["tau"], tau, rtol=1e-7)
        assert np.allclose(d["ess"], chains * steps / tau, rtol=1e-7)
        assert np.all(d["autocorrelation"][0] > 0.3)          # a Metropolis chain is positively correlated
    # a unit Gaussian target after burn-in: the pooled moments are those of the target
    assert np.all(np.abs(d["mean"]) < 0.5) and np.all(np.abs(np.diag(d["covariance"]) - 1.0) < 0.6)


def test_reset_and_reaccumulate():
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 4, 32, seed=3)
    eng.start(np.zeros((32, 4)))
    eng.diag_enable(8)
    eng.step(40)
    first = eng.diag_get()
    eng.diag_reset()
    tr = eng.step_trace(60, want=("points",))
    d = eng.diag_get()
    assert first["samples"] == 32 * 40 and d["samples"] == 32 * 60
    mean, cov, rho, rhat, tau = reference_diagnostics(tr["points"], list(d["lags"]))
    assert np.allclose(d["mean"], mean, rtol=1e-9, atol=1e-12)
    assert np.allclose(d["autocorrelation"], rho, rtol=1e-8, atol=1e-10)


def test_diagnostics_need_enabling():
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 4, 8, seed=3)
    eng.start(np.zeros((8, 4)))
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.diag_get()
