"""Ensemble-pooled adaptation (BASELINE.json config 3: "adaptive covariance
pooled across chains").  The reference has no such mode, so there is nothing
to compare step for step; the checks are the analytic known answers of
SURVEY.md section 4: a Gaussian target with known covariance must be recovered by
the pooled estimate and sampled by the chains."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def correlated_target(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 1, (n, n))
    cov = a @ a.T / n + 0.3 * np.eye(n)
    return cov, np.linalg.inv(cov)


def test_pooled_covariance_recovers_the_target():
    import smcmc_b200
    from smcmc_b200 import binding
    n, E, K, steps = 8, 2048, 10, 800
    cov, err = correlated_target(n, 3)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
    eng.set_error_matrix(err)
    eng.prop_set(binding.PROP_POOLED_EVERY, K)
    assert eng.start(np.zeros(n)).all()
    eng.step(300)
    eng.reset_proposal()                      # forget the burn-in statistics
    tr = eng.step_trace(steps, want=("accepted", "points"))
    assert eng.get("pooled_count")[0] == E * steps
    pooled = np.zeros((n, n))
    pooled[np.tril_indices(n)] = eng.get("pooled_covariance")
    pooled = pooled + np.tril(pooled, -1).T
    assert np.allclose(pooled, cov, rtol=0.15, atol=0.05)
    u = eng.get("pooled_decomposition")
    assert np.allclose(np.tril(u, -1), 0) and np.allclose(u.T @ u, pooled, rtol=1e-10, atol=1e-12)
    assert np.all(np.abs(eng.get("pooled_mean")) < 0.1)
    acc = tr["accepted"][200:].mean()
    assert 0.1 < acc < 0.7
    # the acceptance-driven step size adaptation (:1771-1776) is as slow as in
    # the reference (exponent <= 1/500 per step) but moves the right way:
    # acceptance above the 0.234 target => sigma grows
    assert np.all(eng.get("sigma") > np.sqrt(1.0 / n))
    pts = tr["points"][300:].reshape(-1, n)
    assert np.allclose(np.cov(pts.T), cov, rtol=0.2, atol=0.06)
    assert np.all(eng.get("status") == 0)


def test_pooled_mode_leaves_per_chain_state_alone():
    """Per-chain covariance is neither read nor written in pooled mode."""
    import smcmc_b200
    from smcmc_b200 import binding
    n, E = 5, 64
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=2)
    eng.prop_set(binding.PROP_POOLED_EVERY, 5)
    eng.start(np.zeros(n))
    before = eng.get("covariance").copy()
    eng.step(100)
    assert np.array_equal(eng.get("covariance"), before)
    assert np.all(eng.get("total_steps") == 100)
    # every chain rescaled its sigma to the same pooled trace
    assert len(set(eng.get("sigma_trace"))) == 1


def test_pooled_tensor_path_matches_the_warp_path():
    """For dim >= 128 the pooled proposal x' = x + (sigma z) . U runs as one GEMM
    on the FP64 tensor cores (SMCMC_PROP_POOLED_TENSOR).  Same draws, same U:
    the proposed points agree with the per-warp evaluation to rounding, the
    accept sequences are the same."""
    import smcmc_b200
    from smcmc_b200 import binding
    n, E = 130, 384
    cov, err = correlated_target(n, 11)
    runs = {}
    for tensor in (0, 1):
        eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=21)
        eng.set_error_matrix(err)
        eng.prop_set(binding.PROP_POOLED_EVERY, 8)
        eng.prop_set(binding.PROP_POOLED_TENSOR, tensor)
        eng.set_uniform(5, -3.0, 3.0)                # one uniform dimension: skipped by the contraction
        assert eng.start(np.zeros(n)).all()
        runs[tensor] = eng.step_trace(80, want=("accepted", "points", "step_rms"))
        runs[tensor]["u"] = eng.get("pooled_decomposition")
    assert np.array_equal(runs[0]["accepted"], runs[1]["accepted"])
    # (the pooled statistics are sums of slightly different points, added in atomic order)
    assert np.allclose(runs[0]["u"], runs[1]["u"], rtol=1e-9, atol=1e-12)
    assert np.allclose(runs[0]["points"], runs[1]["points"], rtol=1e-10, atol=1e-12)
    assert np.allclose(runs[0]["step_rms"], runs[1]["step_rms"], rtol=1e-10)
    assert runs[0]["accepted"].mean() > 0.02


@pytest.mark.parametrize("n,E", [(8, 2048), (50, 5001), (130, 384)])
def test_tensor_core_accumulation_matches_the_scalar_kernel(monkeypatch, n, E):
    """kPoolAccumulateDmma (S = Y^T Y on the FP64 tensor cores, lower-triangular
    64 x 64 tile pairs, chains sliced over CTAs) against kPoolAccumulate
    (SMCMC_POOL_ACC_SCALAR=1) and against numpy on the traced points.  One exchange
    at the last step: the chains are the same in both runs."""
    import smcmc_b200
    from smcmc_b200 import binding
    steps = 24
    runs = {}
    for scalar in (0, 1):
        if scalar:
            monkeypatch.setenv("SMCMC_POOL_ACC_SCALAR", "1")
        else:
            monkeypatch.delenv("SMCMC_POOL_ACC_SCALAR", raising=False)
        eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=17)
        eng.prop_set(binding.PROP_POOLED_EVERY, steps)
        x0 = np.random.default_rng(3).normal(0, 1, (E, n))
        assert eng.start(x0).all()
        tr = eng.step_trace(steps, want=("accepted", "points"))
        runs[scalar] = {"points": tr["points"], "count": eng.get("pooled_count")[0], "mean": eng.get("pooled_mean"),
                        "cov": eng.get("pooled_covariance")}
    assert np.array_equal(runs[0]["points"], runs[1]["points"])
    assert runs[0]["count"] == runs[1]["count"] == E * steps
    pts = runs[0]["points"].reshape(-1, n)
    assert np.allclose(runs[0]["mean"], pts.mean(axis=0), rtol=1e-11, atol=1e-13)
    assert np.allclose(runs[0]["mean"], runs[1]["mean"], rtol=1e-11, atol=1e-13)
    scale = np.abs(runs[1]["cov"]).max()
    assert np.allclose(runs[0]["cov"], runs[1]["cov"], rtol=1e-9, atol=1e-11 * scale)


@pytest.mark.parametrize("n,E", [(8, 512), (50, 2048), (130, 384)])
def test_cta_wide_factorisation_equals_the_warp_one(monkeypatch, n, E):
    """kPoolFactorCta (one thread per column) against kPoolFactor (one warp, SMCMC_POOL_FACTOR_WARP=1):
    the same operations in the same order per entry: the shared factor and every chain agree to the
    rounding of the (atomically summed) statistics."""
    import smcmc_b200
    from smcmc_b200 import binding
    runs = {}
    for warp in (0, 1):
        if warp:
            monkeypatch.setenv("SMCMC_POOL_FACTOR_WARP", "1")
        else:
            monkeypatch.delenv("SMCMC_POOL_FACTOR_WARP", raising=False)
        eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=13)
        eng.prop_set(binding.PROP_POOLED_EVERY, 6)
        x0 = np.random.default_rng(2).normal(0, 1, (E, n))
        assert eng.start(x0).all()
        tr = eng.step_trace(30, want=("accepted", "points"))
        runs[warp] = {"points": tr["points"], "u": eng.get("pooled_decomposition"), "cov": eng.get("pooled_covariance")}
    # (the statistics themselves are sums of FP64 atomics: equal to rounding from run to run)
    assert np.allclose(runs[0]["u"], runs[1]["u"], rtol=1e-9, atol=1e-12)
    assert np.allclose(runs[0]["points"], runs[1]["points"], rtol=1e-9, atol=1e-11)
    u = runs[0]["u"]
    assert np.abs(np.triu(u, 1)).max() > 0 and np.allclose(np.tril(u, -1), 0)
