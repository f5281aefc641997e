"""Parity at BASELINE.json's own shapes (the sizes the configs are quoted on, not
scaled-down stand-ins) and the analytic known-answer tests of SURVEY.md section 4:

  C4  TSimpleHMC / TDummyLogLikelihood at n = 500: EXACT mode bit for bit against the
      port, TENSOR (DMMA) mode within 1e-12 of it with identical trajectories;
  C3  pooled adaptation on THorrific and TASym at n = 50; the pooled proposal of every
      kernel path (shared-memory tile, warp, DMMA GEMM at n >= 128 and at n = 500)
      against an FP64 scalar evaluation of x + sum_i (sigma z_i) U(i, .) in the
      reference's order (TSimpleMCMC.H:709-724) -- not against another kernel;
  KAT the Horrific ridge Var(sum x_i) = sigma^2 n / 3 (THorrificLogLikelihood.H:27-36)
      and the as-shipped VERY_CORRELATED covariance (TDummyLogLikelihood.H:78-88),
      65 536 chains each.
"""
import ctypes
import os

import numpy as np
import pytest

from helpers import hmc_error_matrix

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_normals(seed, chain, step, n):
    """The draws of (seed, chain, step), slots 0..n-1, from the HOST build of
    include/smcmc_rng.h (libsmcmc_hostkat.so)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "root-simple-mcmc_b200", "smcmc_b200", "libsmcmc_hostkat.so"))
    lib.smcmc_kat_normals.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_void_p]
    out = np.zeros(n)
    lib.smcmc_kat_normals(seed, chain, step, 1, n, out.ctypes.data)
    return out


def scalar_proposal(x, sigma, u, z):
    """TSimpleMCMC.H:709-724 for Gaussian dimensions: for i ascending, for every j,
    proposal[j] += fSigma * r_i * U(i, j) -- every product and sum rounded separately."""
    p = x.copy()
    for i in range(len(x)):
        p = p + (sigma * z[i]) * u[i]
    return p


# ---------------------------------------------------------------------------------- C4
def test_hmc_500_dimensions_exact_mode_equals_the_port(checkers):
    """n = 500, the analytic gradient, 48 chains: every position, potential, step size and
    trajectory length of 12 steps, and the final state, bit for bit against the port for
    six of the chains (the port is pinned to the reference build by the golden chains)."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    cc = checkers
    n, E, steps, seed = 500, 48, 12, 5
    prec = hmc_error_matrix("spd500")
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=seed)
    eng.set_error_matrix(prec)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_start(np.ones(n))                                   # SimpleHMC.C:45
    tr = eng.hmc_step_trace(steps, 0)
    sc = eng.hmc_get("scalars")
    acc, mom = eng.hmc_get("accepted"), eng.hmc_get("momentum")
    for c in (0, 1, 17, 31, 32, 47):
        o = cc.CpuHmc("orc", cc.LLH_DUMMY, n, True, seed, c)
        o.set_error_matrix(prec)
        o.start(np.ones(n))
        w = o.step(steps, 0)
        assert np.array_equal(tr["leapfrog"][:, c], w["leapfrog"]), c
        assert np.array_equal(tr["mean_epsilon"][:, c], w["epsilon"]), c
        assert np.array_equal(tr["potential"][:, c], w["potential"]), c
        assert np.array_equal(tr["points"][:, c], w["x"]), c
        st = o.state()
        assert np.array_equal(acc[c], st["accepted"]) and np.array_equal(mom[c], st["momentum"])
        for k in ("acceptance", "mean_epsilon", "leapfrog", "accepted_potential", "gradient_count", "potential_count",
                  "cov_trials", "est_cov_trace"):
            assert sc[c][b.HMC_SCALARS.index(k)] == st[k], (c, k)
    assert tr["accepted"].sum() > 0 and np.abs(tr["leapfrog"]).max() >= 2


def test_hmc_500_dimensions_tensor_mode(checkers):
    """TENSOR mode (kDummyContractDmma) at n = 500 with 300 chains: likelihood within 1e-12
    relative of the reference-ordered value (the specification's tolerance), and an HMC
    ensemble whose accept sequences and trajectory lengths are those of the EXACT mode over
    25 steps (potentials to 1e-9)."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    cc = checkers
    n, E = 500, 300
    prec = hmc_error_matrix("spd500")
    pts = np.random.default_rng(500).normal(size=(E, n))
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=2)
    eng.set_error_matrix(prec)
    exact = eng.eval(pts)
    eng.set_dummy_mode(b.DUMMY_TENSOR)
    tensor = eng.eval(pts)
    o = cc.CpuChain("orc", cc.LLH_DUMMY, n, 2, 0)
    o.set_error_matrix(prec)
    ref = np.array([o.llh(p) for p in pts[:12]])
    assert np.array_equal(exact[:12], ref)
    assert np.max(np.abs(tensor / exact - 1.0)) < 1e-12
    runs = {}
    for mode in (b.DUMMY_EXACT, b.DUMMY_TENSOR):
        h = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=6)
        h.set_error_matrix(prec)
        h.set_dummy_mode(mode)
        h.hmc_set(b.HMC_USER_GRADIENT, 1)
        h.hmc_start(np.ones(n))
        runs[mode] = h.hmc_step_trace(25, 0, want=("potential", "leapfrog", "accepted"))
    assert np.array_equal(runs[0]["leapfrog"], runs[1]["leapfrog"])
    assert np.array_equal(runs[0]["accepted"], runs[1]["accepted"])
    assert np.allclose(runs[0]["potential"], runs[1]["potential"], rtol=1e-9)


# ---------------------------------------------------------------------------------- C3, pooled
@pytest.mark.parametrize("kind_name,n,E,tensor", [("horrific", 50, 1024, 0), ("asym", 50, 1024, 0),
                                                  ("dummy", 130, 384, 1), ("dummy", 130, 384, 0), ("dummy", 500, 256, 1)])
def test_pooled_proposal_against_fp64_scalar_evaluation(kind_name, n, E, tensor):
    """After the shared factor has adapted, one step's proposed points of every chain equal
    x + sum_i (sigma z_i) U(i, .) evaluated in FP64 on the host in the reference's order, with
    the draws of the host build of the random stream: bit for bit for the kernels that keep the
    order (tile / warp), to 1e-13 of the scale for the DMMA GEMM."""
    import smcmc_b200
    from smcmc_b200 import binding
    seed, offset = 23, 7
    kind = {"horrific": smcmc_b200.LLH_HORRIFIC, "asym": smcmc_b200.LLH_ASYM, "dummy": smcmc_b200.LLH_DUMMY}[kind_name]
    eng = smcmc_b200.Engine(kind, n, E, seed=seed, chain_offset=offset)
    if kind_name == "dummy":
        eng.set_error_matrix(hmc_error_matrix("spd%d" % n))
    eng.prop_set(binding.PROP_POOLED_EVERY, 8)
    eng.prop_set(binding.PROP_POOLED_TENSOR, tensor)
    if kind_name == "asym":
        # slope 100 below zero: with the default step sqrt(1/n) nothing is accepted for the first
        # ~1000 steps (the reference adapts sigma by at most a factor (a/0.234)^(1/500) per step)
        eng.prop_set(binding.PROP_SIGMA, 0.002)
    assert eng.start(np.full(n, 0.01)).all()
    eng.step(48 if n < 500 else 16)                      # six (two) exchanges: U is no longer the start-up diagonal
    eng.prop_set(binding.PROP_POOLED_EVERY, 1 << 30)     # no exchange during the step under test
    u = eng.get("pooled_decomposition")
    assert np.abs(np.triu(u, 1)).max() > 0               # correlations have been learned
    x = eng.get("accepted")
    step = int(eng.save_state()["step_index"][0])
    eng.step_trace(1, want=("accepted",))
    xp = eng.get("proposed")
    sigma = eng.get("sigma")                             # after UpdateState: the value the proposal used
    worst = 0.0
    for c in list(range(0, E, 37)) + [E - 1]:
        want = scalar_proposal(x[c], sigma[c], u, host_normals(seed, offset + c, step, n))
        if tensor:
            scale = np.abs(x[c]).max() + np.abs(want - x[c]).max()
            worst = max(worst, np.abs(xp[c] - want).max() / scale)
        else:
            assert np.array_equal(xp[c], want), c
    if tensor:
        assert worst < 1e-13, worst


@pytest.mark.parametrize("kind_name", ["horrific", "asym"])
def test_pooled_adaptation_on_the_config3_targets(kind_name):
    """BASELINE.json configs[2] at its own dimension: 50-dim THorrific / TASym, 8192 chains,
    adaptation pooled across the chains.  Every chain stays healthy, the pooled covariance is
    what numpy computes from the traced points, the factor reproduces it, and the targets'
    known structure shows: the Horrific ridge (sum of the coordinates confined to ~0.01
    sqrt(n/3)), the ASym support (x >= 0 up to the slope-100 tail)."""
    import smcmc_b200
    from smcmc_b200 import binding
    n, E, K = 50, 8192, 10
    kind = smcmc_b200.LLH_HORRIFIC if kind_name == "horrific" else smcmc_b200.LLH_ASYM
    eng = smcmc_b200.Engine(kind, n, E, seed=4)
    eng.prop_set(binding.PROP_POOLED_EVERY, K)
    if kind_name == "asym":
        eng.prop_set(binding.PROP_SIGMA, 0.002)          # see test_pooled_proposal_against_fp64_scalar_evaluation
    assert eng.start(np.zeros(n) if kind_name == "horrific" else np.full(n, 0.01)).all()
    eng.step(300)
    eng.reset_proposal()                                 # forget the transient
    tr = eng.step_trace(40, want=("accepted", "points"))
    assert np.all(eng.get("status") == 0)
    assert eng.get("pooled_count")[0] == E * 40
    pts = tr["points"].reshape(-1, n)
    pooled = np.zeros((n, n))
    pooled[np.tril_indices(n)] = eng.get("pooled_covariance")
    pooled = pooled + np.tril(pooled, -1).T
    want = np.cov(pts.T, bias=True)
    assert np.allclose(pooled, want, rtol=1e-9, atol=1e-12 * np.abs(want).max())
    u = eng.get("pooled_decomposition")
    assert np.allclose(u.T @ u, pooled, rtol=1e-9, atol=1e-14)
    acc = tr["accepted"].mean()
    assert 0.02 < acc < 0.9, acc
    if kind_name == "horrific":
        s = pts.sum(axis=1)
        assert np.all(np.abs(pts) <= 1.0)
        assert s.std() < 3.0 * 0.01 * np.sqrt(n / 3.0)   # on the ridge (equilibrium value: 1.0 x)
        assert pts.std(axis=0).min() > 5.0 * s.std() / np.sqrt(n)     # and spreading along it
    else:
        assert pts.min() > -0.2 and pts.mean() > 0.0


# ---------------------------------------------------------------------------------- known answers, 65 536 chains
def test_horrific_ridge_variance_65536_chains():
    """THorrificLogLikelihood.H:27-36: inside the box the likelihood only constrains
    s = sum x_i, with s / sqrt(n/3) ~ N(0, 0.01^2): Var(s) = 0.01^2 n / 3.
    65 536 chains x 50 dimensions, pooled adaptation; the cross-chain variance of s at the end
    (one sample per chain: independent chains)."""
    import smcmc_b200
    from smcmc_b200 import binding
    n, E = 50, 65536
    eng = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, n, E, seed=4)
    eng.prop_set(binding.PROP_POOLED_EVERY, 16)
    assert eng.start(np.zeros(n)).all()
    eng.step(400)
    eng.reset_proposal()
    eng.step(800)
    want = 0.01 ** 2 * n / 3.0
    got = []
    for _ in range(4):
        eng.step(100)
        s = eng.get("accepted").sum(axis=1)
        got.append(s.var())
        assert abs(s.mean()) < 5.0 * np.sqrt(want / E) + 1e-4
    got = np.array(got)
    assert np.all(np.abs(got / want - 1.0) < 0.05), got / want
    assert np.all(eng.get("status") == 0)


def test_very_correlated_covariance_65536_chains(checkers):
    """TDummyLogLikelihood as shipped (dim 100, VERY_CORRELATED: identity plus 0.999999 between
    coordinates 0 and 99, TDummyLogLikelihood.H:78-88): the ensemble reproduces the covariance
    the likelihood was built from -- unit variances, the 0.999999 correlation (i.e. variance
    1e-6 of (x_0 - x_99)/sqrt 2), nothing else correlated."""
    import smcmc_b200
    from smcmc_b200 import binding
    from helpers import golden
    n, E = 100, 65536
    g = golden("chains.npz")
    err, cov = g["dummy100_error"], g["dummy100_covariance"]
    assert cov[0, 99] == 0.999999 and cov[0, 0] == 1.0
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=9)
    eng.set_error_matrix(err)
    eng.prop_set(binding.PROP_POOLED_EVERY, 16)
    assert eng.start(np.zeros(n)).all()                  # SimpleMCMC.C:149
    for _ in range(3):
        eng.step(500)
        eng.reset_proposal()
    eng.step(1000)
    x = eng.get("accepted")
    got = np.cov(x.T)
    se = np.sqrt(2.0 / E)
    assert np.all(np.abs(np.diag(got) - 1.0) < 8 * se + 0.02), np.abs(np.diag(got) - 1.0).max()
    narrow = (x[:, 0] - x[:, 99]) / np.sqrt(2.0)
    assert abs(narrow.var() / 1e-6 - 1.0) < 0.1, narrow.var()
    off = got - np.diag(np.diag(got))
    off[0, 99] = off[99, 0] = 0.0
    assert np.abs(off).max() < 10 / np.sqrt(E) + 0.02, np.abs(off).max()
    assert np.all(np.abs(x.mean(0)) < 8 / np.sqrt(E) + 0.02)
