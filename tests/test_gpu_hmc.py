"""Step-for-step parity of the device TSimpleHMC ensemble with the reference:
TSimpleHMC::Step (TSimpleHMC.H:279-401), LeapFrog (:582-651), the gradient
variants (:417-532), UpdateCovariance / UpdateErrorMatrix (:665-858), driven by
identical draws (include/smcmc_rng.h).

Required: the accepted position and potential after every step, the adapted
mean step size and the trajectory length are those of the reference -- bit for
bit except for the one libm call in the loop (log of the accept draw, :347;
CUDA's log and glibc's can differ in the last ulp, which only matters at a
knife edge) -- and the final adaptive state (running mean, covariance estimate,
error matrix, counters) is identical.
"""
import numpy as np
import pytest
import torch

from helpers import HMC_GOLDEN, HMC_PORT_ONLY, golden, golden_chain, hmc_error_matrix, hmc_scalar_mask

pytestmark = pytest.mark.gpu


def device_hmc(cfg, error, chains=4):
    import smcmc_b200
    from smcmc_b200 import binding as b
    assert torch.cuda.is_available()
    lo = max(0, cfg["chain"] - 1)
    eng = smcmc_b200.Engine(cfg["kind"], cfg["dim"], chains, seed=cfg["seed"], chain_offset=lo)
    if error is not None:
        eng.set_error_matrix(error)
    # THorrificLogLikelihood's gradient functor exists but declines (:41-43):
    # the same as having none
    eng.hmc_set(b.HMC_USER_GRADIENT, 1 if (cfg["grad"] and cfg["kind"] in (1, 6)) else 0)
    if cfg["gtype"] == 2:
        eng.hmc_set(b.HMC_KEEP_ERROR_MATRIX, 1)
    fields = {"alpha": b.HMC_ALPHA, "mean_epsilon": b.HMC_MEAN_EPSILON, "leapfrog": b.HMC_LEAPFROG}
    for f, v in cfg.get("pre", ()):
        eng.hmc_set(fields[f], v)
    eng.hmc_start(np.full(cfg["dim"], cfg["x0"]))
    for f, v in cfg.get("post", ()):
        eng.hmc_set(fields[f], v)
    tr = eng.hmc_step_trace(cfg["nsteps"], cfg["gtype"])
    return eng, tr, cfg["chain"] - lo


def check_against(eng, tr, c, want_tr, want_scalars, want_arrays, keep_error):
    from smcmc_b200 import binding as b
    assert np.array_equal(tr["leapfrog"][:, c], want_tr["leapfrog"])
    assert np.array_equal(tr["mean_epsilon"][:, c], want_tr["epsilon"])
    assert np.array_equal(tr["potential"][:, c], want_tr["potential"])
    assert np.array_equal(tr["points"][:, c], want_tr["x"])
    sc = eng.hmc_get("scalars")[c]
    mask = hmc_scalar_mask(want_scalars, eng.dim)
    assert np.array_equal(sc[mask], want_scalars[mask]), list(zip(b.HMC_SCALARS, sc, want_scalars))
    for dev, ref in (("accepted", "accepted"), ("momentum", "momentum"), ("central", "central"),
                     ("average", "average"), ("covariance", "covariance")):
        assert np.array_equal(eng.hmc_get(dev)[c], want_arrays[ref], equal_nan=True), dev
    if keep_error:
        assert np.array_equal(eng.hmc_get("error_matrix")[c], want_arrays["error"], equal_nan=True)


@pytest.mark.parametrize("defer", ["0", "5", "16"])
@pytest.mark.parametrize("name", sorted(HMC_GOLDEN))
def test_hmc_golden_chain(monkeypatch, name, defer):
    """Chain `id` of a 4-chain ensemble equals the golden run of the reference
    build for that chain id -- with fEXXT rewritten every step (defer 0) and with
    the recorded UpdateCovariance calls applied 5 or 16 at a time (kHmcExxtFlush;
    large ensembles use 16 by default)."""
    monkeypatch.setenv("SMCMC_HMC_DEFER", defer)
    g = golden("hmc.npz")
    want = golden_chain(g, name)
    cfg = HMC_GOLDEN[name]
    err = g["error_" + cfg["error"]] if "error" in cfg else None
    eng, tr, c = device_hmc(cfg, err)
    arrays = {k: want["final_" + k] for k in ("accepted", "momentum", "central", "average", "covariance", "error")}
    check_against(eng, tr, c, want, want["final_scalars"], arrays, cfg["gtype"] == 2)


@pytest.mark.parametrize("name", sorted(HMC_PORT_ONLY))
def test_hmc_against_the_port(name):
    """Dimensions the reference build cannot run: every chain of an 8-chain
    ensemble equals the oracle port's chain with the same id."""
    from oracle import cpu_checkers as cc
    from test_oracle import run_cpu_hmc
    cfg = dict(HMC_PORT_ONLY[name])
    err = hmc_error_matrix(cfg["error"]) if "error" in cfg else None
    eng, tr, _ = device_hmc(cfg, err, chains=8)
    lo = max(0, cfg["chain"] - 1)
    for c in range(8):
        one = dict(cfg, chain=lo + c)
        otr, st = run_cpu_hmc(cc, "orc", one, err)
        scal = np.array([st[k] for k in cc.HMC_STATE_FIELDS])
        check_against(eng, tr, c, otr, scal, st, cfg["gtype"] == 2)


def test_hmc_ensemble_samples_the_target():
    """2048 chains on a correlated 24-dimensional Gaussian: ensemble mean and
    covariance after burn-in agree with the target (posterior mean 0,
    covariance = inverse of the precision matrix)."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    n, E = 24, 2048
    prec = hmc_error_matrix("spd24")
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
    eng.set_error_matrix(prec)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_start(np.zeros(n))
    eng.hmc_step(150)
    pts = []
    for _ in range(10):
        eng.hmc_step(10)
        pts.append(eng.hmc_get("accepted"))
    x = np.concatenate(pts)
    cov = np.linalg.inv(prec)
    assert np.all(np.abs(x.mean(0)) < 0.05)
    got = np.cov(x.T)
    assert np.max(np.abs(got - cov)) < 0.08 * np.max(np.diag(cov))
    sc = eng.hmc_scalars()
    assert np.all(sc["step_count"] == 250)
    assert 0.3 < np.mean(sc["acceptance"]) < 1.0


def test_hmc_errors():
    import smcmc_b200
    from smcmc_b200 import binding as b
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, 4, 2, seed=1)
    with pytest.raises(smcmc_b200.SmcmcError) as ei:
        eng.hmc_step(1)                       # "Must initialize starting point", TSimpleHMC.H:280-284
    assert ei.value.status == -1
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.hmc_set(b.HMC_USER_GRADIENT, 1)   # the documentation likelihood has no gradient functor
    eng.hmc_start(np.zeros(4))
    with pytest.raises(smcmc_b200.SmcmcError) as ei:
        eng.hmc_step(1, 4)                    # type 4 without a user gradient: the reference's bare throw (:521)
    assert ei.value.status == -2
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.hmc_step(1, 2)                    # covariant gradient needs the error matrix
    eng.hmc_step(3)
    assert np.all(eng.hmc_scalars()["step_count"] == 3)


def test_tensor_mode_matches_exact_mode():
    """SMCMC_DUMMY_TENSOR (X . Error^T on the FP64 tensor cores) against the
    reference-ordered evaluation: likelihood values to 1e-12 relative (the
    tolerance of the specification), and an HMC ensemble whose trajectories stay
    on top of the exact ones (same trajectory lengths, potentials to 1e-9)."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    from oracle import cpu_checkers as cc
    for n, E in ((37, 70), (200, 130)):
        prec = hmc_error_matrix("spd%d" % n)
        pts = np.random.default_rng(n).normal(size=(E, n))
        eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=2)
        eng.set_error_matrix(prec)
        exact = eng.eval(pts)
        eng.set_dummy_mode(b.DUMMY_TENSOR)
        tensor = eng.eval(pts)
        o = cc.CpuChain("orc", cc.LLH_DUMMY, n, 2, 0)
        o.set_error_matrix(prec)
        ref = np.array([o.llh(p) for p in pts[:16]])
        assert np.array_equal(exact[:16], ref)                       # exact mode: bit for bit
        assert np.max(np.abs(tensor / exact - 1.0)) < 1e-12          # tensor mode: the specified tolerance
        runs = {}
        for mode in (b.DUMMY_EXACT, b.DUMMY_TENSOR):
            h = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=6)
            h.set_error_matrix(prec)
            h.set_dummy_mode(mode)
            h.hmc_set(b.HMC_USER_GRADIENT, 1)
            h.hmc_start(np.full(n, 0.5))
            runs[mode] = h.hmc_step_trace(60, 0, want=("potential", "leapfrog", "accepted"))
        assert np.array_equal(runs[0]["leapfrog"], runs[1]["leapfrog"])
        assert np.array_equal(runs[0]["accepted"], runs[1]["accepted"])
        assert np.allclose(runs[0]["potential"], runs[1]["potential"], rtol=1e-9)


def test_pooled_covariance_mode():
    """SMCMC_HMC_POOLED_COVARIANCE: one running mean / covariance for the ensemble instead of
    one per chain (not in the reference).  2048 chains on a correlated 24-dimensional Gaussian:
    the pooled estimate converges to the target covariance, UpdateErrorMatrix runs on it (largest
    and smallest eigenvalue from power / inverse iteration through a Cholesky factor: compared
    with numpy's eigenvalues of the same matrix), every chain takes its step size and trajectory
    length from it, and the ensemble still samples the target."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    n, E = 24, 2048
    prec = hmc_error_matrix("spd24")
    cov = np.linalg.inv(prec)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
    eng.set_error_matrix(prec)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_set(b.HMC_POOLED_COVARIANCE, 1)
    eng.hmc_start(np.zeros(n))
    eng.hmc_step(31)
    ps = dict(zip(b.HMC_POOLED_SCALARS, eng.hmc_get("pooled_scalars")))
    assert ps["updates"] == 0 and ps["step_count"] == 31 and ps["trials"] == 31 * E
    eng.hmc_step(1)                                     # step 32: the first UpdateErrorMatrix on the pooled estimate
    ps = dict(zip(b.HMC_POOLED_SCALARS, eng.hmc_get("pooled_scalars")))
    got = eng.hmc_get("pooled_covariance")
    assert ps["updates"] == 1 and ps["repaired"] == 0
    lam = np.linalg.eigvalsh(got)
    assert abs(ps["max_scale"] / max(0.1, np.sqrt(lam[-1])) - 1.0) < 1e-6
    assert abs(ps["min_scale"] / max(0.01, np.sqrt(lam[0])) - 1.0) < 1e-5
    assert abs(ps["est_cov_trace"] / np.trace(got) - 1.0) < 1e-12
    assert abs(ps["orbit_length"] - 2.0 * 3.14 * ps["max_scale"]) < 1e-12
    sc = eng.hmc_scalars()
    target = 0.4 * ps["orbit_length"]
    assert np.all(sc["leapfrog"] >= 2) and np.all(sc["leapfrog"] % 2 == 0)       # :841-842
    assert np.allclose(sc["mean_epsilon"] * sc["leapfrog"], target, rtol=1e-12)  # :847
    with pytest.raises(smcmc_b200.SmcmcError):
        eng.hmc_get("covariance")
    eng.hmc_step(150)
    pts = []
    for _ in range(10):
        eng.hmc_step(10)
        pts.append(eng.hmc_get("accepted"))
    x = np.concatenate(pts)
    assert np.all(np.abs(x.mean(0)) < 0.05)
    assert np.max(np.abs(np.cov(x.T) - cov)) < 0.08 * np.max(np.diag(cov))
    pooled = eng.hmc_get("pooled_covariance")
    assert np.max(np.abs(pooled - cov)) < 0.1 * np.max(np.diag(cov))
    assert np.all(np.abs(eng.hmc_get("pooled_average")) < 0.05)
    sc = eng.hmc_scalars()
    assert np.all(sc["step_count"] == 282) and 0.3 < np.mean(sc["acceptance"]) < 1.0


def test_pooled_covariance_is_the_default_for_large_ensembles_only():
    """n = 500: 300 chains (300 MB of per-chain triangles) pool automatically, 48 chains keep the
    reference's per-chain estimate."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    prec = hmc_error_matrix("spd500")
    for E, pooled in ((48, False), (300, True)):
        eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, 500, E, seed=5)
        eng.set_error_matrix(prec)
        eng.set_dummy_mode(b.DUMMY_TENSOR)
        eng.hmc_set(b.HMC_USER_GRADIENT, 1)
        eng.hmc_start(np.ones(500))
        eng.hmc_step(3)
        if pooled:
            assert eng.hmc_get("pooled_scalars")[0] == 3 * E
        else:
            with pytest.raises(smcmc_b200.SmcmcError):
                eng.hmc_get("pooled_scalars")


@pytest.mark.parametrize("n,E", [(130, 70), (37, 130), (500, 96), (64, 700)])
def test_fused_leapfrog_stage_equals_gradient_plus_kick_drift(monkeypatch, n, E):
    """TENSOR mode runs a leap-frog stage as ONE launch (kHmcLeapDmma: the gradient GEMM with the
    kick, the drift and the U-turn partial sums in its epilogue).  Against the same mode with the
    gradient kernel and kHmcKickDrift as separate launches (SMCMC_HMC_NO_FUSE=1): the same
    arithmetic per element, so positions, potentials, step sizes, trajectory lengths and the
    gradient counters are identical (ragged chain and column tiles, odd and even dimensions)."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    prec = hmc_error_matrix("spd%d" % n)
    runs = {}
    # fused = 2: the fused launches take the chains in order of trajectory length (rows behind the
    # chains that still run skip the GEMM) even where the engine would not bother;
    # fused = 3: every step computes its first gradient instead of taking it from the previous step,
    # and the likelihood of the proposed point by a contraction of its own
    for fused in (1, 0, 2, 3):
        for name in ("SMCMC_HMC_ORDER_ALWAYS", "SMCMC_HMC_NO_GRADIENT_CACHE", "SMCMC_HMC_SEPARATE_POTENTIAL"):
            monkeypatch.delenv(name, raising=False)
        if fused:
            monkeypatch.delenv("SMCMC_HMC_NO_FUSE", raising=False)
            if fused == 2:
                monkeypatch.setenv("SMCMC_HMC_ORDER_ALWAYS", "1")
            if fused == 3:
                monkeypatch.setenv("SMCMC_HMC_NO_GRADIENT_CACHE", "1")
                monkeypatch.setenv("SMCMC_HMC_SEPARATE_POTENTIAL", "1")
        else:
            monkeypatch.setenv("SMCMC_HMC_NO_FUSE", "1")
        h = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=6)
        h.set_error_matrix(prec)
        h.set_dummy_mode(b.DUMMY_TENSOR)
        h.hmc_set(b.HMC_USER_GRADIENT, 1)
        h.hmc_start(np.full(n, 0.5))
        before = h.launch_count()
        tr = h.hmc_step_trace(30 if n < 500 else 12, 0)
        tr["launches"] = h.launch_count() - before
        tr["scalars"] = h.hmc_get("scalars")
        tr["momentum"] = h.hmc_get("momentum")
        runs[fused] = tr
    for k in ("potential", "points", "mean_epsilon", "leapfrog", "accepted", "scalars", "momentum"):
        assert np.array_equal(runs[1][k], runs[0][k]), k
        assert np.array_equal(runs[2][k], runs[0][k]), k
        assert np.array_equal(runs[3][k], runs[0][k]), k
    assert runs[1]["launches"] < 0.62 * runs[0]["launches"]       # one launch per stage instead of two
    assert runs[1]["launches"] < runs[3]["launches"]
    assert np.abs(runs[1]["leapfrog"]).max() > 10 and runs[1]["accepted"].sum() > 0
    if E >= 700:                                                  # the ordering has something to order
        lf = np.abs(runs[1]["leapfrog"])
        assert max(len(np.unique(lf[s])) for s in range(lf.shape[0])) > 1


def test_kept_gradients_are_dropped_when_the_points_or_the_matrix_change(monkeypatch):
    """TENSOR mode takes the first gradient of a step from the previous step (kHmcLeapCached).  Moving
    the chains (SetPosition), restarting them, changing the error matrix or running a step of another
    kind in between must drop what was kept: the run equals the one that computes every gradient
    (SMCMC_HMC_NO_GRADIENT_CACHE=1) bit for bit."""
    import smcmc_b200
    from smcmc_b200 import binding as b
    n, E = 96, 150
    prec, prec2 = hmc_error_matrix("spd%d" % n), hmc_error_matrix("spd%d" % n) * 1.25
    runs = []
    for cache in (1, 0):
        if cache:
            monkeypatch.delenv("SMCMC_HMC_NO_GRADIENT_CACHE", raising=False)
        else:
            monkeypatch.setenv("SMCMC_HMC_NO_GRADIENT_CACHE", "1")
        h = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=9)
        h.set_error_matrix(prec)
        h.set_dummy_mode(b.DUMMY_TENSOR)
        h.hmc_set(b.HMC_USER_GRADIENT, 1)
        h.hmc_start(np.full(n, 0.3))
        out = [h.hmc_step_trace(6, 0)]
        h.hmc_set_position(np.random.default_rng(3).normal(0, 0.4, (E, n)))
        out.append(h.hmc_step_trace(5, 0))
        h.set_error_matrix(prec2)
        out.append(h.hmc_step_trace(5, 0))
        h.hmc_step(2, 5)                                   # gradient type 5 (zero gradient): another kind of step
        out.append(h.hmc_step_trace(5, 0))
        h.hmc_start(np.full(n, -0.2))
        out.append(h.hmc_step_trace(5, 0))
        runs.append(out)
    for a, c in zip(*runs):
        for k in ("potential", "points", "mean_epsilon", "leapfrog", "accepted"):
            assert np.array_equal(a[k], c[k]), k
