"""Step-for-step parity of the device Metropolis chains with the oracle:
TSimpleMCMC::Step (TSimpleMCMC.H:370-496) + TProposeAdaptiveStep (:640-1831),
driven by identical draws (include/smcmc_rng.h).

What is required:
  * the accept/reject sequence is IDENTICAL, step for step, for every chain;
  * with the step size frozen (no pow() in the loop) every accepted point, the
    covariance and its Cholesky factor are BIT-IDENTICAL;
  * with the default adaptive step size the points agree to 1e-12 relative
    (CUDA's pow and the host's differ in the last ulp of sigma).
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CHAINS, bind_likelihood_inputs, configure_golden, golden, golden_chain

pytestmark = pytest.mark.gpu


def _set_field(eng, name, value):
    from smcmc_b200 import binding
    eng.prop_set({"acceptance_rigidity": binding.PROP_ACCEPTANCE_RIGIDITY,
                  "sigma": binding.PROP_SIGMA}[name], value)


def close(a, b, rtol=1e-12):
    """Relative to the scale of the quantity (coordinates of order one pass
    through zero, so a pure relative test would be meaningless there)."""
    scale = max(1.0, float(np.max(np.abs(b))))
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), scale))


@pytest.mark.parametrize("name", sorted(GOLDEN_CHAINS))
def test_golden_chain(name):
    """One engine holds 4 chains; chain `id` must equal the golden run of the
    reference for that chain id, whatever its neighbours do."""
    import smcmc_b200
    assert torch.cuda.is_available()
    kind, dim, seed, chain, nsteps, start = GOLDEN_CHAINS[name]
    g = golden("chains.npz")
    want = golden_chain(g, name)
    lo = max(0, chain - 2)
    eng = smcmc_b200.Engine(kind, dim, 4, seed=seed, chain_offset=lo)
    bind_likelihood_inputs(eng, kind, g)
    configure_golden(name, eng, _set_field)
    x0 = np.zeros(dim) if start is None else np.full(dim, start)
    ok = eng.start(x0)
    assert ok[chain - lo] == int(want["ok"][0])
    tr = eng.step_trace(nsteps)
    c = chain - lo
    assert np.array_equal(tr["accepted"][:, c], want["accepted"])
    exact = name in ("unit9_frozen_sigma",)
    if exact:
        assert np.array_equal(tr["points"][:, c], want["x"])
        assert np.array_equal(tr["llh_accepted"][:, c], want["llh_accepted"])
        assert np.array_equal(tr["llh_proposed"][:, c], want["llh_proposed"])
    else:
        assert close(tr["points"][:, c], want["x"])
        assert close(tr["llh_accepted"][:, c], want["llh_accepted"], 1e-11)
        assert close(tr["sigma"][:, c], want["sigma"])
    n = dim
    cov_full = want["final_cov"]
    packed = np.array([cov_full[i, j] for i in range(n) for j in range(i + 1)])
    got_cov = eng.get("covariance")[c]
    got_dec = eng.get("decomposition")[c]
    if exact:
        assert np.array_equal(got_cov, packed)
        assert np.array_equal(got_dec, want["final_decomp"])
        assert np.array_equal(eng.get("center")[c], want["final_center"])
    else:
        assert close(got_cov, packed, 1e-10)
        assert np.allclose(got_dec, want["final_decomp"], rtol=1e-9, atol=1e-12)
    scal = dict(zip(__import__("oracle.cpu_checkers", fromlist=["x"]).STATE_FIELDS, want["final_scalars"]))
    assert eng.get("trials")[c] == scal["trials"]
    assert eng.get("successes")[c] == scal["successes"]
    assert eng.get("next_update")[c] == scal["next_update"]
    assert eng.get("total_steps")[c] == scal["total_steps"]
    assert eng.get("llh_calls")[c] == scal["llh_calls"]
    assert abs(eng.get("step_rms")[c] - scal["step_rms"]) <= 1e-12 * scal["step_rms"]


def test_every_chain_of_an_ensemble_matches_its_own_oracle_run(checkers):
    """64 chains, frozen step size: all 64 are bit-identical to 64 separate
    oracle runs (different chain ids = different draws)."""
    import smcmc_b200
    from smcmc_b200 import binding
    E, dim, seed, n = 64, 7, 21, 600
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, dim, E, seed=seed)
    eng.prop_set(binding.PROP_ACCEPTANCE_RIGIDITY, -1.0)
    eng.prop_set(binding.PROP_SIGMA, 0.5)
    x0 = np.random.default_rng(1).uniform(-1, 1, (E, dim))
    eng.start(x0)
    tr = eng.step_trace(n)
    for c in range(E):
        o = checkers.CpuChain("orc", checkers.LLH_UNIT_GAUSS, dim, seed, c)
        o.set(checkers.SET_ACCEPTANCE_RIGIDITY, -1.0)
        o.set(checkers.SET_SIGMA, 0.5)
        o.start(x0[c])
        w = o.step(n)
        assert np.array_equal(tr["accepted"][:, c], w["accepted"]), c
        assert np.array_equal(tr["points"][:, c], w["x"]), c
    # and the ensemble samples the target: pooled mean 0, variance 1
    pts = tr["points"][200:].reshape(-1, dim)
    assert np.all(np.abs(pts.mean(0)) < 0.1)
    assert np.all(np.abs(pts.var(0) - 1.0) < 0.15)


def test_fake_likelihood_schedule_matches_golden():
    """example/FakeMCMC.C:93-165 in miniature: burn-in, ResetProposal,
    burn-in, UpdateProposal, run -- accept sequence identical to the reference
    build, points to 1e-12."""
    import smcmc_b200
    g = golden("fake_likelihood.npz")
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 8, seed=4242)
    eng.set_fake_events(g["events"])
    eng.set_fake_data(g["data"], float(g["exposure"]))
    x0 = np.zeros((8, 9))
    for c in range(8):
        x0[c] = np.random.default_rng(100 + c).uniform(-1, 1, 9)
    assert np.array_equal(x0[0], g["chain0_x0"]) and np.array_equal(x0[7], g["chain7_x0"])
    eng.start(x0)
    parts = [eng.step_trace(60)]
    eng.reset_proposal()
    parts.append(eng.step_trace(60))
    eng.update_proposal()
    parts.append(eng.step_trace(120))
    for chain in (0, 7):
        acc = np.concatenate([p["accepted"][:, chain] for p in parts])
        pts = np.concatenate([p["points"][:, chain] for p in parts])
        la = np.concatenate([p["llh_accepted"][:, chain] for p in parts])
        assert np.array_equal(acc, g["chain%d_accepted" % chain])
        assert close(pts, g["chain%d_x" % chain])
        assert close(la, g["chain%d_llh_accepted" % chain])


def test_step_equals_step_trace_and_metropolis_modes(checkers):
    import smcmc_b200
    E, dim, seed = 16, 5, 9
    a = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, dim, E, seed=seed)
    b = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, dim, E, seed=seed)
    a.start(np.zeros(dim))
    b.start(np.zeros(dim))
    a.step(300)
    b.step_trace(300, want=("accepted",))
    assert np.array_equal(a.get("accepted"), b.get("accepted"))
    assert np.array_equal(a.get("sigma"), b.get("sigma"))
    # metropolis == 2 accepts everything (:414-426), == 1 only uphill (:448)
    tr = a.step_trace(50, metropolis=2)
    assert tr["accepted"].all()
    o = checkers.CpuChain("orc", checkers.LLH_UNIT_GAUSS, dim, seed, 3)
    o.start(np.zeros(dim))
    w1 = o.step(300)
    w2 = o.step(50, metropolis=2)
    assert np.allclose(tr["points"][:, 3], w2["x"], rtol=1e-12, atol=1e-15)
    tr = a.step_trace(200, metropolis=1)
    w3 = o.step(200, metropolis=1)
    assert np.array_equal(tr["accepted"][:, 3], w3["accepted"])
    assert np.all(np.diff(tr["llh_accepted"][:, 3]) >= 0)


def test_bad_start_is_reported():
    """Start() returns false when the starting likelihood is not usable
    (TSimpleMCMC.H:265-268): outside Horrific's box the value is -1e30."""
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, 10, 3, seed=1)
    x0 = np.zeros((3, 10))
    x0[1, 4] = 1.5
    ok = eng.start(x0)
    assert list(ok) == [1, 0, 1]
    eng.step(10)
    assert eng.get("total_steps")[1] == 0 and eng.get("total_steps")[0] == 10


def test_conditioning_ladder_keeps_chains_alive():
    """Out-of-range correlation hints (SimpleMCMC.C:107-115): both Cholesky
    attempts fail and the eigen-decomposition stage (TSimpleMCMC.H:1252-1321)
    must leave every chain with a decomposition U whose U^T U is positive
    definite and reproduces the conditioned covariance up to the clamped
    eigenvalues; the chains must keep sampling the target."""
    import smcmc_b200
    E, dim = 32, 6
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, dim, E, seed=3)
    configure_golden("unit6_clamped", eng, _set_field)
    assert eng.start(np.zeros(dim)).all()
    u = eng.get("decomposition")
    assert np.all(np.isfinite(u))
    packed = eng.get("covariance")[0]
    cov = np.zeros((dim, dim))
    cov[np.tril_indices(dim)] = packed
    cov = cov + np.tril(cov, -1).T
    rebuilt = u[0].T @ u[0]
    assert np.all(np.linalg.eigvalsh(rebuilt) > 0)
    # the hinted matrix is indefinite (one eigenvalue of about -4e-3): the clamp
    # at (1-maxCorr)*lambda_0 replaces it, so U^T U matches only to that size
    assert np.allclose(rebuilt, cov, atol=2e-2)
    tr = eng.step_trace(4000, want=("accepted", "points"))
    assert 0.05 < tr["accepted"].mean() < 0.6
    pts = tr["points"][1500:].reshape(-1, dim)
    assert np.all(np.abs(pts.mean(0)) < 0.15)
    # The clamped hint (2,3) leaves the proposal (almost) no width along x2-x3,
    # and the first covariance update needs ~1000 ACCEPTED steps
    # (TSimpleMCMC.H:1693-1697), more than this run has: as in the reference,
    # x2-x3 stays frozen near the start while x2+x3 samples N(0,2).
    free = [0, 1, 4, 5]
    assert np.all(np.abs(pts[:, free].var(0) - 1.0) < 0.25)
    assert abs(((pts[:, 2] + pts[:, 3]) / np.sqrt(2.0)).var() - 1.0) < 0.25
    assert (pts[:, 2] - pts[:, 3]).var() < 0.25
    assert np.all(eng.get("status") == 0)


def test_restore_continues_the_reference_chain():
    """smcmc_restore_state == Restore() + RestoreState(): a fresh engine that
    adopts the state the reference saved after 300 steps continues with the
    reference's accept/reject sequence (TSimpleMCMC.H:282-352, :1501-1610)."""
    import smcmc_b200
    from oracle import cpu_checkers as cc
    want = golden_chain(golden("chains.npz"), "restore7")
    sc = dict(zip(cc.STATE_FIELDS, want["saved_scalars"]))
    n, E = 7, 3
    cov = want["saved_cov"]
    saved = {
        "accepted": np.tile(want["saved_accepted"], (E, 1)), "log_likelihood": np.full(E, sc["accepted_llh"]),
        "total_steps": np.full(E, sc["total_steps"]), "step_rms": np.full(E, sc["step_rms"]),
        "trials": np.full(E, sc["trials"]), "successes": np.full(E, sc["successes"]),
        "next_update": np.full(E, sc["next_update"]), "acceptance": np.full(E, sc["acceptance"]),
        "acceptance_trials": np.full(E, sc["acceptance_trials"]), "sigma": np.full(E, sc["sigma"]),
        "central_point": np.tile(want["saved_center"], (E, 1)), "central_point_trials": np.full(E, sc["center_trials"]),
        "covariance": np.tile(np.array([cov[i, j] for i in range(n) for j in range(i + 1)]), (E, 1)),
        "covariance_trials": np.full(E, sc["covariance_trials"]), "step_index": np.array([300]),
    }
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=31, chain_offset=0)
    eng.start(np.zeros(n))
    mismatch = eng.restore_state(saved)
    assert not mismatch.any()
    tr = eng.step_trace(200)
    c = 2                                              # the golden chain id
    assert np.array_equal(tr["accepted"][:, c], want["accepted"])
    assert close(tr["points"][:, c], want["x"])
    assert close(tr["sigma"][:, c], want["sigma"])
    fin = dict(zip(cc.STATE_FIELDS, want["final_scalars"]))
    assert eng.get("total_steps")[c] == fin["total_steps"] == 500
    assert eng.get("llh_calls")[c] == fin["llh_calls"]
    assert eng.get("trials")[c] == fin["trials"] and eng.get("next_update")[c] == fin["next_update"]


def test_save_then_restore_round_trip():
    """save_state -> restore_state into a new engine -> save_state returns the
    same chain state.  (Restore is not an uninterrupted continuation in the
    reference either: it resets the accept heuristic's memory and the sigma
    trace, TSimpleMCMC.H:1513-1514,1582.)"""
    import smcmc_b200
    from smcmc_b200 import binding
    n, E = 6, 40

    def fresh():
        e = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, E, seed=8)
        e.prop_set(binding.PROP_COVARIANCE_DEWEIGHT, 0.0)      # UpdateProposal at restore must not deweight
        e.prop_set(binding.PROP_ACCEPTANCE_DEWEIGHT, 0.0)
        e.start(np.zeros(n))
        return e
    a = fresh()
    a.step(150)
    saved = a.save_state()
    assert saved["step_index"][0] == 150 and np.all(saved["total_steps"] == 150)
    b = fresh()
    assert not b.restore_state(saved).any()
    again = b.save_state()
    for k in ("accepted", "log_likelihood", "total_steps", "step_rms", "trials", "successes", "acceptance",
              "acceptance_trials", "sigma", "central_point", "central_point_trials", "covariance",
              "covariance_trials", "step_index"):
        assert np.array_equal(saved[k], again[k]), k
    # the restored engine factored the restored covariance
    u = b.get("decomposition")[3]
    cov = np.zeros((n, n))
    cov[np.tril_indices(n)] = again["covariance"][3]
    cov = cov + np.tril(cov, -1).T
    assert np.allclose(u.T @ u, cov, rtol=1e-12, atol=1e-14)
    b.step(50)
    assert np.all(b.get("total_steps") == 200)
