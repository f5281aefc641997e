"""User-written device functors behind the reference's template contract
(reference TSimpleMCMC.H:48-106: the plugin API is "hand TSimpleMCMC your own
functor"; example4/Constrained.C:17-25 uses it that way).

The functor under test is example4's TConstrainedLikelihood written by the
"user" as a device functor (tests/cpp/constrained_functor.cuh) and compiled by
nvcc in a translation unit of its own; libsmcmc_b200 has no built-in kernel
for it.  Checks: the likelihood equals the reference build's values, a chain
reproduces the golden chain of the reference build (tests/test_gpu_chains.py,
"constrained25", runs through the same binding), the compiled C++ program that
mirrors example4/Constrained.C reproduces it too, the ensemble reproduces the
closed-form posterior that example4/ConstrainedCheck.C:19-69 is meant to show,
and TSimpleHMC runs with the functor's own gradient.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import golden, golden_chain, user_functor_library

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "root-simple-mcmc_b200", "smcmc_b200")


def closed_form():
    """Posterior of example4 (TConstrainedLikelihood.H:26-46 with the priors of :55-110):
    precision diag(1/s_i^2) + 1 1^T / S^2, linear term mu_i/s_i^2 + T/S^2."""
    mu = np.array([76.0] * 24 + [80.0])
    sg = np.array([76.0 * 0.08] * 24 + [2.0])
    prec = np.diag(1.0 / sg ** 2) + np.ones((25, 25)) / 16.0 ** 2
    cov = np.linalg.inv(prec)
    mean = cov @ (mu / sg ** 2 + 1902.0 / 16.0 ** 2)
    return mean, cov


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "constrained")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "constrained.cu"), "-L", LIBDIR, "-lsmcmc_b200",
                    "-Xlinker", "-rpath," + LIBDIR], check=True)
    return exe


def user_engine(chains, seed=1, chain_offset=0):
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_USER, 25, chains, seed=seed, chain_offset=chain_offset)
    eng.bind_user_library(user_functor_library(), "user_constrained_bind")
    return eng


def test_unregistered_functor_fails_loudly():
    import smcmc_b200
    eng = smcmc_b200.Engine(smcmc_b200.LLH_USER, 25, 2, seed=1)
    with pytest.raises(smcmc_b200.SmcmcError) as ei:
        eng.start(np.zeros(25))
    assert ei.value.status == -2 and "smcmc_user_set_ops" in str(ei.value)


def test_user_likelihood_equals_the_reference(checkers):
    """Bit for bit against the port (pinned to the reference build by the golden
    chain), for more points than one CTA holds and for a ragged last CTA."""
    eng = user_engine(4)
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.normal(76.0, 6.0, (333, 25)), np.zeros((1, 25)), np.full((1, 25), 1e6)])
    got = eng.eval(pts)
    o = checkers.CpuChain("orc", checkers.LLH_CONSTRAINED, 25, 1, 0)
    want = np.array([o.llh(p) for p in pts])
    assert np.array_equal(got, want)
    mean, cov = closed_form()
    l0 = eng.eval(mean[None, :])[0]
    d = rng.normal(0, 2, (16, 25))
    quad = np.array([l0 - 0.5 * v @ np.linalg.inv(cov) @ v for v in d])
    assert np.allclose(eng.eval(mean + d), quad, rtol=1e-9)


def test_cpp_program_reproduces_the_golden_chain(program):
    """sMCMC::TSimpleMCMC<TConstrainedLikelihood> through the mirror header: the accept
    sequence of the reference build's chain, points to 1e-12."""
    r = subprocess.run([program, "chain", "400"], capture_output=True, text=True, check=True)
    want = golden_chain(golden("chains.npz"), "constrained25")
    s = re.search(r"start llh (\S+) direct (\S+)", r.stdout)
    assert s and s.group(1) == s.group(2)
    rows = re.findall(r"step (\d+) acc (\d) llh (\S+) x0 (\S+) sigma (\S+)", r.stdout)
    assert len(rows) == 400
    assert np.array_equal(np.array([int(x[1]) for x in rows]), want["accepted"][:400])
    assert np.allclose(np.array([float(x[2]) for x in rows]), want["llh_accepted"][:400], rtol=1e-11)
    assert np.allclose(np.array([float(x[3]) for x in rows]), want["x"][:400, 0], rtol=1e-12)
    assert np.allclose(np.array([float(x[4]) for x in rows]), want["sigma"][:400], rtol=1e-12)
    m = re.search(r"entries (\d+) calls (\d+)", r.stdout)
    assert m and int(m.group(1)) == 400 and int(m.group(2)) == 401


def test_ensemble_reproduces_the_closed_form_posterior():
    """4096 chains with pooled adaptation: mean, marginal variances, the induced
    negative correlation and the variance of the SUM (the constraint) agree with the
    closed form -- what ConstrainedCheck.C profiles from the reference's tree."""
    from smcmc_b200 import binding
    E = 4096
    eng = user_engine(E, seed=9)
    eng.prop_set(binding.PROP_POOLED_EVERY, 20)
    mean, cov = closed_form()
    assert eng.start(np.full(25, 76.0)).all()
    eng.step(1500)
    eng.reset_proposal()
    eng.step(1500)
    pts = []
    for _ in range(16):
        eng.step(60)
        pts.append(eng.get("accepted"))
    x = np.concatenate(pts)
    se = np.sqrt(np.diag(cov) / (E * 4))             # generous: autocorrelated samples
    assert np.all(np.abs(x.mean(0) - mean) < 6 * se), np.abs(x.mean(0) - mean) / se
    got = np.cov(x.T)
    assert np.allclose(np.diag(got), np.diag(cov), rtol=0.06)
    assert abs(got[0, 1] - cov[0, 1]) < 0.15 * abs(cov[0, 1]) + 0.05
    s = x.sum(axis=1)
    want_var = np.ones(25) @ cov @ np.ones(25)
    assert abs(s.mean() - mean.sum()) < 0.5
    assert abs(s.var() / want_var - 1.0) < 0.08
    assert np.all(eng.get("status") == 0)


def test_cpp_ensemble_program(program):
    """The per-chain adaptive version through the C++ API (256 chains, the schedule of
    example4/Constrained.C): means within a few standard errors, variance of the sum."""
    r = subprocess.run([program, "ensemble", "256", "40"], capture_output=True, text=True, check=True)
    mean, cov = closed_form()
    rows = re.findall(r"mean (\d+) (\S+) var (\S+)", r.stdout)
    assert len(rows) == 25
    got_mean = np.array([float(x[1]) for x in rows])
    got_var = np.array([float(x[2]) for x in rows])
    assert np.all(np.abs(got_mean - mean) < 0.5)
    assert np.allclose(got_var, np.diag(cov), rtol=0.2)
    m = re.search(r"sum mean (\S+) var (\S+)", r.stdout)
    want_var = np.ones(25) @ cov @ np.ones(25)
    assert abs(float(m.group(1)) - mean.sum()) < 1.5
    assert abs(float(m.group(2)) / want_var - 1.0) < 0.25


def test_hmc_with_the_functors_own_gradient(program):
    """TSimpleHMC<L, L>: the user gradient entry is called (gradient count grows by
    trajectory length + 1 per step, no finite-difference potentials), and the ensemble
    moves to the constrained region."""
    from smcmc_b200 import binding as b
    eng = user_engine(512, seed=4)
    eng.hmc_set(b.HMC_USER_GRADIENT, 1)
    eng.hmc_start(np.full(25, 76.0))
    eng.hmc_step(300, 4)                      # type 4: user gradient or an error (TSimpleHMC.H:521)
    sc = eng.hmc_scalars()
    assert np.all(sc["step_count"] == 300)
    # Start + one per step (+ one per UpdateErrorMatrix, :729): finite differences would add 2 x 25 per gradient
    assert np.all(sc["potential_count"] >= 301) and np.all(sc["potential_count"] < 400)
    assert np.all(sc["gradient_count"] > 300)
    x = eng.hmc_get("accepted")
    mean, cov = closed_form()
    assert abs(x.sum(axis=1).mean() - mean.sum()) < 3.0
    assert np.all(np.abs(x.mean(0) - mean) < 1.5)
    # and the user gradient agrees with finite differences of the user likelihood
    eng2 = user_engine(4, seed=4)
    eng2.hmc_start(np.full(25, 70.0))
    eng2.hmc_step(1, 3)
    r = subprocess.run([program, "hmc", "4", "20"], capture_output=True, text=True, check=True)
    assert re.search(r"hmc potentials 21 gradients (\d+)", r.stdout)
