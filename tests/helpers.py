"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_chain(data, name):
    prefix = name + "__"
    return {k[len(prefix):]: data[k] for k in data.files if k.startswith(prefix)}


# The proposal hints each golden chain was generated with
# (tests/golden/make_golden.py); `c` is anything with the setter names of
# oracle.cpu_checkers.CpuChain / smcmc_b200.Engine.
def configure_golden(name, c, fields):
    if name == "unit5_hints":
        c.set_gaussian(3, 2.0)
        c.set_uniform(4, -5, 5)
        c.set_correlation(3, 4, 0.3)
    elif name == "unit9_frozen_sigma":
        fields(c, "acceptance_rigidity", -1.0)
        fields(c, "sigma", 0.4)
    elif name == "unit6_clamped":
        rng = np.random.default_rng(8)
        for i in range(6):
            for j in range(i + 1, 6):
                c.set_correlation(i, j, float(rng.uniform(-0.2, 0.2)))
        c.set_correlation(2, 3, 2.0)
    elif name == "unit6_infinite":
        # SimpleMCMC.C:107-115 -DIMPOSE_RANDOM_CORRELATIONS with its injected fault c = 1.0/0.0
        rng = np.random.default_rng(9)
        for i in range(6):
            for j in range(i + 1, 6):
                c.set_correlation(i, j, float(rng.uniform(-0.2, 0.2)))
        c.set_correlation(1, 4, float("inf"))
        c.set_correlation(0, 5, float("-inf"))


# name -> (likelihood kind, dim, seed, chain id, steps, start point)
GOLDEN_CHAINS = {
    "unit5_hints": (0, 5, 1, 0, 4000, None),
    "unit5_plain": (0, 5, 1, 2, 4000, None),
    "unit9": (0, 9, 5, 1, 2500, None),
    "unit9_frozen_sigma": (0, 9, 5, 3, 2500, None),
    "horrific75": (2, 75, 4, 11, 2500, None),
    "asym100": (3, 100, 4, 12, 1500, 0.01),
    "dummy100": (1, 100, 9, 0, 1200, None),
    "unit6_clamped": (0, 6, 3, 4, 1500, None),
    "unit6_infinite": (0, 6, 3, 5, 1500, None),
    "hard6": (6, 6, 13, 2, 2500, 0.5),
    # example4/Constrained.C: TConstrainedLikelihood (25 dimensions, a constraint on the sum) -- on
    # the device a USER functor (tests/cpp/constrained_functor.cuh), not a built-in kernel
    "constrained25": (8, 25, 51, 3, 3000, 70.0),
}


_USER_LIB = None


def user_functor_library():
    """tests/cpp/user_functor_lib.cu compiled by nvcc (once per session) into
    tests/cpp/_build/libuser_functor.so: the USER's translation unit, with the
    kernels of include/smcmc_device_functor.cuh instantiated for example4's
    TConstrainedLikelihood written as a device functor."""
    global _USER_LIB
    if _USER_LIB is not None:
        return _USER_LIB
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "root-simple-mcmc_b200", "smcmc_b200")
    out = os.path.join(root, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libuser_functor.so")
    src = [os.path.join(root, "tests", "cpp", f) for f in ("user_functor_lib.cu", "constrained_functor.cuh")]
    src.append(os.path.join(root, "include", "smcmc_device_functor.cuh"))
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                        "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
                        "-I", os.path.join(root, "tests", "cpp"), "-o", so, src[0],
                        "-L", libdir, "-lsmcmc_b200", "-Xlinker", "-rpath," + libdir], check=True)
    _USER_LIB = ctypes.CDLL(so)
    return _USER_LIB


def bind_likelihood_inputs(eng, kind, g=None):
    """What a golden chain's likelihood needs before Start(): the error matrix of
    TDummyLogLikelihood, or -- kind 8 -- the user functor registered with the engine."""
    if kind == 1 and g is not None:
        eng.set_error_matrix(g["dummy100_error"])
    if kind == 8:
        eng.bind_user_library(user_functor_library(), "user_constrained_bind")


def run_debug_modes(c):
    """The call sequence of the golden run "debug9" (tests/golden/make_golden.py):
    regular steps, a forced step taken with metropolis = 2 (the idiom of
    TSimpleMCMC.H:797-808), SetEstimatedCenter, scans of a Gaussian and of a uniform
    dimension (:685-704), a forced step under the normal Metropolis rule, regular
    steps again.  `c` has the methods of oracle.cpu_checkers.CpuChain; returns the
    concatenated trace."""
    parts = [c.step(60)]
    c.force_step(np.linspace(-0.4, 0.4, 9))
    parts.append(c.step(1, 2))
    parts.append(c.step(5))
    c.set_center(np.linspace(0.3, -0.3, 9))
    c.set_scan(3)
    parts.append(c.step(25))
    c.set_scan(6)                       # a SetUniform dimension
    parts.append(c.step(15))
    c.set_scan(-1)
    c.force_step(np.full(9, 0.05))
    parts.append(c.step(1))
    parts.append(c.step(80))
    return {k: np.concatenate([p[k] for p in parts]) for k in ("accepted", "llh_accepted", "llh_proposed", "x", "sigma")}


def configure_debug_modes(c):
    c.set_gaussian(3, 0.7)
    c.set_uniform(6, -1.5, 2.0)


# ---------------------------------------------------------------------------
# TSimpleHMC golden chains (tests/golden/hmc.npz, made by make_golden.py with
# the reference build).  name -> settings; "error" names the matrix stored in
# the file for TDummyLogLikelihood chains.
#   pre / post: (setting name, value) applied before / after Start()
#   (Start() resets the mean epsilon, TSimpleHMC.H:229, so a fixed step size
#   has to be set afterwards, as SimpleAHMC.C does).
# ---------------------------------------------------------------------------
HMC_GOLDEN = {
    # SimpleHMC.C: the as-shipped 100-dimensional TDummyLogLikelihood with its own
    # gradient functor, automatic step size and trajectory length; long enough
    # for UpdateErrorMatrix (covariance trials >= 2 dim, :705) to fire
    "dummy100_user": dict(kind=1, dim=100, grad=True, seed=28, chain=0, nsteps=260, gtype=0, x0=1.0, error="dummy100"),
    # SimpleAHMC.C: the covariant approximate gradient (type 2, :447-454)
    "dummy100_covariant": dict(kind=1, dim=100, grad=True, seed=23, chain=5, nsteps=260, gtype=2, x0=0.3, error="dummy100"),
    # fixed trajectory (SetLeapFrog), fixed step size (negative mean epsilon),
    # partial momentum refresh (SetAlpha), forced user gradient (type 4)
    "dummy100_fixed": dict(kind=1, dim=100, grad=True, seed=24, chain=1, nsteps=200, gtype=4, x0=0.2, error="dummy100",
                           pre=(("leapfrog", 7), ("alpha", 0.3)), post=(("mean_epsilon", -0.11),)),
    # TSimpleHMC<L>: finite-difference gradient (:417-444)
    "dummy100_finite": dict(kind=1, dim=100, grad=False, seed=29, chain=2, nsteps=40, gtype=0, x0=0.4, error="dummy100"),
    "unit6_finite": dict(kind=0, dim=6, grad=False, seed=22, chain=0, nsteps=300, gtype=0, x0=0.5),
    # a box-constrained target whose gradient functor declines (THorrificLogLikelihood.H:41-43):
    # finite differences, proposals that leave the box see -1e30
    "horrific75_finite": dict(kind=2, dim=75, grad=True, seed=25, chain=2, nsteps=120, gtype=0, x0=0.0),
    # zero gradient (type 5), and the single-approximate-step mode SetLeapFrog(0) (:598-611)
    "unit5_flat": dict(kind=0, dim=5, grad=False, seed=26, chain=4, nsteps=200, gtype=5, x0=0.1),
    "unit5_forced": dict(kind=0, dim=5, grad=False, seed=27, chain=6, nsteps=200, gtype=0, x0=0.1,
                         pre=(("leapfrog", 0),)),
    # SimpleHMC.C -DUSE_HARD_LIKELIHOOD: the Rosenbrock valley with its gradient functor
    "hard6_user": dict(kind=6, dim=6, grad=True, seed=41, chain=3, nsteps=400, gtype=0, x0=0.8),
    "hard6_finite": dict(kind=6, dim=6, grad=False, seed=42, chain=0, nsteps=150, gtype=0, x0=0.8),
}

# Chains the reference build cannot run (its TDummyLogLikelihood is hard-wired to
# 100 dimensions): the device is compared with the oracle port, which the golden
# chains above pin to the reference.
HMC_PORT_ONLY = {
    "dummy16_user": dict(kind=1, dim=16, grad=True, seed=31, chain=3, nsteps=400, gtype=0, x0=1.0, error="spd16"),
    "dummy10_covariant": dict(kind=1, dim=10, grad=True, seed=33, chain=5, nsteps=400, gtype=2, x0=0.3, error="spd10"),
    "dummy37_user": dict(kind=1, dim=37, grad=True, seed=34, chain=9, nsteps=300, gtype=0, x0=-0.6, error="spd37"),
    "asym7_finite": dict(kind=3, dim=7, grad=False, seed=35, chain=1, nsteps=200, gtype=0, x0=0.05),
}


def hmc_error_matrix(name):
    """Deterministic SPD precision matrices for the TDummyLogLikelihood chains."""
    n = int(name[3:])
    rng = np.random.default_rng(1000 + n)
    a = rng.normal(size=(n, n))
    m = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))
    return 0.5 * (m + m.T)


def hmc_scalar_mask(scalars, dim):
    """The reference never initialises fCurrentCovarianceTrace and
    fEstimatedOrbitLength (TSimpleHMC.H:130-134, :210-269): they hold garbage
    until UpdateErrorMatrix assigns them (:708, :828).  True where a scalar of
    the HMC_STATE_FIELDS block is defined."""
    names = ["acceptance", "mean_epsilon", "leapfrog", "reversal_len", "accepted_potential",
             "proposed_potential", "central_potential", "potential_count", "gradient_count",
             "step_count", "cov_trials", "average_trials", "est_cov_trace", "cur_cov_trace",
             "orbit_length", "steps_remaining", "steps_since_update"]
    s = dict(zip(names, scalars))
    mask = np.ones(len(names), bool)
    mask[names.index("cur_cov_trace")] = s["leapfrog"] != 0 and s["cov_trials"] >= 2 * dim
    mask[names.index("orbit_length")] = s["steps_remaining"] > 0
    return mask
