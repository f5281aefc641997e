"""The C++ mirror of the reference API (include/TSimpleMCMC.H +
include/smcmc_likelihoods.H): a program written like the reference's
documentation example (reference TSimpleMCMC.H:122-156) compiles against it,
links libsmcmc_b200, and -- on a GPU -- reproduces the reference chain."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from helpers import golden, golden_chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "root-simple-mcmc_b200", "smcmc_b200")


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    import smcmc_b200
    if not os.path.exists(smcmc_b200.library_path()):
        smcmc_b200.build_library()
    exe = str(tmp_path_factory.mktemp("cpp") / "simple_mcmc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "simple_mcmc.cc"),
                    "-L", LIBDIR, "-lsmcmc_b200", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


@pytest.fixture(scope="module")
def hmc_program(tmp_path_factory):
    import smcmc_b200
    if not os.path.exists(smcmc_b200.library_path()):
        smcmc_b200.build_library()
    exe = str(tmp_path_factory.mktemp("cpp") / "simple_hmc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "simple_hmc.cc"),
                    "-L", LIBDIR, "-lsmcmc_b200", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_hmc_program_compiles_and_refuses_to_run_without_gpu(hmc_program):
    r = subprocess.run([hmc_program, "hmc", "1", "3"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_simple_hmc_program_reproduces_reference_chain(hmc_program):
    """SimpleHMC.C through include/TSimpleHMC.H against the golden chain of the
    reference build.  The header builds the as-shipped error matrix itself
    (its own inversion routine), so values agree to rounding, not bit for bit."""
    r = subprocess.run([hmc_program, "hmc", "2", "60"], capture_output=True, text=True, check=True)
    want = golden_chain(golden("hmc.npz"), "dummy100_user")
    rows = re.findall(r"step (\d+) potential (\S+) x0 (\S+) eps (\S+) leap (-?\d+)", r.stdout)
    assert len(rows) == 60
    pot = np.array([float(x[1]) for x in rows])
    x0 = np.array([float(x[2]) for x in rows])
    eps = np.array([float(x[3]) for x in rows])
    assert np.allclose(pot, want["potential"][:60], rtol=1e-7)
    assert np.allclose(x0, want["x"][:60, 0], rtol=1e-7, atol=1e-9)
    assert np.allclose(eps, want["epsilon"][:60], rtol=1e-9)
    m = re.search(r"entries (\d+) expected (\d+) calls (\d+) gradients (\d+)", r.stdout)
    assert m and m.group(1) == m.group(2) == str(2 * 61)
    assert int(m.group(3)) == 61                        # Start + one potential per step
    assert "caught invalid_argument: Must initialize starting point" in r.stdout


@pytest.mark.gpu
def test_restore_through_the_cpp_api(hmc_program):
    """Save 300 steps + the full state to a tree, Restore() a NEW sampler from
    it: it continues with the accept/reject sequence of the reference's own
    restore run (reference TSimpleMCMC.H:282-352, golden chain restore7)."""
    r = subprocess.run([hmc_program, "restore", "3", "300"], capture_output=True, text=True, check=True)
    want = golden_chain(golden("chains.npz"), "restore7")
    m = re.search(r"restored llh (\S+) saved (\S+) x0 (\S+) sigma (\S+) trials (\d+)", r.stdout)
    assert m and float(m.group(1)) == float(m.group(2))
    assert float(m.group(3)) == want["saved_accepted"][0]
    rows = re.findall(r"step (\d+) acc (\d) llh (\S+)", r.stdout)
    assert len(rows) == 50
    assert np.array_equal(np.array([int(x[1]) for x in rows]), want["accepted"][:50])
    assert np.allclose(np.array([float(x[2]) for x in rows]), want["llh_accepted"][:50], rtol=1e-12, atol=1e-13)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compiles_and_refuses_to_run_without_gpu(program):
    r = subprocess.run([program, "unit", "1", "3"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_documentation_example_reproduces_reference_chain(program):
    r = subprocess.run([program, "unit", "1", "300"], capture_output=True, text=True, check=True)
    want = golden_chain(golden("chains.npz"), "unit5_hints")
    rows = re.findall(r"step (\d+) acc (\d) llh (\S+) x0 (\S+) sigma (\S+)", r.stdout)
    assert len(rows) == 300
    acc = np.array([int(x[1]) for x in rows])
    llh = np.array([float(x[2]) for x in rows])
    x0 = np.array([float(x[3]) for x in rows])
    assert np.array_equal(acc, want["accepted"][:300])
    assert np.allclose(llh, want["llh_accepted"][:300], rtol=1e-12, atol=1e-13)
    assert np.allclose(x0, want["x"][:300, 0], rtol=1e-12, atol=1e-13)
    m = re.search(r"entries (\d+) expected (\d+) accepted (\d+) calls (\d+)", r.stdout)
    assert m and m.group(1) == m.group(2) == "301"
    assert int(m.group(4)) == 301                       # Start + 300 steps
    assert int(m.group(3)) == int(want["accepted"][:300].sum())
    assert "last covariance size 15 first covariance size 0" in r.stdout   # full state only on SaveStep()
    assert "caught invalid_argument: Uninitialized starting point" in r.stdout
    assert "start llh 0 direct 0" in r.stdout


@pytest.mark.gpu
def test_fake_likelihood_ensemble_through_the_cpp_api(program):
    r = subprocess.run([program, "fake", "8", "40"], capture_output=True, text=True, check=True)
    m = re.search(r"entries (\d+) expected (\d+) accepted (\d+) calls (\d+)", r.stdout)
    assert m and m.group(1) == m.group(2) == str(8 * 41)
    s = re.search(r"start llh (\S+) direct (\S+)", r.stdout)
    assert s and s.group(1) == s.group(2) and np.isfinite(float(s.group(1)))


@pytest.mark.gpu
def test_example2_likelihood_ensemble_through_the_cpp_api(program):
    """example2's FakeLikelihood (include/smcmc_likelihoods.H::FakeLikelihood2) started at the
    true event counts, 8 chains."""
    r = subprocess.run([program, "fake2", "8", "40"], capture_output=True, text=True, check=True)
    m = re.search(r"entries (\d+) expected (\d+) accepted (\d+) calls (\d+)", r.stdout)
    assert m and m.group(1) == m.group(2) == str(8 * 41)
    s = re.search(r"start llh (\S+) direct (\S+)", r.stdout)
    assert s and s.group(1) == s.group(2) and np.isfinite(float(s.group(1)))
    assert float(s.group(1)) > -2000.0          # near the truth the fit is good


@pytest.mark.gpu
def test_vaat_proposal_through_the_cpp_api(program):
    """SimpleVAAT.C's pairing TSimpleMCMC<L, TProposeVAATStep> through include/TProposeVAATStep.H:
    chain 0 of the program is chain 3 of seed 7, the golden chain "vaat_unit5" of the reference build."""
    r = subprocess.run([program, "vaat", "2", "400"], capture_output=True, text=True, check=True)
    want = golden("vaat.npz")
    rows = re.findall(r"step (\d+) acc (\d) llh (\S+) x0 (\S+) sigma (\S+)", r.stdout)
    assert len(rows) == 400
    acc = np.array([int(x[1]) for x in rows])
    x0 = np.array([float(x[3]) for x in rows])
    sigma = np.array([float(x[4]) for x in rows])
    assert np.array_equal(acc, want["vaat_unit5/accepted"][:400])
    assert np.allclose(x0, want["vaat_unit5/x"][:400, 0], rtol=1e-12, atol=1e-13)
    assert np.allclose(sigma, want["vaat_unit5/sigma"][:400], rtol=1e-12)
    m = re.search(r"entries (\d+) accepted (\d+) trials (\d+) successes (\d+) window (\S+)", r.stdout)
    assert m and int(m.group(1)) == 2 * 400 and int(m.group(3)) == 400 and float(m.group(5)) == 100.0
    assert int(m.group(2)) == int(acc.sum())


@pytest.mark.gpu
def test_debug_modes_through_the_cpp_api(program):
    """ForceStep / SetScanDimension / SetEstimatedCenter / GetCovarianceFrozen through the mirror
    header against the golden run "debug9" of the reference build."""
    r = subprocess.run([program, "debug"], capture_output=True, text=True, check=True)
    want = golden_chain(golden("chains.npz"), "debug9")
    rows = re.findall(r"step (\d+) acc (\d) llh (\S+) x0 (\S+) x3 (\S+) x6 (\S+)", r.stdout)
    assert len(rows) == 187
    assert np.array_equal(np.array([int(x[1]) for x in rows]), want["accepted"])
    for col, dim in ((3, 0), (4, 3), (5, 6)):
        assert np.allclose(np.array([float(x[col]) for x in rows]), want["x"][:, dim], rtol=1e-12, atol=1e-14)
    assert np.allclose(np.array([float(x[2]) for x in rows]), want["llh_accepted"], rtol=1e-11, atol=1e-13)
    assert "frozen 0 calls 188" in r.stdout
    assert "caught invalid_argument: Invalid forced step point." in r.stdout


@pytest.mark.gpu
def test_randomized_restore_through_the_cpp_api(program):
    """Restore(tree, randomize = true) (reference :309-316): each of six chains adopts the
    same entry of its 400-entry tree as the reference build did."""
    r = subprocess.run([program, "restore_random"], capture_output=True, text=True, check=True)
    want = golden("chains.npz")["restore_random__picks"]
    rows = re.findall(r"pick (\d) x (\S+) (\S+) (\S+) (\S+) llh (\S+)", r.stdout)
    assert len(rows) == 6
    for row, w in zip(rows, want):
        got = np.array([float(v) for v in row[1:]])
        assert np.allclose(got[:4], w[1:5], rtol=1e-12, atol=1e-14)
        assert np.isclose(got[4], w[5], rtol=1e-11)


@pytest.mark.gpu
def test_saving_an_ensemble_to_the_tree_reads_the_device_once_per_step(program):
    """4096 chains, Step(true) with a tree attached: one entry per chain and step, and the
    cost per step stays within a small multiple of the step itself (it used to be one device
    read per chain and branch: minutes per step)."""
    r = subprocess.run([program, "tree", "4096", "6", "100"], capture_output=True, text=True, check=True)
    m = re.search(r"ms_per_step_plain (\S+) ms_per_step_tree (\S+) full_save_ms (\S+) entries (\d+)", r.stdout)
    assert m and int(m.group(4)) == 4096 * 7
    assert float(m.group(2)) < 100.0 and float(m.group(3)) < 200.0
