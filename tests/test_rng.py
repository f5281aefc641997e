"""The counter-based stream of include/smcmc_rng.h (host side)."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "root-simple-mcmc_b200")


@pytest.fixture(scope="module")
def kat():
    subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "../smcmc_b200/libsmcmc_hostkat.so"],
                   check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(PKG, "smcmc_b200", "libsmcmc_hostkat.so"))
    d, u32, u64 = ctypes.c_double, ctypes.c_uint32, ctypes.c_uint64
    lib.smcmc_kat_uniform.restype = d
    lib.smcmc_kat_uniform.argtypes = [u64, u32, u32, u32, u32]
    lib.smcmc_kat_normal.restype = d
    lib.smcmc_kat_normal.argtypes = [u64, u32, u32, u32, u32]
    lib.smcmc_kat_normals.argtypes = [u64, u32, u32, u32, u32, ctypes.c_void_p]
    lib.smcmc_kat_det_log.restype = d
    lib.smcmc_kat_det_log.argtypes = [d]
    lib.smcmc_kat_det_cos2pi.restype = d
    lib.smcmc_kat_det_cos2pi.argtypes = [d]
    lib.smcmc_kat_det_sin2pi.restype = d
    lib.smcmc_kat_det_sin2pi.argtypes = [d]
    lib.smcmc_kat_normal_pair.argtypes = [u64, u32, u32, u32, u32, ctypes.c_void_p]
    lib.smcmc_kat_bits_to_open01.restype = d
    lib.smcmc_kat_bits_to_open01.argtypes = [u32, u32]
    lib.smcmc_kat_seq_add.restype = d
    lib.smcmc_kat_seq_add.argtypes = [d, d, u32]
    lib.smcmc_kat_seq_add_naive.restype = d
    lib.smcmc_kat_seq_add_naive.argtypes = [d, d, u32]
    return lib


def philox(lib, ctr, key):
    out = (ctypes.c_uint32 * 4)()
    lib.smcmc_kat_philox((ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), out)
    return list(out)


def test_philox_known_answers(kat):
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert philox(kat, [0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox(kat, [0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox(kat, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_open_interval_and_addressing(kat):
    u = [kat.smcmc_kat_uniform(7, c, s, k, 0) for c in range(4) for s in range(50) for k in range(10)]
    assert min(u) > 0.0 and max(u) < 1.0
    assert len(set(u)) == len(u)                       # distinct counters -> distinct draws
    assert abs(np.mean(u) - 0.5) < 0.02
    assert kat.smcmc_kat_uniform(7, 1, 2, 3, 0) == kat.smcmc_kat_uniform(7, 1, 2, 3, 0)
    assert kat.smcmc_kat_uniform(7, 1, 2, 3, 0) != kat.smcmc_kat_uniform(8, 1, 2, 3, 0)
    assert kat.smcmc_kat_uniform(7, 1, 2, 3, 0) != kat.smcmc_kat_uniform(7, 1, 2, 3, 1)


def test_uniform_end_points_are_exact(kat):
    # 52 random bits k -> (k + 1/2) 2^-52: the all-ones word must stay below 1, the zero word above 0
    assert kat.smcmc_kat_bits_to_open01(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53
    assert kat.smcmc_kat_bits_to_open01(0, 0) == 2.0 ** -53
    assert kat.smcmc_kat_bits_to_open01(0x80000000, 0) == 0.5 + 2.0 ** -53
    assert kat.smcmc_kat_bits_to_open01(0, 0xfff) == 2.0 ** -53          # the low 12 bits are not used
    assert kat.smcmc_kat_bits_to_open01(0, 0x1000) == 1.5 * 2.0 ** -52


def test_normals_come_in_pairs(kat):
    # slots 2k and 2k+1 are the cosine and the sine branch of one Box-Muller block
    out = np.zeros(2)
    for chain, step, pair in [(0, 0, 0), (3, 17, 4), (4095, 1999, 24), (7, 1, 249)]:
        kat.smcmc_kat_normal_pair(11, chain, step, pair, 0, out.ctypes.data)
        assert out[0] == kat.smcmc_kat_normal(11, chain, step, 2 * pair, 0)
        assert out[1] == kat.smcmc_kat_normal(11, chain, step, 2 * pair + 1, 0)
        assert out[0] != out[1]
    # the pair is uncorrelated and each branch is a unit normal
    n_steps = 20000
    g = np.zeros(n_steps * 2)
    kat.smcmc_kat_normals(5, 1, 0, n_steps, 2, g.ctypes.data)
    g = g.reshape(n_steps, 2)
    assert abs(np.mean(g[:, 0] * g[:, 1])) < 0.02
    assert abs(g[:, 1].var() - 1.0) < 0.03 and abs(g[:, 0].var() - 1.0) < 0.03
    r2 = (g ** 2).sum(axis=1)                          # chi-square with 2 dof: mean 2
    assert abs(r2.mean() - 2.0) < 0.05
    # a uniform draw of slot 2k does not share its bits with the normals of the pair (own sub-stream)
    u = kat.smcmc_kat_uniform(5, 1, 0, 0, 0)
    kat.smcmc_kat_normal_pair(5, 1, 0, 0, 0, out.ctypes.data)
    rad2 = out[0] ** 2 + out[1] ** 2
    assert abs(math.exp(-0.5 * rad2) - u) > 1e-9


def test_normal_moments(kat):
    n_steps, n_slots = 4000, 50
    g = np.zeros(n_steps * n_slots)
    kat.smcmc_kat_normals(123, 5, 0, n_steps, n_slots, g.ctypes.data)
    assert abs(g.mean()) < 0.01
    assert abs(g.var() - 1.0) < 0.01
    assert abs(np.mean(g ** 3)) < 0.03
    assert abs(np.mean(g ** 4) - 3.0) < 0.06
    assert np.abs(g).max() < 7.0
    # tail fractions of a unit normal
    assert abs(np.mean(np.abs(g) > 1.959964) - 0.05) < 0.003


def test_deterministic_log_and_cos_accuracy(kat):
    rng = np.random.default_rng(2)
    for x in rng.uniform(0, 1, 20000):
        assert abs(kat.smcmc_kat_det_log(float(x)) - math.log(x)) <= 4e-16 * abs(math.log(x)) + 1e-18
    for x in [1e-300, 5e-324, 2.0 ** -53 * 0.5, 0.5, 0.9999999999999999]:
        assert abs(kat.smcmc_kat_det_log(x) - math.log(x)) <= 4e-16 * abs(math.log(x))
    for x in rng.uniform(0, 1, 20000):
        assert abs(kat.smcmc_kat_det_cos2pi(float(x)) - math.cos(2 * math.pi * x)) < 1.5e-15
        assert abs(kat.smcmc_kat_det_sin2pi(float(x)) - math.sin(2 * math.pi * x)) < 1.5e-15
    for k in range(8):                                  # octant boundaries
        for x in (k / 8.0 + 2.0 ** -53, (k + 1) / 8.0 - 2.0 ** -53, k / 8.0 + 0.0625):
            assert abs(kat.smcmc_kat_det_cos2pi(x) - math.cos(2 * math.pi * x)) < 1.5e-15
            assert abs(kat.smcmc_kat_det_sin2pi(x) - math.sin(2 * math.pi * x)) < 1.5e-15


def test_sequential_sum_emulation_is_exact(kat):
    rng = np.random.default_rng(1)
    for t in range(4000):
        w = float(np.exp(rng.uniform(-6, 3)))
        if t % 3 == 0:                                  # tie-heavy: few significant bits
            w = float(2.0 ** int(rng.integers(-6, 3)) * int(rng.integers(1, 16)))
        s = 0.0 if t % 2 else float(rng.uniform(0, 100))
        n = int(rng.integers(0, 100000))
        assert kat.smcmc_kat_seq_add(s, w, n) == kat.smcmc_kat_seq_add_naive(s, w, n)
    # event counts of the size of the streaming benchmark (16.8 M events in a bin's class): many binades crossed
    for w in (0.3141592653589793, 0.0958270088772365, 1.0, 0.75, 1e-3, 0.05 * 2.0 ** -3, 37.5):
        for n in (16777216, 9999999):
            assert kat.smcmc_kat_seq_add(0.0, w, n) == kat.smcmc_kat_seq_add_naive(0.0, w, n), (w, n)
            assert kat.smcmc_kat_seq_add(123.456, w, n) == kat.smcmc_kat_seq_add_naive(123.456, w, n), (w, n)
    for s, w, n in [(0.0, 0.0, 10), (1.0, 1e-30, 1000), (0.0, 0.1, 0), (0.0, 0.1, 1), (0.0, 0.1, 3),
                    (0.0, float("inf"), 5), (0.0, -0.5, 7), (3.0, 0.09582700887723655, 1000020)]:
        a, b = kat.smcmc_kat_seq_add(s, w, n), kat.smcmc_kat_seq_add_naive(s, w, n)
        assert a == b or (math.isnan(a) and math.isnan(b))


def test_hmc_chains_in_order_of_trajectory_length():
    """csrc/hmc_order.h (host code of the fused HMC stage): chains sorted by trajectory length, longest
    first, equal lengths in chain order, chains without a trajectory last; and per gradient k the row
    tiles that hold every chain taking part in it."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(PKG, "smcmc_b200", "libsmcmc_hostkat.so"))
    lib.smcmc_kat_hmc_order.restype = ctypes.c_longlong
    lib.smcmc_kat_hmc_order.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(11)
    for chains, top, tile in [(1, 3, 64), (63, 5, 64), (700, 9, 64), (16384, 50, 64), (1000, 1, 64), (257, 1500, 16)]:
        steps = rng.integers(-1, top + 1, chains).astype(np.int32)
        steps[rng.integers(0, chains)] = top                       # the longest length is present
        order = np.full(chains, -1, np.int32)
        tiles = np.zeros(top + 1, np.int32)
        scratch = np.zeros(top + 2, np.int32)
        busy = lib.smcmc_kat_hmc_order(steps.ctypes.data, chains, top, tile, order.ctypes.data, tiles.ctypes.data,
                                       scratch.ctypes.data)
        assert sorted(order.tolist()) == list(range(chains))       # a permutation
        running = steps[order] >= 1
        assert not np.any(running[1:] & ~running[:-1])             # chains without a trajectory last ...
        assert np.all(np.diff(order[~running]) > 0)                # ... in chain order
        key = steps[order][running]
        assert np.all(np.diff(key) <= 0)                           # lengths descending
        same = np.diff(key) == 0
        assert np.all(np.diff(order[running])[same] > 0)           # equal lengths in chain order
        want = [-(-int(np.sum(steps >= max(k, 1))) // tile) for k in range(top + 1)]
        assert tiles.tolist() == want and busy == sum(want)
        # the first tiles[k] row tiles hold every chain of gradient k
        for k in (0, 1, top // 2, top):
            active = set(np.nonzero(steps >= max(k, 1))[0].tolist())
            assert active <= set(order[: tiles[k] * tile].tolist())
