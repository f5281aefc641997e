"""The N>1 host logic on CPU: two gloo ranks shard an ensemble, the shards
tile it exactly, the pooled moments equal the single-process moments, and a
chain's draws depend on its global index only (not on the sharding)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "root-simple-mcmc_b200")
TOTAL, DIM = 37, 6


def _draws(seed, chain, nsteps, nslots):
    lib = ctypes.CDLL(os.path.join(PKG, "smcmc_b200", "libsmcmc_hostkat.so"))
    lib.smcmc_kat_normals.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_void_p]
    out = np.zeros(nsteps * nslots)
    lib.smcmc_kat_normals(seed, chain, 0, nsteps, nslots, out.ctypes.data)
    return out


def _ensemble_points():
    rng = np.random.default_rng(5)
    return rng.normal(0, 1, (TOTAL, DIM)) @ rng.normal(0, 1, (DIM, DIM))


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, PKG)
    from smcmc_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offset, count = shard.chain_shard(TOTAL, world, rank)
    pts = _ensemble_points()[offset:offset + count]
    n, mean, cov = shard.pooled_moments(pts, dist)
    # every chain of this shard draws from its GLOBAL index
    draws = np.stack([_draws(9, offset + c, 3, DIM + 1) for c in range(count)])
    gathered = shard.gather_chains(draws, dist)
    # max-over-ranks timing, as bench.py does it
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), offset=offset, count=count, n=n, mean=mean, cov=cov,
             gathered=gathered, tmax=t.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import subprocess
    subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "../smcmc_b200/libsmcmc_hostkat.so"], check=True,
                   stdout=subprocess.DEVNULL)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / ("rank%d.npz" % k)) for k in range(2)]
    # the shards tile the ensemble
    assert int(r[0]["offset"]) == 0 and int(r[0]["count"]) == 19
    assert int(r[1]["offset"]) == 19 and int(r[1]["count"]) == 18
    # pooled moments == single-process moments, identical on both ranks
    pts = _ensemble_points()
    assert float(r[0]["n"]) == TOTAL
    assert np.allclose(r[0]["mean"], pts.mean(0), rtol=1e-12, atol=1e-14)
    assert np.allclose(r[0]["cov"], np.cov(pts.T, bias=True), rtol=1e-10, atol=1e-12)
    assert np.array_equal(r[0]["cov"], r[1]["cov"]) and np.array_equal(r[0]["mean"], r[1]["mean"])
    # draws are a function of the global chain index only
    want = np.stack([_draws(9, c, 3, DIM + 1) for c in range(TOTAL)])
    assert np.array_equal(r[0]["gathered"], want) and np.array_equal(r[1]["gathered"], want)
    assert float(r[0]["tmax"][0]) == 2.0 and float(r[1]["tmax"][0]) == 2.0


def test_shard_arithmetic():
    import sys
    sys.path.insert(0, PKG)
    from smcmc_b200 import shard
    for total in (1, 7, 8, 4096, 262144, 1000003):
        for world in (1, 2, 4, 8):
            blocks = [shard.chain_shard(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (o1, c1), (o2, _) in zip(blocks, blocks[1:]):
                assert o1 + c1 == o2
            assert blocks[-1][0] + blocks[-1][1] == total
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard.chain_shard(10, 2, 2)
