"""Two GPUs, NCCL over NVLink: event sharding (all-reduce of integer event
counts inside an event group) and pooled adaptation statistics (all-reduce over
the world).  Skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu` runs it."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "root-simple-mcmc_b200")


def _worker(rank, world, uid, out_dir):
    sys.path.insert(0, PKG)
    import smcmc_b200
    from smcmc_b200 import binding, synth
    torch.cuda.set_device(rank)
    events, data = synth.fake_inputs(2000, 2000, 10, seed=41)          # 60 000 events
    pts = np.random.default_rng(2).uniform(-1.5, 1.5, (96, 9))
    # --- event sharding: both ranks hold the SAME 96 chains, half the events each
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 96, seed=7, device=rank, chain_offset=0)
    eng.comm_init(uid, world, rank, event_group=world)
    eng.set_fake_events(events[rank::world])
    eng.set_fake_data(data, 0.11)
    llh = eng.eval(pts)
    counts = eng.fake_counts(pts)
    eng.start(pts)
    tr = eng.step_trace(25, want=("accepted", "points", "llh_accepted"))
    np.savez(os.path.join(out_dir, "shard%d.npz" % rank), llh=llh, counts=counts, acc=tr["accepted"],
             pts=tr["points"], la=tr["llh_accepted"])
    # --- the same with enough chains for the reduce-scatter / all-gather exchange (>= 256 per
    # rank; 700 = a full block for rank 0 and a ragged one for rank 1), and example2's finish
    big = np.random.default_rng(5).uniform(-1.5, 1.5, (700, 9))
    for kind, tag in ((smcmc_b200.LLH_FAKE, "big"), (smcmc_b200.LLH_FAKE2, "big2")):
        if kind == smcmc_b200.LLH_FAKE2:
            big = big.copy()
            big[:, 0] = 20000.0 + 100.0 * big[:, 0]          # example2: parameters 0, 1 are the event counts
            big[:, 1] = 40000.0 + 100.0 * big[:, 1]
        be = smcmc_b200.Engine(kind, 9, 700, seed=11, device=rank, chain_offset=0)
        be.comm_init(_MORE_IDS[0 if tag == "big" else 1], world, rank, event_group=world)
        be.set_fake_events(events[rank::world])
        be.set_fake_data(data, 0.11)
        b_llh = be.eval(big)
        b_counts = be.fake_counts(big[:300])
        be.start(big)
        btr = be.step_trace(12, want=("accepted", "llh_accepted"))
        np.savez(os.path.join(out_dir, "%s%d.npz" % (tag, rank)), llh=b_llh, counts=b_counts, acc=btr["accepted"],
                 la=btr["llh_accepted"])
        be.close()
        if rank == 0:
            bf = smcmc_b200.Engine(kind, 9, 700, seed=11, device=rank, chain_offset=0)
            bf.set_fake_events(events)
            bf.set_fake_data(data, 0.11)
            f_llh = bf.eval(big)
            f_counts = bf.fake_counts(big[:300])
            bf.start(big)
            ftr = bf.step_trace(12, want=("accepted", "llh_accepted"))
            np.savez(os.path.join(out_dir, "%sfull.npz" % tag), llh=f_llh, counts=f_counts, acc=ftr["accepted"],
                     la=ftr["llh_accepted"])
            bf.close()
    if rank == 0:
        full = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 96, seed=7, device=rank, chain_offset=0)
        full.set_fake_events(events)
        full.set_fake_data(data, 0.11)
        f_llh = full.eval(pts)
        f_counts = full.fake_counts(pts)
        full.start(pts)
        ftr = full.step_trace(25, want=("accepted", "points", "llh_accepted"))
        np.savez(os.path.join(out_dir, "full.npz"), llh=f_llh, counts=f_counts, acc=ftr["accepted"],
                 pts=ftr["points"], la=ftr["llh_accepted"])
    eng.close()
    # --- the unbinned likelihood: partial log-likelihoods of the event slices add up
    ub = smcmc_b200.Engine(smcmc_b200.LLH_UNBINNED, 9, 96, seed=7, device=rank, chain_offset=0)
    ub.comm_init(_THIRD_ID[0], world, rank, event_group=world)
    ub.set_unbinned_events(events[rank::world])
    u_llh = ub.eval(pts)
    ub.start(pts)
    utr = ub.step_trace(20, want=("accepted", "llh_accepted"))
    np.savez(os.path.join(out_dir, "unb%d.npz" % rank), llh=u_llh, acc=utr["accepted"], la=utr["llh_accepted"])
    ub.close()
    if rank == 0:
        fu = smcmc_b200.Engine(smcmc_b200.LLH_UNBINNED, 9, 96, seed=7, device=rank, chain_offset=0)
        fu.set_unbinned_events(events)
        f_llh = fu.eval(pts)
        fu.start(pts)
        ftr = fu.step_trace(20, want=("accepted", "llh_accepted"))
        np.savez(os.path.join(out_dir, "unbfull.npz"), llh=f_llh, acc=ftr["accepted"], la=ftr["llh_accepted"])
        fu.close()
    # --- pooled adaptation over chain shards: rank r owns chains [r*256, (r+1)*256)
    n = 6
    pe = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, 256, seed=9, device=rank, chain_offset=rank * 256)
    # a second engine needs its own communicator, hence its own unique id
    pe.prop_set(binding.PROP_POOLED_EVERY, 4)
    pe.comm_init(_SECOND_ID[0], world, rank, event_group=1)
    pe.start(np.zeros(n))
    pe.step(40)
    np.savez(os.path.join(out_dir, "pool%d.npz" % rank), count=pe.get("pooled_count"), cov=pe.get("pooled_covariance"),
             mean=pe.get("pooled_mean"), u=pe.get("pooled_decomposition"))
    pe.close()


_SECOND_ID = [None]
_THIRD_ID = [None]
_MORE_IDS = [None, None]


def _entry(rank, world, uid, uid2, uid3, uid4, uid5, out_dir):
    _SECOND_ID[0] = uid2
    _THIRD_ID[0] = uid3
    _MORE_IDS[0], _MORE_IDS[1] = uid4, uid5
    _worker(rank, world, uid, out_dir)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_event_sharding_and_pooled_statistics_over_two_gpus(tmp_path):
    sys.path.insert(0, PKG)
    from smcmc_b200 import binding
    uid, uid2, uid3, uid4, uid5 = [binding.comm_unique_id() for _ in range(5)]
    mp.spawn(_entry, args=(2, uid, uid2, uid3, uid4, uid5, str(tmp_path)), nprocs=2, join=True)
    full = np.load(tmp_path / "full.npz")
    for r in range(2):
        s = np.load(tmp_path / ("shard%d.npz" % r))
        # integer counts add exactly: sharded == unsharded, bit for bit, on every rank
        assert np.array_equal(s["counts"], full["counts"])
        assert np.array_equal(s["llh"], full["llh"])
        assert np.array_equal(s["acc"], full["acc"])
        assert np.array_equal(s["pts"], full["pts"])
        assert np.array_equal(s["la"], full["la"])
    # 700 chains: the count table is reduce-scattered over chains, each rank finishes its block,
    # the log-likelihoods are all-gathered -- still the unsharded numbers bit for bit
    for tag in ("big", "big2"):
        bfull = np.load(tmp_path / ("%sfull.npz" % tag))
        for r in range(2):
            b = np.load(tmp_path / ("%s%d.npz" % (tag, r)))
            for k in ("counts", "llh", "acc", "la"):
                assert np.array_equal(b[k], bfull[k], equal_nan=True), (tag, r, k)
        assert np.isfinite(bfull["llh"]).all() and bfull["acc"].sum() > 0
    # unbinned: a floating-point sum, so the split changes the last bits only; both ranks
    # hold the same all-reduced value
    uf = np.load(tmp_path / "unbfull.npz")
    u0, u1 = np.load(tmp_path / "unb0.npz"), np.load(tmp_path / "unb1.npz")
    assert np.array_equal(u0["llh"], u1["llh"]) and np.array_equal(u0["acc"], u1["acc"])
    assert np.max(np.abs(u0["llh"] / uf["llh"] - 1.0)) < 1e-12
    assert np.array_equal(u0["acc"], uf["acc"])
    assert np.allclose(u0["la"], uf["la"], rtol=1e-11)
    p0, p1 = np.load(tmp_path / "pool0.npz"), np.load(tmp_path / "pool1.npz")
    assert p0["count"][0] == 2 * 256 * 40
    for k in ("count", "cov", "mean", "u"):
        assert np.array_equal(p0[k], p1[k]), k          # every rank factors the same pooled matrix
