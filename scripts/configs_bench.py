#!/usr/bin/env python
"""Per-config figures of SURVEY.md 8(d) for the BASELINE.json configs other than
the headline one (bench.py measures C2):

  C1  1 chain, default adaptive proposal: steps/s (latency bound), next to the
      reference's own code on one host core
  C3  65 536 chains x 50 dimensions (THorrific / TASym), per-chain adaptation
      (33 KB of state traffic per chain-step => HBM bound) and pooled adaptation
  C4  TSimpleHMC, 500-dimensional Gaussian with the analytic gradient,
      E in {1, 1024, 16384}: gradient + likelihood contractions per second
      against the FP64 peak, EXACT and TENSOR (DMMA) modes
  C5  ensemble sweep: chains x events of the event likelihood sharded over the
      GPUs of the node (chain groups x event groups, integer counts all-reduced
      inside an event group over NCCL)

    python scripts/configs_bench.py c1 c3 c4            # one GPU
    python -m torch.distributed.run --nproc-per-node G scripts/configs_bench.py c5 [--event-group K]

Prints one JSON object per measurement; profiles/r01_configs.jsonl keeps a run."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200"))
sys.path.insert(0, ROOT)
import torch
import smcmc_b200
from smcmc_b200 import binding as b, synth

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)


def emit(rec):
    print(json.dumps(rec), flush=True)


def timed(fn, sync):
    sync()
    t = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t


def c1():
    from oracle import cpu_checkers as cc
    have_ref = cc.available("ref")
    err100 = None
    if have_ref:
        _, err100 = cc.ref_dummy_matrices()
    for name, kind, dim in (("unit gauss 5-dim (README example)", 0, 5), ("TDummyLogLikelihood 100-dim as shipped", 1, 100)):
        if kind == 1 and err100 is None:
            continue
        eng = smcmc_b200.Engine(kind, dim, 1, seed=1)
        if kind == 1:
            eng.set_error_matrix(err100)
        eng.start(np.zeros(dim))
        eng.step(500)
        steps = 5000
        os.environ["SMCMC_NO_RESIDENT"] = "1"
        dt3 = timed(lambda: eng.step(steps), eng.sync)
        del os.environ["SMCMC_NO_RESIDENT"]
        dt = timed(lambda: eng.step(steps), eng.sync)
        rec = {"config": "C1", "target": name, "chains": 1, "dim": dim, "steps_per_s": steps / dt,
               "us_per_step": 1e6 * dt / steps, "bound": "latency (one chain)",
               "kernel": "kStepsResident (all steps of the call in one launch)",
               "steps_per_s_three_launch_step": steps / dt3}
        which = "ref" if have_ref else "orc"
        c = cc.CpuChain(which, kind, dim, 1, 0)
        if kind == 1 and which == "orc":
            c.set_error_matrix(err100)
        c.start(np.zeros(dim))
        c.step(500)
        t = time.perf_counter()
        c.step(steps)
        rec["cpu_steps_per_s"] = steps / (time.perf_counter() - t)
        rec["cpu_kind"] = "reference" if have_ref else "port"
        emit(rec)


def c3():
    E, n, steps = 65536, 50, 100
    tri = n * (n + 1) // 2
    for name, kind in (("THorrificLogLikelihood", 2), ("TASymLogLikelihood", 3)):
        # the reference-exact per-chain adaptation with the chain state resident in shared
        # memory over the steps of one call (kStepsResident): not HBM-bound any more
        for call in (100, 1000):
            os.environ["SMCMC_RESIDENT"] = "1"       # beyond one wave of CTAs the engine would not choose it
            eng = smcmc_b200.Engine(kind, n, E, seed=4)
            eng.start(np.zeros(n) if kind == 2 else np.full(n, 0.01))
            eng.step(20)
            dt = timed(lambda: eng.step(call), eng.sync)
            rate = E * call / dt
            per = ((3 * tri + 10 * n) * 8 + 256) / call
            emit({"config": "C3", "target": name, "chains": E, "dim": n, "mode": "per-chain, resident (%d steps per launch)" % call,
                  "kernel": "kStepsResident", "ms_per_step": 1e3 * dt / call, "chain_steps_per_s": rate,
                  "hbm_bytes_per_chain_step": per, "hbm_gbs": rate * per / 1e9,
                  "three_launch_algorithmic_bytes_per_chain_step": (3 * tri + 6 * n) * 8,
                  "equivalent_hbm_frac_of_three_launch_algorithm": rate * (3 * tri + 6 * n) * 8 / 1e9 / HBM,
                  "acceptance": float(eng.get("acceptance").mean())})
            eng.close()
            del os.environ["SMCMC_RESIDENT"]
        for pooled in (0, 16):
            eng = smcmc_b200.Engine(kind, n, E, seed=4)
            if pooled:
                eng.prop_set(b.PROP_POOLED_EVERY, pooled)
            else:
                os.environ["SMCMC_NO_RESIDENT"] = "1"
            eng.start(np.zeros(n) if kind == 2 else np.full(n, 0.01))
            eng.step(20)
            dt = timed(lambda: eng.step(steps), eng.sync)
            os.environ.pop("SMCMC_NO_RESIDENT", None)
            # per chain-step 3 n(n+1)/2 doubles (covariance read + write, U read)
            # + 6n doubles; pooled mode (4n + 8) doubles
            bytes_per = (3 * tri + 6 * n) * 8 if not pooled else (4 * n + 8) * 8
            rate = E * steps / dt
            emit({"config": "C3", "target": name, "chains": E, "dim": n, "mode": "pooled/%d" % pooled if pooled else "per-chain",
                  "ms_per_step": 1e3 * dt / steps, "chain_steps_per_s": rate, "algorithmic_bytes_per_chain_step": bytes_per,
                  "hbm_gbs": rate * bytes_per / 1e9, "hbm_peak_gbs": HBM, "frac": rate * bytes_per / 1e9 / HBM,
                  "acceptance": float(eng.get("acceptance").mean())})
            eng.close()


def precision(n, seed=5):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n))
    m = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))
    return 0.5 * (m + m.T)


def c4():
    n = 500
    fp64 = b.measure_fp64_peak(0)
    prec = precision(n)
    # 32 timed steps: a whole number of deferred-fEXXT periods (16 steps) lies inside the timed region
    for E, steps, burn in ((1, 32, 5), (1024, 32, 5), (16384, 32, 3)):
        for mode in (b.DUMMY_EXACT, b.DUMMY_TENSOR):
            eng = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=5)
            eng.set_error_matrix(prec)
            eng.set_dummy_mode(mode)
            eng.hmc_set(b.HMC_USER_GRADIENT, 1)
            eng.hmc_start(np.ones(n))
            eng.hmc_step(burn)
            eng.sync()
            s0 = eng.hmc_scalars()
            dt = timed(lambda: eng.hmc_step(steps), eng.sync)
            s1 = eng.hmc_scalars()
            evals = float((s1["gradient_count"] - s0["gradient_count"]).sum() + (s1["potential_count"] - s0["potential_count"]).sum())
            flops = evals * 2.0 * n * n / dt
            emit({"config": "C4", "target": "TSimpleHMC, 500-dim dense Gaussian, analytic gradient", "chains": E, "dim": n,
                  "mode": "tensor (DMMA)" if mode else "exact (reference order)", "ms_per_step": 1e3 * dt / steps,
                  "chain_steps_per_s": E * steps / dt, "gradients_plus_likelihoods_per_s": evals / dt,
                  "tflops_2n2_each": flops / 1e12, "fp64_peak_tflops_measured": fp64, "frac": flops / 1e12 / fp64,
                  "mean_leapfrog": float(np.abs(s1["leapfrog"]).mean()), "acceptance": float(s1["acceptance"].mean())})
            eng.close()


def vaat():
    """TProposeVAATStep on the C3 shape (65 536 chains x 50 dimensions)."""
    E, n, steps = 65536, 50, 200
    for name, kind in (("THorrificLogLikelihood", 2), ("TASymLogLikelihood", 3)):
        eng = smcmc_b200.Engine(kind, n, E, seed=4, proposal=smcmc_b200.PROPOSAL_VAAT)
        eng.start(np.zeros(n) if kind == 2 else np.full(n, 0.01))
        eng.step(50)
        dt = timed(lambda: eng.step(steps), eng.sync)
        # per chain-step: the likelihood reads the proposed point (n doubles), the 128-byte scalar
        # record is read and written twice, the proposal touches O(1) entries; accepted rows
        # (a few per cent here) are copied back (2 n)
        bytes_per = n * 8 + 4 * 128
        rate = E * steps / dt
        emit({"config": "VAAT (8f rank 4)", "target": name, "chains": E, "dim": n, "proposal": "TProposeVAATStep",
              "ms_per_step": 1e3 * dt / steps, "chain_steps_per_s": rate, "algorithmic_bytes_per_chain_step": bytes_per,
              "hbm_gbs": rate * bytes_per / 1e9, "hbm_peak_gbs": HBM, "frac": rate * bytes_per / 1e9 / HBM,
              "acceptance": float(eng.get("acceptance").mean())})
        eng.close()


def stream():
    """The streaming regime of the event likelihood: few chains over 16.8 M events, every
    event read once per evaluation (16 bytes in the FP32 tile layout) => HBM-bound for E <= ~4."""
    N = 16777216
    signal = N // 3 + 1
    events = synth.make_mc_sample(signal, N - signal, seed=2)
    data = synth.make_data_histograms(33334, 33334, seed=2)
    for E in (1, 2, 4, 8):
        eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
        eng.set_fake_events(events)
        eng.set_fake_data(data, 0.006)
        x0 = np.zeros((E, 9))
        for c in range(E):
            x0[c] = np.random.default_rng([3, c]).uniform(-1.0, 1.0, 9)
        eng.start(x0)
        eng.step(3)
        eng.enable_kernel_timing(True)
        eng.pair_kernel_stats(reset=True)
        steps = 30
        dt = timed(lambda: eng.step(steps), eng.sync)
        ms, launches = eng.pair_kernel_stats()
        kernel_s = ms / max(launches, 1) / 1e3
        bytes_alg = N * 16.0
        emit({"config": "event likelihood, streaming regime", "chains": E, "events": N, "kernel": "kFakeStream",
              "ms_per_step": 1e3 * dt / steps, "kernel_ms": 1e3 * kernel_s, "launches_timed": launches,
              "algorithmic_bytes_per_launch": bytes_alg, "hbm_gbs": bytes_alg / kernel_s / 1e9, "hbm_peak_gbs": HBM,
              "frac": bytes_alg / kernel_s / 1e9 / HBM, "pair_evals_per_s": E * float(N) / kernel_s,
              "chain_steps_per_s": E * steps / dt})
        eng.close()


def ess():
    """Effective samples per second on C2 (the metric a user cares about):
    the FakeMCMC.C schedule (100 + reset + 100 + update + production, FakeMCMC.C:93-165), then the
    on-device diagnostics over the production steps."""
    E, prod, lag = 4096, 1000, 96
    signal = 333334
    events = synth.make_mc_sample(signal, 1000000 - signal, seed=2)
    data = synth.make_data_histograms(33334, 33334, seed=2)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
    eng.set_fake_events(events)
    expo = synth.exposure_ratio(eng, data)
    eng.set_fake_data(data, expo)
    x0 = np.zeros((E, 9))
    for c in range(E):
        x0[c] = np.random.default_rng([3, c]).uniform(-1.0, 1.0, 9)
    eng.start(x0)
    eng.step(100)
    eng.reset_proposal()
    eng.step(100)
    eng.update_proposal()
    eng.step(500)
    eng.diag_enable(lag)
    dt = timed(lambda: eng.step(prod), eng.sync)
    d = eng.diag_get()
    emit({"config": "C2 effective samples", "chains": E, "events": len(events), "production_steps": prod,
          "ms_per_step_with_diagnostics": 1e3 * dt / prod, "chain_steps_per_s": E * prod / dt,
          "acceptance": float(eng.get("acceptance").mean()), "rhat_max": float(np.max(d["rhat"])),
          "tau": [float(v) for v in d["tau"]], "ess_per_s_min_dim": float(np.min(d["ess"]) / dt),
          "ess_per_s_median_dim": float(np.median(d["ess"]) / dt), "lag1_autocorrelation": [float(v) for v in d["autocorrelation"][0]],
          "posterior_mean": [float(v) for v in d["mean"]]})


def ex2():
    """example2's likelihood on the shape of C2 (4096 chains x 1M events),
    next to the reference's own example2 code on one host core."""
    from oracle import cpu_checkers as cc
    E, steps = 4096, 20
    signal = 333334
    events = synth.make_mc_sample(signal, 1000000 - signal, seed=2)        # the C2 event set
    data = synth.make_data_histograms(33334, 33334, seed=2, variant=2)
    eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE2, 9, E, seed=3)
    eng.set_fake_events(events)
    eng.set_fake_data(data, 1.0)
    x0 = np.zeros((E, 9))
    for c in range(E):
        x0[c] = np.random.default_rng([3, c]).uniform(-1.0, 1.0, 9)
    x0[:, 0] += 33334.0
    x0[:, 1] += 33334.0
    eng.set_gaussian(0, 300.0)
    eng.set_gaussian(1, 300.0)
    eng.start(x0)
    eng.step(3)
    dt = timed(lambda: eng.step(steps), eng.sync)
    rec = {"config": "example2 (8f rank 2)", "target": "example2/FakeLikelihood.H on the C2 shape", "chains": E,
           "events": len(events), "ms_per_step": 1e3 * dt / steps, "chain_steps_per_s": E * steps / dt,
           "pair_evals_per_s": E * steps * float(len(events)) / dt, "acceptance": float(eng.get("acceptance").mean())}
    which = "ref" if cc.available("ref") else "orc"
    c = cc.CpuChain(which, cc.LLH_FAKE2, 9, 3, 0)
    c.set_fake(events, data, 1.0)
    c.set_gaussian(0, 300.0)
    c.set_gaussian(1, 300.0)
    c.start(x0[0])
    t = time.perf_counter()
    c.step(6)
    rec["cpu_steps_per_s_one_core"] = 6 / (time.perf_counter() - t)
    rec["cpu_kind"] = "reference" if which == "ref" else "port"
    emit(rec)


def c5(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eg = args.event_group
    assert world % eg == 0
    chain_groups = world // eg
    chains_total, events_total = args.chains, args.events
    chains = chains_total // chain_groups                     # per chain group
    offset = (rank // eg) * chains
    signal = events_total // 3 + 1
    events = synth.make_mc_sample(signal, events_total - signal, seed=2)
    data = synth.make_data_histograms(33334, 33334, seed=2)
    mine = events[(rank % eg)::eg]                            # this rank's slice of the events
    kind = smcmc_b200.LLH_UNBINNED if args.unbinned else smcmc_b200.LLH_FAKE
    eng = smcmc_b200.Engine(kind, 9, chains, seed=3, device=local, chain_offset=offset)
    if world > 1 and eg > 1:
        uid = [b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], world, rank, event_group=eg)
    if args.unbinned:
        eng.set_unbinned_events(mine)
    else:
        eng.set_fake_events(mine)
        eng.set_fake_data(data, 0.1)
    x0 = np.zeros((chains, 9))
    for c in range(chains):
        x0[c] = np.random.default_rng([3, offset + c]).uniform(-1.0, 1.0, 9)
    eng.start(x0)
    eng.step(1)

    def sync():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dt = timed(lambda: eng.step(args.steps), sync)
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        pairs = float(chains_total) * float(events_total) * args.steps
        target = ("ensemble sweep of the unbinned mixture likelihood (defined in smcmc_b200.h; 3 exp + 1 log1p in FP64 per pair)"
                  if args.unbinned else "ensemble sweep of the event likelihood (binned Poisson, example/FakeLikelihood.H)")
        exchange = ("all-reduce of %d partial log-likelihoods (f64) per step inside each event group" % chains if args.unbinned
                    else "reduce-scatter of %d x %d uint32 event counts + all-gather of the log-likelihoods per step inside each event group" % (450, chains))
        emit({"config": "C5", "target": target,
              "chains": chains_total, "events": events_total, "gpus": world, "chain_groups": chain_groups, "event_group": eg,
              "steps": args.steps, "s_per_step": dt / args.steps, "chain_steps_per_s": chains_total * args.steps / dt,
              "pair_evals_per_s": pairs / dt, "pair_evals_per_s_per_gpu": pairs / dt / world,
              "exchange": exchange if eg > 1 else "none"})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="+")
    ap.add_argument("--event-group", type=int, default=1)
    ap.add_argument("--chains", type=int, default=262144)
    ap.add_argument("--events", type=int, default=16777216)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--unbinned", action="store_true")
    a = ap.parse_args()
    for w in a.which:
        if w == "c5":
            c5(a)
        elif w == "ex2":
            ex2()
        elif w == "vaat":
            vaat()
        elif w == "stream":
            stream()
        elif w == "ess":
            ess()
        else:
            {"c1": c1, "c3": c3, "c4": c4}[w]()
