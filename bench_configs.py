"""bench_configs.py -- `bench.py --config c1|c3|c4`: the bench contract (one JSON
line with value / e2e / roofline / cpu_baseline / clocks / gpu_launches) for the
BASELINE.json configs other than the headline one.

  c1  SimpleMCMC.C / mcmc.exe (reference SimpleMCMC.C:163-256): 1 chain, burn-in
      schedule, then cycles x steps production steps.  Latency-bound: steps/s
      only, no roofline claim.
  c3  THorrificLogLikelihood, 50 dimensions, 65 536 chains per GPU, per-chain
      (reference-exact) adaptation: 33 KB of state traffic per chain-step,
      HBM-bound; pooled adaptation as a sub-object.
  c4  TSimpleHMC, 500-dimensional dense Gaussian with its analytic gradient,
      16 384 chains per GPU, contractions on the FP64 tensor cores (DMMA).

Imported by bench.py only.  The CPU legs use oracle/ as the checker-side baseline
(the one place besides tests/ and smoke() that may)."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def precision_matrix(n, seed=5):
    """a dense random SPD precision matrix (seed 5)."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n))
    m = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))
    return 0.5 * (m + m.T)


# ---- CPU arms: one single-chain process per host core (the reference's own scale-out model) ----
_w = {}


def _cpu_init(spec):
    from oracle import cpu_checkers as cc
    _w["cc"], _w["spec"] = cc, spec


def _cpu_start(chain):
    cc, spec = _w["cc"], _w["spec"]
    which = spec["which"]
    if spec["sampler"] == "hmc":
        c = cc.CpuHmc(which, cc.LLH_DUMMY, spec["dim"], True, spec["seed"], chain)
        c.set_error_matrix(spec["error"])
        c.start(np.ones(spec["dim"]))
    else:
        c = cc.CpuChain(which, spec["kind"], spec["dim"], spec["seed"], chain)
        c.start(np.full(spec["dim"], spec["x0"]))
    _w["chain"] = c
    return chain


def _cpu_step(nsteps):
    c, spec = _w["chain"], _w["spec"]
    if spec["sampler"] == "hmc":
        c.step(nsteps, 0)
        st = c.state()
        _w["leapfrog"] = abs(st["leapfrog"])
        return st["gradient_count"] + st["potential_count"]
    c.step(nsteps, want_x=False)
    return 0.0


def _cpu_leapfrog(_):
    return _w.get("leapfrog", 0.0)


def cpu_arm(spec, steps, warmup, cores=None):
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, _cpu_init, (spec,)) for _ in range(cores)]
    try:
        for r in [p.apply_async(_cpu_start, (i,)) for i, p in enumerate(pools)]:
            r.get()
        before = [r.get() for r in [p.apply_async(_cpu_step, (warmup,)) for p in pools]]
        t0 = time.perf_counter()
        after = [r.get() for r in [p.apply_async(_cpu_step, (steps,)) for p in pools]]
        wall = time.perf_counter() - t0
        if spec["sampler"] == "hmc":
            spec["mean_trajectory_length"] = float(np.mean([p.apply(_cpu_leapfrog, (0,)) for p in pools]))
    finally:
        for p in pools:
            p.terminate()
    return cores * steps / wall, (sum(after) - sum(before)) / wall, cores


def _which():
    from oracle import cpu_checkers as cc
    cc.build(("orc",))
    return cc


# --------------------------------------------------------------------------------------------
def run(args, emit, ClockSampler):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            {"c1": ref_c1, "c3": ref_c3, "c4": ref_c4}[args.config](args, emit)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MCMC step path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"rank": rank, "world": world, "local": local, "torch": torch, "dist": dist,
           "dev": torch.device("cuda", local), "ClockSampler": ClockSampler}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=ctx["dev"])
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx["barrier"], ctx["max"] = barrier, max_over_ranks
    {"c1": gpu_c1, "c3": gpu_c3, "c4": gpu_c4}[args.config](args, emit, ctx)
    if world > 1:
        dist.destroy_process_group()


def _timed(ctx, fn, steps):
    """`steps` calls of fn() between two CUDA events on the engine's stream, barrier + synchronize
    on both sides, clocks sampled meanwhile; max over ranks.  Returns (ms, clocks)."""
    torch = ctx["torch"]
    sampler = ctx["ClockSampler"](ctx["local"])
    sampler.start()
    ctx["barrier"]()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    ctx["barrier"]()
    ms = ctx["max"](t0.elapsed_time(t1))
    return ms, sampler.stop()


# ---- C3 ------------------------------------------------------------------------------------
C3_CHAINS, C3_DIM, C3_SEED = 65536, 50, 4


def gpu_c3(args, emit, ctx):
    import smcmc_b200
    from smcmc_b200 import binding as b
    torch, rank, world, local = ctx["torch"], ctx["rank"], ctx["world"], ctx["local"]
    E, n = C3_CHAINS, C3_DIM
    tri = n * (n + 1) // 2
    offset = rank * E
    stream = torch.cuda.current_stream().cuda_stream
    os.environ["SMCMC_NO_RESIDENT"] = "1"          # the per-step path: one launch set per step, state through HBM
    eng = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, n, E, seed=C3_SEED, device=local, chain_offset=offset)
    eng.set_stream(stream)
    assert eng.start(np.zeros(n)).all()
    warm = max(args.warmup, 3)
    eng.step(warm)
    eng.sync()
    l0 = eng.launch_count()
    ms, clocks = _timed(ctx, lambda: eng.step(1), args.steps)
    launches = eng.launch_count() - l0
    value = world * E * args.steps / (ms * 1e-3)
    # end to end: Start from host points, every step's accepted points + likelihoods back to pinned host memory
    out = {"points": torch.empty((1, E, n), dtype=torch.float64, pin_memory=True).numpy(),
           "llh_accepted": torch.empty((1, E), dtype=torch.float64, pin_memory=True).numpy(),
           "accepted": torch.empty((1, E), dtype=torch.int32, pin_memory=True).numpy()}
    x0 = torch.zeros((E, n), dtype=torch.float64, pin_memory=True).numpy()

    def e2e_pass(steps):
        e = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, n, E, seed=C3_SEED, device=local, chain_offset=offset)
        e.set_stream(stream)
        ctx["barrier"]()
        t0 = time.perf_counter()
        e.start(x0)
        for _ in range(steps):
            e.step_trace(1, want=("points", "llh_accepted", "accepted"), out=out)
        ctx["barrier"]()
        dt = time.perf_counter() - t0
        e.close()
        return dt
    e2e_pass(3)
    e2e_s = ctx["max"](e2e_pass(args.steps))
    # pooled adaptation (BASELINE.json: "adaptive covariance pooled across chains")
    pe = smcmc_b200.Engine(smcmc_b200.LLH_HORRIFIC, n, E, seed=C3_SEED, device=local, chain_offset=offset)
    pe.set_stream(stream)
    pe.prop_set(b.PROP_POOLED_EVERY, 16)
    pe.start(np.zeros(n))
    pe.step(warm)
    pe.sync()
    pms, _ = _timed(ctx, lambda: pe.step(1), args.steps)
    pe.close()
    del os.environ["SMCMC_NO_RESIDENT"]
    if rank != 0:
        return
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    bytes_per = (3 * tri + 6 * n) * 8.0
    rate = E * args.steps / (ms * 1e-3)                # per GPU
    prate = E * args.steps / (pms * 1e-3)
    line = {
        "metric": "MH steps/sec (chains x steps)", "value": value, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "THorrificLogLikelihood 50-dim, %d chains per GPU, TProposeAdaptiveStep per chain "
                               "(BASELINE.json configs[2])" % E, "chains_per_gpu": E, "dim": n,
                   "l2": "inputs larger than L2: 2.1 GB of per-chain state is read and rewritten every step"},
        "roofline": {"kernel": "smcmc::kProposeStaged (UpdateState + proposal, one CTA per chain, TMA-staged rows)",
                     "bound": "hbm", "achieved": rate * bytes_per / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": rate * bytes_per / 1e9 / hbm, "peak_source": src, "traffic": None,
                     "algorithmic": "%.0f B per chain-step (covariance read + write, factor read: 3 n(n+1)/2 doubles, "
                                    "+ 6 n doubles) x %d chains; whole step timed (proposal + likelihood + accept)"
                                    % (bytes_per, E),
                     "pooled": {"ms_per_step": pms / args.steps, "chain_steps_per_s": prate * world,
                                "algorithmic_bytes_per_chain_step": (4 * n + 8) * 8.0,
                                "hbm_frac": prate * (4 * n + 8) * 8.0 / 1e9 / hbm,
                                "note": "one shared factor for the ensemble, statistics exchanged every 16 steps; bound by "
                                        "the Philox / Box-Muller draws and the per-chain scalar chain, not by HBM"}},
        "e2e": {"value": world * E * args.steps / e2e_s, "unit": "steps/s",
                "h2d_bytes_per_step": int(E * n * 8 / args.steps),
                "d2h_bytes_per_step": int(E * (n * 8 + 8 + 4)),
                "includes": "Start from pinned host points, %d x Step(save=true): accepted points, likelihoods and "
                            "accept flags to pinned host memory every step" % args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cc = _which()
        spec = {"which": "orc", "sampler": "mh", "kind": cc.LLH_HORRIFIC, "dim": n, "seed": C3_SEED, "x0": 0.0}
        v, _, cores = cpu_arm(spec, 20000, 2000)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                "sample": "%d chains (one process per host core) x 20000 steps; the reference's "
                                          "THorrificLogLikelihood is hard-wired to 75 dimensions, so the port runs "
                                          "the 50-dimensional case" % cores}
    emit(line)


def ref_c3(args, emit):
    cc = _which()
    n = C3_DIM
    spec = {"which": "orc", "sampler": "mh", "kind": cc.LLH_HORRIFIC, "dim": n, "seed": C3_SEED, "x0": 0.0}
    steps = max(args.steps, 1) * 1000
    v, _, cores = cpu_arm(spec, steps, 2000)
    desc = {"kind": "port", "cores": cores, "value": v, "unit": "steps/s",
            "sample": "%d chains (one per core) x %d steps; 50-dim port of THorrificLogLikelihood" % (cores, steps)}
    emit({"impl": "reference", "metric": "MH steps/sec (chains x steps)", "value": v, "unit": "steps/s",
          "n_gpus": args.gpus, "steps": steps, "warmup": 2000, "ms_per_step": 1e3 * cores / v, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "THorrificLogLikelihood 50-dim, CPU arm runs %d chains (one per core)" % cores,
                     "chains": cores, "dim": n},
          "cpu_baseline": desc,
          "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ---- C4 ------------------------------------------------------------------------------------
C4_CHAINS, C4_DIM, C4_SEED = 16384, 500, 5


def gpu_c4(args, emit, ctx):
    import smcmc_b200
    from smcmc_b200 import binding as b
    torch, rank, world, local = ctx["torch"], ctx["rank"], ctx["world"], ctx["local"]
    E, n = args.c4_chains, C4_DIM
    offset = rank * E
    prec = precision_matrix(n)
    stream = torch.cuda.current_stream().cuda_stream

    def make():
        e = smcmc_b200.Engine(smcmc_b200.LLH_DUMMY, n, E, seed=C4_SEED, device=local, chain_offset=offset)
        e.set_stream(stream)
        e.set_error_matrix(prec)
        e.set_dummy_mode(b.DUMMY_TENSOR)
        e.hmc_set(b.HMC_USER_GRADIENT, 1)
        return e
    eng = make()
    eng.hmc_start(np.ones(n))

    def measure(k):
        s0 = eng.hmc_scalars()
        l0 = eng.launch_count()
        ms, clocks = _timed(ctx, lambda: eng.hmc_step(1), k)
        s1 = eng.hmc_scalars()
        evals = float((s1["gradient_count"] - s0["gradient_count"]).sum()
                      + (s1["potential_count"] - s0["potential_count"]).sum())
        # the chains' counters follow the reference (L + 1 gradients and one likelihood per step); the
        # engine takes the first gradient of a step from the previous step (the point is the same) and the
        # likelihood out of the last gradient's launch: the contractions it really runs are L per step
        ran = evals
        if not os.environ.get("SMCMC_HMC_NO_GRADIENT_CACHE"):
            ran -= float(E * k)
        if not os.environ.get("SMCMC_HMC_SEPARATE_POTENTIAL"):
            ran -= float((s1["potential_count"] - s0["potential_count"]).sum())
        return ms, clocks, evals, eng.launch_count() - l0, s1, ran

    # The sampler tunes its step size and trajectory length while it runs (TSimpleHMC.H:833-847): on this
    # target the length passes through ~44 around step 40 and settles at 6 by step 80.  An "HMC step"
    # therefore costs 45 gradient evaluations in one regime and 7 in the other.  The line reports the
    # SETTLED regime (SimpleHMC.C keeps its 1000 steps after 100 + n burn-in steps), and the tuning
    # transient of the first steps next to it; gradient + likelihood evaluations per second is the
    # figure that does not depend on the regime.
    eng.hmc_step(40)
    eng.sync()
    t_ms, _, t_evals, _, t_s, t_ran = measure(16)
    transient = {"after_steps": 40, "steps": 16, "ms_per_step": t_ms / 16,
                 "steps_per_s": world * E * 16 / (t_ms * 1e-3),
                 "likelihood_evals_per_s": world * t_evals / (t_ms * 1e-3),
                 "tflops": t_evals * 2.0 * n * n / (t_ms * 1e-3) / 1e12,
                 "tflops_executed": t_ran * 2.0 * n * n / (t_ms * 1e-3) / 1e12,
                 "mean_trajectory_length": float(np.abs(t_s["leapfrog"]).mean())}
    warm = max(args.warmup, 120)
    eng.hmc_step(max(0, warm - 56))
    eng.sync()
    steps = max(args.steps, 32) // 16 * 16          # whole periods of the deferred covariance update
    ms, clocks, evals, launches, s1, ran = measure(steps)
    eng.close()
    out = {"points": torch.empty((1, E, n), dtype=torch.float64, pin_memory=True).numpy(),
           "potential": torch.empty((1, E), dtype=torch.float64, pin_memory=True).numpy()}
    x0 = torch.ones((E, n), dtype=torch.float64, pin_memory=True).numpy()

    def e2e_pass(k):
        e = make()
        ctx["barrier"]()
        t0 = time.perf_counter()
        e.hmc_start(x0)
        for _ in range(k):
            e.hmc_step_trace(1, 0, want=("points", "potential"), out=out)
        ctx["barrier"]()
        dt = time.perf_counter() - t0
        e.close()
        return dt
    e2e_pass(2)
    e2e_steps = 16
    e2e_s = min(ctx["max"](e2e_pass(e2e_steps)), ctx["max"](e2e_pass(e2e_steps)))     # the faster of two passes
    if rank != 0:
        return
    dmma = b.measure_dmma_peak(local)
    dfma = b.measure_fp64_peak(local)
    flops = evals * 2.0 * n * n / (ms * 1e-3) / 1e12
    line = {
        "metric": "HMC steps/sec (chains x steps)", "value": world * E * steps / (ms * 1e-3), "unit": "steps/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "TSimpleHMC, 500-dim dense Gaussian (random SPD precision matrix, seed 5), analytic "
                               "gradient, %d chains per GPU, contractions on the FP64 tensor cores "
                               "(BASELINE.json configs[3])" % E, "chains_per_gpu": E, "dim": n,
                   "l2": "inputs larger than L2: 16384 x 500 points, momenta and gradients (~0.4 GB) plus the "
                         "per-chain covariance accumulators stream through every step"},
        "likelihood_evals_per_s": world * evals / (ms * 1e-3),
        "mean_trajectory_length": float(np.abs(s1["leapfrog"]).mean()), "acceptance": float(s1["acceptance"].mean()),
        "tuning_transient": transient,
        "roofline": {"kernel": "smcmc::kDummyContractDmma (X . Error^T, mma.sync.m8n8k4.f64) over the whole HMC step",
                     "bound": "tensor", "achieved": flops, "peak": dmma, "unit": "TFLOP/s", "frac": flops / dmma,
                     "executed": {"achieved": ran * 2.0 * n * n / (ms * 1e-3) / 1e12,
                                  "frac": ran * 2.0 * n * n / (ms * 1e-3) / 1e12 / dmma,
                                  "note": "contractions the engine really runs: the first gradient of a step is the one "
                                          "the previous step took at the same point, the likelihood of the proposed point "
                                          "comes out of the last gradient's launch (both counted in `achieved`, the "
                                          "algorithmic figure: L + 1 gradients and one likelihood per step as in the "
                                          "reference)"},
                     "peak_source": "measured in this run: register-resident DMMA chains on every SM (FP64 tensor "
                                    "cores; the DFMA chain measures %.1f TFLOP/s)" % dfma,
                     "algorithmic": "2 n^2 = %.0f flop per gradient or likelihood x %.4g evaluations in the timed region "
                                    "(gradient and potential counters of the chains)" % (2.0 * n * n, evals),
                     "traffic": None},
        "e2e": {"value": world * E * e2e_steps / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": int(E * n * 8 / e2e_steps),
                "d2h_bytes_per_step": int(E * (n * 8 + 8)),
                "includes": "Start from pinned host points, %d x Step(save=true): accepted points and potentials to "
                            "pinned host memory every step; the faster of two passes (each with a fresh engine: "
                            "allocations included)" % e2e_steps},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        _which()
        spec = {"which": "orc", "sampler": "hmc", "dim": n, "seed": C4_SEED, "error": prec}
        v, ev, cores = cpu_arm(spec, 40, warm)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                "evals_per_s": ev, "mean_trajectory_length": spec.get("mean_trajectory_length"),
                                "sample": "%d TSimpleHMC chains (one process per host core) x 40 steps after %d "
                                          "warm-up steps, like the GPU arm; the reference's TDummyLogLikelihood is "
                                          "hard-wired to 100 dimensions, so the port runs n = 500" % (cores, warm)}
    emit(line)


def ref_c4(args, emit):
    _which()
    n = C4_DIM
    spec = {"which": "orc", "sampler": "hmc", "dim": n, "seed": C4_SEED, "error": precision_matrix(n)}
    steps = max(args.steps, 1) * 2
    warm = max(args.warmup, 120)                    # the settled regime, as the GPU arm (gpu_c4)
    v, ev, cores = cpu_arm(spec, steps, warm)
    desc = {"kind": "port", "cores": cores, "value": v, "unit": "steps/s", "evals_per_s": ev,
            "mean_trajectory_length": spec.get("mean_trajectory_length"),
            "sample": "%d TSimpleHMC chains (one per core) x %d steps after %d warm-up steps, n = 500" % (cores, steps, warm)}
    emit({"impl": "reference", "metric": "HMC steps/sec (chains x steps)", "value": v, "unit": "steps/s",
          "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * cores / v, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "TSimpleHMC 500-dim dense Gaussian, analytic gradient; CPU arm runs %d chains" % cores,
                     "chains": cores, "dim": n},
          "cpu_baseline": desc,
          "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ---- C1 ------------------------------------------------------------------------------------
def _mcmc_exe_schedule(step, step_saved, reset, update, cycles, steps):
    """The call sequence of mcmc.exe (reference SimpleMCMC.C:163-256): `steps` unsaved burn-in
    steps, ResetProposal, four tuning phases of `steps` saved steps each followed by UpdateProposal,
    then cycles x steps production steps with UpdateProposal between cycles."""
    step(steps)
    reset()
    for _ in range(4):
        step_saved(steps)
        update()


def gpu_c1(args, emit, ctx):
    import smcmc_b200
    torch, rank, local = ctx["torch"], ctx["rank"], ctx["local"]
    if rank != 0:
        return                                        # one chain: replicas only
    n, cycles, steps = 5, 50, 1000
    eng = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, 1, seed=1, device=local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.start(np.zeros(n))                            # SimpleMCMC.C:149
    _mcmc_exe_schedule(eng.step, lambda k: eng.step_trace(k, want=("points",)), eng.reset_proposal,
                       eng.update_proposal, cycles, steps)
    eng.sync()
    l0 = eng.launch_count()
    sampler = ctx["ClockSampler"](local)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(cycles):                           # SimpleMCMC.C:204-256 (-DNO_OUTPUT: no tree)
        eng.step(steps)
        eng.update_proposal()
    t1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    launches = eng.launch_count() - l0
    # end to end with the tree-side record: every production step's accepted point to host memory
    e = smcmc_b200.Engine(smcmc_b200.LLH_UNIT_GAUSS, n, 1, seed=1, device=local)
    e.start(np.zeros(n))
    e.step(steps)
    tw = time.perf_counter()
    for _ in range(cycles):
        e.step_trace(steps, want=("points", "llh_accepted"))
        e.update_proposal()
    e2e_s = time.perf_counter() - tw
    total = cycles * steps
    line = {
        "metric": "MH steps/sec (chains x steps)", "value": total / (ms * 1e-3), "unit": "steps/s", "n_gpus": 1,
        "steps": total, "warmup": 5 * steps, "ms_per_step": ms / total, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "SimpleMCMC.C / mcmc.exe 50 1000: 5-dim Gaussian (the documentation example -1/2 sum x^2), "
                               "1 chain, default TProposeAdaptiveStep, burn-in schedule then 50 x 1000 production steps "
                               "(BASELINE.json configs[0])", "chains": 1, "dim": n,
                   "l2": "not applicable: one chain's state (a few hundred bytes) lives in shared memory for the "
                         "1000 steps of a call (kStepsResident)"},
        "roofline": {"kernel": "smcmc::kStepsResident", "bound": "latency",
                     "note": "one chain on 148 SMs: a step is one dependent chain of ~5 700 cycles (pow, divisions, the "
                             "ordered sums); no bandwidth or throughput roofline applies (SURVEY.md 8d: steps/s only)",
                     "achieved": total / (ms * 1e-3), "peak": None, "unit": "steps/s", "frac": None, "traffic": None},
        "e2e": {"value": total / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * 8 + 8,
                "includes": "every production step's accepted point and likelihood copied to host memory "
                            "(smcmc_step_trace, 1000 steps per call)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _c1_cpu(cycles, steps)
    emit(line)


def _c1_cpu(cycles, steps):
    from oracle import cpu_checkers as cc
    which = "ref" if cc.available("ref") else "orc"
    if which == "orc":
        cc.build(("orc",))
    c = cc.CpuChain(which, cc.LLH_UNIT_GAUSS, 5, 1, 0)
    c.start(np.zeros(5))
    _mcmc_exe_schedule(lambda k: c.step(k, want_x=False), lambda k: c.step(k, want_x=False), c.reset_proposal,
                       c.update_proposal, cycles, steps)
    t = time.perf_counter()
    for _ in range(cycles):
        c.step(steps, want_x=False)
        c.update_proposal()
    dt = time.perf_counter() - t
    return {"value": cycles * steps / dt, "unit": "steps/s", "cores": 1,
            "kind": "reference" if which == "ref" else "port",
            "sample": "the whole workload: one chain, the mcmc.exe schedule, %d production steps on one host core "
                      "(the reference is single-threaded)" % (cycles * steps)}


def ref_c1(args, emit):
    d = _c1_cpu(50, 1000)
    emit({"impl": "reference", "metric": "MH steps/sec (chains x steps)", "value": d["value"], "unit": "steps/s",
          "n_gpus": args.gpus, "steps": 50000, "warmup": 5000, "ms_per_step": 1e3 / d["value"], "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "SimpleMCMC.C / mcmc.exe 50 1000, 5-dim Gaussian, 1 chain", "chains": 1, "dim": 5},
          "cpu_baseline": d,
          "e2e": {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
