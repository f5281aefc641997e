"""First contact with the GPU: correctness probes + a rough timing."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT)
import smcmc_b200
from oracle import cpu_checkers as cc
import __graft_entry__ as ge

ge.smoke()
# analytic likelihood chains, bit-exact?
for kind, dim in [(0, 5), (0, 9), (2, 75), (3, 100), (2, 50)]:
    E, seed, N = 8, 11, 1500
    eng = smcmc_b200.Engine(kind, dim, E, seed=seed)
    eng.start(np.zeros((E, dim)))
    tr = eng.step_trace(N)
    o = cc.CpuChain("orc", kind, dim, seed, 5); o.start(np.zeros(dim)); w = o.step(N)
    same_acc = np.array_equal(tr["accepted"][:, 5], w["accepted"])
    same_x = np.array_equal(tr["points"][:, 5], w["x"])
    dx = np.max(np.abs(tr["points"][:, 5] - w["x"]))
    print("kind", kind, "dim", dim, "acc-seq", same_acc, "x-bit", same_x, "maxdx", dx,
          "sigma rel", np.max(np.abs(tr["sigma"][:, 5] / w["sigma"] - 1)), "acc", w["accepted"].mean(), flush=True)
    st = o.state()
    print("   sigma", eng.get("sigma")[5], st["sigma"], "trace", eng.get("covariance_trace")[5], st["covariance_trace"],
          "nextupd", eng.get("next_update")[5], st["next_update"], "succ", eng.get("successes")[5], st["successes"])
# timing of the headline config
E, N = 4096, 1000000
events = smcmc_b200.synth.make_mc_sample(N // 3, N - N // 3, 1)
data = smcmc_b200.synth.make_data_histograms(33334, 33334, 2)
eng = smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3)
t = time.time(); eng.set_fake_events(events); eng.sync(); print("upload s", time.time() - t)
expo = smcmc_b200.synth.exposure_ratio(eng, data); eng.set_fake_data(data, expo); print("expo", expo)
rng = np.random.default_rng(0)
eng.start(rng.uniform(-1, 1, (E, 9)))
eng.enable_kernel_timing(True)
eng.step(3); eng.sync(); eng.pair_kernel_stats(reset=True)
t = time.time(); eng.step(10); eng.sync(); dt = time.time() - t
ms, n = eng.pair_kernel_stats()
print("10 steps wall %.3f s -> %.1f MH steps/s ; pair kernel %.3f ms/launch (%d)" % (dt, E * 10 / dt, ms / n, n))
print("pairs/s %.3e" % (E * N * n / (ms * 1e-3)))
print("acceptance", eng.get("acceptance").mean(), "sigma", eng.get("sigma").mean())
