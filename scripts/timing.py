"""Where the end-to-end time goes: engine creation, event upload + re-layout
(first and second call), Start, single traced steps."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200"))
sys.path.insert(0, ROOT)
import torch
import smcmc_b200
from smcmc_b200 import binding, synth

E, N = 4096, 1000000
events = synth.make_mc_sample(N // 3 + 1, N - N // 3 - 1, 2)
data = synth.make_data_histograms(33334, 33334, 2)
pinned = torch.empty(len(events) * 48, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = events.view(np.uint8)
ev = pinned.numpy().view(binding.EVENT_DTYPE)
x0 = np.random.default_rng(0).uniform(-1, 1, (E, 9))


def clock(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print("%-28s %8.2f ms" % (label, (time.perf_counter() - t) * 1e3), flush=True)
    return r


for rep in range(2):
    eng = clock("create", lambda: smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, E, seed=3))
    clock("set_fake_events (1st)", lambda: eng.set_fake_events(ev))
    clock("set_fake_events (2nd)", lambda: eng.set_fake_events(ev))
    clock("set_fake_data", lambda: eng.set_fake_data(data, 0.1))
    clock("start", lambda: eng.start(x0))
    out = {}
    clock("step_trace(1) first", lambda: eng.step_trace(1, want=("points", "llh_accepted", "accepted"), out=out))
    clock("step_trace(1) x10", lambda: [eng.step_trace(1, want=("points", "llh_accepted", "accepted"), out=out) for _ in range(10)])
    clock("step(10)+sync", lambda: (eng.step(10), eng.sync()))
    eng.close()
