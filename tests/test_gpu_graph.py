"""With SMCMC_GRAPH=1 smcmc_step(nsteps) replays the step as a CUDA graph when nothing in it needs the
host (engine.cu, stepMany): the kernels read the step counter from a device word the
graph increments.  The chains must be exactly the ones the plain launch loop gives."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIELDS = ("accepted", "accepted_llh", "sigma", "trials", "successes", "total_steps", "llh_calls", "step_rms")


def _run(monkeypatch, graph, make, steps):
    import smcmc_b200
    if graph:
        monkeypatch.setenv("SMCMC_GRAPH", "1")
    else:
        monkeypatch.delenv("SMCMC_GRAPH", raising=False)
    eng = make(smcmc_b200)
    for n in steps:                       # several calls: the graph is rebuilt per call
        eng.step(n)
    eng.sync()
    out = {f: eng.get(f) for f in FIELDS}
    out["launches"] = eng.launch_count()
    # and the stream continues correctly afterwards (traced steps never use the graph)
    out["tail"] = eng.step_trace(5, want=("accepted", "points"))["points"]
    return out


def _adaptive(sm):
    eng = sm.Engine(sm.LLH_HORRIFIC, 20, 257, seed=9)
    eng.start(np.zeros(20))
    return eng


def _vaat(sm):
    eng = sm.Engine(sm.LLH_UNIT_GAUSS, 7, 64, seed=4, proposal=sm.PROPOSAL_VAAT)
    eng.start(np.zeros(7))
    return eng


def _events(sm):
    events, data = sm.synth.fake_inputs(80, 80, 10, seed=6)
    eng = sm.Engine(sm.LLH_FAKE, 9, 3, seed=8)                 # streaming kernel
    eng.set_fake_events(events)
    eng.set_fake_data(data, 0.1)
    eng.start(np.random.default_rng(1).uniform(-1, 1, (3, 9)))
    return eng


@pytest.mark.parametrize("make", [_adaptive, _vaat, _events])
def test_graph_replay_equals_the_launch_loop(monkeypatch, make):
    assert torch.cuda.is_available()
    steps = (40, 3, 25)
    a = _run(monkeypatch, True, make, steps)
    b = _run(monkeypatch, False, make, steps)
    for k in FIELDS + ("tail",):
        assert np.array_equal(a[k], b[k]), k
    assert np.all(a["total_steps"] == sum(steps))
    assert a["launches"] >= b["launches"]        # the graph adds one node (the counter) per step
