import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "root-simple-mcmc_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import smcmc_b200
from smcmc_b200 import binding
from helpers import GOLDEN_CHAINS, configure_golden, golden, golden_chain
def _set_field(eng, name, value):
    eng.prop_set({"acceptance_rigidity": binding.PROP_ACCEPTANCE_RIGIDITY, "sigma": binding.PROP_SIGMA}[name], value)
g = golden("chains.npz")
for name in sorted(GOLDEN_CHAINS):
    kind, dim, seed, chain, nsteps, start = GOLDEN_CHAINS[name]
    want = golden_chain(g, name)
    lo = max(0, chain - 2)
    eng = smcmc_b200.Engine(kind, dim, 4, seed=seed, chain_offset=lo)
    if kind == 1: eng.set_error_matrix(g["dummy100_error"])
    configure_golden(name, eng, _set_field)
    x0 = np.zeros(dim) if start is None else np.full(dim, start)
    eng.start(x0); tr = eng.step_trace(nsteps); c = chain - lo
    same = np.array_equal(tr["accepted"][:, c], want["accepted"])
    first = int(np.argmax(tr["accepted"][:, c] != want["accepted"])) if not same else -1
    dx = np.abs(tr["points"][:, c] - want["x"])
    step_bad = np.argmax(dx.max(1) > 1e-12) if (dx.max(1) > 1e-12).any() else -1
    print(name, "accseq", same, first, "max|dx|", dx.max(), "first step >1e-12:", step_bad,
          "sigma rel", np.max(np.abs(tr["sigma"][:, c] / want["sigma"] - 1)),
          "llh rel", np.max(np.abs(tr["llh_accepted"][:, c] - want["llh_accepted"]) / np.maximum(np.abs(want["llh_accepted"]), 1e-3)), flush=True)
