"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_chain(data, name):
    prefix = name + "__"
    return {k[len(prefix):]: data[k] for k in data.files if k.startswith(prefix)}


# The proposal hints each golden chain was generated with
# (tests/golden/make_golden.py); `c` is anything with the setter names of
# oracle.cpu_checkers.CpuChain / smcmc_b200.Engine.
def configure_golden(name, c, fields):
    if name == "unit5_hints":
        c.set_gaussian(3, 2.0)
        c.set_uniform(4, -5, 5)
        c.set_correlation(3, 4, 0.3)
    elif name == "unit9_frozen_sigma":
        fields(c, "acceptance_rigidity", -1.0)
        fields(c, "sigma", 0.4)
    elif name == "unit6_clamped":
        rng = np.random.default_rng(8)
        for i in range(6):
            for j in range(i + 1, 6):
                c.set_correlation(i, j, float(rng.uniform(-0.2, 0.2)))
        c.set_correlation(2, 3, 2.0)


# name -> (likelihood kind, dim, seed, chain id, steps, start point)
GOLDEN_CHAINS = {
    "unit5_hints": (0, 5, 1, 0, 4000, None),
    "unit5_plain": (0, 5, 1, 2, 4000, None),
    "unit9": (0, 9, 5, 1, 2500, None),
    "unit9_frozen_sigma": (0, 9, 5, 3, 2500, None),
    "horrific75": (2, 75, 4, 11, 2500, None),
    "asym100": (3, 100, 4, 12, 1500, 0.01),
    "dummy100": (1, 100, 9, 0, 1200, None),
    "unit6_clamped": (0, 6, 3, 4, 1500, None),
}
