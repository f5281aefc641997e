"""Small and ragged sizes through the kernels added late in round 1 (scripts/sanitize_small.py:
resident steps with odd / even dimensions and a uniform dimension, the fused accept with a
one-chain warp tile, pooled accumulation with and without off-diagonal blocks, the split
finish kernel, the deferred fEXXT update with a ring of 3).  compute-sanitizer is not
available on the GPU pool, so the script's own assertions and a clean exit are the check."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_small_sizes_run_clean():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_small.py")], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sanitize_small ok" in r.stdout
