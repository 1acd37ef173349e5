"""The A/B knobs keep alternative kernels alive: every one of them must still produce the
oracle's bits.  The knobs are read once per process, so each case runs ``__graft_entry__.smoke()`` - front end,
voxel grid, radius outliers, RANSAC ground removal on a 32 768-point scan, compared bit for bit with the CPU
oracle - in a fresh interpreter with the knob set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KNOBS = [
    {"APC_FOLD": "1"},                      # k_begin inside k_dedup_insert, counters written by k_rs_final
    {"APC_RADIUS_CELL": "1"},               # cells of r, 27-cell walk
    {"APC_RADIUS_CELL": "1", "APC_RADIUS_SPLIT": "1"},
    {"APC_RADIUS_COOP": "2,8"},             # warp-cooperative service of the lanes left short
    {"APC_NO_CELL_LIST": "1", "APC_RADIUS_BLOCK": "256"},
    {"APC_STREAM_IN": "1", "APC_VOX_ITEMS": "4", "APC_RS_CH": "10"},
    {"APC_HASH_SLOTS_PER_POINT": "2", "APC_NO_GRID_FUSION": "1"},
]


@pytest.mark.parametrize("knobs", KNOBS, ids=lambda k: "+".join(f"{a}={b}" for a, b in k.items()))
def test_knob_variants_match_the_oracle(knobs):
    env = dict(os.environ, **knobs)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "smoke"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout
