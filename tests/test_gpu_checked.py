"""The launch shapes of the step kernel through the range-checked build of the same source (-DNCG_CHECKED: every table, record
and slot index is checked on the device and a violation traps, csrc/ncg_defs.cuh).  compute-sanitizer is closed on the
measurement pool, so this is the memory-safety evidence that can be produced there; races are covered by the bit-identity of
all launch shapes (test_gpu_api.py) and of rank slices vs one engine (test_gpu_replan.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {
    "default": {},
    "queue_rpl4": {"NCG_RAY_QUEUE": "1", "NCG_RAYS_PER_LANE": "4"},
    "pair": {"NCG_PHYS_WARPS": "2", "NCG_RAY_QUEUE": "1", "NCG_RAYS_PER_LANE": "4"},
    "spread": {"NCG_PHYS_WARPS": "4"},
    "unstaged_three_resident": {"NCG_NO_STAGE": "1", "NCG_MIN_BLOCKS": "3", "NCG_RAYS_PER_LANE": "4"},
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_checked_build_runs_every_launch_shape_without_a_violation(shape):
    so = os.path.join(ROOT, "nascargymnasium_b200", "libncg_b200_checked.so")
    assert os.path.exists(so), "build it with NCG_CHECKED=1 (done by __graft_entry__.build())"
    env = dict(os.environ, NCG_CHECKED="1", **SHAPES[shape])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")], capture_output=True, text=True, env=env, timeout=600)
    assert "NCG_CHECK failed" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0 and "sanitize workload ok" in r.stdout, (r.stdout + r.stderr)[-2000:]
