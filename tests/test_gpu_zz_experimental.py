"""Experimental launch variants that were written after the GPU budget of round 1 was spent (ROADMAP.md): each runs in a
child process with a timeout, must reproduce the default variant bit for bit, and is marked xfail(strict=False) — the product
path does not use them, so a failure here documents the experiment and does not turn the suite red; a pass (XPASS) is the
first validation step of the round-2 plan.  The file sorts after the other GPU tests on purpose: a variant that hangs costs its
child's 75 s timeout (three children in all; scripts/gpu_experimental.sh runs the wider set), and that must not delay the parity suite."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(var, value):
    out = subprocess.run([sys.executable, os.path.join(HERE, "_experimental_child.py"), var, value], capture_output=True, text=True, timeout=75)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"], line
    return line


@pytest.mark.xfail(strict=False, reason="k_topk_fast_grouped has not run on a GPU yet (written without GPU access)")
@pytest.mark.parametrize("groups", ["4"])
def test_grouped_topk_equals_default(groups):
    _run("GDR_TOPK_GROUPS", groups)


@pytest.mark.xfail(strict=False, reason="k_score_topk_fused has not run on a GPU yet (written without GPU access)")
@pytest.mark.parametrize("groups", ["4"])
def test_fused_score_topk_equals_default(groups):
    """ONE launch scoring batch i and selecting the top-k of batch i-1 (gdr_score_fused, two handles) must return exactly what
    gdr_score_topk returns batch by batch."""
    _run("FUSED", groups)


@pytest.mark.xfail(strict=False, reason="the priority launch attribute was wired after the last GPU session; bench.py's autotune is its first run")
def test_launch_priorities_do_not_change_results():
    """GDR_LAUNCH_PRIORITIES only attaches cudaLaunchAttributePriority to the launches (bench.py's autotune may switch it on)."""
    _run("GDR_LAUNCH_PRIORITIES", "1")
