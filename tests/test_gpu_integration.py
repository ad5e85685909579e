"""INTEGRATION.md's edits applied to the live reference (the unmodified sources under baseline/_ref/, shipped to the GPU box by
__graft_entry__.build(); /root/reference in the dev container) and run against the untouched reference on the same inputs:
the vendored beam search with `TreeMask` in place of the Python mask block, and `validation_step_i` with `FineStage` in place of
its gather / score / top-k code (tests/_integration_child.py)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _reference_present():
    return any(os.path.isfile(os.path.join(r, "GDR_model", "main_models.py")) for r in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")))


def _run(extra):
    out = subprocess.run([sys.executable, os.path.join(HERE, "_integration_child.py")] + extra, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"], line
    return line


@pytest.mark.gpu
@pytest.mark.skipif(not _reference_present(), reason="no reference tree (run __graft_entry__.build() in the dev container to ship baseline/_ref/)")
def test_integration_edits_against_the_live_reference():
    line = _run([])
    assert line["beam"]["identical_beams"] and line["beam"]["valid_clusters"] == line["beam"]["rows"]
    assert all(v["identical_inf_index_batch"] for v in line["fine"].values())


@pytest.mark.skipif(not _reference_present(), reason="no reference tree")
def test_integration_patching_machinery_cpu_dry_run():
    """The same child with the CUDA kernels replaced by the oracle: the in-memory edits themselves apply and run (dev container, no GPU)."""
    _run(["--cpu-dry-run"])
