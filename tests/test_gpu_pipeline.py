"""The pipelined schedules and launch options must reproduce the plain gdr_score_topk call BIT FOR BIT (include/gdr_b200.h:
"a different schedule, not a different result").  bench.py's headline number is measured through PipelinedRetriever, so these are
hard tests: every variant runs in a child process with a timeout (a variant that hangs costs its own 90 s, not the session) and a
mismatch turns the suite red."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(mode, value):
    out = subprocess.run([sys.executable, os.path.join(HERE, "_variant_child.py"), mode, value], capture_output=True, text=True, timeout=90)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"], line
    return line


@pytest.mark.parametrize("groups", ["5", "9"])
def test_fused_score_topk_equals_default(groups):
    """ONE launch scoring batch i and selecting the top-k of batch i-1 (gdr_score_fused, two handles; prob + tanh + alpha) must
    return exactly what gdr_score_topk returns batch by batch — five 128-thread groups running the lean select (default) and nine 64-thread groups."""
    _run("FUSED", groups)


@pytest.mark.parametrize("schedule,groups", [("auto", "5"), ("auto", "9"), ("batches", "5")])
def test_pipelined_retriever_equals_default(schedule, groups):
    """gdr_b200.PipelinedRetriever: eager, replayed from a CUDA graph, and through pinned host buffers."""
    _run("PIPELINE_" + schedule.upper(), groups)


@pytest.mark.parametrize("ctas_per_sm", ["2", "1"])
def test_partitioned_schedule_equals_default(ctas_per_sm):
    """SM partition (CUDA green contexts; gdr_partition_*, PipelinedRetriever(schedule="partitioned")): inversion and top-k on one SM
    set, scoring on the other with two 4-stage CTAs per SM (k_score_umma_x2) or one 6-stage CTA — eager, replayed from a CUDA graph,
    and through pinned host buffers."""
    line = _run("PIPELINE_PARTITIONED", ctas_per_sm)
    if line.get("skipped"):
        pytest.skip(line["skipped"])


@pytest.mark.parametrize("groups", ["4", "1"])
def test_grouped_topk_equals_default(groups):
    _run("topk_groups", groups)


def test_two_scoring_ctas_per_sm_equal_default():
    """GDR_OPT_UMMA_CTAS_PER_SM = 2 (k_score_umma_x2: 4-stage ring, two CTAs per SM) on the whole device: same bits as the default kernel."""
    _run("umma_ctas_per_sm", "2")


def test_launch_priorities_do_not_change_results():
    """GDR_OPT_LAUNCH_PRIORITIES only attaches cudaLaunchAttributePriority to the launches."""
    _run("launch_priorities", "1")
