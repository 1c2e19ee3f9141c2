"""The recorded boundary-call schedules (tests/golden/*_step_schedule.json) match the call
counts of SURVEY.md §3.1 and replay end to end on the CPU oracle back-end."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def test_fluid_schedule_counts():
    from tpugan_b200 import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "fluid_step_schedule.json"))
    c = doc["counts"]
    assert (c["knn"], c["group"], c["fps"], c["gather"], c["ball_query"], c["frnn"], c["chamfer"]) == \
        (42, 105, 27, 27, 27, 11, 1)
    assert c["chamfer_bwd"] == 1 and c["group_bwd"] > 0
    assert doc["n_lo"] == 2048 and doc["n_hi"] == 8192


def test_action_schedule_counts():
    from tpugan_b200 import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "action_step_schedule.json"))
    c = doc["counts"]
    assert (c["knn"], c["group"], c["fps"], c["gather"], c["ball_query"], c["frnn"], c["chamfer"]) == \
        (36, 99, 27, 27, 27, 9, 1)


def test_batch_rescale_and_bytes():
    from tpugan_b200 import hotpath_trace as ht

    d2 = ht.load_schedule(os.path.join(GOLDEN, "fluid_step_schedule.json"))
    d8 = ht.load_schedule(os.path.join(GOLDEN, "fluid_step_schedule.json"), 8)
    assert ht.count_queries(d8) == 4 * ht.count_queries(d2)
    b2 = sum(ht.algorithmic_bytes(c) for c in d2["calls"])
    b8 = sum(ht.algorithmic_bytes(c) for c in d8["calls"])
    assert b8 == 4 * b2
    # SURVEY.md §8d worked number: grouping [8,64,2048] x [8,2048,12] = 55.3 MB
    call = {"op": "group", "in": {"f": {"shape": [8, 64, 2048]}, "idx": {"shape": [8, 2048, 12]}}, "out": {}}
    assert abs(ht.algorithmic_bytes(call) - 55.3e6) < 0.1e6


def test_replay_runs_on_the_oracle_backend():
    """Small-shape replay of a truncated schedule through bench.OracleOps (CPU)."""
    import bench
    from tpugan_b200 import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "action_step_schedule.json"), 1)
    rp = ht.TraceReplay(doc, bench.OracleOps(), seed=3)
    loss = rp.run_step()
    assert np.isfinite(loss) and loss > 0
    assert rp.queries == ht.count_queries(doc)
