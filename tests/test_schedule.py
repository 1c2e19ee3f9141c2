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
    import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "fluid_step_schedule.json"))
    c = doc["counts"]
    assert (c["knn"], c["group"], c["fps"], c["gather"], c["ball_query"], c["frnn"], c["chamfer"]) == \
        (42, 105, 27, 27, 27, 11, 1)
    assert c["chamfer_bwd"] == 1 and c["group_bwd"] > 0
    assert doc["n_lo"] == 2048 and doc["n_hi"] == 8192


def test_action_schedule_counts():
    import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "action_step_schedule.json"))
    c = doc["counts"]
    assert (c["knn"], c["group"], c["fps"], c["gather"], c["ball_query"], c["frnn"], c["chamfer"]) == \
        (36, 99, 27, 27, 27, 9, 1)


def test_batch_rescale_and_bytes():
    import hotpath_trace as ht

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
    import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, "action_step_schedule.json"), 1)
    rp = ht.TraceReplay(doc, bench.OracleOps(), seed=3)
    loss = rp.run_step()
    assert np.isfinite(loss) and loss > 0
    assert rp.queries == ht.count_queries(doc)


def test_dependency_tracker_follows_weights_views_and_autograd():
    """oracle.shims.DepTracker (used by tests/golden/make_schedule.py): the producer sets recorded with a
    call must follow data through optimizer updates, writes through views and autograd."""
    import torch

    import oracle.shims as sh

    sh.activate()
    from oracle.shims import _backend as B

    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(4, 3))
    opt = torch.optim.Adam([w], lr=0.1)
    x = torch.randn(2, 64, 3)
    sh.recorder.start(shapes_only=True, track_deps=True)
    try:
        f = torch.einsum("bnc,dc->bdn", x, w).contiguous()
        _, i = B.knn(x, x, 4)                                    # call 0
        g = B.Grouping.apply(f, i.to(torch.int32))               # call 1
        g.sum().backward()                                       # call 2: group_bwd
        opt.step()
        f2 = torch.einsum("bnc,dc->bdn", x, w).contiguous()      # reads the updated weight
        g2 = B.Grouping.apply(f2, i.to(torch.int32))             # call 3
        v = f2.view(2, -1)
        v[:, 0] = g2.reshape(2, -1)[:, 0]                        # write through a view
        B.knn(f2.transpose(1, 2)[..., :3].contiguous(), x, 2)    # call 4
    finally:
        calls = sh.recorder.stop()
    deps = [c[1]["deps"] for c in calls]
    assert [c[0] for c in calls] == ["knn", "group", "group_bwd", "group", "knn"]
    assert deps[0] == {"p1": [], "p2": []}
    assert deps[1] == {"f": [], "idx": [0]}
    assert deps[2]["grad_out"] == [1] and deps[2]["idx"] == [0]
    assert deps[3] == {"f": [2], "idx": [0]}          # through the Adam update of w
    assert deps[4] == {"p1": [2, 3], "p2": []}       # through the view write


@pytest.mark.parametrize("name", ["fluid", "action"])
def test_schedule_dependencies_form_a_dag_with_parallel_chains(name):
    import hotpath_trace as ht

    doc = ht.load_schedule(os.path.join(GOLDEN, f"{name}_step_schedule.json"))
    calls = doc["calls"]
    prod_of_idx = {"knn", "frnn", "ball_query"}
    for n, c in enumerate(calls):
        d = c["in"]["deps"]
        assert all(0 <= x < n for v in d.values() for x in v)
        if c["op"] == "group":   # the neighbour lists come from exactly one search call (or FRNN + kNN fill)
            assert d["idx"] and all(calls[x]["op"] in prod_of_idx for x in d["idx"])
        if c["op"] == "gather":
            assert [calls[x]["op"] for x in d["idx"]] == ["fps"]
        if c["op"] == "ball_query":
            assert [calls[x]["op"] for x in d["new_xyz"]] == ["gather"]
    # the generator passes over frames 0 and 2 start from external inputs only: independent chains exist
    roots = [n for n, c in enumerate(calls) if not any(c["in"]["deps"].values())]
    assert len(roots) >= 3
    rp = ht.TraceReplay.__new__(ht.TraceReplay)
    rp.calls = calls
    for lanes in (2, 8, 32):
        plan = rp.plan_lanes(lanes)
        assert len(plan) == len(calls) and 0 <= min(plan) and max(plan) < lanes
