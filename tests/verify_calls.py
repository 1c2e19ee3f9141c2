"""Check a boundary-call log (tpugan_b200.recording) against the CPU oracle, call by call, on
each call's OWN recorded inputs.  Indices bit-exact; canonical distances bit-exact; gathers exact;
gradients / sums / interpolated values within 1e-5 relative (BASELINE.json north_star tolerances).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-5


def _np(t):
    return None if t is None else (t.detach().cpu().numpy() if hasattr(t, "detach") else t)


def _close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    # per element 1e-5 relative, with an absolute floor of 1e-5 x the tensor's magnitude: gradients are sums
    # of many terms that cancel, so elements near zero carry the rounding of much larger addends
    assert np.allclose(a, b, rtol=rtol, atol=rtol * scale), f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def check_call(oracle, c) -> str:
    """Verify one recorded call; returns a short tag of what was compared."""
    i = {k: _np(v) for k, v in c.inputs.items()}
    o = {k: _np(v) for k, v in c.outputs.items()}
    op = c.op
    if op == "knn":
        d, idx = oracle.knn(i["p1"], i["p2"], i["K"], i.get("lengths1"), i.get("lengths2"))
        assert np.array_equal(o["idx"], idx), "knn idx"
        assert np.array_equal(o["dists"], d), "knn dists (canonical, bit-exact)"
        return "idx+dists exact"
    if op == "frnn":
        d, idx = oracle.frnn(i["p1"], i["p2"], i["K"], i["r"], i.get("lengths1"), i.get("lengths2"))
        assert np.array_equal(o["idx"], idx), "frnn idx"
        assert np.array_equal(o["dists"], d), "frnn dists"
        return "idx+dists exact"
    if op == "ball_query":
        idx = oracle.ball_query(i["radius"], i["nsample"], i["xyz"], i["new_xyz"])
        assert np.array_equal(o["idx"], idx), "ball_query idx"
        return "idx exact"
    if op == "fps":
        idx = oracle.fps(i["xyz"], i["npoint"])
        assert np.array_equal(o["idx"], idx), "fps idx"
        return "idx exact"
    if op == "fps_start":
        want_rows = o.get("rows") is not None
        r = oracle.fps_start(i["pts"], i["k"], i["start"], return_rows=want_rows)
        idx = r[0] if want_rows else r
        assert np.array_equal(o["idx"], idx), "fps_start idx"
        if want_rows:
            assert np.array_equal(o["rows"], r[1]), "fps_start rows"
        return "idx exact"
    if op in ("group", "gather"):
        out = oracle.group_fwd(i["f"], i["idx"], i.get("center"))
        assert np.array_equal(o["out"], out), f"{op} values"
        return "values exact"
    if op in ("group_bwd", "gather_bwd"):
        idx = i["idx"] if i["idx"].ndim == 3 else i["idx"][:, :, None]
        g = i["grad_out"]
        B, C = g.shape[:2]
        ref = oracle.group_bwd(g.reshape(B, C, idx.shape[1], idx.shape[2]), idx, i["N"])
        _close(o["grad_f"], ref, op)
        return "grad 1e-5" + (" (bit-equal)" if np.array_equal(o["grad_f"], ref) else "")
    if op == "group_reduce":
        out, arg = oracle.group_reduce_fwd(i["f"], i["idx"], i["op"])
        assert np.array_equal(o["out"], out), "group_reduce values"
        if o.get("arg") is not None:
            assert np.array_equal(o["arg"], arg), "group_reduce arg"
        return "values+arg exact"
    if op == "group_reduce_bwd":
        ref = oracle.group_reduce_bwd(i["grad_out"], i["idx"], i["arg"], i["N"], i["op"])
        _close(o["grad_f"], ref, op)
        return "grad 1e-5"
    if op == "group_assemble":
        parts = list(zip(i["modes"], [_np(t) for t in c.inputs["srcs"]], [_np(t) for t in c.inputs["centers"]]))
        assert np.array_equal(o["out"], oracle.group_assemble(parts, i["idx"])), "group_assemble values"
        return "values exact"
    if op == "edge_affine":
        assert np.array_equal(o["out"], oracle.edge_affine_fwd(i["p"], i["q"], i["center"], i["idx"], i["slope"])), op
        return "values exact"
    if op == "edge_affine_bwd":
        g2, gc = oracle.edge_affine_bwd(i["grad_out"], i["q"], i["center"], i["idx"], i["slope"])
        assert np.array_equal(o["g2"], g2), "edge_affine_bwd g2"
        if o.get("grad_center") is not None:
            assert np.array_equal(o["grad_center"], gc), "edge_affine_bwd grad_center"
        return "values exact"
    if op == "three_nn":
        d, idx = oracle.three_nn(i["unknown"], i["known"])
        assert np.array_equal(o["idx"], idx), "three_nn idx"
        _close(o["dist"], d, "three_nn dist")
        return "idx exact"
    if op == "three_interpolate":
        _close(o["out"], oracle.three_interpolate_fwd(i["f"], i["idx"], i["w"]), op)
        return "values 1e-5"
    if op == "three_interpolate_bwd":
        _close(o["grad_f"], oracle.three_interpolate_bwd(i["grad_out"], i["idx"], i["w"], i["m"]), op)
        return "grad 1e-5"
    if op == "chamfer":
        r = oracle.chamfer_fwd(i["src"], i["tgt"], i["directions"])
        for k in ("i_src", "i_tgt"):
            if o.get(k) is not None:
                assert np.array_equal(o[k], r[k]), f"chamfer {k}"
        for k in ("d_src", "d_tgt"):
            if o.get(k) is not None:
                assert np.array_equal(o[k], r[k]), f"chamfer {k} (canonical distances)"
        for k in ("sum_src", "sum_tgt"):
            if o.get(k) is not None:
                _close(o[k], r[k], f"chamfer {k}")
        return "idx exact, sums 1e-5"
    if op == "chamfer_bwd":
        gs, gt = oracle.chamfer_bwd(i["src"], i["tgt"], i["i_src"], i["i_tgt"], i["g_src"], i["g_tgt"], i["directions"])
        if o.get("grad_src") is not None:
            _close(o["grad_src"], gs, "chamfer grad_src")
        if o.get("grad_tgt") is not None:
            _close(o["grad_tgt"], gt, "chamfer grad_tgt")
        return "grads 1e-5"
    if op == "cubic_interp":
        _close(o["out"], oracle.cubic_interp(i["query"], i["field"], i["pos"], i["cutoff"]), op)
        return "values 1e-5"
    if op == "gather_rows":
        assert np.array_equal(o["out"], oracle.gather_rows(i["x"], i["idx"])), "gather_rows"
        return "values exact"
    if op == "gather_rows_bwd":
        go, idx = i["grad_out"], i["idx"]
        ref = np.zeros((go.shape[0], i["N"], go.shape[2]), np.float64)
        for b in range(go.shape[0]):
            v = idx[b] >= 0
            np.add.at(ref[b], idx[b][v], go[b][v].astype(np.float64))
        _close(o["grad_x"], ref, op)
        return "grad 1e-5"
    if op == "knn_bwd":
        p1, p2, idx, g = (i[k].astype(np.float64) if k != "idx" else i[k] for k in ("p1", "p2", "idx", "grad_dists"))
        B, P1, K = idx.shape
        g = g.copy()
        g[idx < 0] = 0
        for b in range(B):
            kv = K if i.get("lengths2") is None else min(K, int(i["lengths2"][b]))
            g[b, :, kv:] = 0
            if i.get("lengths1") is not None:
                g[b, int(i["lengths1"][b]):] = 0
        nb = p2[np.arange(B)[:, None, None], np.maximum(idx, 0)]
        diff = 2.0 * g[..., None] * (p1[:, :, None, :] - nb)
        if o.get("grad_p1") is not None:
            _close(o["grad_p1"], diff.sum(2), "knn_bwd grad_p1")
        if o.get("grad_p2") is not None:
            ref = np.zeros_like(p2)
            for b in range(B):
                np.add.at(ref[b], np.maximum(idx[b], 0).reshape(-1), -diff[b].reshape(P1 * K, -1))
            _close(o["grad_p2"], ref, "knn_bwd grad_p2")
        return "grads 1e-5"
    raise AssertionError(f"no checker for op {op!r}")


def check_log(oracle, calls, max_calls=None):
    """Verify every call (or the first `max_calls` of each op); returns {op: (count, tag)}."""
    seen = {}
    for n, c in enumerate(calls):
        k = seen.get(c.op, (0, ""))[0]
        if max_calls is not None and k >= max_calls:
            seen[c.op] = (k + 1, seen[c.op][1])
            continue
        try:
            tag = check_call(oracle, c)
        except AssertionError as e:
            raise AssertionError(f"boundary call #{n} ({c.op}) differs from the oracle on its recorded inputs: {e}") from e
        seen[c.op] = (k + 1, tag)
    return seen


def counts(calls):
    out = {}
    for c in calls:
        out[c.op] = out.get(c.op, 0) + 1
    return out


def grouping_equivalents(calls):
    """Number of `grouping_operation` calls the log stands for: plain groupings plus the gather parts of every
    one-pass assembly (QueryAndGroup = 2 groupings, FlowEmbedding = 2 groupings + a repeat)."""
    n = 0
    for c in calls:
        if c.op == "group":
            n += 1
        elif c.op == "group_assemble":
            n += sum(1 for m in c.inputs["modes"] if m == "gather")
    return n
