"""world_size-2 gloo tests of the N>1 host logic (tpugan_b200/sharding.py): batch shards +
reductions reproduce the single-process result of the path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        import synth
        from tpugan_b200 import sharding

        rng = np.random.default_rng(7)  # same data on every rank, then sharded
        B = 5  # uneven split: 3 + 2
        gt = synth.fluid_cloud(rng, B, 256)
        pred = (gt[:, ::2] + rng.normal(0, 0.003, size=(B, 128, 3))).astype(np.float32)
        tg, tp = sharding.shard_batch([torch.from_numpy(gt), torch.from_numpy(pred)], world, rank)
        lo, hi = sharding.shard_range(B, world, rank)
        assert tg.shape[0] == hi - lo
        # per-cloud ops are shard-local: indices equal the corresponding slice of the full result
        _, idx_local = oracle.knn(tp.numpy(), tg.numpy(), 8)
        _, idx_full = oracle.knn(pred, gt, 8)
        assert np.array_equal(idx_local, idx_full[lo:hi])
        # Chamfer loss: local batch mean, then count-weighted mean over ranks == global batch mean
        local = torch.tensor(float(oracle.chamfer_distance(tg.numpy(), tp.numpy(), bidirectional=True)))
        glob = sharding.reduce_mean_(local.clone(), weight=float(hi - lo))
        full = float(oracle.chamfer_distance(gt, pred, bidirectional=True))
        assert abs(float(glob) - full) <= 1e-5 * abs(full), (float(glob), full)
        # branch flag agreement
        flag = sharding.agree_any(torch.tensor([1.0 if rank == 1 else 0.0]))
        assert float(flag) == 1.0
        # gradient buckets are averaged
        b = [torch.full((10,), float(rank + 1)), torch.full((3,), 2.0 * rank)]
        sharding.allreduce_buckets_(b)
        assert torch.allclose(b[0], torch.full((10,), 1.5)) and torch.allclose(b[1], torch.full((3,), 1.0))
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_path_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_range_partitions():
    from tpugan_b200.sharding import shard_range

    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


# ---- the data-parallel harness of the reference's train step (tools/refstep.py::DataParallel) over gloo ---------------
def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import refstep

        torch.set_num_threads(2)
        # different data per rank (seed), identical initial weights after the harness's broadcast
        ctx = refstep.build("action", B=2, n_lo=64, ratio=16, backend="oracle", device="cpu", seed=1 + rank)
        dp = refstep.DataParallel(ctx, sync_bn=False)
        seen = {}
        for name, net, optim in zip(("G", "tempoD", "spatialD"), ctx.networks(), ctx.optims):
            params = [p for p in net.parameters() if p.requires_grad]
            optim.register_step_pre_hook(lambda _o, _a, _k, name=name, params=params: seen.__setitem__(
                name, torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None]).clone()))
        np.random.seed(3)
        torch.manual_seed(3)
        losses = refstep.step(ctx, 12)
        assert all(np.isfinite(v) for v in losses.values())
        assert dp.bytes_per_step == 4 * sum(refstep.param_counts(ctx)), (dp.bytes_per_step, refstep.param_counts(ctx))
        # every rank stepped its optimisers on the SAME (averaged) gradients and ends with the SAME weights
        for name in ("G", "tempoD", "spatialD"):
            g = seen[name]
            gathered = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(gathered, g)
            assert all(torch.equal(gathered[0], x) for x in gathered), name
        w = torch.cat([p.detach().reshape(-1) for p in ctx.sr_net.parameters()])
        gathered = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        assert torch.equal(gathered[0], gathered[1])
        q.put((rank, "ok"))
    except Exception as e:
        import traceback

        q.put((rank, repr(e) + traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not (os.path.exists(os.path.join(ROOT, "baseline", "_ref", "train_step_final.py")) or
                         os.path.isdir("/root/reference")), reason="reference not installed")
def test_data_parallel_reference_step_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
