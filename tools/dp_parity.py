#!/usr/bin/env python
"""Data-parallel parity of the reference's train step, run under torchrun with 2+ ranks:

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_parity.py [--domain fluid] [--batch 4]

Every rank runs the reference's unmodified step on its shard of the batch with `refstep.DataParallel`
(gradient buckets all-reduced in optimiser pre-step hooks, branch flag agreed, SyncBatchNorm in the
discriminators); rank 0 then runs the SAME step on the whole batch in one process and compares the losses and
the gradients every optimiser saw.  Dropout is switched off and every cloud drops the same number of points
(no (999,999,999) dummies, hence no per-rank numpy re-draw), so both runs are deterministic functions of
the same inputs.  Exit code 0 and a JSON line with the max relative errors when the losses agree to `--loss-rtol`
(1e-4; measured 1e-5) and the gradients to `--rtol` of their largest entry (1e-2; measured 1e-4 .. 8e-3: the
discriminators end in BatchNorm1d over a batch of 4, whose backward amplifies fp32 summation-order differences
between one batch of 4 and two synchronised batches of 2 — a plumbing error, e.g. a wrong average, shows as O(1)).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import refstep  # noqa: E402


def no_dropout(ctx):
    for net in ctx.networks():
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0


def grab_grads(ctx, store):
    for name, net, optim in zip(("G", "tempoD", "spatialD"), ctx.networks(), ctx.optims):
        params = [p for p in net.parameters() if p.requires_grad]

        def hook(_o, _a, _k, name=name, params=params):
            store[name] = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params]).clone()

        optim.register_step_pre_hook(hook)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--domain", default="fluid")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--n-lo", type=int, default=256)
    ap.add_argument("--ratio", type=int, default=4)
    ap.add_argument("--rtol", type=float, default=1e-2, help="max |grad difference| / max |grad|")
    ap.add_argument("--loss-rtol", type=float, default=1e-4)
    a = ap.parse_args()
    # cuDNN runs fp32 convolutions on TF32 tensor cores by default, with a batch-size dependent algorithm choice: the
    # 1e-4 differences that causes between a batch of 4 and two batches of 2 are amplified by the small-batch
    # BatchNorm layers.  Exact fp32 here, so that what is compared is the data-parallel plumbing.
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert a.batch % world == 0
    per = a.batch // world

    def build(B):
        ctx = refstep.build(a.domain, B=B, n_lo=a.n_lo, ratio=a.ratio, backend="cuda", device=dev, seed=1,
                            masked_frac=(0.03, 0.03))
        no_dropout(ctx)
        return ctx

    # one set of frames / weights for everybody: the full batch, built identically on every rank
    full = build(a.batch)
    state = [{k: v.detach().clone() for k, v in net.state_dict().items()} for net in full.networks()]
    # ---- sharded run ----
    shard = build(per)
    for net, st in zip(shard.networks(), state):
        net.load_state_dict(st)
    lo_, hi_ = rank * per, (rank + 1) * per
    shard.lo = [t[lo_:hi_].contiguous() for t in full.lo]
    shard.hi = [t[lo_:hi_].contiguous() for t in full.hi]
    shard.keep = full.keep[lo_:hi_].contiguous() if full.keep is not None else None
    refstep.DataParallel(shard, sync_bn=True)
    got = {}
    grab_grads(shard, got)
    n_sync = sum(isinstance(m, torch.nn.SyncBatchNorm) for net in shard.networks() for m in net.modules())
    n_bn = sum(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and not isinstance(m, torch.nn.SyncBatchNorm)
               for net in shard.networks() for m in net.modules())
    # forward-only probe of both discriminators on real frames (train mode: batch statistics)
    acts_dp = []

    def rec_hooks(net, acts):
        hs = []
        for name, m in net.named_modules():
            if len(list(m.children())) == 0:
                hs.append(m.register_forward_hook(lambda mod, i, o, name=name, acts=acts: acts.append((name, type(mod).__name__, o.detach().clone() if torch.is_tensor(o) else None))))
        return hs

    hs = rec_hooks(shard.spatial_dis, acts_dp)
    with torch.no_grad():
        probe_dp = [shard.spatial_dis(shard.hi[1]).flatten().clone(), shard.tempo_dis(list(shard.hi), 0.1).flatten().clone()]
    for h in hs:
        h.remove()
    for net, st in zip(shard.networks(), state):  # the probe advanced spectral-norm / BN buffers: restore
        net.load_state_dict(st)
    np.random.seed(7)
    torch.manual_seed(7)
    losses_dp = refstep.step(shard, 12)
    torch.cuda.synchronize()
    loss_t = torch.tensor([losses_dp[k] for k in sorted(losses_dp)], dtype=torch.float64, device=dev)
    dist.all_reduce(loss_t, op=dist.ReduceOp.SUM)
    loss_t /= world
    rc = 0
    if rank == 0:
        # ---- single-process run on the whole batch (restore the un-wrapped loss first) ----
        mods = refstep.import_reference("cuda")
        ref = refstep.build(a.domain, B=a.batch, n_lo=a.n_lo, ratio=a.ratio, backend="cuda", device=dev, seed=1,
                            masked_frac=(0.03, 0.03), mods=mods)
        no_dropout(ref)
        for net, st in zip(ref.networks(), state):
            net.load_state_dict(st)
        ref.lo, ref.hi, ref.keep = full.lo, full.hi, full.keep
        want = {}
        grab_grads(ref, want)
        acts_ref = []
        hs = rec_hooks(ref.spatial_dis, acts_ref)
        with torch.no_grad():
            probe_ref = [ref.spatial_dis(ref.hi[1]).flatten().clone(), ref.tempo_dis(list(ref.hi), 0.1).flatten().clone()]
        for h in hs:
            h.remove()
        for (n1, t1, a1), (n2, t2, a2) in zip(acts_dp, acts_ref):
            if a1 is None or a2 is None:
                continue
            a2 = a2[:a1.shape[0]]
            err = float((a1 - a2).abs().max() / a2.abs().max().clamp_min(1e-30)) if a1.shape == a2.shape else -1.0
            if err > 1e-4:
                print(f"layer {n1:45s} {t1:18s} {tuple(a1.shape)} rel err {err:.3e}", file=sys.stderr)
        for net, st in zip(ref.networks(), state):
            net.load_state_dict(st)
        print("probe spatial: dp", probe_dp[0].tolist(), "single", probe_ref[0][:per].tolist(), file=sys.stderr)
        print("probe tempo: dp", probe_dp[1].tolist(), "single", probe_ref[1][:per].tolist(), file=sys.stderr)
        print("sync bn layers", n_sync, "plain bn layers", n_bn, file=sys.stderr)
        np.random.seed(7)
        torch.manual_seed(7)
        losses_ref = refstep.step(ref, 12)
        torch.cuda.synchronize()
        out = {"world": world, "batch": a.batch, "losses_single": losses_ref,
               "losses_dp_mean": dict(zip(sorted(losses_dp), loss_t.tolist())), "grad_rel_err": {}}
        for k in ("G", "tempoD", "spatialD"):
            w, g = want[k].double(), got[k].double()
            err = float((w - g).abs().max() / w.abs().max().clamp_min(1e-30))
            out["grad_rel_err"][k] = err
            out["grad_norm_" + k] = float(w.norm())
            if not err <= a.rtol:
                rc = 1
        for k, v in losses_ref.items():
            d = out["losses_dp_mean"][k]
            if abs(d - v) > a.loss_rtol * max(abs(v), 1e-6):
                rc = 1
        out["ok"] = rc == 0
        print(json.dumps(out), flush=True)
    flag = torch.tensor([rc], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
