"""Record the boundary-call schedule of the reference's UNMODIFIED GAN train steps.

Runs in the build container only (needs /root/reference).  The reference's
`tempo_gan_step` / `tempo_gan_step_no_mask` (train_step_final.py:69-320) is executed on CPU
over the oracle-backed shims (oracle/shims) with the recorder in shapes-only mode; the list
of (op, shapes, parameters, forward/backward links) is written as JSON.  bench.py replays
exactly this schedule on synthetic data (with B rescaled), and tests/test_schedule.py checks
its call counts against SURVEY.md §3.1.

    python tests/golden/make_schedule.py fluid 2 2048 4     -> fluid_step_schedule.json
    python tests/golden/make_schedule.py action 2 128 16    -> action_step_schedule.json

Deviations from a production run, both shape-neutral: CUDA placement calls are no-ops on
this GPU-less box, and the masking loss is scaled to zero *after* it is computed so that the
`ml < 0.1` gate (train_step_final.py:117) takes the GAN branch with an untrained generator.
"""
import json
import os
import sys
import time
import warnings
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle.shims as sh  # noqa: E402

sh.activate()
sys.path.insert(1, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200", "import_stubs"))
sys.path.insert(1, "/root/reference")
warnings.simplefilter("ignore")
torch.Tensor.cuda = lambda self, *a, **k: self  # train_step_final.py:30,156, loss.py:174

import synth  # noqa: E402
import train_step_final  # noqa: E402
from discriminator import ActionSpatialDis, ActionTempoDis, FluidSpatialDis, FluidTempoDis  # noqa: E402
from upsampling_network import NoMaskSRNet, SRNet  # noqa: E402


def main():
    domain, B, n_lo, ratio = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    torch.manual_seed(1)
    np.random.seed(1)
    rng = np.random.default_rng(1)
    n_hi = n_lo * ratio
    if domain == "fluid":
        hi = [torch.from_numpy(synth.fluid_cloud(rng, B, n_hi)) for _ in range(3)]
        # low-res = every ratio-th point + N(0, 0.003^2) jitter (tempo_dataset.py:27,92)
        lo = [h[:, ::ratio].contiguous() + 0.003 * torch.randn(B, n_lo, 3) for h in hi]
        g, sd, td = SRNet(3, 128, upsample_ratio=ratio), FluidSpatialDis(), FluidTempoDis(3)
        opt = Namespace(use_vel=False, in_node_feats=3, R=0.10, cutoff=0.025, w=0.5)
        orig = train_step_final.tpugan_sr_loss

        def gated(*a, **k):
            pl, cd, ml = orig(*a, **k)
            return cd + 0 * ml, cd, ml * 0  # take the GAN branch (see module docstring)

        train_step_final.tpugan_sr_loss = gated
    else:
        hi = [torch.from_numpy(synth.action_cloud(rng, B, n_hi)) for _ in range(3)]
        lo = [h[:, ::ratio].contiguous() for h in hi]
        g, sd, td = NoMaskSRNet(3, 128, ratio), ActionSpatialDis(), ActionTempoDis(3)
        opt = Namespace(R=2.0, w=2.0)
    og, ot, os_ = (torch.optim.Adam(m.parameters(), lr=1e-4) for m in (g, td, sd))
    sh.recorder.start(shapes_only=True, track_deps=True)
    t = time.time()
    if domain == "fluid":
        out = train_step_final.tempo_gan_step(g, sd, td, lo, None, hi, None, 1.0, opt, 12, og, ot, os_)
    else:
        out = train_step_final.tempo_gan_step_no_mask(g, sd, td, lo, hi, opt, 12, og, ot, os_)
    calls = sh.recorder.stop()
    print("step took %.1fs on CPU:" % (time.time() - t), out)
    counts = {}
    for op, _, _ in calls:
        counts[op] = counts.get(op, 0) + 1
    print(counts)
    doc = {
        "source": "reference train_step_final.%s, unmodified, over oracle shims" % (
            "tempo_gan_step" if domain == "fluid" else "tempo_gan_step_no_mask"),
        "domain": domain, "B": B, "n_lo": n_lo, "ratio": ratio, "n_hi": n_hi, "n_iter": 12,
        "counts": counts,
        "calls": [{"op": op, "in": i, "out": o} for op, i, o in calls],
    }
    path = os.path.join(HERE, "%s_step_schedule.json" % domain)
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
