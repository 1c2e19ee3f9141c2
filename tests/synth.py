"""Seeded synthetic clouds shared by tests, smoke() and bench.py."""
import numpy as np

BASE_RADIUS = 0.025  # train_utils.py:10 — SPH particle spacing


def fluid_cloud(rng, B, N, D=3):
    """Uniform i.i.d. points in a cube of side 0.025*N^(1/3), centroid-centred per cloud
    (train_utils.py:214-221) — the density of the reference's fluid frames."""
    L = BASE_RADIUS * (N ** (1.0 / 3.0))
    p = rng.uniform(0.0, L, size=(B, N, D)).astype(np.float32)
    p -= p.mean(axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(p, dtype=np.float32)


def action_cloud(rng, B, N):
    """MSR-Action-like box [-.5,.5]x[-1,1]x[-.25,.25] (msr_dataset.py:81-84)."""
    lo = np.array([-0.5, -1.0, -0.25], np.float32)
    hi = -lo
    return (rng.uniform(size=(B, N, 3)).astype(np.float32) * (hi - lo) + lo).astype(np.float32)


def with_duplicates(rng, p, frac=0.25):
    """Exact duplicates, as produced by masked generator outputs (upsampling_network.py:136-138)
    and by MSR clips that repeat points (msr_dataset.py:72-74)."""
    p = p.copy()
    B, N, _ = p.shape
    n = max(1, int(N * frac))
    for b in range(B):
        dst = rng.choice(N, size=n, replace=False)
        src = rng.choice(N, size=n, replace=True)
        p[b, dst] = p[b, src]
    return p


def with_dummies(rng, p, frac=0.3, value=999.0):
    """A block of (999,999,999) padding points (upsampling_network.py:149)."""
    p = p.copy()
    B, N, _ = p.shape
    n = max(1, int(N * frac))
    for b in range(B):
        p[b, rng.choice(N, size=n, replace=False)] = value
    return p


def lattice_cloud(B, side, spacing=0.025):
    """Regular lattice: massive distance ties (every point has 6 equidistant neighbours)."""
    g = np.arange(side, dtype=np.float32) * np.float32(spacing)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    p = np.stack([x.ravel(), y.ravel(), z.ravel()], -1).astype(np.float32)
    return np.ascontiguousarray(np.broadcast_to(p, (B,) + p.shape))
