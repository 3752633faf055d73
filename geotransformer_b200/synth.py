"""Deterministic synthetic registration pairs (SURVEY.md section 8d / BASELINE.md section 3).

There is no dataset on the GPU box; every test, the smoke run and bench.py draw their clouds here.
Coordinates are continuous (no quantisation) so exact-distance ties -- the one thing the reference's
``radius_neighbors`` orders unreproducibly -- do not occur, and index outputs are bit-comparable.

Output dict mirrors what the reference datasets hand to the collate function
(reference ``geotransformer/datasets/registration/threedmatch/dataset.py:131-135``):
``ref_points, src_points (n,3) f32``, ``ref_feats, src_feats (n,1) f32 ones``, ``transform (4,4) f32`` with
``ref = R @ src + t``.
"""
import numpy as np


def _room(rng, n, L, sigma, off):
    face = rng.integers(0, 3, size=n)
    uv = rng.uniform(0.0, L, size=(n, 2))
    pts = np.zeros((n, 3), dtype=np.float64)
    for f in range(3):
        m = face == f
        axes = [a for a in range(3) if a != f]
        pts[m, axes[0]] = uv[m, 0]
        pts[m, axes[1]] = uv[m, 1]
    pts += rng.normal(0.0, sigma, size=(n, 3))
    return pts + off


def _sphere(rng, n, R, sigma):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * R + rng.normal(0.0, sigma, size=(n, 3))


def _ground(rng, n, R, sigma):
    r = R * np.sqrt(rng.uniform(size=n))
    th = rng.uniform(0.0, 2 * np.pi, size=n)
    z = rng.normal(0.0, sigma, size=n)
    k = n // 5
    z[:k] = rng.uniform(0.0, 3.0, size=k)
    return np.stack([r * np.cos(th), r * np.sin(th), z], axis=1)


def _rodrigues(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


# name -> (model config, generator, kwargs, translation scale)
WORKLOADS = {
    'demo2k': ('3dmatch', _room, dict(n=2048, L=0.90, sigma=0.003, off=0.1), 1.0),       # BASELINE configs[0]
    'modelnet717': ('modelnet', _sphere, dict(n=717, R=0.60, sigma=0.01), 1.0),          # configs[1]
    '3dmatch20k': ('3dmatch', _room, dict(n=20000, L=1.85, sigma=0.003, off=0.1), 1.0),  # configs[2] (and [4])
    'kitti60k': ('kitti', _ground, dict(n=60000, R=60.0, sigma=0.05), 10.0),             # configs[3]
    'kitti20k': ('kitti', _ground, dict(n=20000, R=35.0, sigma=0.05), 10.0),             # configs[3] density, oracle-runnable in ~30 s
    'kitti4k': ('kitti', _ground, dict(n=4000, R=18.0, sigma=0.05), 10.0),               # configs[3] shape at a reference-runnable size
}


def make_pair(workload, index=0, base_seed=7351):
    """Pair ``index`` of a workload; seed = base_seed + index (reference config.py:13 seed)."""
    cfg_name, gen, kw, tscale = WORKLOADS[workload]
    rng = np.random.default_rng(base_seed + index)
    ref = gen(rng, **kw)
    world = gen(rng, **kw)  # independent resample of the same surface
    axis = rng.normal(size=3)
    angle = rng.uniform(0.0, np.pi / 4)
    R = _rodrigues(axis, angle)
    t = rng.uniform(-0.5, 0.5, size=3) * tscale
    src = (world - t) @ R  # so that ref ~= R @ src + t
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    n = ref.shape[0]
    return {
        'ref_points': ref.astype(np.float32),
        'src_points': src.astype(np.float32),
        'ref_feats': np.ones((n, 1), dtype=np.float32),
        'src_feats': np.ones((n, 1), dtype=np.float32),
        'transform': T.astype(np.float32),
        'config': cfg_name,
    }
