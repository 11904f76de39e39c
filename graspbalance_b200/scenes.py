"""Seeded synthetic inputs for tests and bench.py.

"tabletop": 20000 fp32 points in metres -- a noisy table plane plus eight surface-sampled objects (boxes, spheres,
cylinders of 3-12 cm), shuffled, with the last `n_dup` points exact duplicates of earlier points (the real loader pads
short clouds by duplication, DataProcessing/graspnet_dataset.py:130-133, which is what exercises the FPS tie rule).
"uniform": U(-0.5,0.5)^3, the worst case for early exit in the candidate scans.

numpy only: the generator is shared by the CPU oracle tests and the GPU path.
"""
import numpy as np


def _box(rng, n, size):
    face = rng.integers(0, 6, n)
    p = rng.uniform(-0.5, 0.5, (n, 3))
    ax = face // 2
    p[np.arange(n), ax] = np.where(face % 2 == 0, -0.5, 0.5)
    return p * size


def _sphere(rng, n, r):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
    return v * r


def _cylinder(rng, n, r, h):
    th = rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(-0.5, 0.5, n) * h
    return np.stack([r * np.cos(th), r * np.sin(th), z], axis=1)


def tabletop_scene(seed, n=20000, n_dup=500, n_objects=8):
    """One scene [n,3] float32."""
    rng = np.random.default_rng(seed)
    n_dup = min(n_dup, n // 8)
    n_unique = n - n_dup
    n_obj_pts = (n_unique * 2) // 5 // max(n_objects, 1)
    n_plane = n_unique - n_obj_pts * n_objects
    parts = [np.stack([rng.uniform(-0.35, 0.35, n_plane), rng.uniform(-0.25, 0.25, n_plane),
                       0.5 + rng.normal(0, 0.002, n_plane)], axis=1)]
    for o in range(n_objects):
        kind = o % 3
        c = np.array([rng.uniform(-0.28, 0.28), rng.uniform(-0.18, 0.18), 0.0])
        if kind == 0:
            size = rng.uniform(0.03, 0.12, 3)
            p = _box(rng, n_obj_pts, size)
            c[2] = 0.5 - size[2] / 2
        elif kind == 1:
            r = rng.uniform(0.015, 0.06)
            p = _sphere(rng, n_obj_pts, r)
            c[2] = 0.5 - r
        else:
            r, h = rng.uniform(0.015, 0.05), rng.uniform(0.03, 0.12)
            p = _cylinder(rng, n_obj_pts, r, h)
            c[2] = 0.5 - h / 2
        parts.append(p + c)
    pts = np.concatenate(parts, axis=0)
    pts = pts[rng.permutation(pts.shape[0])]
    if n_dup > 0:
        pts = np.concatenate([pts, pts[rng.integers(0, n_unique, n_dup)]], axis=0)
    return np.ascontiguousarray(pts, dtype=np.float32)


def uniform_scene(seed, n=20000):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)


def scene_batch(seeds, n=20000, kind="tabletop"):
    f = tabletop_scene if kind == "tabletop" else uniform_scene
    return np.stack([f(int(s), n) for s in seeds], axis=0)


def random_rotations(rng, shape):
    """Proper rotations (det=+1) from the QR of N(0,1) matrices; float64 [*shape,3,3]."""
    a = rng.normal(size=tuple(shape) + (3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[..., None, :]
    det = np.linalg.det(q)
    q[..., :, 2] *= det[..., None]
    return q


def viewpoint_rotations(towards, angle):
    """Rotation matrices from approach vectors and in-plane angles -- the construction of
    loss_utils.batch_viewpoint_params_to_matrix (loss_utils.py:33-49): x axis = approach direction,
    y axis = (-a_y, a_x, 0) normalised ((0,1,0) when degenerate), z = x cross y, times a rotation about x."""
    towards = np.asarray(towards, dtype=np.float32)
    angle = np.asarray(angle, dtype=np.float32)
    ax = towards
    zeros = np.zeros_like(ax[..., 0])
    ay = np.stack([-ax[..., 1], ax[..., 0], zeros], axis=-1)
    deg = np.linalg.norm(ay, axis=-1) == 0
    ay[deg] = np.array([0, 1, 0], dtype=np.float32)
    ax = ax / np.linalg.norm(ax, axis=-1, keepdims=True)
    ay = ay / np.linalg.norm(ay, axis=-1, keepdims=True)
    az = np.cross(ax, ay)
    s, c = np.sin(angle), np.cos(angle)
    ones = np.ones_like(s)
    r1 = np.stack([ones, zeros, zeros, zeros, c, -s, zeros, s, c], axis=-1).reshape(angle.shape + (3, 3))
    r2 = np.stack([ax, ay, az], axis=-1)
    return np.matmul(r2, r1).astype(np.float32)


def grasp_set(seed, scene_points, g=1024):
    """cfg1 grasps: T = random scene points, R = random rotations, heights 0.02, depths in {.01,.02,.03,.04},
    widths U(0.01,0.1); all float64 (the GraspGroup attributes `detect` reads, collision_detector.py:19-22)."""
    rng = np.random.default_rng(seed)
    pts = np.asarray(scene_points, dtype=np.float64)
    T = pts[rng.integers(0, pts.shape[0], g)].copy()
    R = random_rotations(rng, (g,))
    heights = np.full(g, 0.02)
    depths = rng.choice(np.array([0.01, 0.02, 0.03, 0.04]), g)
    widths = rng.uniform(0.01, 0.1, g)
    return dict(translations=T, rotation_matrices=R, heights=heights, depths=depths, widths=widths)


class GraspGroupStandIn:
    """Duck-typed stand-in for graspnetAPI.GraspGroup: only the five attributes `detect` reads."""

    def __init__(self, translations, rotation_matrices, heights, depths, widths):
        self.translations = translations
        self.rotation_matrices = rotation_matrices
        self.heights = heights
        self.depths = depths
        self.widths = widths
