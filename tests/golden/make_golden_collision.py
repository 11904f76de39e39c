#!/usr/bin/env python
"""Golden vectors for the collision test, produced by the REFERENCE's own collision_detector.py.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_collision.py
Writes tests/golden/collision_ref.npz.  open3d / graspnetAPI are absent here, so `open3d` is stubbed in
sys.modules with a PointCloud whose voxel_down_sample is the identity: the scene handed to the detector is already
down-sampled (oracle.voxel_down_sample), so that what this fixture pins is `detect` itself
(collision_detector.py:16-64), which never touches open3d.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("GB_REFERENCE_ROOT", "/root/reference")

o3d = types.ModuleType("open3d")


class _PC:
    def __init__(self):
        self.points = None

    def voxel_down_sample(self, v):
        return self


o3d.geometry = types.SimpleNamespace(PointCloud=_PC)
o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, dtype=np.float64))
sys.modules["open3d"] = o3d
sys.path.insert(0, REF)
import collision_detector as ref_cd  # noqa: E402  (the unmodified reference file)

import oracle  # noqa: E402
from graspbalance_b200 import scenes  # noqa: E402


def main():
    out = {}
    for tag, (seed, n, g, voxel) in {"a": (11, 6000, 192, 0.01), "b": (12, 4000, 128, 0.005)}.items():
        raw = scenes.tabletop_scene(seed, n).astype(np.float64)
        pts = oracle.voxel_down_sample(raw, voxel)
        gs = scenes.grasp_set(seed + 100, pts, g)
        det = ref_cd.ModelFreeCollisionDetector(pts, voxel_size=voxel)
        assert det.scene_points.shape == pts.shape
        gg = scenes.GraspGroupStandIn(**gs)
        plain = det.detect(gg, approach_dist=0.05, collision_thresh=0.01)
        full = det.detect(gg, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True,
                          empty_thresh=0.01, return_ious=True)
        assert (plain == full[0]).all()
        out.update({f"{tag}_points": pts, f"{tag}_voxel": np.float64(voxel),
                    **{f"{tag}_{k}": v for k, v in gs.items()},
                    f"{tag}_collision": full[0], f"{tag}_empty": full[1], f"{tag}_ious": np.stack(full[2])})
        print(tag, pts.shape, "collisions", int(full[0].sum()), "empty", int(full[1].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collision_ref.npz"), **out)


if __name__ == "__main__":
    main()
