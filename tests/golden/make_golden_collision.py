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
    # a, b: small; c: BASELINE config 1 as written (20k-point scene, voxel 0.01, 1024 grasps); d: a float32 grasp group (what
    # graspnetAPI builds from network output: the reference then evaluates thresholds and volumes in float32)
    for tag, (seed, n, g, voxel, dtype) in {"a": (11, 6000, 192, 0.01, np.float64), "b": (12, 4000, 128, 0.005, np.float64),
                                            "c": (13, 20000, 1024, 0.01, np.float64), "d": (14, 8000, 256, 0.01, np.float32)}.items():
        raw = scenes.tabletop_scene(seed, n).astype(np.float64)
        pts = oracle.voxel_down_sample(raw, voxel)
        gs = {k: np.ascontiguousarray(v.astype(dtype)) for k, v in scenes.grasp_set(seed + 100, pts, g).items()}
        det = ref_cd.ModelFreeCollisionDetector(pts, voxel_size=voxel)
        assert det.scene_points.shape == pts.shape
        gg = scenes.GraspGroupStandIn(**gs)
        plain = det.detect(gg, approach_dist=0.05, collision_thresh=0.01)
        full = det.detect(gg, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True,
                          empty_thresh=0.01, return_ious=True)
        assert (plain == full[0]).all()
        if tag in "cd":  # keep the fixture small: the scene is regenerated from its seed by the test
            out.update({f"{tag}_seed": np.int64(seed), f"{tag}_n": np.int64(n), f"{tag}_points_shape": np.array(pts.shape),
                        f"{tag}_points_sum": pts.sum(0)})
        else:
            out[f"{tag}_points"] = pts
        out.update({f"{tag}_voxel": np.float64(voxel),
                    **{f"{tag}_{k}": v for k, v in gs.items()},
                    f"{tag}_collision": full[0], f"{tag}_empty": full[1], f"{tag}_ious": np.stack(full[2])})
        print(tag, pts.shape, "collisions", int(full[0].sum()), "empty", int(full[1].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collision_ref.npz"), **out)


if __name__ == "__main__":
    main()
