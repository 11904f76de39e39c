#!/usr/bin/env python
"""Small invocations of every kernel family added late in the round, meant to run under
    python tests/ubench/sanitize_small.py   (compute-sanitizer is closed on the shared pool; elsewhere: under --tool memcheck)
(out-of-bounds / misaligned accesses in the multi-radius read-out, the segmented FPS, the batched collision test, the voxel
means, the degree-sorted backward, the aligned forward partition, three_nn + weights)."""
import os, sys
import numpy as np
import torch

import oracle  # noqa: E402 (test infrastructure)
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, pointnet2_utils as pu, scenes
from graspbalance_b200.collision_detector import ModelFreeCollisionDetector, collision_counts_batched
from graspbalance_b200.modules import GraspWidthGrouping, multi_scale_group

dev = torch.device("cuda:0")
B, N, m = 2, 6000, 64
xyz = torch.from_numpy(scenes.scene_batch(range(B), N, "tabletop")).to(dev)
inds, new_xyz = pu.furthest_point_sample_xyz(xyz, m)
rng = np.random.default_rng(0)
rot = torch.from_numpy(scenes.viewpoint_rotations(-rng.normal(size=(B, m, 3)).astype(np.float32), np.zeros((B, m), np.float32))).to(dev)
mods = [GraspWidthGrouping(16, 3, cylinder_radius=0.08 * s, hmax_list=[0.01, 0.02, 0.03, 0.04], mlps=torch.nn.Identity()) for s in (0.25, 0.5, 0.75, 1.0)]
out = multi_scale_group(mods, new_xyz, xyz, rot)
small = xyz[:, :1500].contiguous()
out2 = multi_scale_group(mods, new_xyz, small, rot)  # full-scan path
local = pu.furthest_point_sample_segments(xyz[0], [5, 1000, 0, 4995], [5, 64, 0, 128])
_, i3, w = pu.three_nn_weights(xyz, new_xyz)
f = torch.randn((B, 32, m), device=dev, requires_grad=True)
up = pu.three_interpolate(f, i3, w)
up.backward(torch.randn_like(up))
sub = xyz[:, :2048].contiguous()
idx = pu.ball_query(0.08, 32, sub, sub)
feats = torch.randn((B, 24, 2048), device=dev, requires_grad=True)
g = pu.grouping_operation(feats, idx)
g.backward(torch.randn_like(g))
dets = [ModelFreeCollisionDetector(xyz[b].double().cpu().numpy(), voxel_size=0.01, device=dev) for b in range(B)]
gs = [scenes.grasp_set(b, dets[b].scene_points, 32) for b in range(B)]
thr = [oracle.collision_thresholds(gs[b]["heights"], gs[b]["depths"], gs[b]["widths"], 0.03) for b in range(B)]
T_, R_, H_ = (torch.from_numpy(np.ascontiguousarray(np.stack(a))).to(dev) for a in ([x["translations"] for x in gs], [x["rotation_matrices"] for x in gs], thr))
cnt = collision_counts_batched([d._scene_dev for d in dets], T_, R_, H_)
torch.cuda.synchronize()
print("sanitize_small ok", [tuple(o.shape) for o in out], tuple(local.shape), tuple(cnt.shape))
