"""Drop-in for the grasp-crop grouping of TrainModel/modules.py (the caller right above the hot path).

GraspWidthGrouping (modules.py:87-124) builds one CylinderQueryAndGroup per entry of hmax_list and calls them in a loop with
the same seeds, approach rotations, radius and hmin -- four scans of the cloud whose cylinders are nested -- then stacks
the four [B,3,num_seed,nsample] results along a new depth axis and views them as [B,3,num_seed*num_depth,nsample].
Here the loop is ONE multi-depth scan (gb_cylinder_query_multi, index lists laid out [B,num_seed,num_depth,nsample]) and
ONE grouped-coordinate launch that writes the stacked tensor directly (gb_group_xyz with nsample' = num_depth*nsample):
bit-identical to the loop, 2 launches instead of 4 x (query + group + the stack copy).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as pu


def _shared_mlp(dims):
    """Conv2d(1x1, no bias) + BatchNorm2d + ReLU per layer: what pt_utils.SharedMLP(dims, bn=True) builds (pytorch_utils.py)."""
    layers = []
    for i in range(len(dims) - 1):
        layers += [nn.Conv2d(dims[i], dims[i + 1], kernel_size=1, bias=False), nn.BatchNorm2d(dims[i + 1]), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class GraspWidthGrouping(nn.Module):
    """modules.py:87-124.  `groupers` keeps the reference's per-depth CylinderQueryAndGroup list (same constructor
    arguments) for callers that index it; forward() uses the fused path whenever the inputs allow it."""

    def __init__(self, nsample, seed_feature_dim, cylinder_radius=0.05, hmin=-0.02, hmax_list=(0.01, 0.02, 0.03, 0.04), mlps=None):
        super().__init__()
        self.nsample = nsample
        self.in_dim = seed_feature_dim
        self.cylinder_radius = cylinder_radius
        self.hmin = hmin
        self.hmax_list = list(hmax_list)
        self.groupers = [pu.CylinderQueryAndGroup(cylinder_radius, hmin, hmax, nsample, use_xyz=True) for hmax in self.hmax_list]
        self.mlps = _shared_mlp([self.in_dim, 64, 128, 256]) if mlps is None else mlps

    def group(self, seed_xyz, pointcloud, vp_rot):
        """The grouped coordinates of all depths, [B, 3, num_seed*num_depth, nsample] (modules.py:107-117)."""
        B, num_seed = vp_rot.shape[0], vp_rot.shape[1]
        D = len(self.hmax_list)
        if 1 <= D <= 4 and pu._fusable(pointcloud, seed_xyz, vp_rot):
            rot = vp_rot.reshape(B, num_seed, 9)
            idx = pu.cylinder_query_multi(self.cylinder_radius, self.hmin, self.hmax_list, self.nsample, pointcloud, seed_xyz, rot)
            g = pu._FusedQueryGroup.apply(pointcloud, seed_xyz, idx.view(B, num_seed, D * self.nsample), rot, None, None)
            return g.view(B, 3, num_seed * D, self.nsample)
        grouped = torch.stack([grouper(pointcloud, seed_xyz, vp_rot) for grouper in self.groupers], dim=3)
        return grouped.view(B, -1, num_seed * D, self.nsample)

    def forward(self, seed_xyz, pointcloud, vp_rot):
        B, num_seed = vp_rot.shape[0], vp_rot.shape[1]
        vp_features = self.mlps(self.group(seed_xyz, pointcloud, vp_rot))
        vp_features = F.max_pool2d(vp_features, kernel_size=[1, vp_features.size(3)])
        return vp_features.view(B, -1, num_seed, len(self.hmax_list))


def multi_scale_group(width_groups, seed_xyz, pointcloud, vp_rot):
    """The grouped coordinates of several GraspWidthGrouping modules that see the same inputs -- WidthGroup1..4 of
    GraspPoseStage2_seed_features_multi_scale.forward (TrainModel/graspbalance.py:104-107), which differ in the cylinder radius
    only.  One scan of the cloud with the largest radius serves all radii x depths (gb_cylinder_query_multi_radius), then one
    grouped-coordinate launch per module.  Returns [m.group(seed_xyz, pointcloud, vp_rot) for m in width_groups], bit for bit."""
    g0 = width_groups[0]
    same = all(m.hmin == g0.hmin and m.hmax_list == g0.hmax_list and m.nsample == g0.nsample for m in width_groups)
    D, R = len(g0.hmax_list), len(width_groups)
    if not (same and 1 <= D <= 4 and 2 <= R <= 4 and pu._fusable(pointcloud, seed_xyz, vp_rot)):
        return [m.group(seed_xyz, pointcloud, vp_rot) for m in width_groups]
    B, num_seed = vp_rot.shape[0], vp_rot.shape[1]
    rot = vp_rot.reshape(B, num_seed, 9)
    idx = pu.cylinder_query_multi_radius([m.cylinder_radius for m in width_groups], g0.hmin, g0.hmax_list, g0.nsample, pointcloud,
                                         seed_xyz, rot)
    return [pu._FusedQueryGroup.apply(pointcloud, seed_xyz, idx[k].view(B, num_seed, D * g0.nsample), rot, None, None)
            .view(B, 3, num_seed * D, g0.nsample) for k in range(R)]


def balance_plan(labels, counts, num_seed=1024):
    """The sampling plan of one scene (modules.py:190-193,199-205): labels = the sorted unique segment labels, counts = their
    point counts.  Every non-background object t gets num_seed // num_objects seeds, the last one the remainder as well, with
    num_objects = len(labels) - 1 exactly as upstream (label 0 = background is assumed present).  Returns (first position in
    the label-sorted point order, points, seeds) per object, in label order.  Host logic only."""
    num_objects = len(labels) - 1
    if num_objects <= 0:
        raise ZeroDivisionError("ObjectBalanceSampling needs at least one object besides the background (as upstream)")
    per_obj = [num_seed // num_objects] * num_objects
    per_obj[-1] += num_seed % num_objects
    plan, start, t = [], 0, 0
    for lab, c in zip(labels, counts):
        if lab != 0:
            plan.append((start, c, per_obj[t]))  # IndexError when label 0 is absent, as upstream
            t += 1
        start += c
    return plan


def ObjectBalanceSampling(end_points, num_seed=1024):
    """TrainModel/modules.py:177-223: every segmented object of a scene contributes num_seed // num_objects seeds (the last
    one takes the remainder), chosen by FPS among the object's points; the seeds' indices, coordinates and up-sampled
    features replace fp2_inds / fp2_xyz / fp2_features.  The reference loops over scenes and objects with one B=1
    furthest_point_sample per object; here all objects of all scenes are sampled by ONE segmented launch
    (gb_fps_segments), with identical picks.  Label 0 is background, as in the reference."""
    batch_seg_res = end_points["seed_cluster"]
    batch_points = end_points["point_clouds"]
    batch_features = end_points["up_sample_features"].permute(0, 2, 1)
    B, N = batch_seg_res.shape
    order = torch.argsort(batch_seg_res, dim=1, stable=True)          # points of a scene grouped by label, index order kept
    packed_src, counts, ks, scene_of_obj = [], [], [], []
    for i in range(B):
        labels, cnt = torch.unique(batch_seg_res[i], return_counts=True)  # sorted labels (one host sync per scene, as upstream)
        for start, c, k in balance_plan(labels.tolist(), cnt.tolist(), num_seed):
            packed_src.append(order[i, start:start + c])
            counts.append(c)
            ks.append(k)
            scene_of_obj.append(i)
    src = torch.cat(packed_src)                                         # scene-local index of every packed point
    scene_ids = torch.repeat_interleave(torch.tensor(scene_of_obj, device=src.device), torch.tensor(counts, device=src.device))
    packed_xyz = batch_points[scene_ids, src].contiguous().float()
    local = pu.furthest_point_sample_segments(packed_xyz, counts, ks).long()
    firsts = torch.tensor([0] + counts[:-1], device=src.device).cumsum(0)
    pick = src[torch.repeat_interleave(firsts, torch.tensor(ks, device=src.device)) + local]  # scene-local indices, scene-major
    fp2_inds = pick.view(B, num_seed)
    end_points["fp2_inds_fps"] = end_points["fp2_inds"]
    end_points["fp2_inds"] = fp2_inds.int()
    end_points["fp2_xyz"] = torch.gather(batch_points, 1, fp2_inds.unsqueeze(-1).expand(-1, -1, 3))
    end_points["fp2_features"] = torch.gather(batch_features, 1, fp2_inds.unsqueeze(-1).expand(-1, -1, batch_features.shape[-1])).permute(0, 2, 1)
    return end_points
