"""TEST INFRASTRUCTURE (never imported by graspbalance_b200): the op chain of BASELINE config 5 driven through the
UNMODIFIED reference extensions compiled into oracle/_ref/ by oracle/build_ref.py -- module A = pointnet2._ext
(PointNet/_ext_src), module B = pointnet2_batch_cuda (pointnet2_batch/src) -- with exactly the torch glue the reference's
Python layer puts around them:

  SA modules   PointnetSAModuleVotes.forward (PointNet/pointnet2_modules.py:148-188) -> furthest_point_sample,
               gather_operation, QueryAndGroup.forward (PointNet/pointnet2_utils.py:164-207: ball_query, grouping_operation
               of the transposed coordinates, subtract the centre, divide by the radius, grouping_operation of the features,
               torch.cat), backward through GroupingOperation (pointnet2_utils.py:119-137);
  InvResMLP    group.QueryAndGroup.forward (ModifiedNetTools/group.py:167-180) with module B's out-parameter wrappers
               (group.py:62-89,128-145): zero-filled idx, ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper;
  FP / up      PointnetFPModule.forward (pointnet2_modules.py:407-435): three_nn, sqrt, 1/(d+1e-8), normalise,
               three_interpolate, backward through ThreeInterpolate (pointnet2_utils.py:94-116);
  crops        CylinderQueryAndGroup.forward (pointnet2_utils.py:261-308) once per radius x depth
               (TrainModel/modules.py:104-113, graspbalance.py:104-107): cylinder_query, group, subtract, permute, matmul,
               permute.

It consumes the stand-in feature / gradient tensors of a graspbalance_b200.pipeline.OpPipeline so that both chains see the
same inputs.  Used by tests/test_pipeline_gpu.py (parity of the whole chain, tensor by tensor) and by bench.py's
`gpu_baseline` leg (the reference kernels recompiled for sm_100a, timed outside the timed region of the product).
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_refs():
    """(module A, module B, module C) from oracle/_ref, None for what is not built."""
    mods = []
    for name in ("gbref_pointnet2_ext", "gbref_pointnet2_batch", "gbref_knn"):
        path = os.path.join(_HERE, "_ref", name + ".so")
        if not os.path.exists(path):
            mods.append(None)
            continue
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def _group_a(ref_a, xyz, new_xyz, features, radius, nsample):
    """QueryAndGroup(radius, nsample, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True).forward."""
    idx = ref_a.ball_query(new_xyz, xyz, radius, nsample)
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = ref_a.group_points(xyz_trans, idx)
    grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
    grouped_xyz /= radius
    if features is not None:
        grouped_features = ref_a.group_points(features, idx)
        new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
    else:
        new_features = grouped_xyz
    return idx, new_features


def _group_b(ref_b, query_xyz, support_xyz, features, radius, nsample):
    """group.QueryAndGroup(radius, nsample).forward with module B's wrappers."""
    B, N, _ = support_xyz.shape
    m = query_xyz.shape[1]
    idx = torch.zeros((B, m, nsample), dtype=torch.int32, device=support_xyz.device)
    ref_b.ball_query_wrapper(B, N, m, radius, nsample, query_xyz, support_xyz, idx)
    xyz_t = support_xyz.transpose(1, 2).contiguous()
    grouped_xyz = torch.empty((B, 3, m, nsample), dtype=torch.float32, device=idx.device)
    ref_b.group_points_wrapper(B, 3, N, m, nsample, xyz_t, idx, grouped_xyz)
    grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
    C = features.shape[1]
    fj = torch.empty((B, C, m, nsample), dtype=torch.float32, device=idx.device)
    ref_b.group_points_wrapper(B, C, N, m, nsample, features, idx, fj)
    return idx, grouped_xyz, fj


def run(pipe, xyz, view_rot, ref_a, ref_b, collect=None, backward=True):
    """One forward(+backward) pass of the chain.  `collect`: dict that receives every index tensor, forward tensor and
    gradient under the names OpPipeline.run(collect=...) uses."""
    from graspbalance_b200.pipeline import CROP_HMAX, CROP_HMIN, CROP_RADII, IRM_SPECS, SA_SPECS
    c = collect
    B = xyz.shape[0]
    cur, levels = xyz, []
    for lvl, (npoint, radius, nsample, c_in) in enumerate(SA_SPECS):
        inds = ref_a.furthest_point_sampling(cur, npoint)
        new_xyz = ref_a.gather_points(cur.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        feats = pipe.sa_in_feats[lvl]
        f = None if feats is None else feats.detach()
        idx, grouped = _group_a(ref_a, cur, new_xyz, f, radius, nsample)
        grad = None
        if backward and f is not None:
            go = pipe.sa_grads[lvl]
            grad = ref_a.group_points_grad(go[:, 3:].contiguous(), idx, cur.shape[1])
        if c is not None:
            c[f"sa{lvl}_inds"], c[f"sa{lvl}_xyz"], c[f"sa{lvl}_idx"], c[f"sa{lvl}_grouped"] = inds, new_xyz, idx, grouped
            if grad is not None:
                c[f"sa{lvl}_grad"] = grad
        blocks, C, r2, nsb = IRM_SPECS[lvl]
        fi = pipe.irm_feats[lvl].detach()
        gsum = torch.zeros_like(fi) if backward else None
        for blk in range(blocks):
            idx_b, dp, fj = _group_b(ref_b, new_xyz, new_xyz, fi, r2, nsb)
            if backward:
                ref_b.group_points_grad_wrapper(B, C, npoint, npoint, nsb, pipe.irm_grads[lvl], idx_b, gsum)
        if c is not None:
            c[f"irm{lvl}_idx"], c[f"irm{lvl}_dp"], c[f"irm{lvl}_fj"] = idx_b, dp, fj
            if backward:
                c[f"irm{lvl}_grad"] = gsum
        cur = new_xyz
        levels.append(new_xyz)
    for i, (unknown, known) in enumerate(((levels[2], levels[3]), (levels[1], levels[2]), (xyz, levels[1]))):
        dist2, idx3 = ref_a.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        dist_recip = 1.0 / (dist + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        weight = dist_recip / norm
        f = pipe.fp_feats[i].detach()
        out = ref_a.three_interpolate(f, idx3, weight)
        if c is not None:
            c[f"fp{i}_idx"], c[f"fp{i}_weight"], c[f"fp{i}_out"] = idx3, weight, out
        if backward:
            g = ref_a.three_interpolate_grad(pipe.fp_grads[i].contiguous(), idx3, weight, f.shape[2])
            if c is not None:
                c[f"fp{i}_grad"] = g
    seed_xyz = levels[1]
    xyz_trans = xyz.transpose(1, 2).contiguous()
    rot9 = view_rot.reshape(B, seed_xyz.shape[1], 9).contiguous()
    rot33 = view_rot.reshape(B, seed_xyz.shape[1], 3, 3)
    for k, radius in enumerate(CROP_RADII):
        for d, hmax in enumerate(CROP_HMAX):
            idx = ref_a.cylinder_query(seed_xyz, xyz, rot9, radius, CROP_HMIN, hmax, 64)
            grouped_xyz = ref_a.group_points(xyz_trans, idx)
            grouped_xyz -= seed_xyz.transpose(1, 2).unsqueeze(-1)
            grouped_xyz_ = grouped_xyz.permute(0, 2, 3, 1).contiguous()
            grouped_xyz_ = torch.matmul(grouped_xyz_, rot33)
            grouped_xyz = grouped_xyz_.permute(0, 3, 1, 2).contiguous()
            if c is not None:
                c[f"crop{k}_{d}_idx"], c[f"crop{k}_{d}_xyz"] = idx, grouped_xyz
    return levels
