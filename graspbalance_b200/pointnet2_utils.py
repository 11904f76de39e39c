"""Drop-in for the reference's PointNet/pointnet2_utils.py: the seven autograd Function aliases and the grouper modules,
same names, call signatures, dtypes and index layouts, running on graspbalance_b200._ext (libgbops.so).

    furthest_point_sample(xyz, npoint)                        pointnet2_utils.py:46-56
    gather_operation(features, idx)                           :59-76
    three_nn(unknown, known) -> (dist, idx)                   :79-91   (dist is the sqrt of the native squared distance)
    three_interpolate(features, idx, weight)                  :94-116
    grouping_operation(features, idx)                         :119-137
    ball_query(radius, nsample, xyz, new_xyz)                 :140-150 (note the native order: new_xyz, xyz, radius, nsample)
    cylinder_query(radius, hmin, hmax, nsample, xyz, new_xyz, rot)   :235-244
    QueryAndGroup / GroupAll / CylinderQueryAndGroup / RandomDropout :35-43,152-232,247-308
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        idx = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, _, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        idx = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class CylinderQuery(Function):
    @staticmethod
    def forward(ctx, radius, hmin, hmax, nsample, xyz, new_xyz, rot):
        idx = _ext.cylinder_query(new_xyz, xyz, rot, radius, hmin, hmax, nsample)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return (None,) * 7


cylinder_query = CylinderQuery.apply


class RandomDropout(nn.Module):
    """pointnet2_utils.py:35-43.  The reference calls pt_utils.feature_dropout_no_scaling, which its pytorch_utils.py
    does not define; this keeps the constructor and applies an unscaled whole-channel dropout with rate U(0, p)."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p, self.inplace = p, inplace

    def forward(self, X):
        theta = torch.empty(1).uniform_(0, self.p).item()
        if not self.training or theta == 0:
            return X
        keep = (torch.rand(X.shape[:2] + (1,) * (X.dim() - 2), device=X.device) >= theta).to(X.dtype)
        return X.mul_(keep) if self.inplace else X * keep


def _resample_uniformly(idx, nsample):
    """The reference's `sample_uniformly` post-pass (pointnet2_utils.py:167-176,270-279): per region keep the unique
    indices and pad by random re-draws of them; returns the per-region unique counts as well (host loop, as upstream)."""
    unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
    for b in range(idx.shape[0]):
        for r in range(idx.shape[1]):
            uniq = torch.unique(idx[b, r, :])
            k = uniq.shape[0]
            unique_cnt[b, r] = k
            draw = torch.randint(0, k, (nsample - k,), dtype=torch.long)
            idx[b, r, :] = torch.cat((uniq, uniq[draw]))
    return unique_cnt


class _GroupBase(nn.Module):
    """Shared tail of QueryAndGroup / CylinderQueryAndGroup: group xyz, centre, optional scale / rotation, group
    features, concatenate, and the ret_* tuple convention (pointnet2_utils.py:178-207,281-308)."""

    def _finish(self, idx, xyz, new_xyz, features, rot=None):
        unique_cnt = _resample_uniformly(idx, self.nsample) if self.sample_uniformly else None
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B, 3, npoint, nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz /= self.radius
        if rot is not None:
            g = torch.matmul(grouped_xyz.permute(0, 2, 3, 1).contiguous(), rot)
            grouped_xyz = g.permute(0, 3, 1, 2).contiguous()
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class QueryAndGroup(_GroupBase):
    """pointnet2_utils.py:152-207: ball query, then grouped (relative xyz, features) as (B, 3+C, npoint, nsample)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False, sample_uniformly=False,
                 ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz, self.normalize_xyz = ret_grouped_xyz, normalize_xyz
        self.sample_uniformly, self.ret_unique_cnt = sample_uniformly, ret_unique_cnt
        if ret_unique_cnt:
            assert sample_uniformly

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        return self._finish(idx, xyz, new_xyz, features)


class GroupAll(nn.Module):
    """pointnet2_utils.py:210-232 (no native call)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz  # the reference forgets to store this (:213-214) and raises in forward

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            new_features = grouped_xyz
        else:
            f = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, f], dim=1) if self.use_xyz else f
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz
        return new_features


class CylinderQueryAndGroup(_GroupBase):
    """pointnet2_utils.py:247-308: cylinder query in each seed's gripper frame, grouped xyz rotated into that frame."""

    def __init__(self, radius, hmin, hmax, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 rotate_xyz=True, sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.hmin, self.hmax = radius, nsample, hmin, hmax
        self.use_xyz, self.ret_grouped_xyz, self.normalize_xyz = use_xyz, ret_grouped_xyz, normalize_xyz
        self.rotate_xyz, self.sample_uniformly, self.ret_unique_cnt = rotate_xyz, sample_uniformly, ret_unique_cnt
        if ret_unique_cnt:
            assert sample_uniformly

    def forward(self, xyz, new_xyz, rot, features=None):
        B, npoint, _ = new_xyz.size()
        idx = cylinder_query(self.radius, self.hmin, self.hmax, self.nsample, xyz, new_xyz, rot.view(B, npoint, 9))
        return self._finish(idx, xyz, new_xyz, features, rot if self.rotate_xyz else None)
