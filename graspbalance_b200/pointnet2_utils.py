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
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext
from . import _lib


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


class FurthestPointSamplingXYZ(Function):
    """furthest_point_sample + the gather_operation PointnetSAModuleVotes runs on its result
    (pointnet2_modules.py:151-158): returns (inds [B,npoint] i32, new_xyz [B,npoint,3]) from one launch.  new_xyz carries no
    gradient to xyz on this path (callers that need one use gather_operation)."""

    @staticmethod
    def forward(ctx, xyz, npoint, max_cluster=0):
        inds, new_xyz = _ext.furthest_point_sampling_xyz(xyz, npoint, max_cluster)
        ctx.mark_non_differentiable(inds, new_xyz)
        return inds, new_xyz

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


def furthest_point_sample_xyz(xyz, npoint, max_cluster=0):
    """max_cluster > 0: background sampling -- at most that many CTAs per scene (gb_fps_xyz_hint); same picks."""
    return FurthestPointSamplingXYZ.apply(xyz, npoint, max_cluster)


def furthest_point_sample_segments(points, counts, nsamples):
    """FPS of many point sets of different sizes in one launch -- the per-object loop of ObjectBalanceSampling
    (TrainModel/modules.py:186-213).  points [total,3] f32 CUDA = the sets back to back; counts / nsamples = per-set sizes
    and sample counts (host sequences).  Returns the concatenated set-local indices ([sum(nsamples)] i32), each set sampled
    exactly as furthest_point_sample(set.unsqueeze(0), k)[0] samples it."""
    rows, total_points, total_out = segment_table(counts, nsamples)
    assert total_points <= points.shape[0]
    # one CTA holds a segment in registers (up to SEGMENT_MAX_POINTS points); an object that owns more of the cloud than that
    # is sampled by the cluster kernel on its own, as the reference's per-object call does (modules.py:207)
    small = [r for r in rows if r[1] <= SEGMENT_MAX_POINTS]
    seg = torch.tensor(small, dtype=torch.int32).reshape(-1, 4).to(points.device)
    out = _ext.furthest_point_sampling_segments(points, seg, max((r[1] for r in small), default=0), max((r[2] for r in small), default=0),
                                                total_out)
    for first, c, k, slot in rows:
        if c > SEGMENT_MAX_POINTS and k > 0:
            out[slot:slot + k] = _ext.furthest_point_sampling(points[first:first + c].unsqueeze(0).contiguous(), k)[0]
    return out


SEGMENT_MAX_POINTS = 10240  # gb_fps_segments: 512 threads x 20 points


def segment_table(counts, nsamples):
    """Rows (first point, points, samples, first output slot) of gb_fps_segments for point sets stored back to back; also
    the number of points and of output slots they span.  Host logic only."""
    counts, nsamples = [int(c) for c in counts], [int(k) for k in nsamples]
    if len(counts) != len(nsamples) or any(c < 0 for c in counts) or any(k < 0 for k in nsamples):
        raise ValueError("counts and nsamples must be non-negative sequences of equal length")
    rows, first, slot = [], 0, 0
    for c, k in zip(counts, nsamples):
        rows.append((first, c, k, slot))
        first += c
        slot += k
    return rows, first, slot


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


class ThreeNNWeights(Function):
    """three_nn and the inverse-distance weights every caller derives from it (pointnet2_modules.py:413-416,
    TrainModel/graspbalance.py:37-41) in one launch: returns (dist, idx, weight), weight = (1/(dist+1e-8)) normalised over
    the three neighbours -- bit-identical to the torch ops it replaces.  Like three_nn, nothing here is differentiable."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist, idx, weight = _ext.three_nn_weights(unknown, known)
        ctx.mark_non_differentiable(dist, idx, weight)
        return dist, idx, weight

    @staticmethod
    def backward(ctx, a=None, b=None, c=None):
        return None, None


three_nn_weights = ThreeNNWeights.apply


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


class ThreeInterpolation(Function):
    """three_nn -> inverse-distance weights -> three_interpolate as ONE launch: PointnetFPModule.forward
    (pointnet2_modules.py:413-420), upsampling.three_interpolation (upsampling.py:67-74), graspbalance.py:37-41.  dist / idx /
    weight [B,n,3] are not materialised; idx and weight are kept (written by the same launch) only when the features need a
    gradient.  Values are bit-identical to the three separate steps; coordinates get no gradient, as in the reference
    (three_nn's outputs are non-differentiable there too)."""

    @staticmethod
    def forward(ctx, unknown, known, features):
        need = features.requires_grad
        res = _ext.three_interpolation(unknown, known, features, need)
        if res is None:  # shape the fused kernel does not take: the two launches
            _, idx, weight = _ext.three_nn_weights(unknown, known)
            res = (_ext.three_interpolate(features, idx, weight), idx, weight)
        out, idx, weight = res
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return None, None, _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m)


def three_interpolation(unknown, known, features):
    """Fused feature propagation; falls back to the unfused reference expression for inputs the kernels do not take."""
    if _fusable(unknown, known) and features.is_cuda and features.dtype == torch.float32 and features.is_contiguous():
        return ThreeInterpolation.apply(unknown, known, features)
    dist, idx = three_nn(unknown, known)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    return three_interpolate(features, idx, weight)


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


class GroupingMax(Function):
    """grouping_operation + max over nsample as one pass: what the 'max' pooling of PointnetSAModuleVotes(_WOMLP)
    (pointnet2_modules.py:173-175, 324-335) computes from the grouped tensor when no MLP sits in between --
    F.max_pool2d(grouping_operation(features, idx), kernel_size=[1, nsample]).squeeze(-1), bit for bit, without the
    [B,C,npoint,nsample] tensor.  Backward sends each gradient to the source that won the maximum (max_pool2d's backward
    followed by GroupingOperation's)."""

    @staticmethod
    def forward(ctx, features, idx):
        res = _ext.group_points_max(features, idx, features.requires_grad)
        if res is None:
            raise RuntimeError("grouping_max: rows of more than 12800 points are not taken; use grouping_operation + max_pool2d")
        out, arg = res
        ctx.for_backwards = (arg, features.size(2))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        arg, N = ctx.for_backwards
        return _ext.group_points_max_grad(grad_out.contiguous(), arg, N), None


def grouping_max(features, idx):
    """max over the samples of grouping_operation(features, idx): one pass (GroupingMax) for clouds of up to 12800 points,
    whose rows fit shared memory four at a time; the two reference steps otherwise."""
    if features.is_cuda and features.dtype == torch.float32 and features.is_contiguous() and features.size(2) * 16 <= 200 * 1024:
        return GroupingMax.apply(features, idx)
    import torch.nn.functional as F
    return F.max_pool2d(grouping_operation(features, idx), kernel_size=[1, idx.size(2)]).squeeze(-1)


class QueryGroupMaxPool(nn.Module):
    """QueryAndGroup(radius, nsample, use_xyz, normalize_xyz) followed by max pooling over the samples, as
    PointnetSAModuleVotes_WOMLP.forward runs them (pointnet2_modules.py:310-335), in three small launches: ball_query and one
    grouping_max each for coordinates and features.  The maximum of the centred (and scaled) coordinates is taken on the raw
    coordinates first -- subtracting the centre and multiplying by 1/radius are monotone, so the values are bit-identical.
    Returns [B, 3+C, npoint] (or [B, C, npoint] without use_xyz)."""

    def __init__(self, radius, nsample, use_xyz=True, normalize_xyz=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz, self.normalize_xyz = radius, nsample, use_xyz, normalize_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        parts = []
        if self.use_xyz or features is None:
            pooled = grouping_max(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2)
            if self.normalize_xyz:
                pooled = pooled / self.radius
            parts.append(pooled)
        if features is not None:
            parts.append(grouping_max(features, idx))
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


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


class CylinderQueryMulti(Function):
    """All depths of GraspWidthGrouping's loop over hmax_list (TrainModel/modules.py:104-113) in one scan:
    idx [B,npoint,D,nsample] with idx[:, :, d] == cylinder_query(radius, hmin, hmax_list[d], nsample, xyz, new_xyz, rot)."""

    @staticmethod
    def forward(ctx, radius, hmin, hmax_list, nsample, xyz, new_xyz, rot):
        idx = _ext.cylinder_query_multi(new_xyz, xyz, rot, radius, hmin, list(hmax_list), nsample)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return (None,) * 7


cylinder_query_multi = CylinderQueryMulti.apply


class CylinderQueryMultiRadius(Function):
    """All radii x depths of the four GraspWidthGrouping modules of GraspPoseStage2_seed_features_multi_scale.forward
    (TrainModel/graspbalance.py:104-107) in one scan: idx [R,B,npoint,D,nsample] with
    idx[k, :, :, d] == cylinder_query(radii[k], hmin, hmax_list[d], nsample, xyz, new_xyz, rot)."""

    @staticmethod
    def forward(ctx, radii, hmin, hmax_list, nsample, xyz, new_xyz, rot):
        idx = _ext.cylinder_query_multi_radius(new_xyz, xyz, rot, list(radii), hmin, list(hmax_list), nsample)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return (None,) * 7


cylinder_query_multi_radius = CylinderQueryMultiRadius.apply


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


def _fusable(*tensors):
    """The fused grouper kernels take CUDA fp32 contiguous inputs and do not differentiate w.r.t. coordinates."""
    return all(t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and not t.requires_grad) for t in tensors)


class _FusedQueryGroup(Function):
    """(B, 3[+C], npoint, nsample) result of QueryAndGroup / CylinderQueryAndGroup written in place: gb_group_xyz (gather + centre
    + scale + rotate) into rows 0..2 and gb_group_fwd_strided into rows 3.. (one launch, gb_group_xyz_feat, when there are
    features and no rotation) -- instead of
    transpose, group, subtract, divide, permute, matmul, permute, group, cat (pointnet2_utils.py:178-207, 281-308).
    Backward reads the feature rows' gradient from the same slice (gb_group_bwd_strided); coordinates get no gradient
    (callers that need one take the unfused path)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, idx, rot, features, inv_radius):
        B, npoint, nsample = idx.shape
        N = xyz.shape[1]
        C = 0 if features is None else features.shape[1]
        per = npoint * nsample
        out = torch.empty((B, 3 + C, npoint, nsample), dtype=torch.float32, device=xyz.device)
        stride = (3 + C) * per
        if C and rot is None and features.shape[2] == N:
            # one launch: every CTA of the feature kernel first writes its share of the coordinate rows
            _lib.call("gb_group_xyz_feat", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), out.data_ptr(), stride,
                      float(inv_radius or 0.0), 1 if inv_radius is not None else 0, features.data_ptr(), out.data_ptr() + 12 * per,
                      stride, B, C, N, npoint, nsample)
        else:
            _lib.call("gb_group_xyz", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), None if rot is None else rot.data_ptr(),
                      out.data_ptr(), B, N, npoint, nsample, float(inv_radius or 0.0), 1 if inv_radius is not None else 0, stride)
            if C:
                _lib.call("gb_group_fwd_strided", features, features.data_ptr(), idx.data_ptr(), out.data_ptr() + 12 * per, B, C,
                          features.shape[2], npoint, nsample, stride)
        ctx.for_backwards = (idx, None if features is None else features.shape[2], C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N, C = ctx.for_backwards
        if not C or not ctx.needs_input_grad[4]:
            return None, None, None, None, None, None
        B, npoint, nsample = idx.shape
        per = npoint * nsample
        grad_out = grad_out.contiguous()
        grad = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
        _lib.call("gb_group_bwd_strided", grad_out, grad_out.data_ptr() + 12 * per, idx.data_ptr(), grad.data_ptr(), B, C, N, npoint,
                  nsample, (3 + C) * per, 1)
        return None, None, None, None, grad, None


class _GroupBase(nn.Module):
    """Shared tail of QueryAndGroup / CylinderQueryAndGroup: group xyz, centre, optional scale / rotation, group
    features, concatenate, and the ret_* tuple convention (pointnet2_utils.py:178-207,281-308)."""

    def _finish(self, idx, xyz, new_xyz, features, rot=None):
        unique_cnt = _resample_uniformly(idx, self.nsample) if self.sample_uniformly else None
        if self.use_xyz and _fusable(xyz, new_xyz, rot) and (features is None or (features.is_cuda and features.dtype == torch.float32
                                                                                   and features.is_contiguous())) and idx.is_contiguous():
            # ATen evaluates `grouped_xyz /= radius` as a multiply by the fp32 reciprocal
            inv = float(np.float32(1.0) / np.float32(self.radius)) if self.normalize_xyz else None
            new_features = _FusedQueryGroup.apply(xyz, new_xyz, idx, None if rot is None else rot.reshape(rot.shape[0], rot.shape[1], 9),
                                                  features, inv)
            ret = [new_features]
            if self.ret_grouped_xyz:
                ret.append(new_features[:, :3])
            if self.ret_unique_cnt:
                ret.append(unique_cnt)
            return ret[0] if len(ret) == 1 else tuple(ret)
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
