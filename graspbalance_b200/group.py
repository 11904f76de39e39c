"""Drop-in for the reference's ModifiedNetTools/group.py (the PointNeXt-style groupers over pointnet2_batch_cuda).

Public names kept: KNN, DenseDilated, DilatedKNN, GroupingOperation / grouping_operation, torch_grouping_operation,
GatherOperation / gather_operation, BallQuery / ball_query, QueryAndGroup, GroupAll, KNNGroup,
get_aggregation_feautres (sic), create_grouper.  The three Functions call graspbalance_b200.pointnet2_batch_cuda with the
reference's out-parameter convention (group.py:62-145); KNN / KNNGroup stay torch.cdist + topk as in the reference
(group.py:15-24: torch-internal tie order, not a parity target).
"""
import copy
import logging

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from . import pointnet2_batch_cuda as pointnet2_cuda


class KNN(nn.Module):
    def __init__(self, neighbors, transpose_mode=True):
        super().__init__()
        self.neighbors = neighbors

    @torch.no_grad()
    def forward(self, support, query):
        k_dist = torch.cdist(support, query).topk(k=self.neighbors, dim=1, largest=False)
        return k_dist.values, k_dist.indices.transpose(1, 2).contiguous().int()


class DenseDilated(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            pick = torch.randperm(self.k * self.dilation)[:self.k]
            return edge_index[:, :, pick].contiguous()
        return edge_index[:, :, ::self.dilation].contiguous()


class DilatedKNN(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.knn = KNN(k * self.dilation, transpose_mode=True)

    def forward(self, query):
        _, idx = self.knn(query, query)
        return self._dilated(idx)


class GroupingOperation(Function):
    """group.py:62-86: features (B,C,N) f32, idx (B,npoint,nsample) i32 -> (B,C,npoint,nsample)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features, idx):
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=features.device)
        pointnet2_cuda.group_points_wrapper(B, C, N, npoint, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        # group.py:83-85 zero-fills and accumulates; the _set entry writes every element, same values
        grad_features = torch.empty([B, C, N], dtype=torch.float, device=grad_out.device)
        pointnet2_cuda.group_points_grad_set(B, C, N, npoint, nsample, grad_out.detach().contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


def torch_grouping_operation(features, idx):
    """group.py:92-96: the same gather with torch.gather (the one self-check the reference carries)."""
    B, C = features.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(-1, C, -1).long()
    return features.gather(2, flat).reshape(B, C, idx.shape[1], idx.shape[2])


class GatherOperation(Function):
    """group.py:99-122: features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        pointnet2_cuda.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros([B, C, N], dtype=torch.float, device=grad_out.device)
        pointnet2_cuda.gather_points_grad_wrapper(B, C, N, npoint, grad_out.detach().contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class BallQuery(Function):
    """group.py:128-142: (radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> idx (B,npoint,nsample) i32."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        # group.py:136 zero-fills because the reference kernel leaves the slots of an empty neighbourhood untouched; gb_ball_query
        # writes every slot (zeros when there is no hit), so the fill pass is skipped
        idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        pointnet2_cuda.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class _GroupXyzFeatures(Function):
    """Tail of QueryAndGroup.forward (group.py:167-179) as one launch: grouped_xyz = (support_xyz[idx] - query_xyz) (* 1/radius)
    as (B,3,npoint,nsample) and grouping_operation(features, idx) as (B,C,npoint,nsample).  Bit-identical to gb_group_xyz +
    GroupingOperation; coordinates carry no gradient (the caller checked that none is required), features get
    GroupingOperation's backward."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, support_xyz, query_xyz, idx, features, inv_radius, use_scale):
        B, npoint, nsample = idx.shape
        _, C, N = features.shape
        per = npoint * nsample
        grouped_xyz = torch.empty((B, 3, npoint, nsample), dtype=torch.float32, device=idx.device)
        grouped = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=idx.device)
        _lib.call("gb_group_xyz_feat", features, support_xyz.data_ptr(), query_xyz.data_ptr(), idx.data_ptr(), grouped_xyz.data_ptr(),
                  3 * per, float(inv_radius), int(use_scale), features.data_ptr(), grouped.data_ptr(), C * per, B, C, N, npoint, nsample)
        ctx.mark_non_differentiable(grouped_xyz)
        ctx.set_materialize_grads(False)  # no zero-filled [B,3,npoint,nsample] gradient for the coordinate output
        ctx.for_backwards = (idx, N)
        return grouped_xyz, grouped

    @staticmethod
    def backward(ctx, _grad_xyz, grad_out):
        if grad_out is None:
            return None, None, None, None, None, None
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.empty([B, C, N], dtype=torch.float, device=grad_out.device)
        pointnet2_cuda.group_points_grad_set(B, C, N, npoint, nsample, grad_out.detach().contiguous(), idx, grad_features)
        return None, None, None, grad_features, None, None


class QueryAndGroup(nn.Module):
    """group.py:147-180: returns (grouped_xyz, grouped_features); argument order is (query_xyz, support_xyz, features)."""

    def __init__(self, radius, nsample, relative_xyz=True, normalize_dp=False, normalize_by_std=False,
                 normalize_by_allstd=False, normalize_by_allstd2=False, return_only_idx=False, **kwargs):
        super().__init__()
        self.radius, self.nsample = radius, nsample
        self.normalize_dp, self.normalize_by_std = normalize_dp, normalize_by_std
        self.normalize_by_allstd, self.normalize_by_allstd2 = normalize_by_allstd, normalize_by_allstd2
        assert self.normalize_dp + self.normalize_by_std + self.normalize_by_allstd < 2
        self.relative_xyz, self.return_only_idx = relative_xyz, return_only_idx

    def forward(self, query_xyz, support_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, support_xyz, query_xyz)
        if self.return_only_idx:
            return idx
        fusable = all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and not t.requires_grad
                      for t in (query_xyz, support_xyz))
        if self.relative_xyz and fusable and features is not None and features.is_cuda and features.is_contiguous() \
                and features.dim() == 3 and features.is_floating_point() and features.shape[2] == support_xyz.shape[1]:
            # coordinates and features of the neighbourhoods in ONE launch (gb_group_xyz_feat); same values as below
            inv = float(np.float32(1.0) / np.float32(self.radius)) if self.normalize_dp else 0.0
            return _GroupXyzFeatures.apply(support_xyz, query_xyz, idx, features, inv, 1 if self.normalize_dp else 0)
        if self.relative_xyz and fusable:
            # gather + centre (+ scale) in one launch instead of transpose, group, subtract, divide (group.py:171-176);
            # ATen evaluates `/= radius` as a multiply by the fp32 reciprocal
            B, npoint, nsample = idx.shape
            grouped_xyz = torch.empty((B, 3, npoint, nsample), dtype=torch.float32, device=idx.device)
            inv = float(np.float32(1.0) / np.float32(self.radius)) if self.normalize_dp else 0.0
            _lib.call("gb_group_xyz", support_xyz, support_xyz.data_ptr(), query_xyz.data_ptr(), idx.data_ptr(), None,
                      grouped_xyz.data_ptr(), B, support_xyz.shape[1], npoint, nsample, inv, 1 if self.normalize_dp else 0,
                      3 * npoint * nsample)
        else:
            grouped_xyz = grouping_operation(support_xyz.transpose(1, 2).contiguous(), idx)
            if self.relative_xyz:
                grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
                if self.normalize_dp:
                    grouped_xyz /= self.radius
        grouped_features = grouping_operation(features, idx) if features is not None else None
        return grouped_xyz, grouped_features


class GroupAll(nn.Module):
    def forward(self, new_xyz, xyz, features=None):
        return xyz.transpose(1, 2).unsqueeze(2), (features.unsqueeze(2) if features is not None else None)


class KNNGroup(nn.Module):
    """group.py:193-224."""

    def __init__(self, nsample, relative_xyz=True, normalize_dp=False, return_only_idx=False, **kwargs):
        super().__init__()
        self.nsample = nsample
        self.knn = KNN(nsample, transpose_mode=True)
        self.relative_xyz, self.normalize_dp, self.return_only_idx = relative_xyz, normalize_dp, return_only_idx

    def forward(self, query_xyz, support_xyz, features=None):
        _, idx = self.knn(support_xyz, query_xyz)
        if self.return_only_idx:
            return idx
        idx = idx.int()
        grouped_xyz = grouping_operation(support_xyz.transpose(1, 2).contiguous(), idx)
        if self.relative_xyz:
            grouped_xyz -= query_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_dp:
            grouped_xyz /= torch.amax(torch.sqrt(torch.sum(grouped_xyz ** 2, dim=1)), dim=(1, 2)).view(-1, 1, 1, 1)
        return grouped_xyz, (grouping_operation(features, idx) if features is not None else None)


def get_aggregation_feautres(p, dp, f, fj, feature_type='dp_fj'):
    """group.py:226-237 (name misspelt upstream; kept)."""
    if feature_type == 'dp_fj':
        return torch.cat([dp, fj], 1)
    df = fj - f.unsqueeze(-1)
    if feature_type == 'dp_fj_df':
        return torch.cat([dp, fj, df], 1)
    if feature_type == 'pi_dp_fj_df':
        pi = p.transpose(1, 2).unsqueeze(-1).expand(-1, -1, -1, df.shape[-1])
        return torch.cat([pi, dp, fj, df], 1)
    if feature_type == 'dp_df':
        return torch.cat([dp, df], 1)
    return fj


def create_grouper(group_args):
    """group.py:239-253."""
    args = copy.deepcopy(group_args)
    method = args.pop('NAME', 'ballquery')
    radius = args.pop('radius', 0.1)
    nsample = args.pop('nsample', 20)
    logging.info(group_args)
    if nsample is None:
        return GroupAll()
    if method == 'ballquery':
        return QueryAndGroup(radius, nsample, **args)
    if method == 'knn':
        return KNNGroup(nsample, **args)
