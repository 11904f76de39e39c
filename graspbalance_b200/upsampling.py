"""Drop-in for the reference's ModifiedNetTools/upsampling.py: ThreeNN / three_nn, ThreeInterpolate / three_interpolate,
three_interpolation (upsampling.py:13-74) over pointnet2_batch_cuda's out-parameter calls."""
import torch
from torch.autograd import Function

from . import pointnet2_batch_cuda as pointnet2_cuda


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        pointnet2_cuda.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features, idx, weight):
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        pointnet2_cuda.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        # upsampling.py:56-60 zero-fills and accumulates; the _set entry writes every element, same values
        grad_features = torch.empty([B, c, m], dtype=torch.float32, device=grad_out.device)
        pointnet2_cuda.three_interpolate_grad_set(B, c, n, m, grad_out.detach().contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


def three_interpolation(unknown_xyz, known_xyz, know_feat):
    """upsampling.py:67-74: inverse-distance weights over the three nearest known points.  The neighbour search and the
    weight arithmetic (sqrt, +1e-8, reciprocal, sum, divide -- five elementwise passes in the reference) are one launch
    (gb_three_nn_weights, bit-identical values); three_nn's outputs carry no gradient in the reference either.
    pointnet2_utils.three_interpolation is the single-launch variant that writes nothing but the output."""
    from . import pointnet2_utils as pu
    if (unknown_xyz.is_cuda and unknown_xyz.dtype == torch.float32 and known_xyz.dtype == torch.float32 and unknown_xyz.is_contiguous()
            and known_xyz.is_contiguous()):
        _, idx, weight = pu.three_nn_weights(unknown_xyz, known_xyz)
        return three_interpolate(know_feat, idx, weight)
    dist, idx = three_nn(unknown_xyz, known_xyz)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    return three_interpolate(know_feat, idx, weight)
