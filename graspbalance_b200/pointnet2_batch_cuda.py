"""Drop-in for the reference's native module B, `pointnet2_batch_cuda` (pointnet2_batch/src/pointnet2_api.cpp:10-24).

Same nine out-parameter functions with all sizes passed explicitly; outputs are written into caller-allocated tensors.
The reference launches on the legacy default stream (e.g. ball_query_gpu.cu:52); this module launches on torch's current
stream, which is the same stream whenever the caller has not switched streams (DESIGN.md "Streams").  The reference
checks only ball_query's inputs (ball_query.cpp:7-19, fprintf + exit(-1)); here every function checks that its tensors
are contiguous CUDA tensors of the right dtype and raises RuntimeError instead of exiting.
"""
import torch

from . import _lib


def _chk(t, name, dtype):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}")


_f32, _i32 = torch.float32, torch.int32


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    """ball_query.cpp:22-32.  idx [b,m,nsample] i32 is fully written (the reference needs it zero-filled, group.py:136)."""
    _chk(new_xyz, "new_xyz_tensor", _f32); _chk(xyz, "xyz_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _lib.call("gb_ball_query", xyz, new_xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(), b, n, m, float(radius), int(nsample))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    """group_points.cpp:21-30."""
    _chk(points, "points_tensor", _f32); _chk(idx, "idx_tensor", _i32); _chk(out, "out_tensor", _f32)
    _lib.call("gb_group_fwd", points, points.data_ptr(), idx.data_ptr(), out.data_ptr(), b, c, n, npoints, nsample)
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    """group_points.cpp:8-19.  Accumulates into grad_points (callers zero it, group.py:83)."""
    _chk(grad_out, "grad_out_tensor", _f32); _chk(idx, "idx_tensor", _i32); _chk(grad_points, "grad_points_tensor", _f32)
    _lib.call("gb_group_bwd", grad_out, grad_out.data_ptr(), idx.data_ptr(), grad_points.data_ptr(), b, c, n, npoints, nsample)
    return 1


def group_points_grad_set(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    """Not in the reference module: group_points_grad_wrapper into an UNINITIALISED grad_points (fully overwritten), used by
    graspbalance_b200.group.GroupingOperation.backward to skip the zero fill of group.py:83 and its read-back."""
    _chk(grad_out, "grad_out_tensor", _f32); _chk(idx, "idx_tensor", _i32); _chk(grad_points, "grad_points_tensor", _f32)
    _lib.call("gb_group_bwd_set", grad_out, grad_out.data_ptr(), idx.data_ptr(), grad_points.data_ptr(), b, c, n, npoints, nsample)
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    """sampling.cpp:9-18."""
    _chk(points, "points_tensor", _f32); _chk(idx, "idx_tensor", _i32); _chk(out, "out_tensor", _f32)
    _lib.call("gb_gather_fwd", points, points.data_ptr(), idx.data_ptr(), out.data_ptr(), b, c, n, npoints)
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    """sampling.cpp:21-30.  Accumulates into grad_points."""
    _chk(grad_out, "grad_out_tensor", _f32); _chk(idx, "idx_tensor", _i32); _chk(grad_points, "grad_points_tensor", _f32)
    _lib.call("gb_gather_bwd", grad_out, grad_out.data_ptr(), idx.data_ptr(), grad_points.data_ptr(), b, c, n, npoints)
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    """sampling.cpp:32-41.  Variant B: no norm skip, 1024-thread tie order; temp [b,n] holds the running distances
    (filled with 1e10 by the caller, subsample.py:77) and receives their final values, as in the reference."""
    _chk(points, "points_tensor", _f32); _chk(temp, "temp_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _lib.call("gb_fps", points, points.data_ptr(), temp.data_ptr(), idx.data_ptr(), b, n, m, 1)
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    """interpolate.cpp:14-22.  dist2 receives SQUARED distances."""
    _chk(unknown, "unknown_tensor", _f32); _chk(known, "known_tensor", _f32)
    _chk(dist2, "dist2_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _lib.call("gb_three_nn", unknown, unknown.data_ptr(), known.data_ptr(), dist2.data_ptr(), idx.data_ptr(), b, n, m)


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    """interpolate.cpp:25-36."""
    _chk(points, "points_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _chk(weight, "weight_tensor", _f32); _chk(out, "out_tensor", _f32)
    _lib.call("gb_three_interp_fwd", points, points.data_ptr(), idx.data_ptr(), weight.data_ptr(), out.data_ptr(), b, c, m, n)


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    """interpolate.cpp:38-51.  Accumulates into grad_points."""
    _chk(grad_out, "grad_out_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _chk(weight, "weight_tensor", _f32); _chk(grad_points, "grad_points_tensor", _f32)
    _lib.call("gb_three_interp_bwd", grad_out, grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), grad_points.data_ptr(), b, c, n, m)


def three_interpolate_grad_set(b, c, n, m, grad_out, idx, weight, grad_points):
    """Not in the reference module: three_interpolate_grad_wrapper into an UNINITIALISED grad_points (fully overwritten)."""
    _chk(grad_out, "grad_out_tensor", _f32); _chk(idx, "idx_tensor", _i32)
    _chk(weight, "weight_tensor", _f32); _chk(grad_points, "grad_points_tensor", _f32)
    _lib.call("gb_three_interp_bwd_set", grad_out, grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), grad_points.data_ptr(), b, c, n, m)
