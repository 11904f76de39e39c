"""Drop-in for the reference's native module A, `pointnet2._ext` (PointNet/_ext_src/src/bindings.cpp:12-27).

Same ten function names, argument order, dtypes, layouts and ownership (outputs are allocated here and returned) as the
pybind11 module; the input checks raise RuntimeError with the reference's TORCH_CHECK messages (_ext_src/include/utils.h:
10-30, "CPU not supported" for CPU tensors).  The work is done by libgbops.so through the C ABI (include/gbops.h) on
torch's current stream.  Unlike the reference, a CUDA failure raises instead of calling exit(-1) (cuda_utils.h:38-47).
"""
import torch

from . import _lib


def _contig(x, name):
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")


def _is_float(x, name):
    if x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def _is_int(x, name):
    if x.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")


def _cuda(x, name):
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("CPU not supported")


def gather_points(points, idx):
    """sampling.cpp:20-43.  points [B,C,N] f32, idx [B,M] i32 -> [B,C,M]."""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    if points.is_cuda:
        _cuda(idx, "idx")
    _need_cuda(points)
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    _lib.call("gb_gather_fwd", points, points.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, N, M)
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:45-69.  grad_out [B,C,M], idx [B,M] -> [B,C,n]."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    if grad_out.is_cuda:
        _cuda(idx, "idx")
    _need_cuda(grad_out)
    B, C, M = grad_out.shape
    out = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
    _lib.call("gb_gather_bwd", grad_out, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, int(n), M)
    return out


def furthest_point_sampling(points, nsamples):
    """sampling.cpp:70-91.  points [B,N,3] f32 -> [B,nsamples] i32 (variant A: norm skip, 512-thread tie order)."""
    _contig(points, "points"); _is_float(points, "points")
    _need_cuda(points)
    B, N = points.shape[0], points.shape[1]
    out = torch.empty((B, int(nsamples)), dtype=torch.int32, device=points.device)  # every index is written
    _lib.call("gb_fps", points, points.data_ptr(), None, out.data_ptr(), B, N, int(nsamples), 0)
    return out


def furthest_point_sampling_xyz(points, nsamples, max_cluster=0):
    """furthest_point_sampling + the coordinates of the samples (what gather_points fetches from the transposed cloud,
    pointnet2_modules.py:151-158) in one launch: ([B,nsamples] i32, [B,nsamples,3] f32).  max_cluster > 0: at most that many
    CTAs per scene (background sampling beside other kernels: fewer SMs touched, slower rounds, same picks)."""
    _contig(points, "points"); _is_float(points, "points")
    _need_cuda(points)
    B, N = points.shape[0], points.shape[1]
    out = torch.empty((B, int(nsamples)), dtype=torch.int32, device=points.device)
    new_xyz = torch.empty((B, int(nsamples), 3), dtype=torch.float32, device=points.device)
    if max_cluster:
        _lib.call("gb_fps_xyz_hint", points, points.data_ptr(), None, out.data_ptr(), new_xyz.data_ptr(), B, N, int(nsamples), 0, int(max_cluster))
    else:
        _lib.call("gb_fps_xyz", points, points.data_ptr(), None, out.data_ptr(), new_xyz.data_ptr(), B, N, int(nsamples), 0)
    return out, new_xyz


def furthest_point_sampling_segments(points, seg, max_n, max_m, total_out):
    """Segmented FPS (gb_fps_segments): points [total,3] f32 packed segments; seg [S,4] i32 CUDA rows (first point, points,
    samples, first output slot).  Returns segment-local indices [total_out] i32, each segment sampled exactly as
    furthest_point_sampling samples it alone."""
    _contig(points, "points"); _is_float(points, "points"); _contig(seg, "seg"); _is_int(seg, "seg")
    _need_cuda(points); _cuda(seg, "seg")
    out = torch.zeros((int(total_out),), dtype=torch.int32, device=points.device)
    if seg.shape[0]:
        _lib.call("gb_fps_segments", points, points.data_ptr(), seg.data_ptr(), out.data_ptr(), None, int(seg.shape[0]), int(max_n), int(max_m), 0)
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:19-45.  unknowns [B,n,3], knows [B,m,3] -> [dist2 [B,n,3] f32 (squared), idx [B,n,3] i32]."""
    _contig(unknowns, "unknowns"); _contig(knows, "knows"); _is_float(unknowns, "unknowns"); _is_float(knows, "knows")
    if unknowns.is_cuda:
        _cuda(knows, "knows")
    _need_cuda(unknowns)
    B, n = unknowns.shape[0], unknowns.shape[1]
    m = knows.shape[1]
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    _lib.call("gb_three_nn", unknowns, unknowns.data_ptr(), knows.data_ptr(), dist2.data_ptr(), idx.data_ptr(), B, n, m)
    return [dist2, idx]


def three_nn_weights(unknowns, knows):
    """three_nn + inverse-distance weights in one launch: (dist [B,n,3] = sqrt of the squared distances, idx [B,n,3] i32,
    weight [B,n,3]) with weight = (1/(dist+1e-8)) / sum_k(1/(dist_k+1e-8)) (pointnet2_modules.py:413-416)."""
    _contig(unknowns, "unknowns"); _contig(knows, "knows"); _is_float(unknowns, "unknowns"); _is_float(knows, "knows")
    if unknowns.is_cuda:
        _cuda(knows, "knows")
    _need_cuda(unknowns)
    B, n, m = unknowns.shape[0], unknowns.shape[1], knows.shape[1]
    dist = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    weight = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    _lib.call("gb_three_nn_weights", unknowns, unknowns.data_ptr(), knows.data_ptr(), dist.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, n, m)
    return dist, idx, weight


def three_interpolation(unknowns, knows, points, keep_for_backward):
    """three_nn + weights + three_interpolate in one launch (gb_three_interpolation): unknowns [B,n,3], knows
    [B,m,3], points [B,C,m] -> out [B,C,n]; idx and weight [B,n,3] are written only when keep_for_backward (else None).
    Returns None when the fused kernel does not take the shape (the caller runs the two launches)."""
    for t, name in ((unknowns, "unknowns"), (knows, "knows"), (points, "points")):
        _contig(t, name); _is_float(t, name)
    if unknowns.is_cuda:
        _cuda(knows, "knows"); _cuda(points, "points")
    _need_cuda(unknowns)
    B, n, m = unknowns.shape[0], unknowns.shape[1], knows.shape[1]
    C = points.shape[1]
    if m < 1 or m > 4096 or n == 0 or B == 0:
        return None
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=points.device) if keep_for_backward else None
    weight = torch.empty((B, n, 3), dtype=torch.float32, device=points.device) if keep_for_backward else None
    _lib.call("gb_three_interpolation", points, unknowns.data_ptr(), knows.data_ptr(), points.data_ptr(), out.data_ptr(),
              None if idx is None else idx.data_ptr(), None if weight is None else weight.data_ptr(), B, C, n, m)
    return out, idx, weight


def three_interpolate(points, idx, weight):
    """interpolate.cpp:47-75.  points [B,C,m], idx/weight [B,n,3] -> [B,C,n]."""
    _contig(points, "points"); _contig(idx, "idx"); _contig(weight, "weight")
    _is_float(points, "points"); _is_int(idx, "idx"); _is_float(weight, "weight")
    if points.is_cuda:
        _cuda(idx, "idx"); _cuda(weight, "weight")
    _need_cuda(points)
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    _lib.call("gb_three_interp_fwd", points, points.data_ptr(), idx.data_ptr(), weight.data_ptr(), out.data_ptr(), B, C, m, n)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:76-104.  grad_out [B,C,n] -> [B,C,m]."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _contig(weight, "weight")
    _is_float(grad_out, "grad_out"); _is_int(idx, "idx"); _is_float(weight, "weight")
    if grad_out.is_cuda:
        _cuda(idx, "idx"); _cuda(weight, "weight")
    _need_cuda(grad_out)
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)  # fully written by the _set entry
    _lib.call("gb_three_interp_bwd_set", grad_out, grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), out.data_ptr(), B, C, n, int(m))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:13-37.  new_xyz [B,m,3], xyz [B,N,3] -> idx [B,m,nsample] i32."""
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz")
    if new_xyz.is_cuda:
        _cuda(xyz, "xyz")
    _need_cuda(new_xyz)
    B, m = new_xyz.shape[0], new_xyz.shape[1]
    N = xyz.shape[1]
    idx = torch.empty((B, m, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    _lib.call("gb_ball_query", new_xyz, new_xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(), B, N, m, float(radius), int(nsample))
    return idx


def cylinder_query(new_xyz, xyz, rot, radius, hmin, hmax, nsample):
    """cylinder_query.cpp:10-47.  rot [B,m,9] row-major 3x3 per query."""
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _contig(rot, "rot")
    _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz"); _is_float(rot, "rot")
    if new_xyz.is_cuda:
        _cuda(xyz, "xyz"); _cuda(rot, "rot")
    _need_cuda(new_xyz)
    B, m = new_xyz.shape[0], new_xyz.shape[1]
    N = xyz.shape[1]
    idx = torch.empty((B, m, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    _lib.call("gb_cylinder_query", new_xyz, new_xyz.data_ptr(), xyz.data_ptr(), rot.data_ptr(), idx.data_ptr(), B, N, m, float(radius), float(hmin), float(hmax), int(nsample))
    return idx


def cylinder_query_multi(new_xyz, xyz, rot, radius, hmin, hmax_list, nsample):
    """The depth loop of GraspWidthGrouping (TrainModel/modules.py:104-113) in one scan: len(hmax_list) <= 4 nested
    cylinders per query.  Returns idx [B,m,D,nsample] i32 with idx[:, :, d] == cylinder_query(..., hmax_list[d], ...)."""
    import ctypes
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _contig(rot, "rot")
    _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz"); _is_float(rot, "rot")
    if new_xyz.is_cuda:
        _cuda(xyz, "xyz"); _cuda(rot, "rot")
    _need_cuda(new_xyz)
    D = len(hmax_list)
    if not 1 <= D <= 4:
        raise RuntimeError("cylinder_query_multi takes 1 to 4 depths")
    B, m = new_xyz.shape[0], new_xyz.shape[1]
    N = xyz.shape[1]
    idx = torch.empty((B, m, D, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    hm = (ctypes.c_float * D)(*[float(h) for h in hmax_list])
    _lib.call("gb_cylinder_query_multi", new_xyz, new_xyz.data_ptr(), xyz.data_ptr(), rot.data_ptr(), idx.data_ptr(), B, N, m, float(radius),
              float(hmin), hm, D, int(nsample))
    return idx


def cylinder_query_multi_radius(new_xyz, xyz, rot, radii, hmin, hmax_list, nsample):
    """The four GraspWidthGrouping modules of graspbalance.py:104-107 (same seeds, rotations, hmin and depths; four radii)
    in one scan.  Returns idx [R,B,m,D,nsample] i32 with idx[k, :, :, d] == cylinder_query(..., radii[k], hmin, hmax_list[d], ...)."""
    import ctypes
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _contig(rot, "rot")
    _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz"); _is_float(rot, "rot")
    if new_xyz.is_cuda:
        _cuda(xyz, "xyz"); _cuda(rot, "rot")
    _need_cuda(new_xyz)
    R, D = len(radii), len(hmax_list)
    if not (1 <= D <= 4 and 1 <= R <= 4):
        raise RuntimeError("cylinder_query_multi_radius takes 1 to 4 radii and 1 to 4 depths")
    B, m = new_xyz.shape[0], new_xyz.shape[1]
    N = xyz.shape[1]
    idx = torch.empty((R, B, m, D, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    rr = (ctypes.c_float * R)(*[float(r) for r in radii])
    hm = (ctypes.c_float * D)(*[float(h) for h in hmax_list])
    _lib.call("gb_cylinder_query_multi_radius", new_xyz, new_xyz.data_ptr(), xyz.data_ptr(), rot.data_ptr(), idx.data_ptr(), B, N, m, rr, R,
              float(hmin), hm, D, int(nsample))
    return idx


def group_points(points, idx):
    """group_points.cpp:21-47.  points [B,C,N], idx [B,npoints,nsample] -> [B,C,npoints,nsample]."""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    if points.is_cuda:
        _cuda(idx, "idx")
    _need_cuda(points)
    B, C, N = points.shape
    npoints, nsample = idx.shape[1], idx.shape[2]
    out = torch.empty((B, C, npoints, nsample), dtype=torch.float32, device=points.device)
    _lib.call("gb_group_fwd", points, points.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, N, npoints, nsample)
    return out


def group_points_max(points, idx, need_arg=True):
    """group_points followed by the max over nsample in one pass (gb_group_max_fwd): points [B,C,N], idx [B,npoint,nsample]
    -> (out [B,C,npoint], arg [B,C,npoint] i32 source index of each maximum, or None).  Returns None for rows too long for
    the kernel's shared-memory staging (the caller runs group_points + max_pool2d)."""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    if points.is_cuda:
        _cuda(idx, "idx")
    _need_cuda(points)
    B, C, N = points.shape
    npoint, nsample = idx.shape[1], idx.shape[2]
    if N * 16 > 200 * 1024 or nsample < 1:
        return None
    out = torch.empty((B, C, npoint), dtype=torch.float32, device=points.device)
    arg = torch.empty((B, C, npoint), dtype=torch.int32, device=points.device) if need_arg else None
    _lib.call("gb_group_max_fwd", points, points.data_ptr(), idx.data_ptr(), out.data_ptr(), None if arg is None else arg.data_ptr(),
              B, C, N, npoint, nsample)
    return out, arg


def group_points_max_grad(grad_out, arg, n):
    """Backward of group_points_max: grad_out [B,C,npoint] scattered to the winning sources -> [B,C,n]."""
    _contig(grad_out, "grad_out"); _contig(arg, "arg"); _is_float(grad_out, "grad_out"); _is_int(arg, "arg")
    _need_cuda(grad_out)
    B, C, npoint = grad_out.shape
    out = torch.zeros((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    _lib.call("gb_group_max_bwd", grad_out, grad_out.data_ptr(), arg.data_ptr(), out.data_ptr(), B, C, int(n), npoint)
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:49-75.  grad_out [B,C,npoints,nsample] -> [B,C,n]."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    if grad_out.is_cuda:
        _cuda(idx, "idx")
    _need_cuda(grad_out)
    B, C = grad_out.shape[0], grad_out.shape[1]
    npoints, nsample = idx.shape[1], idx.shape[2]
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)  # fully written by the _set entry
    _lib.call("gb_group_bwd_set", grad_out, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, int(n), npoints, nsample)
    return out
