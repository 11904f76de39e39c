"""CPU oracle for the point-cloud operator hot path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers over oracle/gb_oracle.c (the C restatement of the reference kernels; every C
function cites the reference file:line it follows) plus a numpy restatement of
collision_detector.ModelFreeCollisionDetector.detect (collision_detector.py:16-64).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker or the timed CPU baseline.  graspbalance_b200/ never does.

Parity pinning (see DESIGN.md "Oracle"):
  * CUDA ops: pinned by tests/golden/ref_gpu_*.npz -- outputs of the UNMODIFIED reference extensions
    (oracle/_ref/*.so, built by oracle/build_ref.py) run on a B200 by tests/golden/make_golden_gpu.py -- and,
    on the GPU box, live against those same extensions (tests/test_parity_ref_gpu.py).
  * collision: pinned by tests/golden/collision_ref.npz, produced by importing the reference's
    collision_detector.py in the build container (tests/golden/make_golden_collision.py).
  * voxel_down_sample (open3d, un-vendored third party): PARITY UNPINNED.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgb_oracle.so")
    src = os.path.join(_HERE, "gb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libgb_oracle.so"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.gbo_opt_n_threads.restype = ctypes.c_int
        _LIB.gbo_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().gbo_num_threads()


def set_num_threads(t):
    lib().gbo_set_num_threads(int(t))


def opt_n_threads(work_size, cap):
    return lib().gbo_opt_n_threads(int(work_size), int(cap))


def furthest_point_sample(xyz, npoint, variant="A"):
    """xyz [B,N,3] f32 -> idx [B,npoint] i32.  variant A = pointnet2._ext, B = pointnet2_batch_cuda."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32)
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib().gbo_fps(_p(xyz), _p(temp), _p(idx), B, N, int(npoint), 0 if variant == "A" else 1)
    return idx


def gather_operation(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    m = idx.shape[1]
    out = np.empty((B, C, m), dtype=np.float32)
    lib().gbo_gather_fwd(_p(features), _p(idx), _p(out), B, C, N, m)
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().gbo_gather_bwd(_p(grad_out), _p(idx), _p(out), B, C, int(N), m)
    return out


def furthest_point_sample_segments(points, counts, nsamples, variant="A"):
    """The per-object FPS loop of ObjectBalanceSampling (TrainModel/modules.py:201-209): one furthest_point_sample call per
    point set; returns the concatenated set-local indices."""
    points = _f32(points)
    out, first = [], 0
    for c, k in zip(counts, nsamples):
        if k > 0:
            out.append(furthest_point_sample(points[None, first:first + c], k, variant)[0] if c > 0 else np.zeros(k, np.int32))
        first += c
    return np.concatenate(out) if out else np.zeros(0, np.int32)


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((B, m, nsample), dtype=np.int32)
    lib().gbo_ball_query(_p(new_xyz), _p(xyz), _p(idx), B, N, m, ctypes.c_float(radius), int(nsample))
    return idx


def cylinder_query(radius, hmin, hmax, nsample, xyz, new_xyz, rot):
    xyz, new_xyz, rot = _f32(xyz), _f32(new_xyz), _f32(rot)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((B, m, nsample), dtype=np.int32)
    lib().gbo_cylinder_query(_p(new_xyz), _p(xyz), _p(rot), _p(idx), B, N, m, ctypes.c_float(radius),
                             ctypes.c_float(hmin), ctypes.c_float(hmax), int(nsample))
    return idx


def cylinder_query_multi(radius, hmin, hmax_list, nsample, xyz, new_xyz, rot):
    """The loop over hmax_list of GraspWidthGrouping.forward (TrainModel/modules.py:104-113), one cylinder_query per depth;
    stacked along a depth axis: idx [B, m, D, nsample]."""
    return np.stack([cylinder_query(radius, hmin, h, nsample, xyz, new_xyz, rot) for h in hmax_list], axis=2)


def cylinder_query_multi_radius(radii, hmin, hmax_list, nsample, xyz, new_xyz, rot):
    """WidthGroup1..4 of GraspPoseStage2_seed_features_multi_scale.forward (TrainModel/graspbalance.py:104-107): one
    GraspWidthGrouping (= one cylinder_query per depth) per radius; idx [R, B, m, D, nsample]."""
    return np.stack([cylinder_query_multi(r, hmin, hmax_list, nsample, xyz, new_xyz, rot) for r in radii], axis=0)


def grouping_operation(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    _, m, ns = idx.shape
    out = np.empty((B, C, m, ns), dtype=np.float32)
    lib().gbo_group_fwd(_p(features), _p(idx), _p(out), B, C, N, m, ns)
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m, ns = grad_out.shape
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().gbo_group_bwd(_p(grad_out), _p(idx), _p(out), B, C, int(N), m, ns)
    return out


def three_nn_dist2(unknown, known):
    """Native semantics: SQUARED distances + indices (interpolate.cpp:19-45)."""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((B, n, 3), dtype=np.float32)
    idx = np.empty((B, n, 3), dtype=np.int32)
    lib().gbo_three_nn(_p(unknown), _p(known), _p(d2), _p(idx), B, n, m)
    return d2, idx


def three_nn(unknown, known):
    """Python-API semantics: sqrt(dist2), idx (pointnet2_utils.py:82-84)."""
    d2, idx = three_nn_dist2(unknown, known)
    return np.sqrt(d2), idx


def three_nn_weights(unknown, known):
    """three_nn followed by the weight arithmetic of its callers (pointnet2_modules.py:413-416, upsampling.py:69-72,
    graspbalance.py:37-41), every step rounded to fp32 as the torch op does: dist_recip = 1 / (dist + 1e-8);
    weight = dist_recip / sum(dist_recip, dim=2)."""
    dist, idx = three_nn(unknown, known)
    with np.errstate(divide="ignore", invalid="ignore"):
        recip = (np.float32(1.0) / (dist + np.float32(1e-8))).astype(np.float32)
        # torch.sum over the 3-element inner dimension adds (r0 + r2) + r1 (two interleaved accumulators; torch 2.11, CUDA)
        norm = ((recip[..., 0] + recip[..., 2]).astype(np.float32) + recip[..., 1]).astype(np.float32)
        weight = (recip / norm[..., None]).astype(np.float32)
    return dist, idx, weight


def three_interpolate(features, idx, weight):
    features, idx, weight = _f32(features), _i32(idx), _f32(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), dtype=np.float32)
    lib().gbo_three_interp_fwd(_p(features), _p(idx), _p(weight), _p(out), B, C, m, n)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, m), dtype=np.float32)
    lib().gbo_three_interp_bwd(_p(grad_out), _p(idx), _p(weight), _p(out), B, C, n, int(m))
    return out


def knn(ref, query, k):
    """ref [B,D,R], query [B,D,Q] -> idx [B,k,Q] int64, 1-based (knn.cu semantics)."""
    ref, query = _f32(ref), _f32(query)
    B, D, R = ref.shape
    Q = query.shape[2]
    idx = np.empty((B, k, Q), dtype=np.int64)
    lib().gbo_knn(_p(ref), _p(query), _p(idx), B, D, R, Q, int(k))
    return idx


# ------------------------------------------------------------------------------------------------
# collision (collision_detector.py:16-64)
# ------------------------------------------------------------------------------------------------
FINGER_WIDTH = 0.01   # collision_detector.py:8
FINGER_LENGTH = 0.06  # collision_detector.py:9


def collision_thresholds(heights, depths, widths, approach_dist):
    """The ten per-grasp half-space thresholds, evaluated with the reference's own numpy expressions
    (collision_detector.py:26-35) so that they are bit-identical to what `detect` compares against."""
    # the arrays keep their dtype (float32 for a graspnetAPI GraspGroup built from network output): numpy then evaluates the
    # expressions in that dtype, as it does for the reference; the results widen to float64 exactly
    heights = np.asarray(heights)[:, np.newaxis]
    depths = np.asarray(depths)[:, np.newaxis]
    widths = np.asarray(widths)[:, np.newaxis]
    fw, fl = FINGER_WIDTH, FINGER_LENGTH
    thr = np.concatenate([
        -heights / 2, heights / 2,
        depths - fl, depths,
        -(widths / 2 + fw), -widths / 2,
        (widths / 2 + fw), widths / 2,
        depths - fl - fw,
        depths - fl - fw - approach_dist], axis=1)
    return np.ascontiguousarray(thr, dtype=np.float64)


def collision_counts(scene_points, T, R, heights, depths, widths, approach_dist, fma_mode=1):
    pts = np.ascontiguousarray(scene_points, dtype=np.float64)
    T = np.ascontiguousarray(T, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    thr = collision_thresholds(heights, depths, widths, approach_dist)
    G = T.shape[0]
    counts = np.zeros((G, 6), dtype=np.int64)
    lib().gbo_collision_counts(_p(pts), pts.shape[0], _p(T), _p(R), _p(thr), G, int(fma_mode), _p(counts))
    return counts


def collision_finish(counts, heights, depths, widths, voxel_size, approach_dist, collision_thresh=0.05,
                     return_empty_grasp=False, empty_thresh=0.01, return_ious=False):
    """collision_detector.py:43-64: volumes, IoUs, thresholds and the return-shape convention, from the counts."""
    fw, fl = FINGER_WIDTH, FINGER_LENGTH
    heights = np.asarray(heights)[:, np.newaxis]  # dtype kept: the reference's volumes are float32 for float32 grasp groups
    widths = np.asarray(widths)[:, np.newaxis]
    left_right_volume = (heights * fl * fw / (voxel_size ** 3)).reshape(-1)
    bottom_volume = (heights * (widths + 2 * fw) * fw / (voxel_size ** 3)).reshape(-1)
    shifting_volume = (heights * (widths + 2 * fw) * approach_dist / (voxel_size ** 3)).reshape(-1)
    volume = left_right_volume * 2 + bottom_volume + shifting_volume
    global_iou = counts[:, 0] / (volume + 1e-6)
    collision_mask = (global_iou > collision_thresh)
    if not (return_empty_grasp or return_ious):
        return collision_mask
    ret_value = [collision_mask, ]
    if return_empty_grasp:
        inner_volume = (heights * fl * widths / (voxel_size ** 3)).reshape(-1)
        empty_mask = (counts[:, 5] / inner_volume < empty_thresh)
        ret_value.append(empty_mask)
    if return_ious:
        left_iou = counts[:, 1] / (left_right_volume + 1e-6)
        right_iou = counts[:, 2] / (left_right_volume + 1e-6)
        bottom_iou = counts[:, 3] / (bottom_volume + 1e-6)
        shifting_iou = counts[:, 4] / (shifting_volume + 1e-6)
        ret_value.append([global_iou, left_iou, right_iou, bottom_iou, shifting_iou])
    return ret_value


def collision_detect(scene_points, voxel_size, T, R, heights, depths, widths, approach_dist=0.03,
                     collision_thresh=0.05, return_empty_grasp=False, empty_thresh=0.01, return_ious=False,
                     fma_mode=1):
    """C-counts restatement of ModelFreeCollisionDetector.detect on already down-sampled scene_points."""
    approach_dist = max(approach_dist, FINGER_WIDTH)  # collision_detector.py:17
    counts = collision_counts(scene_points, T, R, heights, depths, widths, approach_dist, fma_mode)
    return collision_finish(counts, heights, depths, widths, voxel_size, approach_dist, collision_thresh,
                            return_empty_grasp, empty_thresh, return_ious)


def collision_detect_numpy(scene_points, voxel_size, T, R, heights, depths, widths, approach_dist=0.03,
                           collision_thresh=0.05, return_empty_grasp=False, empty_thresh=0.01, return_ious=False):
    """Whole-array numpy restatement of detect (same temporaries as the reference: [G,N,3] f64 targets and ten
    [G,N] bool masks) -- this is the form bench.py times as the reference's CPU path."""
    approach_dist = max(approach_dist, FINGER_WIDTH)
    fw, fl = FINGER_WIDTH, FINGER_LENGTH
    pts = np.asarray(scene_points, dtype=np.float64)
    h = np.asarray(heights)[:, np.newaxis]
    d = np.asarray(depths)[:, np.newaxis]
    w = np.asarray(widths)[:, np.newaxis]
    t = np.matmul(pts[np.newaxis, :, :] - np.asarray(T)[:, np.newaxis, :], np.asarray(R))
    m1 = (t[:, :, 2] > -h / 2) & (t[:, :, 2] < h / 2)
    m2 = (t[:, :, 0] > d - fl) & (t[:, :, 0] < d)
    m3 = t[:, :, 1] > -(w / 2 + fw)
    m4 = t[:, :, 1] < -w / 2
    m5 = t[:, :, 1] < (w / 2 + fw)
    m6 = t[:, :, 1] > w / 2
    m7 = (t[:, :, 0] <= d - fl) & (t[:, :, 0] > d - fl - fw)
    m8 = (t[:, :, 0] <= d - fl - fw) & (t[:, :, 0] > d - fl - fw - approach_dist)
    left, right = m1 & m2 & m3 & m4, m1 & m2 & m5 & m6
    bottom, shifting = m1 & m3 & m5 & m7, m1 & m3 & m5 & m8
    counts = np.stack([(left | right | bottom | shifting).sum(axis=1), left.sum(axis=1), right.sum(axis=1),
                       bottom.sum(axis=1), shifting.sum(axis=1), (m1 & m2 & (~m4) & (~m6)).sum(axis=1)], axis=1)
    return collision_finish(counts, heights, depths, widths, voxel_size, approach_dist, collision_thresh,
                            return_empty_grasp, empty_thresh, return_ious)


def voxel_down_sample(points, voxel_size):
    """Restatement of open3d.geometry.PointCloud.voxel_down_sample (called at collision_detector.py:13).
    open3d 0.9.0.0 is an un-vendored third-party dependency: PARITY UNPINNED.  Published algorithm: voxel index
    floor((p - (min_bound - 0.5*voxel)) / voxel); each occupied voxel yields the mean of its points (fp64 running sum
    in input order / count).  Output order here is first-occurrence order (open3d's is unordered_map order; `detect`
    is order-invariant)."""
    pts = np.asarray(points, dtype=np.float64)
    if pts.shape[0] == 0:
        return pts.reshape(0, 3)
    min_bound = pts.min(axis=0) - voxel_size * 0.5
    vox = np.floor((pts - min_bound) / voxel_size).astype(np.int64)
    _, first, inv = np.unique(vox, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    nv = first.shape[0]
    sums = np.zeros((nv, 3), dtype=np.float64)
    np.add.at(sums, inv, pts)   # in input order, like the reference's running sum
    cnt = np.bincount(inv, minlength=nv).astype(np.float64)
    out = sums / cnt[:, None]
    order = np.argsort(first, kind="stable")
    return out[order]
