"""Drop-in for the reference's collision_detector.py: ModelFreeCollisionDetector(scene_points, voxel_size).detect(...).

Same constructor and `detect` signature, defaults, return conventions (a bare bool array, or a list
[collision_mask, empty_mask?, [5 IoU arrays]?], collision_detector.py:49-64) and fp64 arithmetic.  The grasp x point
occupancy test -- the part that costs the reference ~1.9 s per scene in numpy -- runs on the GPU (libgbops.so,
gb_collision_counts) and returns six integer counts per grasp; the volumes, IoUs and thresholds are then evaluated on the
host with the reference's own numpy expressions, so every returned value is bit-identical given equal counts.

The ten half-space thresholds are also evaluated on the host with the reference's expressions (:26-35), so the kernel
compares against exactly the doubles numpy would.

`__init__` keeps the reference's voxel down-sampling step: it uses open3d when that package is importable (as the
reference does, :11-14); otherwise the restatement of open3d's voxel_down_sample runs on the GPU (voxel_down_sample_gpu:
torch sort + gb_voxel_means, same voxel index and the same sequential fp64 means as the numpy restatement
voxel_down_sample below, which stays for CPU devices).  Parity of this step is unpinned: open3d is an un-vendored
third-party dependency of the reference; `detect` does not depend on the point order.
"""
import numpy as np
import torch

from . import _lib


def voxel_down_sample(points, voxel_size):
    """numpy restatement of open3d.geometry.PointCloud.voxel_down_sample: voxel index floor((p - (min - voxel/2)) /
    voxel); every occupied voxel yields the mean of its points (fp64 sums accumulated in input order)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if pts.shape[0] == 0:
        return pts
    origin = pts.min(axis=0) - 0.5 * voxel_size
    cell = np.floor((pts - origin) / voxel_size).astype(np.int64)
    _, first, inverse = np.unique(cell, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    sums = np.zeros((first.shape[0], 3), dtype=np.float64)
    np.add.at(sums, inverse, pts)
    means = sums / np.bincount(inverse, minlength=first.shape[0]).astype(np.float64)[:, None]
    return means[np.argsort(first, kind="stable")]


def voxel_down_sample_gpu(points_dev, voxel_size):
    """voxel_down_sample on the GPU: points_dev [N,3] f64 CUDA -> [V,3] f64 CUDA, the same voxel index and the same
    sequential fp64 means as the host restatement (voxels come out in key order instead of first-occurrence order; `detect`
    does not depend on the order).  Keys, stable sort and segment boundaries are torch ops; the ordered per-voxel sums are
    gb_voxel_means.  Returns None when a voxel coordinate would not fit the 21-bit key fields (the caller falls back)."""
    pts = points_dev.reshape(-1, 3)
    if pts.shape[0] == 0:
        return pts
    origin = pts.min(dim=0).values - 0.5 * voxel_size
    cell = torch.floor((pts - origin) / voxel_size).to(torch.int64)
    if not bool(((cell >= 0) & (cell < (1 << 21))).all()):  # NaN / inf coordinates or an extent of more than 2M voxels
        return None
    key = (cell[:, 0] << 42) | (cell[:, 1] << 21) | cell[:, 2]
    skey, order = torch.sort(key, stable=True)
    _, counts = torch.unique_consecutive(skey, return_counts=True)
    V = counts.shape[0]
    seg = torch.zeros(V + 1, dtype=torch.int64, device=pts.device)
    torch.cumsum(counts, 0, out=seg[1:])
    out = torch.empty((V, 3), dtype=torch.float64, device=pts.device)
    pts = pts.contiguous()
    _lib.call("gb_voxel_means", pts, pts.data_ptr(), order.data_ptr(), seg.data_ptr(), out.data_ptr(), V)
    return out


def _down_sample(scene_points, voxel_size):
    try:
        import open3d as o3d  # the reference's path (collision_detector.py:11-14)
    except ImportError:
        return voxel_down_sample(scene_points, voxel_size)
    cloud = o3d.geometry.PointCloud()
    cloud.points = o3d.utility.Vector3dVector(scene_points)
    return np.array(cloud.voxel_down_sample(voxel_size).points)


def collision_counts(scene_points_dev, T, R, thr):
    """Device-side core: scene_points_dev [N,3], T [G,3], R [G,3,3], thr [G,10] (all fp64 CUDA tensors, contiguous)
    -> counts [G,6] int64 CUDA tensor {global, left, right, bottom, shifting, inner}."""
    for t, name in ((scene_points_dev, "scene_points"), (T, "translations"), (R, "rotation_matrices"), (thr, "thresholds")):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
            raise RuntimeError(f"{name} must be a contiguous float64 CUDA tensor")
    G = T.shape[0]
    counts = torch.empty((G, 6), dtype=torch.int64, device=T.device)
    _lib.call("gb_collision_counts", T, scene_points_dev.data_ptr(), scene_points_dev.shape[0], T.data_ptr(), R.data_ptr(),
              thr.data_ptr(), G, counts.data_ptr())
    return counts


def collision_counts_batched(scene_points_list, T, R, thr):
    """collision_counts for several scenes in one launch: scene_points_list = per-scene [N'_s,3] f64 CUDA
    tensors; T [S,G,3], R [S,G,3,3], thr [S,G,10] f64 CUDA -> counts [S,G,6] int64.  Row s equals
    collision_counts(scene_points_list[s], T[s], R[s], thr[s])."""
    for t, name in ((T, "translations"), (R, "rotation_matrices"), (thr, "thresholds")):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
            raise RuntimeError(f"{name} must be a contiguous float64 CUDA tensor")
    S, G = T.shape[0], T.shape[1]
    sizes = [int(p.shape[0]) for p in scene_points_list]
    assert len(sizes) == S
    packed = torch.cat([p.reshape(-1, 3) for p in scene_points_list]).contiguous() if S else T.new_zeros((0, 3))
    if packed.dtype != torch.float64 or not packed.is_cuda:
        raise RuntimeError("scene_points must be float64 CUDA tensors")
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64).to(T.device)
    counts = torch.empty((S, G, 6), dtype=torch.int64, device=T.device)
    _lib.call("gb_collision_counts_batched", T, packed.data_ptr(), off.data_ptr(), S, max(sizes, default=0), T.data_ptr(), R.data_ptr(),
              thr.data_ptr(), G, counts.data_ptr())
    return counts


class ModelFreeCollisionDetector():
    def __init__(self, scene_points, voxel_size=0.005, device="cuda"):
        self.finger_width = 0.01
        self.finger_length = 0.06
        self.voxel_size = voxel_size
        self.device = torch.device(device)
        self._scene_dev = None
        try:
            import open3d  # noqa: F401  (the reference's path, collision_detector.py:11-14, when the package exists)
            have_o3d = True
        except ImportError:
            have_o3d = False
        if not have_o3d and self.device.type == "cuda":
            # down-sample on the GPU: the cloud goes up once, the voxel means stay there for detect()
            raw = scene_points if torch.is_tensor(scene_points) else torch.from_numpy(np.ascontiguousarray(scene_points, dtype=np.float64))
            self._scene_dev = voxel_down_sample_gpu(raw.to(self.device, dtype=torch.float64), voxel_size)
        if self._scene_dev is not None:
            self.scene_points = self._scene_dev.cpu().numpy()
        else:
            host = scene_points.detach().cpu().numpy() if torch.is_tensor(scene_points) else scene_points
            self.scene_points = _down_sample(host, voxel_size)
            self._scene_dev = torch.as_tensor(np.ascontiguousarray(self.scene_points, dtype=np.float64)).to(self.device)

    def _thresholds(self, heights, depths, widths, approach_dist):
        fw, fl = self.finger_width, self.finger_length
        return np.ascontiguousarray(np.concatenate([
            -heights / 2, heights / 2,
            depths - fl, depths,
            -(widths / 2 + fw), -widths / 2,
            (widths / 2 + fw), widths / 2,
            depths - fl - fw,
            depths - fl - fw - approach_dist], axis=1), dtype=np.float64)

    def detect(self, grasp_group, approach_dist=0.03, collision_thresh=0.05, return_empty_grasp=False, empty_thresh=0.01,
               return_ious=False):
        approach_dist = max(approach_dist, self.finger_width)
        T = np.ascontiguousarray(grasp_group.translations, dtype=np.float64)
        R = np.ascontiguousarray(grasp_group.rotation_matrices, dtype=np.float64)
        heights = np.asarray(grasp_group.heights, dtype=np.float64)[:, np.newaxis]
        depths = np.asarray(grasp_group.depths, dtype=np.float64)[:, np.newaxis]
        widths = np.asarray(grasp_group.widths, dtype=np.float64)[:, np.newaxis]
        thr = self._thresholds(heights, depths, widths, approach_dist)

        dev = self.device
        counts = collision_counts(self._scene_dev, torch.from_numpy(T).to(dev), torch.from_numpy(R).to(dev),
                                  torch.from_numpy(thr).to(dev)).cpu().numpy()

        fw, fl, v3 = self.finger_width, self.finger_length, self.voxel_size ** 3
        left_right_volume = (heights * fl * fw / v3).reshape(-1)
        bottom_volume = (heights * (widths + 2 * fw) * fw / v3).reshape(-1)
        shifting_volume = (heights * (widths + 2 * fw) * approach_dist / v3).reshape(-1)
        volume = left_right_volume * 2 + bottom_volume + shifting_volume
        global_iou = counts[:, 0] / (volume + 1e-6)
        collision_mask = (global_iou > collision_thresh)
        if not (return_empty_grasp or return_ious):
            return collision_mask
        ret_value = [collision_mask, ]
        if return_empty_grasp:
            inner_volume = (heights * fl * widths / v3).reshape(-1)
            ret_value.append(counts[:, 5] / inner_volume < empty_thresh)
        if return_ious:
            ret_value.append([global_iou,
                              counts[:, 1] / (left_right_volume + 1e-6), counts[:, 2] / (left_right_volume + 1e-6),
                              counts[:, 3] / (bottom_volume + 1e-6), counts[:, 4] / (shifting_volume + 1e-6)])
        return ret_value
