"""Drop-in for the reference's collision_detector.py: ModelFreeCollisionDetector(scene_points, voxel_size).detect(...).

Same constructor and `detect` signature, defaults, return conventions (a bare bool array, or a list
[collision_mask, empty_mask?, [5 IoU arrays]?], collision_detector.py:49-64) and arithmetic.  All of `detect` runs on the
GPU in one call (libgbops.so, gb_collision_detect): the ten half-space thresholds, the grasp x point occupancy test -- the
part that costs the reference ~1.9 s per scene in numpy --, the gripper volumes, the IoUs and the masks.  Thresholds and
volumes are evaluated in the dtype of the grasp arrays (float32 for a graspnetAPI GraspGroup built from network output,
float64 otherwise) exactly as numpy evaluates the reference's expressions; comparisons and IoUs are fp64.  A host call
costs one upload of the grasp rows and one download of the masks; grasp arrays that already live on the device (the
[Ns,17] tensors of modules.pred_decode) never touch the host: `detect_device`.

`__init__` keeps the reference's voxel down-sampling step: it uses open3d when that package is importable (as the
reference does, :11-14); otherwise the restatement of open3d's voxel_down_sample runs on the GPU (voxel_down_sample_gpu:
torch sort + gb_voxel_means).  Parity of this step is unpinned: open3d is an un-vendored third-party dependency of the
reference; `detect` does not depend on the point order.  There is no CPU path: the device must be CUDA.
"""
import ctypes

import numpy as np
import torch

from . import _lib

# column offsets (translation, rotation, height, depth, width) and row length of the two grasp-row layouts
_PACKED = (0, 3, 12, 13, 14, 15)      # [G,15] built by detect() from a duck-typed GraspGroup
_GRASP_ARRAY = (13, 4, 2, 3, 1, 17)   # graspnetAPI / pred_decode: score, width, height, depth, R(9), T(3), object id


def voxel_down_sample_gpu(points_dev, voxel_size):
    """voxel_down_sample on the GPU: points_dev [N,3] f64 CUDA -> [V,3] f64 CUDA, the same voxel index and the same
    sequential fp64 means as open3d's per-voxel running sums (voxels come out in key order; `detect` does not depend on the
    order).  Restated from the published algorithm: voxel index floor((p - (min - voxel/2)) / voxel), mean of the voxel's
    points accumulated in input order.  Keys, stable sort and segment boundaries are torch ops; the ordered per-voxel sums are
    gb_voxel_means.  Raises when a voxel coordinate does not fit the 21-bit key fields (NaN / inf coordinates or an extent of
    more than 2M voxels per axis)."""
    pts = points_dev.reshape(-1, 3)
    if pts.shape[0] == 0:
        return pts
    origin = pts.min(dim=0).values - 0.5 * voxel_size
    cell = torch.floor((pts - origin) / voxel_size).to(torch.int64)
    if not bool(((cell >= 0) & (cell < (1 << 21))).all()):
        raise RuntimeError("voxel_down_sample_gpu: non-finite coordinates or more than 2^21 voxels along an axis")
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


_OFFSETS = {}  # (device, scene sizes) -> device tensor of packed-array offsets (a step's scenes keep their sizes)


def collision_counts_batched(scene_points_list, T, R, thr):
    """collision_counts for several scenes in one launch: scene_points_list = per-scene [N'_s,3] f64 CUDA
    tensors; T [S,G,3], R [S,G,3,3], thr [S,G,10] f64 CUDA -> counts [S,G,6] int64.  Row s equals
    collision_counts(scene_points_list[s], T[s], R[s], thr[s])."""
    for t, name in ((T, "translations"), (R, "rotation_matrices"), (thr, "thresholds")):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
            raise RuntimeError(f"{name} must be a contiguous float64 CUDA tensor")
    S, G = T.shape[0], T.shape[1]
    sizes = tuple(int(p.shape[0]) for p in scene_points_list)
    assert len(sizes) == S
    packed = torch.cat([p.reshape(-1, 3) for p in scene_points_list]).contiguous() if S else T.new_zeros((0, 3))
    if packed.dtype != torch.float64 or not packed.is_cuda:
        raise RuntimeError("scene_points must be float64 CUDA tensors")
    key = (T.device, sizes)
    off = _OFFSETS.get(key)
    if off is None:  # one small upload per distinct batch layout (not per step: a captured CUDA graph replays without it)
        if len(_OFFSETS) > 64:
            _OFFSETS.clear()
        off = _OFFSETS[key] = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64).to(T.device)
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
        if self.device.type != "cuda":
            raise RuntimeError("ModelFreeCollisionDetector: CPU not supported (graspbalance_b200 has no CPU path)")
        try:
            import open3d as o3d  # the reference's own down-sampling (collision_detector.py:11-14) when the package exists
        except ImportError:
            o3d = None
        if o3d is not None:
            host = scene_points.detach().cpu().numpy() if torch.is_tensor(scene_points) else scene_points
            cloud = o3d.geometry.PointCloud()
            cloud.points = o3d.utility.Vector3dVector(host)
            self._scene_dev = torch.as_tensor(np.ascontiguousarray(np.array(cloud.voxel_down_sample(voxel_size).points),
                                                                   dtype=np.float64)).to(self.device)
        else:
            # down-sample on the GPU: the cloud goes up once, the voxel means stay there for detect()
            raw = scene_points if torch.is_tensor(scene_points) else torch.from_numpy(np.ascontiguousarray(scene_points, dtype=np.float64))
            self._scene_dev = voxel_down_sample_gpu(raw.to(self.device, dtype=torch.float64), voxel_size)
        self._scene_host = None

    @property
    def scene_points(self):
        """[N',3] float64 numpy array, as the reference attribute (copied from the device on first use)."""
        if self._scene_host is None:
            self._scene_host = self._scene_dev.cpu().numpy()
        return self._scene_host

    def _params(self, approach_dist, collision_thresh, empty_thresh):
        approach_dist = max(approach_dist, self.finger_width)
        return (ctypes.c_double * 7)(self.finger_width, self.finger_length, approach_dist, self.voxel_size ** 3, 2 * self.finger_width,
                                     collision_thresh, empty_thresh)

    def detect_device(self, rows, layout=_GRASP_ARRAY, approach_dist=0.03, collision_thresh=0.05, return_empty_grasp=False,
                      empty_thresh=0.01, return_ious=False, return_counts=False):
        """`detect` on grasp rows that live on the device: rows [G, row_len] float32 / float64 CUDA tensor (contiguous), by
        default in the graspnetAPI / pred_decode layout [score, width, height, depth, R(9), T(3), object id].  Returns CUDA
        tensors with the reference's conventions (bool masks, float64 IoUs); nothing is copied to the host."""
        if not (torch.is_tensor(rows) and rows.is_cuda and rows.is_contiguous() and rows.dim() == 2 and
                rows.dtype in (torch.float32, torch.float64)):
            raise RuntimeError("grasp rows must be a contiguous 2-D float32 / float64 CUDA tensor")
        oT, oR, oH, oD, oW, row_len = layout
        if rows.shape[1] != row_len:
            raise RuntimeError(f"grasp rows must have {row_len} columns")
        G = rows.shape[0]
        dev = rows.device
        mask = torch.empty(G, dtype=torch.uint8, device=dev)
        empty = torch.empty(G, dtype=torch.uint8, device=dev) if return_empty_grasp else None
        ious = torch.empty((5, G), dtype=torch.float64, device=dev) if return_ious else None
        counts = torch.empty((G, 6), dtype=torch.int64, device=dev) if return_counts else None
        _lib.call("gb_collision_detect", rows, self._scene_dev.data_ptr(), self._scene_dev.shape[0], rows.data_ptr(),
                  1 if rows.dtype == torch.float64 else 0, G, row_len, oT, oR, oH, oD, oW,
                  self._params(approach_dist, collision_thresh, empty_thresh), mask.data_ptr(),
                  None if empty is None else empty.data_ptr(), None if ious is None else ious.data_ptr(),
                  None if counts is None else counts.data_ptr())
        ret = [mask.bool()]
        if return_empty_grasp:
            ret.append(empty.bool())
        if return_ious:
            ret.append([ious[i] for i in range(5)])
        if return_counts:
            ret.append(counts)
        return ret[0] if len(ret) == 1 else ret

    def detect(self, grasp_group, approach_dist=0.03, collision_thresh=0.05, return_empty_grasp=False, empty_thresh=0.01,
               return_ious=False):
        """collision_detector.py:16-64.  grasp_group: anything with .translations [G,3], .rotation_matrices [G,3,3], .heights,
        .depths, .widths [G] (numpy, float32 or float64 -- the reference's arithmetic follows their dtype -- or CUDA tensors);
        a graspnetAPI GraspGroup's own [G,17] `grasp_group_array` is uploaded as it is.  numpy in -> numpy out (one upload, one
        download); CUDA tensors in -> CUDA tensors out."""
        arr = getattr(grasp_group, "grasp_group_array", None)
        on_device = torch.is_tensor(grasp_group.translations) and grasp_group.translations.is_cuda
        if on_device:
            G = grasp_group.translations.shape[0]
            cols = [grasp_group.translations.reshape(G, 3), grasp_group.rotation_matrices.reshape(G, 9),
                    grasp_group.heights.reshape(G, 1), grasp_group.depths.reshape(G, 1), grasp_group.widths.reshape(G, 1)]
            dt = torch.float32 if all(c.dtype == torch.float32 for c in cols) else torch.float64
            rows, layout = torch.cat([c.to(dt) for c in cols], dim=1).contiguous(), _PACKED
        elif isinstance(arr, np.ndarray) and arr.ndim == 2 and arr.shape[1] == 17 and arr.dtype in (np.float32, np.float64):
            rows, layout = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device), _GRASP_ARRAY
        else:
            parts = [np.asarray(grasp_group.translations), np.asarray(grasp_group.rotation_matrices), np.asarray(grasp_group.heights),
                     np.asarray(grasp_group.depths), np.asarray(grasp_group.widths)]
            G = parts[0].shape[0]
            # numpy's own promotion: float32 only when every array is float32 (then the reference computes in float32 too)
            dt = np.float32 if all(p.dtype == np.float32 for p in parts) else np.float64
            packed = np.empty((G, 15), dtype=dt)
            packed[:, 0:3], packed[:, 3:12] = parts[0].reshape(G, 3), parts[1].reshape(G, 9)
            packed[:, 12], packed[:, 13], packed[:, 14] = parts[2].reshape(G), parts[3].reshape(G), parts[4].reshape(G)
            rows, layout = torch.from_numpy(packed).to(self.device), _PACKED
        ret = self.detect_device(rows, layout, approach_dist, collision_thresh, return_empty_grasp, empty_thresh, return_ious)
        if on_device:
            return ret
        if not (return_empty_grasp or return_ious):
            return ret.cpu().numpy()
        out = [ret[0].cpu().numpy()]
        k = 1
        if return_empty_grasp:
            out.append(ret[k].cpu().numpy())
            k += 1
        if return_ious:
            ious = torch.stack(ret[k]).cpu().numpy()
            out.append([ious[i] for i in range(5)])
        return out
