"""Drop-in for the reference's native module C, `KNN._C` (KNN/Pytorch_CUDA_KNN/vision.cpp:3-5, knn.h:11-59).

`knn(ref, query, idx)` fills the caller's idx [B,k,Q] int64 with the 1-based indices of the k nearest reference points
of every query, ascending by (distance, index), and returns 1.  ref [B,D,R] and query [B,D,Q] are channel-first fp32.
CUDA tensors only: the reference's CPU path (cpu/knn_cpu.cpp) is not reproduced -- there is no CPU fallback here.
"""
import torch

from . import _lib


def knn(ref, query, idx):
    for t, name, dt in ((ref, "ref", torch.float32), (query, "query", torch.float32), (idx, "idx", torch.int64)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor (graspbalance_b200 has no CPU path)")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
        if t.dtype != dt:
            raise RuntimeError(f"{name} must have dtype {dt}")
    B, D, R = ref.shape
    Q = query.shape[2]
    k = idx.shape[1]
    if query.shape[0] != B or query.shape[1] != D or idx.shape[0] != B or idx.shape[2] != Q:
        raise RuntimeError("knn: inconsistent shapes")
    _lib.call("gb_knn", ref, ref.data_ptr(), query.data_ptr(), idx.data_ptr(), B, D, R, Q, k)
    return 1
