"""Drop-in for the reference's KNN/knn_modules.py: `knn` (the native entry) and `myknn(ref, query, k=1)`.

As in the reference (knn_modules.py:11-18) myknn allocates idx with ONE row whatever k is passed, so it always returns
the single nearest reference per query, 1-based, shape [B,1,Q] int64 (callers subtract 1: label_generation.py:58,84).
`knn_k(ref, query, k)` is the general-k call the native entry supports."""
import torch

from .knn_C import knn


def myknn(ref, query, k=1):
    device = ref.device
    ref = ref.float().to(device).contiguous()
    query = query.float().to(device).contiguous()
    inds = torch.empty(query.shape[0], 1, query.shape[2], dtype=torch.int64, device=device)
    knn(ref, query, inds)
    return inds


def knn_k(ref, query, k):
    ref = ref.float().contiguous()
    query = query.float().contiguous()
    inds = torch.empty(query.shape[0], int(k), query.shape[2], dtype=torch.int64, device=ref.device)
    knn(ref, query, inds)
    return inds
