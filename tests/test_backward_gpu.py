"""GPU parity tests of the warp-private group backward (csrc/scatter_private.cu) -- the kernel that carries the largest
share of a pipeline step -- at the backbone's shapes on real ball-query indices against the UNMODIFIED reference module B
(pointnet2_batch/src/group_points_gpu.cu:9-37), and on index tensors that leave its fast path against an fp64 scatter_add.
Gradients within 1e-5 relative (north_star); run-to-run bit-identical (fixed summation order)."""
import numpy as np
import pytest
import torch

from graspbalance_b200 import _ext as gb_a
from graspbalance_b200 import _lib, scenes
from graspbalance_b200 import pointnet2_batch_cuda as gb_b

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-5


@pytest.fixture()
def force_private():
    _lib.set_tuning("scatter_mode", 16)  # the private path whatever the number of (scene, channel group) tasks
    yield
    _lib.set_tuning("scatter_mode", 0)


import contextlib


@contextlib.contextmanager
def force_sorted():
    _lib.set_tuning("scatter_mode", 8)  # never the private path: the sorted backward (seg_sort_dense + seg_dense) answers
    try:
        yield
    finally:
        _lib.set_tuning("scatter_mode", 0)


def _close(got, want):
    scale = max(want.abs().max().item(), 1.0)
    err = (got.double() - want.double()).abs().max().item()
    assert err <= GRAD_RTOL * scale, f"max abs diff {err} vs scale {scale}"


def _levels(dev, B):
    xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
    fidx = gb_a.furthest_point_sampling(xyz, 2048).long()
    return torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()  # FPS prefix property: [:, :k] = FPS to k


# (label, targets n, queries m, nsample, C, radius): the InvResMLP / SA backward launches of BASELINE config 5 (drp.py:161-247)
BACKBONE = [("irm0", 2048, 2048, 64, 128, 0.08), ("irm1", 1024, 1024, 32, 256, 0.2), ("irm2", 512, 512, 16, 256, 0.4),
            ("irm3", 256, 256, 16, 256, 0.6), ("sa2", 2048, 1024, 32, 128, 0.1), ("sa3", 1024, 512, 16, 256, 0.2)]


@pytest.mark.parametrize("label,n,m,ns,C,r", BACKBONE)
def test_group_bwd_backbone_shapes_vs_reference_b(dev, ref_b, force_private, label, n, m, ns, C, r):
    B = 3
    lv0 = _levels(dev, B)
    tgt, qry = lv0[:, :n].contiguous(), lv0[:, :m].contiguous()
    idx = gb_a.ball_query(qry, tgt, r, ns)
    g = torch.Generator(device="cpu").manual_seed(3)
    gout = torch.randn((B, C, m, ns), generator=g).to(dev)
    want = torch.zeros((B, C, n), device=dev)
    ref_b.group_points_grad_wrapper(B, C, n, m, ns, gout, idx, want)
    l0 = _lib.launch_count()
    got = gb_a.group_points_grad(gout, idx, n)
    assert _lib.launch_count() - l0 == 1, "expected the single-launch private path"
    _close(got, want)
    assert torch.equal(got, gb_a.group_points_grad(gout, idx, n)), "backward must be bit-reproducible"
    acc = torch.full((B, C, n), 0.25, device=dev)  # module B accumulates into the caller's tensor
    gb_b.group_points_grad_wrapper(B, C, n, m, ns, gout, idx, acc)
    _close(acc - 0.25, want)
    # and the sorted path (the fallback for shapes the private path does not take) still agrees
    _lib.set_tuning("scatter_mode", 8)
    _close(gb_a.group_points_grad(gout, idx, n), want)


def _make_idx(kind, B, n, m, ns, g):
    if kind == "random":
        return torch.randint(0, n, (B, m, ns), generator=g, dtype=torch.int32)
    if kind == "oob":
        return torch.randint(-5, n + 5, (B, m, ns), generator=g, dtype=torch.int32)
    rows = []
    for _ in range(B * m):
        cnt = ns if kind == "sorted" else int(torch.randint(1, ns + 1, (1,), generator=g))
        r = torch.randperm(n, generator=g)[:cnt].sort().values
        if kind == "knn":  # distinct but not ascending
            r = r[torch.randperm(cnt, generator=g)]
        rows.append(torch.cat([r, r[:1].expand(ns - cnt)]))
    return torch.stack(rows).reshape(B, m, ns).int()


@pytest.mark.parametrize("B,C,n,m,ns,kind", [(3, 7, 300, 64, 32, "random"), (2, 4, 2048, 128, 64, "random"), (2, 9, 1000, 50, 16, "random"),
                                             (2, 17, 500, 40, 8, "random"), (2, 8, 1024, 96, 32, "padded"), (2, 8, 2048, 33, 64, "padded"),
                                             (2, 5, 700, 77, 16, "padded"), (2, 6, 512, 64, 32, "sorted"), (2, 12, 256, 64, 64, "oob"),
                                             (1, 4, 2400, 10, 128, "random"), (2, 130, 1024, 64, 32, "sorted"), (2, 8, 640, 48, 32, "knn"),
                                             (2, 8, 640, 48, 16, "knn")])
def test_group_bwd_private_any_index_tensor(dev, force_private, B, C, n, m, ns, kind):
    g = torch.Generator(device="cpu").manual_seed(1)
    idx = _make_idx(kind, B, n, m, ns, g).to(dev)
    gout = torch.randn((B, C + 3, m, ns), generator=g).to(dev)
    safe = idx.long().reshape(B, 1, m * ns)
    safe = torch.where((safe < 0) | (safe >= n), torch.full_like(safe, n), safe)  # out-of-range targets are dropped
    want = torch.zeros((B, C, n + 1), dtype=torch.float64, device=dev)
    want.scatter_add_(2, safe.expand(-1, C, -1), gout[:, 3:].double().reshape(B, C, m * ns))
    want = want[:, :, :n]
    for strided in (False, True):
        src = gout if strided else gout[:, 3:].contiguous()
        ptr = gout.data_ptr() + 12 * m * ns if strided else src.data_ptr()
        stride = (C + 3) * m * ns if strided else C * m * ns
        for overwrite in (1, 0):
            grad = torch.full((B, C, n), float("nan") if overwrite else 0.5, device=dev)
            l0 = _lib.launch_count()
            _lib.call("gb_group_bwd_strided", gout, ptr, idx.data_ptr(), grad.data_ptr(), B, C, n, m, ns, stride, overwrite)
            assert _lib.launch_count() - l0 == 1
            _close(grad, want + (0.0 if overwrite else 0.5))


def test_group_bwd_dispatch_small_batches_keep_the_sorted_path(dev):
    """Few (scene, channel group) tasks cannot fill the GPU with one warp each: the sorted path answers (more launches)."""
    g = torch.Generator(device="cpu").manual_seed(2)
    idx = torch.randint(0, 512, (1, 64, 32), generator=g, dtype=torch.int32).to(dev)
    gout = torch.randn((1, 16, 64, 32), generator=g).to(dev)
    l0 = _lib.launch_count()
    got = gb_a.group_points_grad(gout, idx, 512)
    assert _lib.launch_count() - l0 > 1
    want = torch.zeros((1, 16, 512), dtype=torch.float64, device=dev)
    want.scatter_add_(2, idx.long().reshape(1, 1, -1).expand(-1, 16, -1), gout.double().reshape(1, 16, -1))
    _close(got, want)


@pytest.mark.parametrize("B,C,n,m,ns", [(16, 3, 20000, 1024, 64), (50, 1, 700, 33, 3), (24, 2, 5000, 128, 16)])
def test_group_bwd_few_channels_single_launch(dev, B, C, n, m, ns):
    """c < 4 (the grouped coordinates' gradient) with at least 48 output rows: one launch with the row sums in shared memory,
    set and accumulate forms, 16-byte aligned and ragged entry counts, against an fp64 scatter_add."""
    g = torch.Generator(device="cpu").manual_seed(11)
    idx = torch.randint(0, n, (B, m, ns), generator=g, dtype=torch.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]  # padded rows: many copies of one target
    gout = torch.randn((B, C, m, ns), generator=g)
    want = torch.zeros((B, C, n), dtype=torch.float64)
    want.scatter_add_(2, idx.long().reshape(B, 1, -1).expand(-1, C, -1), gout.double().reshape(B, C, -1))
    idx, gout = idx.to(dev), gout.to(dev)
    l0 = _lib.launch_count()
    got = gb_a.group_points_grad(gout, idx, n)
    assert _lib.launch_count() - l0 == 1
    _close(got.cpu(), want.float())
    acc = torch.full((B, C, n), -0.5, device=dev)
    gb_b.group_points_grad_wrapper(B, C, n, m, ns, gout, idx, acc)
    _close(acc.cpu() + 0.5, want.float())


@pytest.mark.parametrize("rows_mode", [1, 2])
@pytest.mark.parametrize("B,C,n,m,ns,kind", [(2, 8, 512, 64, 16, "sorted"), (2, 6, 1024, 128, 16, "padded"), (3, 5, 700, 32, 16, "random"),
                                             (2, 12, 600, 64, 8, "padded"), (2, 4, 512, 48, 16, "oob"), (2, 9, 256, 64, 16, "knn")])
def test_group_bwd_private_short_rows_both_layouts(dev, force_private, rows_mode, B, C, n, m, ns, kind):
    """nsample 8 / 16: channel planes sharing a row (priv_rows = 1) and several rows per unit (priv_rows = 2) give the same sums."""
    g = torch.Generator(device="cpu").manual_seed(5)
    idx = _make_idx(kind, B, n, m, ns, g).to(dev)
    gout = torch.randn((B, C, m, ns), generator=g).to(dev)
    safe = idx.long().reshape(B, 1, m * ns)
    safe = torch.where((safe < 0) | (safe >= n), torch.full_like(safe, n), safe)
    want = torch.zeros((B, C, n + 1), dtype=torch.float64, device=dev)
    want.scatter_add_(2, safe.expand(-1, C, -1), gout.double().reshape(B, C, m * ns))
    _lib.set_tuning("priv_rows", rows_mode)
    try:
        l0 = _lib.launch_count()
        got = gb_a.group_points_grad(gout, idx, n)
        assert _lib.launch_count() - l0 == 1
        _close(got, want[:, :, :n])
        assert torch.equal(got, gb_a.group_points_grad(gout, idx, n))
    finally:
        _lib.set_tuning("priv_rows", 0)


@pytest.mark.parametrize("B,C,n,m", [(4, 256, 20000, 1024), (2, 64, 1024, 512), (3, 16, 700, 40)])
def test_interpolate_backward_is_bit_reproducible(dev, B, C, n, m):
    """The sorted backward orders every target's entries by entry number (seg_sort_dense_kernel ranks them after the
    cursor scatter), so the interpolation backward returns the same bits on every run -- the reference's float atomics
    (interpolate_gpu.cu:127-149) do not."""
    g = torch.Generator(device="cpu").manual_seed(7)
    idx = torch.randint(0, m, (B, n, 3), generator=g, dtype=torch.int32).to(dev)
    w = torch.rand((B, n, 3), generator=g).to(dev)
    gout = torch.randn((B, C, n), generator=g).to(dev)
    first = gb_a.three_interpolate_grad(gout, idx, w, m)
    want = torch.zeros((B, C, m), dtype=torch.float64, device=dev)
    want.scatter_add_(2, idx.long().reshape(B, 1, -1).expand(-1, C, -1),
                      (gout.double().unsqueeze(-1) * w.double().unsqueeze(1)).reshape(B, C, -1))
    _close(first, want)
    for _ in range(4):
        assert torch.equal(gb_a.three_interpolate_grad(gout, idx, w, m), first)


def test_sorted_group_backward_is_bit_reproducible(dev):
    """Small batches take the sorted path (too few tasks for one warp each): same bits run to run there too."""
    g = torch.Generator(device="cpu").manual_seed(8)
    for (B, C, n, m, ns) in ((1, 16, 512, 64, 32), (2, 32, 2048, 2048, 64)):
        idx = torch.randint(0, max(n // 8, 1), (B, m, ns), generator=g, dtype=torch.int32).to(dev)  # heavy targets
        gout = torch.randn((B, C, m, ns), generator=g).to(dev)
        with force_sorted():
            first = gb_a.group_points_grad(gout, idx, n)
            for _ in range(4):
                assert torch.equal(gb_a.group_points_grad(gout, idx, n), first)
        want = torch.zeros((B, C, n), dtype=torch.float64, device=dev)
        want.scatter_add_(2, idx.long().reshape(B, 1, -1).expand(-1, C, -1), gout.double().reshape(B, C, -1))
        _close(first, want)


def test_sorted_backward_with_slices_beyond_the_ranking_limit(dev):
    """Targets that receive more than 256 entries of one tile (every neighbourhood empty: ball_query answers index 0) keep the
    cursor order inside seg_sort_dense_kernel; the sums stay within tolerance for the group and the interpolation backward."""
    g = torch.Generator(device="cpu").manual_seed(9)
    idx = torch.randint(0, 3, (2, 1024, 32), generator=g, dtype=torch.int32).to(dev)
    idx[0, :512] = 0
    gout = torch.randn((2, 16, 1024, 32), generator=g).to(dev)
    with force_sorted():
        got = gb_a.group_points_grad(gout, idx, 700)
    want = torch.zeros((2, 16, 700), dtype=torch.float64, device=dev)
    want.scatter_add_(2, idx.long().reshape(2, 1, -1).expand(-1, 16, -1), gout.double().reshape(2, 16, -1))
    assert (got.double() - want).abs().max().item() <= 1e-4 * max(want.abs().max().item(), 1.0)  # ~10^4 addends per target
    idx3 = torch.randint(0, 2, (2, 5000, 3), generator=g, dtype=torch.int32).to(dev)
    w = torch.rand((2, 5000, 3), generator=g).to(dev)
    go = torch.randn((2, 8, 5000), generator=g).to(dev)
    got3 = gb_a.three_interpolate_grad(go, idx3, w, 64)
    want3 = torch.zeros((2, 8, 64), dtype=torch.float64, device=dev)
    want3.scatter_add_(2, idx3.long().reshape(2, 1, -1).expand(-1, 8, -1), (go.double().unsqueeze(-1) * w.double().unsqueeze(1)).reshape(2, 8, -1))
    scale = max(want3.abs().max().item(), 1.0)
    assert (got3.double() - want3).abs().max().item() <= 1e-4 * scale  # thousands of addends per target
