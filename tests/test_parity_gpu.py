"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every op is called through the reference-facing Python API of graspbalance_b200 (which goes through the C ABI of
libgbops.so) and compared
  * with the CPU oracle (oracle/gb_oracle.c) on seeded inputs the oracle finishes in seconds, and
  * with the UNMODIFIED reference extensions compiled into oracle/_ref/ on the BASELINE.json shapes.
Indices, masks and forward fp32 values must be BIT-EXACT; scatter-add gradients within 1e-5 relative (the reference's
own float atomics are order-nondeterministic).
"""
import numpy as np
import pytest
import torch

import oracle
from graspbalance_b200 import _lib, scenes
from graspbalance_b200 import group as gb_group
from graspbalance_b200 import knn_modules as gb_knn
from graspbalance_b200 import pointnet2_batch_cuda as gb_b
from graspbalance_b200 import pointnet2_utils as pu
from graspbalance_b200 import subsample as gb_sub
from graspbalance_b200 import upsampling as gb_up
from graspbalance_b200 import _ext as gb_a
from graspbalance_b200.collision_detector import ModelFreeCollisionDetector

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-5  # north_star: gradients within 1e-5 relative in fp32


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def assert_grad_close(a, b, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.abs(b).max() if scale is None else scale
    assert np.abs(a - b).max() <= GRAD_RTOL * max(s, 1e-30), f"max abs diff {np.abs(a - b).max()} vs scale {s}"


def scene_with_origin_points(seed, n):
    """tabletop scene with 32 points inside the |p|^2 <= 1e-3 skip ball of FPS variant A (incl. index 0)."""
    xyz = scenes.tabletop_scene(seed, n)
    rng = np.random.default_rng(seed + 7)
    where = np.concatenate([[0], rng.choice(np.arange(1, n), 31, replace=False)])
    xyz[where] = rng.uniform(-0.015, 0.015, (32, 3)).astype(np.float32)
    return xyz


# --------------------------------------------------------------------------------------------------------------- FPS
@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("B,N,m,kind", [(2, 20000, 256, "tabletop"), (3, 2048, 512, "tabletop"), (2, 777, 300, "uniform"),
                                        (1, 5, 9, "uniform"), (2, 1, 3, "uniform"), (1, 1500, 1500, "tabletop"),
                                        (2, 4099, 64, "uniform")])
def test_fps_vs_oracle(dev, variant, B, N, m, kind):
    xyz = scenes.scene_batch(range(B), N, kind)
    want = oracle.furthest_point_sample(xyz, m, variant)
    x = T(xyz, dev)
    got = pu.furthest_point_sample(x, m) if variant == "A" else gb_sub.furthest_point_sample(x, m)
    assert got.dtype == torch.int32 and tuple(got.shape) == (B, m)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_fps_norm_skip_and_ties(dev):
    xyz = np.stack([scene_with_origin_points(s, 6000) for s in range(2)])
    xyz[1, 3000:] = xyz[1, :3000]  # half the cloud duplicated: every distance ties
    x = T(xyz, dev)
    for variant, fn in (("A", pu.furthest_point_sample), ("B", gb_sub.furthest_point_sample)):
        np.testing.assert_array_equal(fn(x, 700).cpu().numpy(), oracle.furthest_point_sample(xyz, 700, variant))
    # everything skipped: variant A returns index 0 throughout
    tiny = (np.random.default_rng(0).uniform(-0.01, 0.01, (1, 300, 3))).astype(np.float32)
    np.testing.assert_array_equal(pu.furthest_point_sample(T(tiny, dev), 20).cpu().numpy(), 0)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("threads", [128, 256, 512, 1024])
def test_fps_every_decomposition_gives_the_same_picks(dev, cluster, threads):  # (128 threads: honoured for 4 and 8 CTAs only)
    xyz = scenes.scene_batch([5, 6], 20000, "tabletop")
    want = oracle.furthest_point_sample(xyz, 300, "A")
    _lib.set_tuning("fps_cluster", cluster)
    _lib.set_tuning("fps_threads", threads)
    try:
        got = pu.furthest_point_sample(T(xyz, dev), 300).cpu().numpy()
        gotb = gb_sub.furthest_point_sample(T(xyz, dev), 300).cpu().numpy()
    finally:
        _lib.set_tuning("fps_cluster", 0)
        _lib.set_tuning("fps_threads", 0)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(gotb, oracle.furthest_point_sample(xyz, 300, "B"))


@pytest.mark.parametrize("B", [1, 4, 8])
def test_fps_small_shards_take_the_128_thread_shapes_and_match_the_reference(dev, ref_a, B):
    """Shards of a few scenes (BASELINE config 5 split over 4-8 GPUs) launch eight CTAs of 128 threads per scene, four for the
    2048-point level: the whole sampling chain against the unmodified reference module A."""
    cur = T(scenes.scene_batch(range(20, 20 + B), 20000, "tabletop"), dev)
    for m in (2048, 1024, 512, 256):
        want = ref_a.furthest_point_sampling(cur, m)
        got, new_xyz = pu.furthest_point_sample_xyz(cur, m)
        assert torch.equal(got, want)
        cur = new_xyz


def test_fps_full_size_vs_reference(dev, ref_a, ref_b):
    xyz = T(scenes.scene_batch(range(4), 20000, "tabletop"), dev)
    got = pu.furthest_point_sample(xyz, 1024)
    want = ref_a.furthest_point_sampling(xyz, 1024)
    assert torch.equal(got, want)
    got2048 = pu.furthest_point_sample(xyz, 2048)
    assert torch.equal(got2048, ref_a.furthest_point_sampling(xyz, 2048))
    temp = torch.full((4, 20000), 1e10, device=dev)
    out = torch.empty((4, 1024), dtype=torch.int32, device=dev)
    ref_b.furthest_point_sampling_wrapper(4, 20000, 1024, xyz, temp, out)
    temp2 = torch.full((4, 20000), 1e10, device=dev)
    out2 = torch.empty((4, 1024), dtype=torch.int32, device=dev)
    gb_b.furthest_point_sampling_wrapper(4, 20000, 1024, xyz, temp2, out2)
    assert torch.equal(out2, out)
    assert torch.equal(temp2, temp)  # the running distances left in `temp` match too
    # FPS chain of the backbone: 2048 -> 1024 -> 512 -> 256 on the sampled clouds
    cur = pu.gather_operation(xyz.transpose(1, 2).contiguous(), got2048).transpose(1, 2).contiguous()
    for m in (1024, 512, 256):
        a, b = pu.furthest_point_sample(cur, m), ref_a.furthest_point_sampling(cur, m)
        assert torch.equal(a, b)
        cur = pu.gather_operation(cur.transpose(1, 2).contiguous(), a).transpose(1, 2).contiguous()


# ------------------------------------------------------------------------------------------------- ball / cylinder
def _queries(xyz, m, seed):
    rng = np.random.default_rng(seed)
    B, N, _ = xyz.shape
    pick = np.stack([rng.choice(N, m, replace=m > N) for _ in range(B)])
    return np.take_along_axis(xyz, pick[..., None], axis=1)


@pytest.mark.parametrize("B,N,m,r,ns,kind", [(2, 20000, 256, 0.05, 64, "tabletop"), (2, 2048, 300, 0.08, 64, "tabletop"),
                                             (1, 1001, 77, 0.2, 16, "uniform"), (2, 37, 5, 0.5, 7, "uniform"),
                                             (1, 5000, 128, 1e-4, 32, "uniform"), (1, 4032, 64, 10.0, 100, "uniform"),
                                             (3, 20000, 96, 0.02, 1, "tabletop")])
def test_ball_query_vs_oracle(dev, B, N, m, r, ns, kind):
    xyz = scenes.scene_batch(range(B), N, kind)
    new_xyz = _queries(xyz, m, 3)
    want = oracle.ball_query(r, ns, xyz, new_xyz)
    got = pu.ball_query(r, ns, T(xyz, dev), T(new_xyz, dev))
    assert got.dtype == torch.int32
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    got_b = gb_group.ball_query(r, ns, T(xyz, dev), T(new_xyz, dev))
    np.testing.assert_array_equal(got_b.cpu().numpy(), want)


@pytest.mark.parametrize("qpw", [1, 2, 4])
def test_ball_query_queries_per_warp(dev, qpw):
    xyz = scenes.scene_batch(range(2), 6000, "tabletop")
    new_xyz = _queries(xyz, 203, 9)
    want = oracle.ball_query(0.06, 48, xyz, new_xyz)
    _lib.set_tuning("query_qpw", qpw)
    try:
        got = pu.ball_query(0.06, 48, T(xyz, dev), T(new_xyz, dev)).cpu().numpy()
    finally:
        _lib.set_tuning("query_qpw", 0)
    np.testing.assert_array_equal(got, want)


def _cyl_inputs(B, N, m, seed, kind="tabletop"):
    xyz = scenes.scene_batch(range(seed, seed + B), N, kind)
    new_xyz = _queries(xyz, m, seed)
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(B, m, 3)).astype(np.float32)
    rot = scenes.viewpoint_rotations(-v, rng.uniform(0, np.pi, (B, m)).astype(np.float32))
    return xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9))


@pytest.mark.parametrize("B,N,m,ns,hmax", [(2, 20000, 128, 64, 0.04), (1, 3001, 50, 16, 0.01), (2, 2048, 64, 32, 0.02)])
def test_cylinder_query_vs_oracle(dev, B, N, m, ns, hmax):
    xyz, new_xyz, rot = _cyl_inputs(B, N, m, 21)
    want = oracle.cylinder_query(0.05, -0.02, hmax, ns, xyz, new_xyz, rot)
    got = pu.cylinder_query(0.05, -0.02, hmax, ns, T(xyz, dev), T(new_xyz, dev), T(rot, dev))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_queries_full_size_vs_reference(dev, ref_a, ref_b):
    xyz = T(scenes.scene_batch(range(4), 20000, "tabletop"), dev)
    fidx = pu.furthest_point_sample(xyz, 1024)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    for r, ns in ((0.05, 64), (0.04, 64), (0.1, 32), (0.3, 16)):
        want = ref_a.ball_query(new_xyz, xyz, r, ns)
        assert torch.equal(pu.ball_query(r, ns, xyz, new_xyz), want)
        idx_b = torch.zeros((4, 1024, ns), dtype=torch.int32, device=dev)
        ref_b.ball_query_wrapper(4, 20000, 1024, r, ns, new_xyz, xyz, idx_b)
        assert torch.equal(gb_group.ball_query(r, ns, xyz, new_xyz), idx_b)
    uni = T(scenes.scene_batch(range(2), 20000, "uniform"), dev)
    q = uni[:, :1024].contiguous()
    assert torch.equal(pu.ball_query(0.05, 64, uni, q), ref_a.ball_query(q, uni, 0.05, 64))
    # cylinder: 12 in-plane angles x 4 depths (BASELINE config 4), one scene pair
    rng = np.random.default_rng(0)
    v = rng.normal(size=(4, 1024, 3)).astype(np.float32)
    for a in range(0, 12, 5):
        rot = T(scenes.viewpoint_rotations(-v, np.full((4, 1024), a * np.pi / 12, np.float32)).reshape(4, 1024, 9), dev)
        for hmax in (0.01, 0.02, 0.03, 0.04):
            want = ref_a.cylinder_query(new_xyz, xyz, rot, 0.05, -0.02, hmax, 64)
            assert torch.equal(pu.cylinder_query(0.05, -0.02, hmax, 64, xyz, new_xyz, rot), want)


def _special_cloud(kind, N, seed):
    rng = np.random.default_rng(seed)
    xyz = scenes.scene_batch([seed, seed + 1], N, "tabletop" if kind != "uniform" else "uniform")
    if kind == "nan_inf":  # non-finite points never hit and must not disturb the others
        xyz[0, rng.choice(N, 40, replace=False)] = np.nan
        xyz[1, rng.choice(N, 40, replace=False), 1] = np.inf
        xyz[1, rng.choice(N, 10, replace=False), 2] = -np.inf
    elif kind == "planar":  # zero extent along z: a one-cell-thick grid
        xyz[..., 2] = 0.25
    elif kind == "dups":  # one heavily duplicated point: a cell holding a third of the cloud
        xyz[:, N // 3: 2 * N // 3] = xyz[:, :1]
    elif kind == "outlier":  # one far point stretches the bounding box
        xyz[0, 7] = (40.0, -30.0, 25.0)
    return xyz


@pytest.mark.parametrize("mode", [1, 2])  # 1 = full scan only, 2 = cell grid forced
@pytest.mark.parametrize("kind,N,m,r,ns", [("tabletop", 6000, 200, 0.05, 64), ("nan_inf", 5000, 150, 0.06, 32),
                                           ("planar", 4100, 100, 0.04, 16), ("dups", 3000, 64, 0.03, 64),
                                           ("outlier", 5000, 100, 0.05, 32), ("uniform", 4096, 100, 0.3, 24),
                                           ("tabletop", 2500, 64, -0.05, 16), ("tabletop", 2500, 64, 0.0, 16),
                                           ("tabletop", 2500, 64, 1e30, 40), ("uniform", 70, 9, 0.4, 5)])
def test_ball_query_grid_and_scan_paths(dev, mode, kind, N, m, r, ns):
    xyz = _special_cloud(kind, N, 11)
    new_xyz = _queries(xyz, m, 4)
    new_xyz[:, :3] += 5.0  # queries far outside the cloud
    if kind == "nan_inf":
        new_xyz[0, 5] = np.nan
    want = oracle.ball_query(r, ns, xyz, new_xyz)
    _lib.set_tuning("query_mode", mode)
    try:
        got = pu.ball_query(r, ns, T(xyz, dev), T(new_xyz, dev)).cpu().numpy()
    finally:
        _lib.set_tuning("query_mode", 0)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("kind,N,m,ns,hmin,hmax,rotkind", [("tabletop", 6000, 200, 64, -0.02, 0.04, "ortho"),
                                                         ("nan_inf", 5000, 100, 32, -0.02, 0.02, "ortho"),
                                                         ("planar", 4100, 100, 16, -0.02, 0.01, "ortho"),
                                                         ("tabletop", 5000, 120, 32, -0.02, 0.04, "scaled"),
                                                         ("tabletop", 5000, 120, 32, -0.02, 0.04, "random"),
                                                         ("tabletop", 3000, 60, 16, 0.03, -0.01, "ortho"),
                                                         ("outlier", 5000, 100, 32, -0.5, 0.5, "ortho")])
def test_cylinder_query_grid_and_scan_paths(dev, mode, kind, N, m, ns, hmin, hmax, rotkind):
    xyz = _special_cloud(kind, N, 13)
    new_xyz = _queries(xyz, m, 6)
    rng = np.random.default_rng(17)
    v = rng.normal(size=(2, m, 3)).astype(np.float32)
    rot = scenes.viewpoint_rotations(-v, rng.uniform(0, np.pi, (2, m)).astype(np.float32)).reshape(2, m, 9)
    if rotkind == "scaled":  # not orthonormal: the search region is no longer inside the bounding sphere
        rot = rot * rng.uniform(0.2, 3.0, (2, m, 1)).astype(np.float32)
    elif rotkind == "random":
        rot = rng.normal(size=(2, m, 9)).astype(np.float32)
        rot[0, 3] = np.nan
    rot = np.ascontiguousarray(rot.astype(np.float32))
    want = oracle.cylinder_query(0.05, hmin, hmax, ns, xyz, new_xyz, rot)
    _lib.set_tuning("query_mode", mode)
    try:
        got = pu.cylinder_query(0.05, hmin, hmax, ns, T(xyz, dev), T(new_xyz, dev), T(rot, dev)).cpu().numpy()
    finally:
        _lib.set_tuning("query_mode", 0)
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------------------------------------- group / gather
@pytest.mark.parametrize("B,C,N,m,ns", [(2, 3, 20000, 256, 64), (2, 128, 20000, 64, 64), (1, 131, 2048, 128, 32),
                                        (2, 5, 100, 7, 3), (1, 259, 1024, 512, 16), (1, 1, 50, 4, 4), (2, 16, 60000, 16, 8)])
def test_group_forward_backward_vs_oracle(dev, B, C, N, m, ns):
    rng = np.random.default_rng(B * 1000 + C)
    feats = rng.normal(size=(B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, m, ns)).astype(np.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]  # padded tails, like a ball query with few hits
    gout = rng.normal(size=(B, C, m, ns)).astype(np.float32)
    for mod in (pu, gb_group):
        f = T(feats, dev).requires_grad_(True)
        out = mod.grouping_operation(f, T(idx, dev))
        np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.grouping_operation(feats, idx))
        out.backward(T(gout, dev))
        assert_grad_close(f.grad.cpu().numpy(), oracle.grouping_operation_grad(gout, idx, N))
    assert torch.equal(gb_group.torch_grouping_operation(T(feats, dev), T(idx, dev)), out.detach())


@pytest.mark.parametrize("B,C,N,m,ns,pattern", [(2, 8, 20000, 64, 64, "zeros"), (2, 12, 3000, 128, 64, "one_per_query"),
                                                (1, 16, 50, 256, 32, "random"), (2, 6, 7, 100, 48, "random"),
                                                (1, 9, 2048, 2048, 8, "identity_runs"), (2, 4, 1, 77, 16, "zeros"),
                                                (1, 5, 40000, 300, 64, "two_values")])
def test_group_backward_long_runs(dev, B, C, N, m, ns, pattern):
    """Index patterns whose sorted tiles are dominated by long runs of equal targets (what ball_query returns for empty
    and sparse neighbourhoods): every boundary-shift / shared-target case of the sorted backward."""
    rng = np.random.default_rng(N + m)
    if pattern == "zeros":
        idx = np.zeros((B, m, ns), np.int32)
    elif pattern == "one_per_query":
        idx = np.repeat(rng.integers(0, N, (B, m, 1)), ns, axis=2).astype(np.int32)
    elif pattern == "identity_runs":
        idx = np.repeat((np.arange(m) % N)[None, :, None], ns, axis=2).astype(np.int32) * np.ones((B, 1, 1), np.int32)
    elif pattern == "two_values":
        idx = np.where(rng.random((B, m, ns)) < 0.5, 3, N - 1).astype(np.int32)
    else:
        idx = rng.integers(0, N, (B, m, ns)).astype(np.int32)
    gout = rng.normal(size=(B, C, m, ns)).astype(np.float32)
    want = oracle.grouping_operation_grad(gout, idx, N)
    for mod in (pu, gb_group):
        f = torch.zeros((B, C, N), device=dev, requires_grad=True)
        mod.grouping_operation(f, T(idx, dev)).backward(T(gout, dev))
        assert_grad_close(f.grad.cpu().numpy(), want, scale=max(np.abs(want).max(), 1.0))
    # module B's accumulate-into-caller-tensor entry adds to what is already there
    base = rng.normal(size=(B, C, N)).astype(np.float32)
    acc = T(base, dev)
    gb_b.group_points_grad_wrapper(B, C, N, m, ns, T(gout, dev), T(idx, dev), acc)
    assert_grad_close(acc.cpu().numpy(), base + want, scale=max(np.abs(want).max(), 1.0))


def test_gather_forward_backward_vs_oracle(dev):
    rng = np.random.default_rng(5)
    feats = rng.normal(size=(3, 7, 5000)).astype(np.float32)
    idx = rng.integers(0, 5000, (3, 333)).astype(np.int32)
    gout = rng.normal(size=(3, 7, 333)).astype(np.float32)
    for mod in (pu, gb_group, gb_sub):
        f = T(feats, dev).requires_grad_(True)
        out = mod.gather_operation(f, T(idx, dev))
        np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.gather_operation(feats, idx))
        out.backward(T(gout, dev))
        assert_grad_close(f.grad.cpu().numpy(), oracle.gather_operation_grad(gout, idx, 5000))
    # the one relation the reference itself asserts (subsample.py:145-157)
    ref = torch.gather(T(feats, dev), 2, T(idx, dev).long().unsqueeze(1).expand(-1, 7, -1))
    assert torch.equal(out.detach(), ref)


def test_group_full_size_vs_reference(dev, ref_a, ref_b):
    B, N, m, ns = 4, 20000, 1024, 64
    xyz = T(scenes.scene_batch(range(B), N, "tabletop"), dev)
    fidx = pu.furthest_point_sample(xyz, m)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    idx = pu.ball_query(0.05, ns, xyz, new_xyz)
    g = torch.Generator(device="cpu").manual_seed(1)
    feats = torch.randn((B, 128, N), generator=g).to(dev)
    xyz_t = xyz.transpose(1, 2).contiguous()
    for f in (xyz_t, feats):
        want = ref_a.group_points(f, idx)
        assert torch.equal(gb_a.group_points(f, idx), want)
        out_b = torch.empty_like(want)
        ref_b.group_points_wrapper(B, f.shape[1], N, m, ns, f, idx, out_b)
        assert torch.equal(out_b, want)
        gout = torch.randn(want.shape, generator=g).to(dev)
        want_g = ref_a.group_points_grad(gout, idx, N)
        got_g = gb_a.group_points_grad(gout, idx, N)
        assert_grad_close(got_g.cpu().numpy(), want_g.cpu().numpy())


@pytest.mark.parametrize("B,N,m", [(2, 20000, 1024), (3, 2048, 512), (1, 700, 700), (2, 40000, 64), (1, 100000, 16)])
def test_fps_xyz_is_fps_plus_gather(dev, B, N, m):
    """gb_fps_xyz == furthest_point_sample followed by the gather of pointnet2_modules.py:151-158 (cluster and global paths)."""
    xyz_np = scenes.scene_batch(range(B), N, "tabletop" if N >= 1000 else "uniform")
    xyz = T(xyz_np, dev)
    inds, new_xyz = pu.furthest_point_sample_xyz(xyz, m)
    want = pu.furthest_point_sample(xyz, m)
    assert torch.equal(inds, want)
    want_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), want).transpose(1, 2).contiguous()
    assert torch.equal(new_xyz, want_xyz)
    if N <= 20000:
        np.testing.assert_array_equal(inds.cpu().numpy(), oracle.furthest_point_sample(xyz_np, m, "A"))


def test_fps_segments_match_one_call_per_segment(dev):
    """gb_fps_segments == furthest_point_sample on every segment alone (the per-object loop of ObjectBalanceSampling,
    TrainModel/modules.py:201-209): sizes from 1 to 9000 points, more samples than points, duplicated points (ties),
    points near the origin (variant A's norm skip), an empty segment."""
    rng = np.random.default_rng(7)
    counts = [1, 2, 33, 700, 512, 513, 4096, 9000, 0, 257, 64, 3000]
    ks = [1, 5, 33, 128, 512, 100, 341, 343, 0, 300, 64, 1]
    parts = []
    for c in counts:
        p = rng.uniform(-0.3, 0.3, (c, 3)).astype(np.float32) + np.array([0, 0, 0.5], np.float32)
        if c >= 64:
            p[c // 2:c // 2 + 16] = p[:16]          # exact duplicates: equal distances, the tie rule decides
        if c == 700:
            p[100:110] = rng.uniform(-0.01, 0.01, (10, 3)).astype(np.float32)  # |p|^2 <= 1e-3: skipped by variant A
        parts.append(p)
    packed = np.concatenate(parts)
    got = pu.furthest_point_sample_segments(T(packed, dev), counts, ks).cpu().numpy()
    want = oracle.furthest_point_sample_segments(packed, counts, ks)
    np.testing.assert_array_equal(got, want)
    first = slot = 0
    for c, k in zip(counts, ks):  # and against the product's own single-cloud entry point
        if c > 0 and k > 0:
            alone = pu.furthest_point_sample(T(packed[None, first:first + c], dev), k)[0].cpu().numpy()
            np.testing.assert_array_equal(got[slot:slot + k], alone)
        first += c
        slot += k
    # the C entry point refuses a segment that does not fit the registers of one CTA (it never mis-samples) ...
    with pytest.raises(RuntimeError):
        seg = torch.tensor([[0, 20000, 8, 0]], dtype=torch.int32, device=dev)
        gb_a.furthest_point_sampling_segments(T(np.zeros((20000, 3), np.float32), dev), seg, 20000, 8, 8)


def test_fps_segments_with_an_object_larger_than_one_cta_holds(dev):
    """... and the Python entry routes such a segment through the cluster kernel: an object that owns more than half of a
    20k-point cloud (> 10240 points) is sampled like every other one (the reference's per-object call has no size limit,
    TrainModel/modules.py:207)."""
    rng = np.random.default_rng(9)
    counts, ks = [300, 12000, 40, 15000, 0, 1024], [17, 600, 40, 424, 0, 100]
    packed = (rng.uniform(-0.3, 0.3, (sum(counts), 3)).astype(np.float32) + np.array([0, 0, 0.5], np.float32))
    packed[400:420] = packed[5000:5020]  # ties inside the oversize segment
    got = pu.furthest_point_sample_segments(T(packed, dev), counts, ks).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.furthest_point_sample_segments(packed, counts, ks))


def reference_object_balance_sampling(end_points):
    """TrainModel/modules.py:177-223, literally, on top of the product's single-cloud furthest_point_sample."""
    batch_seg_res = end_points["seed_cluster"]
    batch_points = end_points["point_clouds"]
    batch_features = end_points["up_sample_features"].permute(0, 2, 1)
    B, N = batch_seg_res.shape
    new_xyz, new_feat, new_inds = [], [], []
    for i in range(B):
        seg_res, points, features = batch_seg_res[i], batch_points[i], batch_features[i]
        idxs = torch.unique(seg_res)
        num_objects = len(idxs) - 1
        per = [1024 // num_objects for _ in range(num_objects)]
        per[-1] += 1024 % num_objects
        li, lp, lf, t = [], [], [], 0
        for j in idxs:
            if j == 0:
                continue
            inds = torch.where(seg_res == j)[0]
            object_points = points[seg_res == j]
            sample = pu.furthest_point_sample(object_points.unsqueeze(0).contiguous(), per[t])[0].long()
            t += 1
            sel = torch.gather(inds, 0, sample)
            li.append(sel)
            lp.append(torch.gather(points, 0, sel.unsqueeze(1).expand(-1, 3)))
            lf.append(torch.gather(features, 0, sel.unsqueeze(1).expand(-1, features.shape[1])))
        new_inds.append(torch.cat(li, 0)); new_xyz.append(torch.cat(lp, 0)); new_feat.append(torch.cat(lf, 0))
    return torch.stack(new_inds, 0).int(), torch.stack(new_xyz, 0), torch.stack(new_feat, 0).permute(0, 2, 1)


def test_object_balance_sampling_matches_the_reference_loop(dev):
    from graspbalance_b200.modules import ObjectBalanceSampling
    B, N, C = 3, 20000, 32
    rng = np.random.default_rng(11)
    xyz = scenes.scene_batch(range(B), N, "tabletop")
    labels = np.zeros((B, N), np.int64)
    for b in range(B):  # 3..7 "objects": nearest of a few random centres within 8 cm, everything else background
        nobj = 3 + 2 * b
        centres = xyz[b, rng.choice(N, nobj, replace=False)]
        d = np.linalg.norm(xyz[b][:, None] - centres[None], axis=-1)
        near = d.min(1) < 0.08
        labels[b, near] = d.argmin(1)[near] + 1 + b  # label values differ between scenes
    ep = {"seed_cluster": T(labels, dev), "point_clouds": T(xyz, dev), "up_sample_features": T(rng.normal(size=(B, C, N)).astype(np.float32), dev),
          "fp2_inds": torch.zeros((B, 1024), dtype=torch.int32, device=dev)}
    want_inds, want_xyz, want_feat = reference_object_balance_sampling(ep)
    out = ObjectBalanceSampling(dict(ep))
    assert torch.equal(out["fp2_inds"], want_inds) and out["fp2_inds"].dtype == torch.int32
    assert torch.equal(out["fp2_xyz"], want_xyz)
    assert torch.equal(out["fp2_features"], want_feat)
    assert out["fp2_inds_fps"] is ep["fp2_inds"]


# ------------------------------------------------------------------------------------------- three_nn / interpolate
@pytest.mark.parametrize("B,n,m", [(2, 20000, 1024), (1, 513, 2), (2, 1000, 3), (1, 64, 1), (2, 4000, 2500)])
def test_three_nn_vs_oracle(dev, B, n, m):
    unknown = scenes.scene_batch(range(B), n, "tabletop" if n >= 1000 else "uniform")
    known = _queries(unknown, m, 4)
    want_d2, want_i = oracle.three_nn_dist2(unknown, known)
    d2, idx = gb_a.three_nn(T(unknown, dev), T(known, dev))
    np.testing.assert_array_equal(idx.cpu().numpy(), want_i)
    np.testing.assert_array_equal(d2.cpu().numpy(), want_d2)
    for mod in (pu, gb_up):
        dist, idx2 = mod.three_nn(T(unknown, dev), T(known, dev))
        np.testing.assert_array_equal(idx2.cpu().numpy(), want_i)
        np.testing.assert_array_equal(dist.cpu().numpy(), np.sqrt(want_d2))


@pytest.mark.parametrize("B,n,m", [(2, 20000, 1024), (1, 513, 2), (2, 1000, 3), (1, 64, 1), (2, 4000, 2500), (1, 300, 300)])
def test_three_nn_weights_matches_the_torch_ops_it_replaces(dev, B, n, m):
    """gb_three_nn_weights == three_nn followed by pointnet2_modules.py:413-416 (sqrt, +1e-8, reciprocal, sum, divide):
    bit for bit against the torch ops on the GPU and against the oracle's fp32 numpy restatement."""
    unknown = scenes.scene_batch(range(B), n, "tabletop" if n >= 1000 else "uniform")
    known = _queries(unknown, m, 4)
    if n == 300:
        known = unknown.copy()  # every unknown coincides with a known point: dist 0, weight 1 on the first neighbour
    u, k = T(unknown, dev), T(known, dev)
    dist, idx, weight = pu.three_nn_weights(u, k)
    dist0, idx0 = pu.three_nn(u, k)
    recip = 1.0 / (dist0 + 1e-8)
    weight0 = recip / torch.sum(recip, dim=2, keepdim=True)
    assert torch.equal(idx, idx0) and torch.equal(dist, dist0)
    assert torch.equal(weight.view(torch.int32), weight0.view(torch.int32))  # NaN-safe bit comparison (m < 3: inf / inf)
    wd, wi, ww = oracle.three_nn_weights(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), wi)
    np.testing.assert_array_equal(dist.cpu().numpy(), wd)
    np.testing.assert_array_equal(weight.cpu().numpy().view(np.int32), ww.view(np.int32))
    # the B-side helper takes the fused path and gives what the reference's op sequence gives
    feats = T(np.random.default_rng(0).normal(size=(B, 6, m)).astype(np.float32), dev)
    if m >= 3:
        assert torch.equal(gb_up.three_interpolation(u, k, feats), gb_up.three_interpolate(feats, idx0, weight0.contiguous()))


@pytest.mark.parametrize("B,C,m,n", [(2, 256, 1024, 20000), (1, 7, 50, 333), (2, 64, 256, 512), (1, 130, 512, 1024),
                                     (1, 8, 4, 5000), (2, 16, 1, 999), (1, 5, 30000, 4096), (2, 4, 3, 2)])
def test_three_interpolate_vs_oracle(dev, B, C, m, n):
    rng = np.random.default_rng(C)
    feats = rng.normal(size=(B, C, m)).astype(np.float32)
    idx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (B, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    gout = rng.normal(size=(B, C, n)).astype(np.float32)
    for mod in (pu, gb_up):
        f = T(feats, dev).requires_grad_(True)
        out = mod.three_interpolate(f, T(idx, dev), T(w, dev))
        np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.three_interpolate(feats, idx, w))
        out.backward(T(gout, dev))
        assert_grad_close(f.grad.cpu().numpy(), oracle.three_interpolate_grad(gout, idx, w, m))


def test_fp_chain_full_size_vs_reference(dev, ref_a, ref_b):
    B, n, m, C = 4, 20000, 1024, 256
    xyz = T(scenes.scene_batch(range(B), n, "tabletop"), dev)
    fidx = pu.furthest_point_sample(xyz, m)
    known = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    d2, idx = gb_a.three_nn(xyz, known)
    rd2, ridx = ref_a.three_nn(xyz, known)
    assert torch.equal(idx, ridx) and torch.equal(d2, rd2)
    dist = torch.sqrt(d2)
    recip = 1.0 / (dist + 1e-8)
    w = (recip / recip.sum(dim=2, keepdim=True)).contiguous()
    g = torch.Generator(device="cpu").manual_seed(2)
    feats = torch.randn((B, C, m), generator=g).to(dev)
    want = ref_a.three_interpolate(feats, idx, w)
    assert torch.equal(gb_a.three_interpolate(feats, idx, w), want)
    out_b = torch.empty_like(want)
    ref_b.three_interpolate_wrapper(B, C, m, n, feats, idx, w, out_b)
    assert torch.equal(out_b, want)
    gout = torch.randn(want.shape, generator=g).to(dev)
    assert_grad_close(gb_a.three_interpolate_grad(gout, idx, w, m).cpu().numpy(),
                      ref_a.three_interpolate_grad(gout, idx, w, m).cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------- KNN
@pytest.mark.parametrize("B,D,R,Q,k", [(2, 3, 300, 300, 1), (1, 3, 5000, 200, 1), (2, 3, 2000, 100, 64), (1, 5, 700, 33, 8),
                                       (1, 3, 64, 10, 64), (1, 20, 1000, 17, 3), (1, 3, 20000, 64, 64)])
def test_knn_vs_oracle(dev, B, D, R, Q, k):
    rng = np.random.default_rng(R + k)
    ref = rng.uniform(-1, 1, (B, D, R)).astype(np.float32)
    ref[:, :, R // 2:R // 2 + 20] = ref[:, :, :20]  # exact duplicates: (distance, index) tie order
    query = rng.uniform(-1, 1, (B, D, Q)).astype(np.float32)
    want = oracle.knn(ref, query, k)
    got = gb_knn.knn_k(T(ref, dev), T(query, dev), k)
    assert got.dtype == torch.int64 and tuple(got.shape) == (B, k, Q)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    if k == 1:
        np.testing.assert_array_equal(gb_knn.myknn(T(ref, dev), T(query, dev)).cpu().numpy(), want)


def test_knn_full_size_vs_reference(dev, ref_c):
    xyz = T(scenes.scene_batch(range(2), 20000, "tabletop"), dev)
    ref = xyz.transpose(1, 2).contiguous()
    query = ref[:, :, ::19][:, :, :1024].contiguous() + 0.001
    for k in (1, 64):
        want = torch.empty((2, k, 1024), dtype=torch.int64, device=dev)
        ref_c.knn(ref, query, want)
        assert torch.equal(gb_knn.knn_k(ref, query, k), want)


# ---------------------------------------------------------------------------------------------------------- collision
def _detector_on(pts, voxel, dev):
    """A detector whose (already down-sampled) scene is given: what the goldens pin is `detect` itself."""
    det = ModelFreeCollisionDetector.__new__(ModelFreeCollisionDetector)
    det.finger_width, det.finger_length, det.voxel_size, det.device = 0.01, 0.06, voxel, dev
    det._scene_dev, det._scene_host = T(pts, dev), pts
    return det


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_collision_vs_golden_and_oracle(dev, tag):
    """detect() against the outputs of the reference's own collision_detector.py (tests/golden/make_golden_collision.py):
    a, b small scenes; c = BASELINE config 1 as written (20k-point scene, voxel 0.01, 1024 grasps); d = a float32 grasp group,
    for which the reference evaluates thresholds and volumes in float32.  Masks and all five IoU arrays bit-identical, through
    the duck-typed attributes, through a graspnetAPI-style [G,17] grasp_group_array, and with the arrays on the device."""
    import os
    from test_oracle_cpu import golden_scene
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "collision_ref.npz"))
    pts, voxel = golden_scene(z, tag), float(z[tag + "_voxel"])
    det = _detector_on(pts, voxel, dev)
    Tr, Rm, h, d, w = (z[tag + k] for k in ("_translations", "_rotation_matrices", "_heights", "_depths", "_widths"))
    gg = scenes.GraspGroupStandIn(Tr, Rm, h, d, w)
    plain = det.detect(gg, approach_dist=0.05, collision_thresh=0.01)
    assert isinstance(plain, np.ndarray) and plain.dtype == np.bool_ and (plain == z[tag + "_collision"]).all()

    def check(full):
        assert (np.asarray(full[0]) == z[tag + "_collision"]).all() and (np.asarray(full[1]) == z[tag + "_empty"]).all()
        assert full[0].dtype == np.bool_ and full[1].dtype == np.bool_
        for a, b in zip(full[2], z[tag + "_ious"]):
            assert a.dtype == np.float64
            np.testing.assert_array_equal(a, b)
    check(det.detect(gg, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True))
    # graspnetAPI layout: [score, width, height, depth, R(9), T(3), object id], uploaded as it is
    G = Tr.shape[0]
    arr = np.zeros((G, 17), dtype=Tr.dtype)
    arr[:, 1], arr[:, 2], arr[:, 3], arr[:, 4:13], arr[:, 13:16], arr[:, 16] = w, h, d, Rm.reshape(G, 9), Tr, -1
    gg17 = scenes.GraspGroupStandIn(Tr, Rm, h, d, w)
    gg17.grasp_group_array = arr
    check(det.detect(gg17, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True))
    # network output that never leaves the device (modules.pred_decode emits this array): CUDA tensors in, CUDA tensors out
    on_dev = det.detect_device(T(arr, dev), approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True)
    assert on_dev[0].is_cuda and on_dev[0].dtype == torch.bool
    check([on_dev[0].cpu().numpy(), on_dev[1].cpu().numpy(), [x.cpu().numpy() for x in on_dev[2]]])
    ggd = scenes.GraspGroupStandIn(T(Tr, dev), T(Rm, dev), T(h, dev), T(d, dev), T(w, dev))
    on_dev = det.detect(ggd, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True)
    check([on_dev[0].cpu().numpy(), on_dev[1].cpu().numpy(), [x.cpu().numpy() for x in on_dev[2]]])


def test_collision_detector_refuses_cpu_devices():
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ModelFreeCollisionDetector(np.zeros((10, 3)), voxel_size=0.01, device="cpu")


def test_collision_full_size_vs_oracle(dev):
    raw = scenes.tabletop_scene(3, 20000).astype(np.float64)
    det = ModelFreeCollisionDetector(raw, voxel_size=0.01, device=dev)
    gs = scenes.grasp_set(4, det.scene_points, 1024)
    gg = scenes.GraspGroupStandIn(**gs)
    got = det.detect(gg, approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True)
    want = oracle.collision_detect(det.scene_points, 0.01, gs["translations"], gs["rotation_matrices"], gs["heights"],
                                   gs["depths"], gs["widths"], approach_dist=0.05, collision_thresh=0.01,
                                   return_empty_grasp=True, return_ious=True)
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
    for a, b in zip(got[2], want[2]):
        np.testing.assert_array_equal(a, b)
    assert 0 < got[0].sum() < 1024


def test_collision_counts_batched_equals_one_launch_per_scene(dev):
    from graspbalance_b200.collision_detector import collision_counts, collision_counts_batched
    dets, Ts, Rs, thrs = [], [], [], []
    for sid, n in ((3, 20000), (4, 6000), (5, 300), (6, 20000)):  # scenes of different sizes after down-sampling
        det = ModelFreeCollisionDetector(scenes.tabletop_scene(sid, n).astype(np.float64), voxel_size=0.01, device=dev)
        gs = scenes.grasp_set(sid, det.scene_points, 96)
        thr = oracle.collision_thresholds(gs["heights"], gs["depths"], gs["widths"], 0.03)
        dets.append(det); Ts.append(gs["translations"]); Rs.append(gs["rotation_matrices"]); thrs.append(thr)
    Td, Rd, thd = (T(np.ascontiguousarray(np.stack(a)), dev) for a in (Ts, Rs, thrs))
    got = collision_counts_batched([d._scene_dev for d in dets], Td, Rd, thd)
    assert got.shape == (4, 96, 6)
    for s_, d in enumerate(dets):
        assert torch.equal(got[s_], collision_counts(d._scene_dev, Td[s_].contiguous(), Rd[s_].contiguous(), thd[s_].contiguous()))
    assert int(got[..., 0].sum()) > 0


@pytest.mark.parametrize("n,voxel,kind", [(20000, 0.01, "tabletop"), (20000, 0.005, "tabletop"), (5000, 0.05, "uniform"), (1, 0.01, "uniform"),
                                           (3000, 1e-4, "uniform")])
def test_voxel_down_sample_gpu_matches_the_host_restatement(dev, n, voxel, kind):
    """voxel_down_sample_gpu (torch sort + gb_voxel_means) == the oracle's numpy restatement of open3d's voxel_down_sample: the
    same set of voxel means, bit for bit (sequential fp64 sums in input order); only the order of the voxels differs."""
    from graspbalance_b200.collision_detector import voxel_down_sample_gpu
    voxel_down_sample = oracle.voxel_down_sample
    pts = scenes.scene_batch([9], n, kind)[0].astype(np.float64)
    pts[n // 2:n // 2 + min(50, n // 3)] = pts[:min(50, n // 3)]  # duplicates share a voxel
    got = voxel_down_sample_gpu(T(pts, dev), voxel).cpu().numpy()
    want = voxel_down_sample(pts, voxel)
    assert got.shape == want.shape
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    np.testing.assert_array_equal(key(got), key(want))
    np.testing.assert_array_equal(key(want), key(oracle.voxel_down_sample(pts, voxel)))
    det = ModelFreeCollisionDetector(pts, voxel_size=voxel, device=dev)       # the constructor takes the GPU path ...
    np.testing.assert_array_equal(key(det.scene_points), key(want))
    det_t = ModelFreeCollisionDetector(T(pts, dev), voxel_size=voxel, device=dev)  # ... and accepts a tensor already on the device
    np.testing.assert_array_equal(key(det_t.scene_points), key(want))


# ------------------------------------------------------------------------------------------------- sizes at the edges
def test_empty_inputs_return_empty_outputs(dev):
    """Zero queries / samples / scenes: the entry points return without launching (the reference would launch empty grids)."""
    xyz = T(scenes.scene_batch(range(2), 500, "uniform"), dev)
    none = xyz[:, :0].contiguous()
    assert tuple(pu.ball_query(0.1, 8, xyz, none).shape) == (2, 0, 8)
    assert tuple(pu.furthest_point_sample(xyz, 0).shape) == (2, 0)
    feats = torch.randn((2, 5, 500), device=dev)
    assert tuple(pu.gather_operation(feats, torch.zeros((2, 0), dtype=torch.int32, device=dev)).shape) == (2, 5, 0)
    assert tuple(pu.grouping_operation(feats, torch.zeros((2, 0, 4), dtype=torch.int32, device=dev)).shape) == (2, 5, 0, 4)
    d, i = pu.three_nn(none, xyz)
    assert tuple(d.shape) == (2, 0, 3) and tuple(i.shape) == (2, 0, 3)
    assert tuple(pu.ball_query(0.1, 8, xyz[:0].contiguous(), xyz[:0, :7].contiguous()).shape) == (0, 7, 8)
    g = gb_a.group_points_grad(torch.zeros((2, 5, 0, 4), device=dev), torch.zeros((2, 0, 4), dtype=torch.int32, device=dev), 500)
    assert tuple(g.shape) == (2, 5, 500) and float(g.abs().max()) == 0.0


def test_dataset_maximum_cloud_size(dev):
    """50000 points per cloud is the reference loader's cap (graspnet_dataset.py:19): FPS, ball / cylinder query through the
    cell grid, grouping forward + sorted backward at that size, against the oracle."""
    B, N, m, ns = 1, 50000, 200, 32
    xyz = scenes.scene_batch([77], N, "tabletop")
    x = T(xyz, dev)
    inds = pu.furthest_point_sample(x, m)
    np.testing.assert_array_equal(inds.cpu().numpy(), oracle.furthest_point_sample(xyz, m, "A"))
    new_xyz = np.take_along_axis(xyz, inds.cpu().numpy().astype(np.int64)[..., None], axis=1)
    q = T(new_xyz, dev)
    idx = pu.ball_query(0.04, ns, x, q)
    np.testing.assert_array_equal(idx.cpu().numpy(), oracle.ball_query(0.04, ns, xyz, new_xyz))
    rng = np.random.default_rng(1)
    v = rng.normal(size=(B, m, 3)).astype(np.float32)
    rot = np.ascontiguousarray(scenes.viewpoint_rotations(-v, np.zeros((B, m), np.float32)).reshape(B, m, 9).astype(np.float32))
    cidx = pu.cylinder_query(0.05, -0.02, 0.04, ns, x, q, T(rot, dev))
    np.testing.assert_array_equal(cidx.cpu().numpy(), oracle.cylinder_query(0.05, -0.02, 0.04, ns, xyz, new_xyz, rot))
    feats = rng.normal(size=(B, 6, N)).astype(np.float32)
    f = T(feats, dev).requires_grad_(True)
    out = pu.grouping_operation(f, idx)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.grouping_operation(feats, idx.cpu().numpy()))
    gout = rng.normal(size=tuple(out.shape)).astype(np.float32)
    out.backward(T(gout, dev))
    assert_grad_close(f.grad.cpu().numpy(), oracle.grouping_operation_grad(gout, idx.cpu().numpy(), N))


def test_non_contiguous_and_wrong_dtype_inputs_raise(dev):
    xyz = T(scenes.scene_batch(range(2), 300, "uniform"), dev)
    with pytest.raises((RuntimeError, AssertionError)):
        pu.ball_query(0.1, 4, xyz.transpose(1, 2), xyz)
    with pytest.raises((RuntimeError, AssertionError)):
        pu.furthest_point_sample(xyz.double(), 8)
    with pytest.raises((RuntimeError, AssertionError)):
        pu.grouping_operation(torch.randn((2, 3, 300), device=dev), torch.zeros((2, 4, 4), dtype=torch.int64, device=dev))


@pytest.mark.parametrize("B,n,m,C", [(2, 20000, 1024, 256), (3, 1024, 512, 256), (2, 512, 256, 64), (1, 777, 5, 7), (2, 2050, 1, 4),
                                     (1, 3, 2, 1), (2, 5000, 3000, 12)])
def test_fused_feature_propagation_equals_the_three_steps(dev, ref_a, B, n, m, C):
    """gb_three_interpolation (search + weights + gather in one launch) == the reference's three_nn, the weight arithmetic
    of pointnet2_modules.py:413-416 and three_interpolate, bit for bit (values, and the idx / weight it keeps for backward);
    gradient == three_interpolate_grad of the reference within 1e-5."""
    unknown = T(scenes.scene_batch(range(B), n, "tabletop" if n >= 1000 else "uniform"), dev)
    known = T(scenes.scene_batch(range(50, 50 + B), m, "tabletop" if m >= 1000 else "uniform"), dev)
    g = torch.Generator(device="cpu").manual_seed(4)
    feats = torch.randn((B, C, m), generator=g).to(dev)
    d2, idx = ref_a.three_nn(unknown, known)
    r = 1.0 / (torch.sqrt(d2) + 1e-8)
    w = r / torch.sum(r, dim=2, keepdim=True)
    want = ref_a.three_interpolate(feats, idx, w)
    got_inf = pu.three_interpolation(unknown, known, feats)  # inference: nothing but the output is written
    assert torch.equal(got_inf, want)
    f = feats.clone().requires_grad_(True)
    l0 = _lib.launch_count()
    got = pu.three_interpolation(unknown, known, f)
    assert _lib.launch_count() - l0 == 1
    assert torch.equal(got.detach(), want)
    go = torch.randn(want.shape, generator=g).to(dev)
    got.backward(go)
    assert_grad_close(f.grad.cpu().numpy(), ref_a.three_interpolate_grad(go, idx, w, m).cpu().numpy())
    assert torch.equal(gb_up.three_interpolation(unknown, known, feats), want)  # the ModifiedNetTools entry point (two launches)


def test_collision_pack_culling_never_changes_the_counts(dev):
    """The collision kernel skips packs of 32 points whose bounds the gripper cannot reach.  Counts must equal the exhaustive
    test (the C oracle) for clouds in any order (voxel-key order, shuffled), for rotations that are not orthonormal (scaled,
    sheared: every pack is tested then), for grasps far outside the cloud, and for NaN / inf in centres and points."""
    from graspbalance_b200.collision_detector import collision_counts
    rng = np.random.default_rng(21)
    det = ModelFreeCollisionDetector(scenes.tabletop_scene(15, 20000).astype(np.float64), voxel_size=0.01, device=dev)
    pts_sorted = det.scene_points
    G = 256
    gs = scenes.grasp_set(16, pts_sorted, G)
    Tm, Rm = gs["translations"].copy(), gs["rotation_matrices"].copy()
    Rm[10:40] *= 1.7                                          # scaled
    Rm[40:70] += rng.normal(scale=0.2, size=(30, 3, 3))       # sheared
    Tm[70:90] += 5.0                                          # far outside the cloud
    Tm[90, 0], Tm[91, 1], Rm[92, 1, 1] = np.nan, np.inf, np.nan
    thr = oracle.collision_thresholds(gs["heights"], gs["depths"], gs["widths"], 0.05)
    for pts in (pts_sorted, pts_sorted[rng.permutation(pts_sorted.shape[0])], np.concatenate([pts_sorted[:777], [[np.nan, 0.1, 0.5], [np.inf, 0.0, 0.5]]])):
        got = collision_counts(T(pts, dev), T(Tm, dev), T(Rm, dev), T(thr, dev)).cpu().numpy()
        want = oracle.collision_counts(pts, Tm, Rm, gs["heights"], gs["depths"], gs["widths"], 0.05)
        np.testing.assert_array_equal(got, want)
    assert int(want[:, 0].sum()) > 0


@pytest.mark.parametrize("B,N,m", [(3, 20000, 512), (2, 2048, 300), (4, 9000, 64)])
def test_fps_footprint_hint_never_changes_the_picks(dev, B, N, m):
    """gb_fps_xyz_hint: fewer, fuller CTAs per scene for a sampling chain that runs beside other kernels -- any cap gives the
    oracle's picks and coordinates (the tie order does not depend on the decomposition)."""
    xyz_np = scenes.scene_batch(range(B), N, "tabletop")
    xyz_np[:, N // 2:N // 2 + 200] = xyz_np[:, :200]  # duplicated points: ties
    xyz = T(xyz_np, dev)
    want = oracle.furthest_point_sample(xyz_np, m, "A")
    for cap in (0, 1, 2, 4, 8, 16):
        inds, new_xyz = pu.furthest_point_sample_xyz(xyz, m, cap)
        np.testing.assert_array_equal(inds.cpu().numpy(), want)
        assert torch.equal(new_xyz, torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)))
