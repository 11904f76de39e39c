"""GPU tests of the grouper modules: QueryAndGroup / CylinderQueryAndGroup of module A
(PointNet/pointnet2_utils.py:152-308) and QueryAndGroup of module B (ModifiedNetTools/group.py:147-180).

The product runs them as fused launches (gb_group_xyz + gb_group_fwd_strided, backward gb_group_bwd_strided).  They are
checked against a literal torch restatement of the reference modules' forward (transpose, index gather, subtract, divide,
permute, matmul, permute, cat -- the op sequence of the cited lines, with torch.gather standing in for
grouping_operation, the one equivalence the reference itself asserts, group.py:92-96) and against the product's own
unfused composition of the individual operators.
  * indices, grouped features and centred / scaled coordinates: BIT-EXACT;
  * rotated coordinates: |diff| <= 1e-6 (torch.matmul's accumulation order is cuBLAS's business);
  * feature gradients: 1e-5 relative (float summation order), as for grouping_operation.
"""
import numpy as np
import pytest
import torch

import oracle
from graspbalance_b200 import group as gb_group
from graspbalance_b200 import pointnet2_utils as pu
from graspbalance_b200 import scenes

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def torch_group(features, idx):
    """grouping_operation as an index gather (group.py:92-96 torch_grouping_operation)."""
    B, C, N = features.shape
    _, m, ns = idx.shape
    return torch.gather(features, 2, idx.long().reshape(B, 1, m * ns).expand(-1, C, -1)).reshape(B, C, m, ns)


def reference_query_and_group(idx, xyz, new_xyz, features, radius, normalize_xyz, use_xyz=True, rot=None):
    """pointnet2_utils.py:178-207 / 281-308 with the index tensor given."""
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = torch_group(xyz_trans, idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    if rot is not None:
        g = grouped_xyz.permute(0, 2, 3, 1).contiguous()
        g = torch.matmul(g, rot)
        grouped_xyz = g.permute(0, 3, 1, 2).contiguous()
    if features is not None:
        gf = torch_group(features, idx)
        new_features = torch.cat([grouped_xyz, gf], dim=1) if use_xyz else gf
    else:
        new_features = grouped_xyz
    return new_features, grouped_xyz


def _inputs(dev, B, N, m, C, seed, kind="tabletop"):
    xyz = scenes.scene_batch(range(seed, seed + B), N, kind)
    rng = np.random.default_rng(seed)
    pick = np.stack([rng.choice(N, m, replace=m > N) for _ in range(B)])
    new_xyz = np.take_along_axis(xyz, pick[..., None], axis=1)
    feats = rng.normal(size=(B, C, N)).astype(np.float32) if C else None
    return xyz, new_xyz, feats


@pytest.mark.parametrize("B,N,m,C,r,ns,norm", [(2, 20000, 128, 16, 0.05, 64, True), (2, 2048, 96, 5, 0.1, 32, False),
                                              (1, 777, 33, 0, 0.2, 16, True), (2, 300, 20, 3, 0.3, 6, True),
                                              (1, 4000, 64, 130, 0.04, 64, True)])
def test_query_and_group_module_a(dev, B, N, m, C, r, ns, norm):
    torch.backends.cuda.matmul.allow_tf32 = False
    xyz, new_xyz, feats = _inputs(dev, B, N, m, C, 31)
    x, q = T(xyz, dev), T(new_xyz, dev)
    f = T(feats, dev).requires_grad_(True) if C else None
    mod = pu.QueryAndGroup(r, ns, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=norm)
    out, gxyz = mod(x, q, f)
    idx_want = oracle.ball_query(r, ns, xyz, new_xyz)
    f_ref = T(feats, dev).requires_grad_(True) if C else None
    want, want_xyz = reference_query_and_group(T(idx_want, dev), x, q, f_ref, r, norm)
    assert out.shape == want.shape == (B, 3 + C, m, ns)
    assert torch.equal(out, want)  # coordinates (subtract, multiply by 1/r) and gathered features: bit-exact
    assert torch.equal(gxyz, want_xyz)
    if C:
        g = torch.randn(out.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        out.backward(g)
        want.backward(g)
        scale = f_ref.grad.abs().max().item()
        assert (f.grad - f_ref.grad).abs().max().item() <= 1e-5 * max(scale, 1e-30)
    # the unfused composition (taken when a coordinate tensor needs a gradient) gives the same result
    xg = x.clone().requires_grad_(True)
    out2, _ = mod(xg, q, T(feats, dev) if C else None)
    assert torch.equal(out2.detach(), out.detach())


@pytest.mark.parametrize("B,N,m,C,ns,hmax,rotate", [(2, 20000, 128, 0, 64, 0.04, True), (2, 3000, 50, 8, 16, 0.02, True),
                                                   (1, 2048, 64, 4, 32, 0.01, False), (2, 500, 9, 0, 5, 0.04, True)])
def test_cylinder_query_and_group(dev, B, N, m, C, ns, hmax, rotate):
    torch.backends.cuda.matmul.allow_tf32 = False
    xyz, new_xyz, feats = _inputs(dev, B, N, m, C, 41)
    rng = np.random.default_rng(5)
    v = rng.normal(size=(B, m, 3)).astype(np.float32)
    rot = scenes.viewpoint_rotations(-v, rng.uniform(0, np.pi, (B, m)).astype(np.float32)).astype(np.float32)  # [B,m,3,3]
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot, dev)
    f = T(feats, dev).requires_grad_(True) if C else None
    mod = pu.CylinderQueryAndGroup(0.05, -0.02, hmax, ns, use_xyz=True, ret_grouped_xyz=True, rotate_xyz=rotate)
    res = mod(x, q, R, f)
    out, gxyz = res
    idx_want = oracle.cylinder_query(0.05, -0.02, hmax, ns, xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9)))
    f_ref = T(feats, dev).requires_grad_(True) if C else None
    want, want_xyz = reference_query_and_group(T(idx_want, dev), x, q, f_ref, 0.05, False, rot=R if rotate else None)
    assert out.shape == want.shape
    if rotate:
        assert (out[:, :3] - want[:, :3]).abs().max().item() <= 1e-6
        assert torch.equal(out[:, 3:], want[:, 3:])
    else:
        assert torch.equal(out, want)
    assert (gxyz - want_xyz).abs().max().item() <= 1e-6
    if C:
        g = torch.randn(out.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
        out.backward(g)
        want.backward(g)
        assert (f.grad - f_ref.grad).abs().max().item() <= 1e-5 * max(f_ref.grad.abs().max().item(), 1e-30)


@pytest.mark.parametrize("B,N,C,r,ns,norm", [(2, 2048, 32, 0.08, 64, False), (2, 1024, 7, 0.2, 32, True), (1, 333, 4, 0.4, 10, False)])
def test_query_and_group_module_b(dev, B, N, C, r, ns, norm):
    xyz, _, feats = _inputs(dev, B, N, 1, C, 51)
    x = T(xyz, dev)
    f = T(feats, dev).requires_grad_(True)
    mod = gb_group.QueryAndGroup(r, ns, normalize_dp=norm)
    dp, fj = mod(x, x, f)  # InvResMLP: queries are the support points themselves (drp.py:62-67)
    idx = T(oracle.ball_query(r, ns, xyz, xyz), dev)
    f_ref = T(feats, dev).requires_grad_(True)
    want, want_xyz = reference_query_and_group(idx, x, x, f_ref, r, norm, use_xyz=False)
    assert torch.equal(dp, want_xyz)
    assert torch.equal(fj, want)
    g = torch.randn(fj.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    fj.backward(g)
    want.backward(g)
    assert (f.grad - f_ref.grad).abs().max().item() <= 1e-5 * max(f_ref.grad.abs().max().item(), 1e-30)
    # unfused composition
    xg = x.clone().requires_grad_(True)
    dp2, fj2 = mod(xg, xg, T(feats, dev))
    assert torch.equal(dp2.detach(), dp) and torch.equal(fj2, fj.detach())


def test_strided_group_entries_match_the_contiguous_ones(dev):
    """gb_group_fwd_strided / gb_group_bwd_strided against gb_group_fwd / gb_group_bwd on odd shapes (generic paths too)."""
    from graspbalance_b200 import _ext as A, _lib
    rng = np.random.default_rng(0)
    for (B, C, N, m, ns, pad) in ((2, 6, 500, 17, 4, 3), (1, 5, 100, 7, 3, 2), (2, 16, 5000, 64, 16, 3), (1, 4, 20000, 32, 64, 5)):
        feats = T(rng.normal(size=(B, C, N)).astype(np.float32), dev)
        idx = T(rng.integers(0, N, (B, m, ns)).astype(np.int32), dev)
        per = m * ns
        want = A.group_points(feats, idx)
        big = torch.full((B, pad + C, m, ns), -7.0, device=dev)
        _lib.call("gb_group_fwd_strided", feats, feats.data_ptr(), idx.data_ptr(), big.data_ptr() + 4 * pad * per, B, C, N, m, ns,
                  (pad + C) * per)
        assert torch.equal(big[:, pad:], want) and bool((big[:, :pad] == -7.0).all())
        gout = torch.randn((B, pad + C, m, ns), device=dev)
        want_g = A.group_points_grad(gout[:, pad:].contiguous(), idx, N)
        got = torch.empty((B, C, N), device=dev)
        _lib.call("gb_group_bwd_strided", gout, gout.data_ptr() + 4 * pad * per, idx.data_ptr(), got.data_ptr(), B, C, N, m, ns,
                  (pad + C) * per, 1)
        assert (got - want_g).abs().max().item() <= 1e-5 * max(want_g.abs().max().item(), 1e-30)


@pytest.mark.parametrize("B,C,N,m,ns,scale", [
    (32, 128, 2048, 2048, 64, None),   # InvResMLP level 0 of the bench step: aligned partition, many waves
    (4, 256, 1024, 1024, 32, 5.0),     # 16-channel fills
    (2, 5, 2048, 96, 32, 10.0),        # flattened equal split, ragged channel chunk
    (3, 2, 700, 40, 8, None),          # V = 2 kernel
    (1, 1, 300, 8, 4, 2.0),            # V = 1 kernel
    (2, 16, 20000, 128, 64, 25.0),     # rows beyond shared memory four at a time: coordinates as a launch of their own
    (2, 7, 500, 17, 6, 4.0),           # nsample % 4 != 0: generic kernels
])
def test_group_xyz_feat_is_the_two_launches_in_one(dev, B, C, N, m, ns, scale):
    """gb_group_xyz_feat against gb_group_xyz + gb_group_fwd_strided, bit for bit, in the layouts of both grouper variants:
    variant A writes rows 0..2 and 3.. of one [B,3+C,m,ns] tensor, variant B two tensors (group.py:167-179)."""
    from graspbalance_b200 import _lib
    rng = np.random.default_rng(B * 1000 + C)
    xyz = T(rng.uniform(-1, 1, (B, N, 3)).astype(np.float32), dev)
    new_xyz = T(rng.uniform(-1, 1, (B, m, 3)).astype(np.float32), dev)
    feats = T(rng.normal(size=(B, C, N)).astype(np.float32), dev)
    idx = T(rng.integers(0, N, (B, m, ns)).astype(np.int32), dev)
    per = m * ns
    sc, use = (float(scale), 1) if scale is not None else (0.0, 0)
    # variant A layout
    want = torch.full((B, 3 + C, m, ns), -7.0, device=dev)
    _lib.call("gb_group_xyz", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), None, want.data_ptr(), B, N, m, ns, sc, use, (3 + C) * per)
    _lib.call("gb_group_fwd_strided", feats, feats.data_ptr(), idx.data_ptr(), want.data_ptr() + 12 * per, B, C, N, m, ns, (3 + C) * per)
    got = torch.full((B, 3 + C, m, ns), -9.0, device=dev)
    n0 = _lib.launch_count()
    _lib.call("gb_group_xyz_feat", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), got.data_ptr(), (3 + C) * per, sc, use,
              feats.data_ptr(), got.data_ptr() + 12 * per, (3 + C) * per, B, C, N, m, ns)
    launches = _lib.launch_count() - n0
    assert torch.equal(got, want)
    staged = ns % 4 == 0 and N * 16 <= 200 * 1024
    assert launches == (1 if staged else 2)
    # variant B layout
    gx, gf = torch.full((B, 3, m, ns), -9.0, device=dev), torch.full((B, C, m, ns), -9.0, device=dev)
    _lib.call("gb_group_xyz_feat", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), gx.data_ptr(), 3 * per, sc, use,
              feats.data_ptr(), gf.data_ptr(), C * per, B, C, N, m, ns)
    assert torch.equal(gx, want[:, :3]) and torch.equal(gf, want[:, 3:])
    # no features: coordinates only; empty batch: nothing
    gx.fill_(-9.0)
    _lib.call("gb_group_xyz_feat", xyz, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), gx.data_ptr(), 3 * per, sc, use,
              None, None, 0, B, 0, N, m, ns)
    assert torch.equal(gx, want[:, :3])
    _lib.call("gb_group_xyz_feat", xyz, None, None, None, None, 0, sc, use, None, None, 0, 0, C, N, m, ns)


# ---- multi-depth grasp crop (TrainModel/modules.py:87-124) -------------------------------------------------
def _crop_inputs(dev, B, N, m, seed, kind="tabletop"):
    xyz, new_xyz, _ = _inputs(dev, B, N, m, 0, seed, kind)
    rng = np.random.default_rng(seed + 1)
    v = rng.normal(size=(B, m, 3)).astype(np.float32)
    rot = scenes.viewpoint_rotations(-v, rng.uniform(0, np.pi, (B, m)).astype(np.float32)).astype(np.float32)
    return xyz, new_xyz, rot


@pytest.mark.parametrize("B,N,m,ns,radius,hmaxs", [
    (2, 20000, 128, 64, 0.05, [0.01, 0.02, 0.03, 0.04]),       # cell-grid path, the model's depths
    (2, 20000, 64, 16, 0.08, [0.04, 0.01, 0.03, 0.02]),        # unsorted depths, lists that fill up early
    (1, 20000, 40, 64, 0.02, [0.01, 0.04]),                    # two depths, sparse hits (padding with the first hit)
    (2, 3000, 50, 16, 0.05, [0.01, 0.02, 0.03, 0.04]),         # full-scan path (n < 4096)
    (1, 500, 9, 5, 0.05, [0.02]),                              # one depth
    (2, 6000, 33, 32, 0.05, [0.03, 0.03, float("nan")]),       # duplicate and NaN depths
    (1, 5000, 16, 8, 0.05, [-0.05, 0.0, 0.01, 0.04]),          # hmax below hmin: an empty cylinder
    (1, 8000, 24, 64, 0.6, [0.01, 0.02, 0.03, 0.04]),          # radius comparable to the scene: the grid declines
])
def test_cylinder_query_multi_matches_one_call_per_depth(dev, B, N, m, ns, radius, hmaxs):
    xyz, new_xyz, rot = _crop_inputs(dev, B, N, m, 61)
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot.reshape(B, m, 9), dev)
    got = pu.cylinder_query_multi(radius, -0.02, hmaxs, ns, x, q, R)
    assert got.shape == (B, m, len(hmaxs), ns) and got.dtype == torch.int32
    want = oracle.cylinder_query_multi(radius, -0.02, hmaxs, ns, xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9)))
    assert np.array_equal(got.cpu().numpy(), want)
    for d, h in enumerate(hmaxs):  # and against the product's own single-depth entry point
        assert torch.equal(got[:, :, d], pu.cylinder_query(radius, -0.02, h, ns, x, q, R))


def test_cylinder_query_multi_uniform_and_degenerate_rotations(dev):
    B, N, m, ns = 2, 12000, 48, 32
    xyz, new_xyz, rot = _crop_inputs(dev, B, N, m, 71, kind="uniform")
    rot[0, :8] *= 1.5          # not orthonormal: the grid path searches every cell for these seeds
    rot[1, 3] = np.nan
    xyz[1, 100] = np.inf
    hmaxs = [0.01, 0.02, 0.03, 0.04]
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot.reshape(B, m, 9), dev)
    got = pu.cylinder_query_multi(0.1, -0.02, hmaxs, ns, x, q, R)
    want = oracle.cylinder_query_multi(0.1, -0.02, hmaxs, ns, xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9)))
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("B,N,m,ns", [(2, 20000, 128, 64), (1, 2500, 30, 16)])
def test_grasp_width_grouping_matches_the_reference_loop(dev, B, N, m, ns):
    """GraspWidthGrouping.group() == torch.stack of the per-depth CylinderQueryAndGroup results viewed as
    [B, 3, num_seed*num_depth, nsample] (modules.py:107-117); the product's per-depth groupers give the loop."""
    from graspbalance_b200.modules import GraspWidthGrouping
    torch.backends.cuda.matmul.allow_tf32 = False
    xyz, new_xyz, rot = _crop_inputs(dev, B, N, m, 81)
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot, dev)
    mod = GraspWidthGrouping(ns, 3, cylinder_radius=0.05, hmin=-0.02, hmax_list=[0.01, 0.02, 0.03, 0.04]).to(dev)
    got = mod.group(q, x, R)
    loop = torch.stack([g(x, q, R) for g in mod.groupers], dim=3).view(B, -1, m * 4, ns)
    assert got.shape == loop.shape == (B, 3, m * 4, ns)
    assert torch.equal(got, loop)  # same kernels, same arithmetic: bit-exact
    # and against the literal restatement of the reference module on the oracle's indices
    for d, h in enumerate(mod.hmax_list):
        idx = T(oracle.cylinder_query(0.05, -0.02, h, ns, xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9))), dev)
        want, _ = reference_query_and_group(idx, x, q, None, 0.05, False, rot=R)
        assert (got.view(B, 3, m, 4, ns)[:, :, :, d] - want).abs().max().item() <= 1e-6
    out = mod.eval()(q, x, R)
    assert out.shape == (B, 256, m, 4)


@pytest.mark.parametrize("B,N,m,ns,radii,hmaxs", [
    (2, 20000, 128, 64, [0.02, 0.04, 0.06, 0.08], [0.01, 0.02, 0.03, 0.04]),   # the model's four scales x four depths
    (2, 20000, 48, 16, [0.08, 0.02, 0.05], [0.04, 0.01]),                      # unsorted radii and depths
    (2, 3000, 40, 16, [0.02, 0.04, 0.06, 0.08], [0.01, 0.02, 0.03, 0.04]),     # full-scan path (n < 4096)
    (1, 6000, 20, 32, [0.05, 0.05, float("nan"), 0.0], [0.03, float("nan")]),  # duplicate, NaN and zero radii
    (1, 9000, 16, 64, [0.7, 0.05], [0.01, 0.04]),                              # one radius as large as the scene: the grid declines
])
def test_cylinder_query_multi_radius_matches_one_call_per_cylinder(dev, B, N, m, ns, radii, hmaxs):
    xyz, new_xyz, rot = _crop_inputs(dev, B, N, m, 91)
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot.reshape(B, m, 9), dev)
    got = pu.cylinder_query_multi_radius(radii, -0.02, hmaxs, ns, x, q, R)
    assert got.shape == (len(radii), B, m, len(hmaxs), ns) and got.dtype == torch.int32
    want = oracle.cylinder_query_multi_radius(radii, -0.02, hmaxs, ns, xyz, new_xyz, np.ascontiguousarray(rot.reshape(B, m, 9)))
    assert np.array_equal(got.cpu().numpy(), want)
    for k, r in enumerate(radii):
        for d, h in enumerate(hmaxs):
            assert torch.equal(got[k, :, :, d], pu.cylinder_query(r, -0.02, h, ns, x, q, R))


def test_multi_scale_group_matches_the_four_modules(dev):
    from graspbalance_b200.modules import GraspWidthGrouping, multi_scale_group
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, m, ns = 2, 20000, 96, 64
    xyz, new_xyz, rot = _crop_inputs(dev, B, N, m, 95)
    x, q, R = T(xyz, dev), T(new_xyz, dev), T(rot, dev)
    mods = [GraspWidthGrouping(ns, 3, cylinder_radius=0.08 * s, hmin=-0.02, hmax_list=[0.01, 0.02, 0.03, 0.04], mlps=torch.nn.Identity())
            for s in (0.25, 0.5, 0.75, 1.0)]
    got = multi_scale_group(mods, q, x, R)
    for g, mod in zip(got, mods):
        assert torch.equal(g, mod.group(q, x, R))


def test_empty_batches_and_empty_clouds(dev):
    """Shapes the reference accepts: an empty batch (even with N = 0) is a no-op; a query against an empty cloud finds nothing,
    so every slot keeps the zero of the reference's zero-filled output (ball_query.cpp:25-27, cylinder_query.cpp:27-29)."""
    from graspbalance_b200 import _ext as gb_a, knn_modules
    e3 = torch.zeros((0, 0, 3), device=dev)
    assert tuple(pu.furthest_point_sample(e3, 0).shape) == (0, 0)
    assert tuple(gb_a.ball_query(e3, e3, 0.1, 8).shape) == (0, 0, 8)
    assert tuple(gb_a.three_nn(e3, e3)[1].shape) == (0, 0, 3)
    assert tuple(knn_modules.knn_k(torch.zeros((0, 3, 5), device=dev), torch.zeros((0, 3, 0), device=dev), 1).shape) == (0, 1, 0)
    new = torch.rand((2, 5, 3), device=dev)
    none = torch.zeros((2, 0, 3), device=dev)
    rot = torch.eye(3, device=dev).reshape(1, 1, 9).repeat(2, 5, 1).contiguous()
    for idx in (gb_a.ball_query(new, none, 0.1, 8), gb_a.cylinder_query(new, none, rot, 0.1, -0.02, 0.04, 8)):
        assert tuple(idx.shape) == (2, 5, 8) and idx.dtype == torch.int32 and int(idx.abs().sum()) == 0
    with pytest.raises(RuntimeError):  # samples of an empty cloud are undefined
        pu.furthest_point_sample(none, 4)


def _reference_pred_decode(end_points):
    """TrainModel/graspbalance.py:139-192 and loss_utils.py:33-49, literally (per-scene loop), as the check of decode.pred_decode."""
    def to_matrix(batch_towards, batch_angle):
        axis_x = batch_towards
        ones = torch.ones(axis_x.shape[0], dtype=axis_x.dtype, device=axis_x.device)
        zeros = torch.zeros(axis_x.shape[0], dtype=axis_x.dtype, device=axis_x.device)
        axis_y = torch.stack([-axis_x[:, 1], axis_x[:, 0], zeros], dim=-1)
        mask_y = (torch.norm(axis_y, dim=-1) == 0)
        axis_y[mask_y, 1] = 1
        axis_x = axis_x / torch.norm(axis_x, dim=-1, keepdim=True)
        axis_y = axis_y / torch.norm(axis_y, dim=-1, keepdim=True)
        axis_z = torch.cross(axis_x, axis_y, dim=-1)
        sin, cos = torch.sin(batch_angle), torch.cos(batch_angle)
        R1 = torch.stack([ones, zeros, zeros, zeros, cos, -sin, zeros, sin, cos], dim=-1).reshape([-1, 3, 3])
        return torch.matmul(torch.stack([axis_x, axis_y, axis_z], dim=-1), R1)
    preds = []
    for i in range(len(end_points['point_clouds'])):
        objectness_score = end_points['objectness_score'][i].float()
        grasp_score = end_points['grasp_score_pred'][i].float()
        grasp_center = end_points['fp2_xyz'][i].float()
        approaching = -end_points['grasp_top_view_xyz'][i].float()
        grasp_width = torch.clamp(1.2 * end_points['grasp_width_pred'][i], min=0, max=0.1)
        grasp_tolerance = end_points['grasp_tolerance_pred'][i]
        cls = torch.argmax(end_points['grasp_angle_cls_pred'][i], 0)
        grasp_angle = cls.float() / 12 * np.pi
        cls_ = cls.unsqueeze(0)
        grasp_score = torch.gather(grasp_score, 0, cls_).squeeze(0)
        grasp_width = torch.gather(grasp_width, 0, cls_).squeeze(0)
        grasp_tolerance = torch.gather(grasp_tolerance, 0, cls_).squeeze(0)
        dcls = torch.argmax(grasp_score, 1, keepdims=True)
        grasp_depth = (dcls.float() + 1) * 0.01
        grasp_score, grasp_angle = torch.gather(grasp_score, 1, dcls), torch.gather(grasp_angle, 1, dcls)
        grasp_width, grasp_tolerance = torch.gather(grasp_width, 1, dcls), torch.gather(grasp_tolerance, 1, dcls)
        mask = torch.argmax(objectness_score, 0) == 1
        grasp_score = grasp_score * torch.softmax(objectness_score, dim=0)[1, :].unsqueeze(1)
        grasp_score, grasp_width, grasp_depth = grasp_score[mask], grasp_width[mask], grasp_depth[mask]
        approaching, grasp_angle, grasp_center, grasp_tolerance = approaching[mask], grasp_angle[mask], grasp_center[mask], grasp_tolerance[mask]
        grasp_score = grasp_score * grasp_tolerance / 0.05
        Ns = grasp_angle.size(0)
        rot = to_matrix(approaching.view(Ns, 3), grasp_angle.view(Ns)).view(Ns, 9)
        preds.append(torch.cat([grasp_score, grasp_width, 0.02 * torch.ones_like(grasp_score), grasp_depth, rot, grasp_center,
                                -1 * torch.ones_like(grasp_score)], axis=-1))
    return preds


def test_pred_decode_and_on_device_collision_masks(dev):
    """decode.pred_decode == the reference's per-scene loop (graspbalance.py:139-192) on random head outputs, and its float32
    rows fed to the collision test on the device give the masks the reference flow gives (rows -> host -> GraspGroup arrays ->
    numpy detect, here the dtype-faithful oracle)."""
    from graspbalance_b200 import decode
    from graspbalance_b200.collision_detector import ModelFreeCollisionDetector
    B, Ns, A, D = 2, 1024, 12, 4
    g = torch.Generator(device="cpu").manual_seed(6)
    xyz_np = scenes.scene_batch([31, 32], 20000, "tabletop")
    seeds = torch.from_numpy(xyz_np[:, :Ns].copy()).to(dev)
    views = torch.randn((B, Ns, 3), generator=g).to(dev)
    views[0, :3] = torch.tensor([0.0, 0.0, 1.0])  # approach along z: the degenerate y axis of loss_utils.py:38-39
    ep = {"point_clouds": torch.from_numpy(xyz_np).to(dev), "objectness_score": torch.randn((B, 2, Ns), generator=g).to(dev),
          "grasp_score_pred": torch.rand((B, A, Ns, D), generator=g).to(dev), "fp2_xyz": seeds, "grasp_top_view_xyz": views,
          "grasp_angle_cls_pred": torch.randn((B, A, Ns, D), generator=g).to(dev),
          "grasp_width_pred": (0.1 * torch.rand((B, A, Ns, D), generator=g)).to(dev),
          "grasp_tolerance_pred": (0.05 * torch.rand((B, A, Ns, D), generator=g)).to(dev)}
    got, want = decode.pred_decode(ep), _reference_pred_decode(ep)
    for a, b in zip(got, want):
        assert a.shape == b.shape and a.dtype == torch.float32 and 100 < a.shape[0] < Ns
        assert torch.equal(a, b)
    masks = decode.collision_masks(got, [xyz_np[b].astype(np.float64) for b in range(B)], voxel_size=0.01)
    for b in range(B):
        rows = got[b].cpu().numpy()  # float32, as the arrays of the GraspGroup the reference builds from them
        det = ModelFreeCollisionDetector(xyz_np[b].astype(np.float64), voxel_size=0.01, device=dev)
        ref = oracle.collision_detect(det.scene_points, 0.01, rows[:, 13:16], rows[:, 4:13].reshape(-1, 3, 3), rows[:, 2], rows[:, 3],
                                      rows[:, 1], approach_dist=0.05, collision_thresh=0.01)
        assert masks[b].is_cuda and masks[b].dtype == torch.bool
        np.testing.assert_array_equal(masks[b].cpu().numpy(), ref)
        assert 0 < int(ref.sum()) < ref.shape[0]


@pytest.mark.parametrize("B,N,m,ns,C,r", [(2, 20000, 1024, 64, 35, 0.05), (3, 2048, 1024, 32, 128, 0.1), (2, 1024, 300, 16, 7, 0.3), (1, 500, 64, 5, 4, 0.2)])
def test_grouping_max_equals_group_then_max_pool(dev, B, N, m, ns, C, r):
    """gb_group_max_fwd / bwd == grouping_operation + F.max_pool2d over the samples (pointnet2_modules.py:173-175, 324-335):
    values bit for bit, gradients equal those autograd sends through max_pool2d and GroupingOperation (padded neighbourhoods
    repeat their first index: the first maximum wins in both), and the fused WOMLP-style module equals QueryAndGroup + pooling."""
    import torch.nn.functional as F
    xyz = torch.from_numpy(scenes.scene_batch(range(B), N, "tabletop" if N >= 1000 else "uniform")).to(dev)
    new_xyz = xyz[:, :m].contiguous()
    idx = pu.ball_query(r, ns, xyz, new_xyz)
    g = torch.Generator(device="cpu").manual_seed(8)
    feats = torch.randn((B, C, N), generator=g).to(dev)
    k7 = (N - 1) // 7
    feats[:, :, 0:7 * k7:7] = feats[:, :, 1:7 * k7 + 1:7]  # equal values at different sources: ties
    f1, f2 = feats.clone().requires_grad_(True), feats.clone().requires_grad_(True)
    want = F.max_pool2d(pu.grouping_operation(f1, idx), kernel_size=[1, ns]).squeeze(-1)
    got = pu.grouping_max(f2, idx)
    assert torch.equal(got, want)
    go = torch.randn(want.shape, generator=g).to(dev)
    want.backward(go), got.backward(go)
    scale = float(f1.grad.abs().max())
    assert float((f2.grad - f1.grad).abs().max()) <= 1e-5 * scale  # float atomics on both sides: summation order differs
    for norm in (False, True):
        ref = pu.QueryAndGroup(r, ns, use_xyz=True, normalize_xyz=norm)(xyz, new_xyz, feats)
        ref = F.max_pool2d(ref, kernel_size=[1, ns]).squeeze(-1)
        assert torch.equal(pu.QueryGroupMaxPool(r, ns, use_xyz=True, normalize_xyz=norm)(xyz, new_xyz, feats), ref)
