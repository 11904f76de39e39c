"""CPU-only tests of the oracle (oracle/gb_oracle.c + oracle/__init__.py): golden vectors from the reference, and
independent restatements of each op's definition on small inputs."""
import os

import numpy as np
import pytest

import oracle
from graspbalance_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- golden: the reference's own numpy collision detector (tests/golden/make_golden_collision.py) ----------------
def golden_scene(z, tag):
    """The down-sampled scene of a golden case: stored (a, b) or regenerated from its seed and checked (c, d)."""
    if tag + "_points" in z.files:
        return z[tag + "_points"]
    pts = oracle.voxel_down_sample(scenes.tabletop_scene(int(z[tag + "_seed"]), int(z[tag + "_n"])).astype(np.float64), float(z[tag + "_voxel"]))
    assert tuple(pts.shape) == tuple(z[tag + "_points_shape"]) and np.array_equal(pts.sum(0), z[tag + "_points_sum"])
    return pts


# a, b: small scenes; c: BASELINE config 1 as written (20k-point scene, voxel 0.01, 1024 grasps); d: a float32 grasp group
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
@pytest.mark.parametrize("fn", ["c", "numpy"])
def test_collision_oracle_matches_reference_detector(tag, fn):
    z = np.load(os.path.join(GOLDEN, "collision_ref.npz"))
    if fn == "numpy" and tag == "c":
        pytest.skip("the whole-array form needs a 143 MB temporary per mask at this size; the C form covers it")
    f = oracle.collision_detect if fn == "c" else oracle.collision_detect_numpy
    pts = golden_scene(z, tag)
    r = f(pts, float(z[tag + "_voxel"]), z[tag + "_translations"], z[tag + "_rotation_matrices"], z[tag + "_heights"],
          z[tag + "_depths"], z[tag + "_widths"], approach_dist=0.05, collision_thresh=0.01, return_empty_grasp=True, return_ious=True)
    assert (r[0] == z[tag + "_collision"]).all()
    assert (r[1] == z[tag + "_empty"]).all()
    for a, b in zip(r[2], z[tag + "_ious"]):
        np.testing.assert_array_equal(a, b)
    plain = f(pts, float(z[tag + "_voxel"]), z[tag + "_translations"], z[tag + "_rotation_matrices"], z[tag + "_heights"],
              z[tag + "_depths"], z[tag + "_widths"], approach_dist=0.05, collision_thresh=0.01)
    assert isinstance(plain, np.ndarray) and plain.dtype == np.bool_ and (plain == z[tag + "_collision"]).all()


# ---- golden: the reference's CUDA extensions run on a B200 (tests/golden/make_golden_gpu.py) ---------------------
@pytest.fixture(scope="module")
def gpu_golden():
    p = os.path.join(GOLDEN, "ref_gpu_small.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/ref_gpu_small.npz not generated yet (needs one GPU run of make_golden_gpu.py)")
    return np.load(p)


def test_oracle_matches_reference_cuda_outputs(gpu_golden):
    z = gpu_golden
    xyz, q = z["fps_xyz"], z["q_xyz"]
    np.testing.assert_array_equal(oracle.furthest_point_sample(xyz, 400, "A"), z["fps_a"])
    np.testing.assert_array_equal(oracle.furthest_point_sample(xyz, 400, "B"), z["fps_b"])
    np.testing.assert_array_equal(oracle.furthest_point_sample(z["fps_small_xyz"], 60, "A"), z["fps_small_a"])
    np.testing.assert_array_equal(oracle.ball_query(0.05, 16, xyz, q), z["ball_a"])
    np.testing.assert_array_equal(oracle.ball_query(0.05, 16, xyz, q), z["ball_b"])
    np.testing.assert_array_equal(oracle.cylinder_query(0.05, -0.02, 0.04, 16, xyz, q, z["cyl_rot"]), z["cyl_a"])
    d2, i3 = oracle.three_nn_dist2(xyz[:, :500], q)
    np.testing.assert_array_equal(i3, z["nn_idx"])
    np.testing.assert_array_equal(d2, z["nn_d2"])
    np.testing.assert_array_equal(oracle.three_interpolate(z["interp_f"], z["nn_idx"], z["interp_w"]), z["interp_out"])
    g = oracle.three_interpolate_grad(z["interp_go"], z["nn_idx"], z["interp_w"], 64)
    assert np.abs(g - z["interp_grad"]).max() <= 1e-5 * np.abs(z["interp_grad"]).max()
    np.testing.assert_array_equal(oracle.grouping_operation(z["group_f"], z["group_idx"]), z["group_out"])
    gg = oracle.grouping_operation_grad(z["group_go"], z["group_idx"], 3000)
    assert np.abs(gg - z["group_grad"]).max() <= 1e-5 * np.abs(z["group_grad"]).max()
    for k in (1, 8):
        np.testing.assert_array_equal(oracle.knn(z["knn_ref"], z["knn_query"], k), z[f"knn_k{k}"])


# ---- definitions, restated independently in numpy ------------------------------------------------------------------
def _d2(a, b):
    """fp32 squared distance with the reference's contraction: fma(dz,dz, fma(dx,dx, dy*dy)), via float64 emulation."""
    dx, dy, dz = (a[..., 0] - b[..., 0]).astype(np.float32), (a[..., 1] - b[..., 1]).astype(np.float32), (a[..., 2] - b[..., 2]).astype(np.float32)
    t = (dy * dy).astype(np.float32)                                    # FMUL
    t = (dx.astype(np.float64) * dx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)   # FFMA (exact product, one rounding)
    return (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)


def test_ball_query_definition():
    xyz = scenes.scene_batch([3], 1500, "uniform")
    q = xyz[:, :40] + np.float32(0.01)
    got = oracle.ball_query(0.12, 8, xyz, q)
    r2 = np.float32(0.12) * np.float32(0.12)
    for j in range(40):
        hits = np.nonzero(_d2(q[0, j][None], xyz[0]) < r2)[0]
        want = np.zeros(8, np.int32)
        if len(hits):
            want[:] = hits[0]
            want[:min(8, len(hits))] = hits[:8]
        np.testing.assert_array_equal(got[0, j], want)
    far = oracle.ball_query(1e-6, 4, xyz, q + 5)
    assert (far == 0).all()  # no hit: zeros


def test_three_nn_definition_and_degenerate_sizes():
    rng = np.random.default_rng(0)
    u, k = rng.uniform(-1, 1, (1, 200, 3)).astype(np.float32), rng.uniform(-1, 1, (1, 50, 3)).astype(np.float32)
    k[0, 25:30] = k[0, :5]  # ties -> lowest index first
    d2, idx = oracle.three_nn_dist2(u, k)
    for j in range(200):
        d = _d2(u[0, j][None], k[0])
        order = np.lexsort((np.arange(50), d))[:3]
        np.testing.assert_array_equal(idx[0, j], order)
        np.testing.assert_array_equal(d2[0, j], d[order])
    d2, idx = oracle.three_nn_dist2(u, k[:, :2])
    assert np.isinf(d2[..., 2]).all() and (idx[..., 2] == 0).all()  # m < 3: (float)1e40 = inf, index 0
    dist, _ = oracle.three_nn(u, k)
    np.testing.assert_array_equal(dist, np.sqrt(oracle.three_nn_dist2(u, k)[0]))


def _fps_by_key_rule(xyz, m, variant):
    """FPS restated through the closed-form tie rule: maximise the running distance; among equals the
    smallest bit-reversed (k mod BS), then the smallest k -- independent of the oracle's literal tree emulation."""
    n = xyz.shape[0]
    cap = 512 if variant == "A" else 1024
    bs = min(1 << int(np.floor(np.log2(n))), cap)
    L = int(np.log2(bs))
    k = np.arange(n)
    rev = np.array([int(format(v, f"0{L}b")[::-1], 2) if L else 0 for v in (k % bs)])
    tie = rev.astype(np.int64) * (1 << 32) + (k // bs)
    ok = np.ones(n, bool)
    if variant == "A":
        x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        t = (y * y).astype(np.float32)
        t = (x.astype(np.float64) ** 2 + t).astype(np.float32)
        mag = (z.astype(np.float64) ** 2 + t).astype(np.float32)
        ok = ~(mag.astype(np.float64) <= 1e-3)
    temp = np.full(n, 1e10, np.float32)
    out, old = [0], 0
    for _ in range(1, m):
        d = _d2(xyz, xyz[old][None])
        temp[ok] = np.minimum(d[ok], temp[ok])
        if not ok.any():
            old = 0
        else:
            best = temp[ok].max()
            cand = np.nonzero(ok & (temp == best))[0]
            old = int(cand[np.argmin(tie[cand])])
        out.append(old)
    return np.array(out, np.int32)


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("n,m", [(700, 120), (33, 40), (1, 3), (2050, 64)])
def test_fps_literal_emulation_equals_closed_form_tie_rule(variant, n, m):
    xyz = scenes.uniform_scene(n + m, n)
    if n > 20:
        xyz[n // 2:] = xyz[: n - n // 2]  # duplicated half: ties everywhere
        xyz[3] = [0.001, -0.002, 0.003]   # inside variant A's skip ball
    got = oracle.furthest_point_sample(xyz[None], m, variant)[0]
    np.testing.assert_array_equal(got, _fps_by_key_rule(xyz, m, variant))


def test_knn_definition():
    rng = np.random.default_rng(1)
    ref, qry = rng.uniform(-1, 1, (1, 3, 300)).astype(np.float32), rng.uniform(-1, 1, (1, 3, 20)).astype(np.float32)
    ref[0, :, 100:110] = ref[0, :, :10]
    got = oracle.knn(ref, qry, 5)
    for q in range(20):
        ssd = np.zeros(300, np.float32)
        for d in range(3):
            t = (ref[0, d] - qry[0, d, q]).astype(np.float32)
            ssd = (t.astype(np.float64) * t.astype(np.float64) + ssd.astype(np.float64)).astype(np.float32)
        order = np.lexsort((np.arange(300), ssd))[:5]
        np.testing.assert_array_equal(got[0, :, q], order + 1)  # 1-based


def test_group_gather_interpolate_definitions():
    rng = np.random.default_rng(2)
    f = rng.normal(size=(2, 3, 50)).astype(np.float32)
    idx = rng.integers(0, 50, (2, 7, 4)).astype(np.int32)
    out = oracle.grouping_operation(f, idx)
    want = np.stack([f[b][:, idx[b]] for b in range(2)])
    np.testing.assert_array_equal(out, want)
    go = rng.normal(size=out.shape).astype(np.float32)
    grad = oracle.grouping_operation_grad(go, idx, 50)
    want_g = np.zeros((2, 3, 50), np.float64)
    for b in range(2):
        for c in range(3):
            np.add.at(want_g[b, c], idx[b].reshape(-1), go[b, c].reshape(-1))
    np.testing.assert_allclose(grad, want_g, rtol=1e-5, atol=1e-6)
    gi = rng.integers(0, 50, (2, 9)).astype(np.int32)
    np.testing.assert_array_equal(oracle.gather_operation(f, gi), np.stack([f[b][:, gi[b]] for b in range(2)]))
    w = rng.uniform(0, 1, (2, 11, 3)).astype(np.float32)
    i3 = rng.integers(0, 50, (2, 11, 3)).astype(np.int32)
    got = oracle.three_interpolate(f, i3, w)
    ref = sum(np.stack([f[b][:, i3[b, :, t]] * w[b, :, t] for b in range(2)]).astype(np.float64) for t in range(3))
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_voxel_down_sample_restatement():
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 0.1, (500, 3))
    out = oracle.voxel_down_sample(pts, 0.02)
    origin = pts.min(0) - 0.01
    cells = np.floor((pts - origin) / 0.02).astype(np.int64)
    uniq = np.unique(cells, axis=0)
    assert out.shape[0] == uniq.shape[0]
    for c in uniq[:10]:
        mean = pts[(cells == c).all(1)].mean(0)
        assert np.abs(out - mean).sum(1).min() < 1e-12


# ---- the composite operations of the callers above the hot path: restated as the reference's own loops -----
def test_composite_oracles_are_the_reference_loops():
    """cylinder_query_multi = one cylinder_query per depth (modules.py:104-113); three_nn_weights = three_nn + the weight
    arithmetic of pointnet2_modules.py:413-416; furthest_point_sample_segments = one FPS per object (modules.py:201-209)."""
    rng = np.random.default_rng(3)
    xyz = scenes.scene_batch([1, 2], 1500, "tabletop")
    q = xyz[:, :40].copy()
    rot = scenes.viewpoint_rotations(-rng.normal(size=(2, 40, 3)).astype(np.float32), np.zeros((2, 40), np.float32)).reshape(2, 40, 9)
    depths = [0.01, 0.02, 0.03, 0.04]
    multi = oracle.cylinder_query_multi(0.05, -0.02, depths, 8, xyz, q, rot)
    assert multi.shape == (2, 40, 4, 8)
    for d, h in enumerate(depths):
        np.testing.assert_array_equal(multi[:, :, d], oracle.cylinder_query(0.05, -0.02, h, 8, xyz, q, rot))
    # nested cylinders: without truncation (nsample >= all hits) the hits of a shallower depth are hits of the deeper one
    full = oracle.cylinder_query_multi(0.05, -0.02, depths, 512, xyz, q, rot)
    assert all(set(full[b, j, 0]) <= set(full[b, j, 3]) or not full[b, j, 0].any() for b in range(2) for j in range(40))

    dist, idx, w = oracle.three_nn_weights(xyz, q)
    d0, i0 = oracle.three_nn(xyz, q)
    np.testing.assert_array_equal(idx, i0)
    np.testing.assert_array_equal(dist, d0)
    np.testing.assert_allclose(w.sum(-1), 1.0, rtol=0, atol=3e-7)
    r = 1.0 / (d0.astype(np.float64) + 1e-8)
    np.testing.assert_allclose(w, r / r.sum(-1, keepdims=True), rtol=2e-6)

    counts, ks = [5, 300, 0, 64], [5, 40, 0, 80]
    packed = xyz[0, :sum(counts)]
    got = oracle.furthest_point_sample_segments(packed, counts, ks)
    assert got.shape == (sum(ks),)
    np.testing.assert_array_equal(got[5:45], oracle.furthest_point_sample(packed[None, 5:305], 40, "A")[0])
    np.testing.assert_array_equal(got[45:], oracle.furthest_point_sample(packed[None, 305:369], 80, "A")[0])
