#!/usr/bin/env python
"""Golden vectors from the UNMODIFIED reference CUDA extensions (oracle/_ref/*.so), generated on a B200:

    gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu_small.npz     (then copied to tests/golden/)

Small seeded inputs are stored beside the reference outputs so the CPU-only suite can pin oracle/gb_oracle.c against
what the reference kernels really produce (tests/test_oracle_cpu.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import _load_ref  # noqa: E402
from graspbalance_b200 import scenes  # noqa: E402


def main(out_path):
    dev = torch.device("cuda:0")
    rA, rB, rC = _load_ref("gbref_pointnet2_ext"), _load_ref("gbref_pointnet2_batch"), _load_ref("gbref_knn")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    rng = np.random.default_rng(123)
    out = {}
    # FPS: duplicates + near-origin points, both variants, n not a power of two
    xyz = np.stack([scenes.tabletop_scene(s, 3000, n_dup=300) for s in (1, 2)])
    xyz[0, [0, 17, 900]] = rng.uniform(-0.01, 0.01, (3, 3)).astype(np.float32)
    out["fps_xyz"] = xyz
    out["fps_a"] = rA.furthest_point_sampling(T(xyz), 400).cpu().numpy()
    temp = torch.full((2, 3000), 1e10, device=dev)
    idx = torch.empty((2, 400), dtype=torch.int32, device=dev)
    rB.furthest_point_sampling_wrapper(2, 3000, 400, T(xyz), temp, idx)
    out["fps_b"] = idx.cpu().numpy()
    small = rng.uniform(-1, 1, (1, 40, 3)).astype(np.float32)
    small[0, 20:] = small[0, :20]
    out["fps_small_xyz"] = small
    out["fps_small_a"] = rA.furthest_point_sampling(T(small), 60).cpu().numpy()
    # ball / cylinder
    q = xyz[:, ::47][:, :64].copy()
    out["q_xyz"] = q
    out["ball_a"] = rA.ball_query(T(q), T(xyz), 0.05, 16).cpu().numpy()
    ib = torch.zeros((2, 64, 16), dtype=torch.int32, device=dev)
    rB.ball_query_wrapper(2, 3000, 64, 0.05, 16, T(q), T(xyz), ib)
    out["ball_b"] = ib.cpu().numpy()
    rot = scenes.viewpoint_rotations(-rng.normal(size=(2, 64, 3)).astype(np.float32),
                                     rng.uniform(0, np.pi, (2, 64)).astype(np.float32)).reshape(2, 64, 9)
    out["cyl_rot"] = rot
    out["cyl_a"] = rA.cylinder_query(T(q), T(xyz), T(rot), 0.05, -0.02, 0.04, 16).cpu().numpy()
    # three_nn / interpolate
    d2, i3 = rA.three_nn(T(xyz[:, :500]), T(q))
    out["nn_d2"], out["nn_idx"] = d2.cpu().numpy(), i3.cpu().numpy()
    w = rng.uniform(0.1, 1, (2, 500, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    f = rng.normal(size=(2, 4, 64)).astype(np.float32)
    out["interp_w"], out["interp_f"] = w, f
    out["interp_out"] = rA.three_interpolate(T(f), i3, T(w)).cpu().numpy()
    go = rng.normal(size=(2, 4, 500)).astype(np.float32)
    out["interp_go"] = go
    out["interp_grad"] = rA.three_interpolate_grad(T(go), i3, T(w), 64).cpu().numpy()
    # group grad (atomics: tolerance compare)
    gi = rng.integers(0, 3000, (2, 32, 8)).astype(np.int32)
    gf = rng.normal(size=(2, 5, 3000)).astype(np.float32)
    gg = rng.normal(size=(2, 5, 32, 8)).astype(np.float32)
    out["group_idx"], out["group_f"], out["group_go"] = gi, gf, gg
    out["group_out"] = rA.group_points(T(gf), T(gi)).cpu().numpy()
    out["group_grad"] = rA.group_points_grad(T(gg), T(gi), 3000).cpu().numpy()
    # knn (CUDA path), k=1 and k=8, with duplicate references
    ref = rng.uniform(-1, 1, (2, 3, 500)).astype(np.float32)
    ref[:, :, 250:260] = ref[:, :, :10]
    qry = rng.uniform(-1, 1, (2, 3, 50)).astype(np.float32)
    out["knn_ref"], out["knn_query"] = ref, qry
    for k in (1, 8):
        o = torch.empty((2, k, 50), dtype=torch.int64, device=dev)
        rC.knn(T(ref), T(qry), o)
        out[f"knn_k{k}"] = o.cpu().numpy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_gpu_small.npz"))
