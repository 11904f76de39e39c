"""GPU test of the op pipeline (graspbalance_b200/pipeline.py): the multi-stream schedule (sampling chain and collision
tests on side streams) returns what the single-stream schedule returns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_overlapped_schedule_matches_single_stream(dev):
    import bench
    from graspbalance_b200 import pipeline
    B = 2
    host, offs = bench.make_host_inputs([3, 4], pin=False)
    outs = []
    for overlap in (False, True):
        pipe = pipeline.OpPipeline(B, bench.N_POINTS, dev, seed=0, backward=True, overlap=overlap)
        xyz, rot, grasps = bench.to_device(host, offs, dev)
        for _ in range(2):  # twice: the second pass reuses cached allocations across streams
            o = pipe.run(xyz, rot, grasps)
        torch.cuda.synchronize()
        outs.append({k: v.detach().cpu().numpy() for k, v in o.items()})
    a, b = outs
    np.testing.assert_array_equal(a["sa1_inds"], b["sa1_inds"])
    np.testing.assert_array_equal(a["seed_inds"], b["seed_inds"])
    np.testing.assert_array_equal(a["collision_counts"], b["collision_counts"])
    for k in ("up_checksum", "crop_checksum", "grad_checksum"):
        assert abs(float(a[k]) - float(b[k])) <= 1e-4 * max(1.0, abs(float(a[k]))), k


def test_fused_grasp_crops_match_the_sixteen_separate_calls(dev):
    import bench
    from graspbalance_b200 import pipeline
    host, offs = bench.make_host_inputs([5, 6], pin=False)
    sums = []
    for fused in (False, True):
        pipe = pipeline.OpPipeline(2, bench.N_POINTS, dev, seed=0, backward=False, overlap=False, fused_crops=fused, fused_sampling=fused)
        xyz, rot, grasps = bench.to_device(host, offs, dev)
        o = pipe.run(xyz, rot, None)
        torch.cuda.synchronize()
        sums.append((float(o["crop_checksum"]), o["sa1_inds"].cpu().numpy(), float(o["up_checksum"])))
    assert abs(sums[0][0] - sums[1][0]) <= 1e-4 * max(1.0, abs(sums[0][0]))
    np.testing.assert_array_equal(sums[0][1], sums[1][1])  # FPS + gather in one launch: same samples
    assert abs(sums[0][2] - sums[1][2]) <= 1e-4 * max(1.0, abs(sums[0][2]))


def test_whole_chain_matches_the_reference_extensions_tensor_by_tensor(dev, ref_a, ref_b):
    """BASELINE config 5's op chain (SA1-4, 15 InvResMLP groupings, FP1/FP2/up-sampling, 16 grasp crops; forward + backward)
    through OpPipeline and through the UNMODIFIED reference extensions with the reference's own torch glue
    (oracle/ref_chain.py), on the same scenes and the same stand-in features / gradients: every index tensor and forward
    tensor bit-equal (rotated crop coordinates to 1e-6: cuBLAS owns that sum), every gradient within 1e-5 relative."""
    import bench
    from graspbalance_b200 import pipeline
    from oracle import ref_chain
    B = 2
    host, offs = bench.make_host_inputs([11, 12], pin=False)
    xyz, rot, _ = bench.to_device(host, offs, dev)
    want = {}
    pipe = pipeline.OpPipeline(B, bench.N_POINTS, dev, seed=5, backward=True, overlap=False)
    ref_chain.run(pipe, xyz, rot, ref_a, ref_b, collect=want)
    for fused in (True, False):
        pipe.fused_crops = pipe.fused_sampling = fused
        got = {}
        pipe.run(xyz, rot, None, collect=got)
        torch.cuda.synchronize()
        assert set(want) - set(got) <= {k for k in want if k.startswith("crop") and k.endswith("_idx")}, sorted(set(want) - set(got))
        for k, w in want.items():
            if k not in got:
                continue
            g = got[k]
            assert g.shape == w.shape and g.dtype == w.dtype, (k, g.shape, w.shape, g.dtype, w.dtype)
            if k.endswith("_grad"):
                scale = max(float(w.abs().max()), 1e-30)
                assert float((g - w).abs().max()) <= 1e-5 * scale, (k, float((g - w).abs().max()), scale)
            elif k.startswith("crop"):
                assert float((g - w).abs().max()) <= 1e-6, (k, float((g - w).abs().max()))
            else:
                assert torch.equal(g, w), k
    # the crops' index lists, which the fused path does not return: the product's own cylinder_query against the reference's
    from graspbalance_b200 import pointnet2_utils as pu
    seed = want["sa1_xyz"]
    rot9 = rot.reshape(B, -1, 9).contiguous()
    for k, radius in enumerate(pipeline.CROP_RADII):
        for d, hmax in enumerate(pipeline.CROP_HMAX):
            assert torch.equal(pu.cylinder_query(radius, pipeline.CROP_HMIN, hmax, 64, xyz, seed, rot9), want[f"crop{k}_{d}_idx"])


def test_cross_step_sampling_prefetch_returns_the_same_results(dev):
    """Step k launching step k + 1's sampling chain on the side stream (OpPipeline.run(samples=, prefetch=)) changes the
    schedule only: seeds, collision counts and checksums equal the plain run's, step after step, with different scenes in
    consecutive steps."""
    import bench
    from graspbalance_b200 import pipeline
    B = 2
    batches = []
    for ids in ([21, 22], [23, 24], [25, 26]):
        host, offs = bench.make_host_inputs(ids, pin=False)
        batches.append(bench.to_device(host, offs, dev))
    pipe = pipeline.OpPipeline(B, bench.N_POINTS, dev, seed=1, backward=True, overlap=True)
    plain = []
    for inp in batches:
        o = pipe.run(*inp)
        torch.cuda.synchronize()
        plain.append({k: v.detach().cpu().numpy() for k, v in o.items()})
    bufs = [pipe.alloc_samples(), pipe.alloc_samples()]
    pipe.sampling_chain(batches[0][0], bufs[0])
    for k, inp in enumerate(batches):
        nxt = batches[(k + 1) % len(batches)][0]
        o = pipe.run(*inp, samples=bufs[k % 2], prefetch=(nxt, bufs[(k + 1) % 2]))
        torch.cuda.synchronize()
        got = {key: v.detach().cpu().numpy() for key, v in o.items()}
        for key in ("sa1_inds", "seed_inds", "collision_counts"):
            np.testing.assert_array_equal(got[key], plain[k][key])
        for key in ("up_checksum", "crop_checksum", "grad_checksum"):
            assert abs(float(got[key]) - float(plain[k][key])) <= 1e-4 * max(1.0, abs(float(plain[k][key]))), key
