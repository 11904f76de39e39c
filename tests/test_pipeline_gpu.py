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
