"""world_size-2 gloo test of the scene-sharding path on CPU: each rank runs its own block of scenes (here through the
CPU oracle, since there is no GPU), no collective on the op path, one all_gather of the per-scene outputs at the end;
the gathered result must equal the single-process result in scene order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_rank, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import oracle
    from graspbalance_b200 import scenes, sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = sharding.scene_ids_for_rank(rank, world, per_rank)
        xyz = scenes.scene_batch(ids, 1500, "tabletop")
        inds = torch.from_numpy(oracle.furthest_point_sample(xyz, 64, "A")).to(torch.int64)
        gathered = sharding.gather_scene_outputs(inds, world)
        if rank == 0:
            ret.put(gathered.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_scenes_and_gather_in_scene_order():
    import oracle
    from graspbalance_b200 import scenes
    world, per_rank = 2, 3
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.furthest_point_sample(scenes.scene_batch(range(world * per_rank), 1500, "tabletop"), 64, "A")
    np.testing.assert_array_equal(got, want)
