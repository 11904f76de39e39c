"""Scene sharding across the GPUs of one node: every op is independent per scene, so ranks own
disjoint blocks of scenes and no collective sits on the op path; one all_gather of the small per-scene outputs ends a
step.  Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def scene_ids_for_rank(rank, world, scenes_per_rank, first=0):
    """Rank r owns scenes [first + r*S, first + (r+1)*S)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(first + rank * scenes_per_rank, first + (rank + 1) * scenes_per_rank))


def shard_batch(total_scenes, world):
    """Split `total_scenes` as evenly as possible: returns [(start, stop)] per rank (ranks < remainder get one more)."""
    base, rem = divmod(total_scenes, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def gather_scene_outputs(local, world=None, out=None):
    """all_gather equally-shaped per-scene outputs [S, ...] from every rank into [world*S, ...], rank-major (= scene id
    order under scene_ids_for_rank)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    local = local.contiguous()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out
