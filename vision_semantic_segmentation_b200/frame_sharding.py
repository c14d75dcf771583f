"""Frame sharding across GPUs.

``update_map`` only ever ADDS frame-determined constants to the grid (``src/mapping_replay.py:281,294``),
so frames are independent units: rank r integrates the contiguous block ``[r*F/n, (r+1)*F/n)`` into
its own full-size grid and the grids are summed once at the end -- the only data-path collective.
Counts are integer-valued doubles (the sum is exact in any order); log-likelihood grids are summed
in a different order than the sequential reference (<= 1e-5 relative, in practice ~1e-15).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU tests) is
the transport.  A frame is never split across ranks: the per-frame (cell, class) de-duplication is
local to a frame.
"""

__all__ = ["rank_and_world", "shard_range", "sum_grids", "init_from_env"]


def rank_and_world():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_frames, rank, world):
    """Contiguous block of frame indices owned by ``rank``."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    return range((rank * n_frames) // world, ((rank + 1) * n_frames) // world)


def sum_grids(grid, group=None):
    """In-place all-reduce(sum) of the per-rank grids (a torch tensor, CUDA for NCCL)."""
    import torch.distributed as dist
    dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    return grid


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT) and bind this process to its GPU.  Returns (rank, world, local_rank)."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank
