"""Frame sharding across GPUs.

``update_map`` only ever ADDS frame-determined constants to the grid (``src/mapping_replay.py:281,294``),
so frames are independent units: rank r integrates the contiguous block ``[r*F/n, (r+1)*F/n)`` into
its own full-size grid and the grids are summed once at the end -- the only data-path collective.
Counts are integer-valued doubles (the sum is exact in any order); log-likelihood grids are summed
in a different order than the sequential reference (<= 1e-5 relative, in practice ~1e-15).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU tests) is
the transport.  A frame is never split across ranks: the per-frame (cell, class) de-duplication is
local to a frame.

Large maps (a 10^4 x 10^4 x 19 grid is 15 GB): instead of all-reducing the whole grid and rendering it on
every rank, ``sum_grid_row_tile`` reduce-scatters it by rows -- rank r ends up with the summed rows of its own
tile plus a one-row halo from each neighbour -- ``render_row_tile`` filters and renders that tile
(BORDER_REFLECT_101 applies at the true map edges only; the halo rows supply the 3x3 box across tile seams) and
``gather_rgb_rows`` assembles the image.  Half the collective traffic, 1/n of the render work per rank.
"""

__all__ = ["rank_and_world", "shard_range", "sum_grids", "init_from_env", "row_tile", "sum_grid_row_tile",
           "render_row_tile", "gather_rgb_rows", "backend_is_nccl"]


def backend_is_nccl(group=None):
    """True when torch.distributed runs over NCCL: the grids are then summed by the library's own exchange under the
    C ABI (``smap_exchange_async`` / ``smap_allreduce``: touched window only, counts packed); other backends (gloo in the
    CPU tests) go through ``sum_grids`` / ``sum_grid_row_tile`` below."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and "nccl" in str(dist.get_backend(group))


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


def row_tile(n_rows, rank, world):
    """Rows [r0, r1) of the map owned by ``rank``: equal tiles of ceil(n_rows / world) rows (the last ones may be
    short or empty), which is the layout reduce-scatter produces."""
    per = -(-n_rows // world)
    return min(rank * per, n_rows), min((rank + 1) * per, n_rows)


def sum_grid_row_tile(grid, group=None):
    """Reduce-scatter(sum) of the per-rank grids (MH, MW, C) by rows.  Returns ``(tile, r0, r1, top, bottom)``:
    ``tile`` holds the summed rows ``[r0 - top, r1 + bottom)`` where top / bottom (0 or 1) say whether a halo row
    from the neighbouring tile is present.  Single process: the grid itself."""
    import torch
    import torch.distributed as dist
    rank, world = rank_and_world()
    mh = grid.shape[0]
    if world == 1:
        return grid, 0, mh, 0, 0
    per = -(-mh // world)
    r0, r1 = row_tile(mh, rank, world)
    if dist.get_backend(group) == "nccl":
        src = grid
        if per * world != mh:   # pad with zero rows so that every rank gets an equal chunk
            src = torch.zeros((per * world,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
            src[:mh] = grid
        own = torch.empty((per,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
        dist.reduce_scatter_tensor(own, src.contiguous(), op=dist.ReduceOp.SUM, group=group)
    else:                       # gloo has no reduce-scatter: all-reduce a copy, keep the own rows
        full = grid.clone()
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        own = torch.zeros((per,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
        own[:r1 - r0] = full[r0:r1]
    # halo exchange: everybody publishes its first and last valid row (tiny), neighbours pick what they need
    edge = torch.zeros((2,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
    if r1 > r0:
        edge[0] = own[0]
        edge[1] = own[r1 - r0 - 1]
    edges = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(edges, edge, group=group)
    top = 1 if r0 > 0 and r1 > r0 else 0
    bottom = 1 if r1 < mh and r1 > r0 else 0
    parts = []
    if top:
        parts.append(edges[(r0 - 1) // per][1:2])      # last row of the tile above
    parts.append(own[:r1 - r0])
    if bottom:
        parts.append(edges[r1 // per][0:1])            # first row of the tile below
    return torch.cat(parts, dim=0), r0, r1, top, bottom


def render_row_tile(tile, top, bottom, label_colors, return_filtered=False):
    """apply_filter + render_bev_map of a row tile that carries ``top`` / ``bottom`` halo rows; the halo rows are
    dropped from the result.  Rows next to a halo see their real neighbours, rows at a true map edge are
    reflected (BORDER_REFLECT_101) exactly as in the whole-map render."""
    from .renderer import filter_and_render
    if tile.shape[0] == 0:
        import torch
        rgb = torch.empty((0, tile.shape[1], 3), dtype=torch.uint8, device=tile.device)
        return (rgb, tile) if return_filtered else rgb
    out = filter_and_render(tile.contiguous(), label_colors, return_filtered=return_filtered)
    rgb, filtered = out if return_filtered else (out, None)
    end = tile.shape[0] - bottom
    rgb = rgb[top:end]
    return (rgb, filtered[top:end]) if return_filtered else rgb


def gather_rgb_rows(rgb_tile, n_rows, group=None):
    """All-gather the rendered row tiles into the (MH, MW, 3) image (on every rank)."""
    import torch
    import torch.distributed as dist
    rank, world = rank_and_world()
    if world == 1:
        return rgb_tile
    per = -(-n_rows // world)
    mine = torch.zeros((per,) + tuple(rgb_tile.shape[1:]), dtype=rgb_tile.dtype, device=rgb_tile.device)
    mine[:rgb_tile.shape[0]] = rgb_tile
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.cat(parts, dim=0)[:n_rows]


def bind_to_gpu_cpus(local_rank):
    """One process per GPU launches ~10^5 kernels per second: keep its threads on the CPU cores next to its GPU (NVML's
    ideal affinity: the NUMA node the GPU's PCIe root hangs off).  On the 8-GPU box the ranks whose GPUs sit on the other
    socket queued frames 10 - 40 % slower than the rest (profiles/r2w_*).  Best effort: silently a no-op when NVML or
    the affinity call is not available.  Returns the CPU set, or None."""
    import os
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        h = nv.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = cpus & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return cpus or None
    except Exception:
        return None


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
        if world > 1:
            bind_to_gpu_cpus(local_rank)
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
