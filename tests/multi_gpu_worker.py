"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU): frame-sharded mapping over NCCL must
reproduce the single-process golden result -- bit for bit for count grids (integer-valued), to 1e-12 for log-likelihood
grids (the ranks' partial sums are added in a different order than the sequential reference)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests.common import Case, sha  # noqa: E402
from vision_semantic_segmentation_b200 import frame_sharding, synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.config.base_cfg import get_cfg_defaults  # noqa: E402
from vision_semantic_segmentation_b200.mapping_replay import SemanticMapping  # noqa: E402

rank, world, local_rank = frame_sharding.init_from_env()
case = Case("cfg1_c5_count")
cfg = get_cfg_defaults()
cfg.OUTPUT_DIR = tempfile.mkdtemp()
sm = SemanticMapping(cfg, device=local_rank)
frames = [syn.synthetic_frame(case.spec["seed"], f, case.spec["n_points"], blocky=(f in case.spec["blocky_frames"]))
          for f in range(case.spec["frames"])]

# ---- the API: streaming exchange under the C ABI (smap_exchange_async), one exchange for the whole replay
color_map = sm.mapping_replay(frames, "sharded", write_image=(rank == 0))
assert np.array_equal(color_map, case.arrays["rgb"]), "rank %d: rendered map differs" % rank
filtered_full = sm.map
assert sha(filtered_full) == case.spec["filtered_sha"], "rank %d: filtered grid differs" % rank
info = sm.device_mapper.comm_info()
assert info["n_ranks"] == world and info["pack"] == "u16" and 0 < info["bytes"] < info["grid_bytes"] // 8, info

# ---- an exchange after every frame, frames fed one at a time: the double-buffered increments, several exchanges in flight
sm.EXCHANGE_EVERY, sm.FEED_BATCH, sm._feeder = 1, 1, None
before = sm.device_mapper.comm_info()["exchanges"]
color_map = sm.mapping_replay(frames, "sharded_stream", write_image=False)
assert np.array_equal(color_map, case.arrays["rgb"]), "rank %d: rendered map differs (exchange per frame)" % rank
assert sha(sm.map) == case.spec["filtered_sha"]
assert sm.device_mapper.comm_info()["exchanges"] - before == -(-len(frames) // world)

# ---- large-map variant: every rank filters + renders its own row tile (one-row halos), image all-gathered
color_map = sm.mapping_replay(frames, "sharded_tiles", write_image=False, row_tiles=True)
assert np.array_equal(color_map, case.arrays["rgb"]), "rank %d: row-tiled rendered map differs" % rank
r0, r1 = frame_sharding.row_tile(case.mh, rank, world)
want = np.zeros_like(filtered_full)
want[r0:r1] = filtered_full[r0:r1]
assert np.array_equal(sm.map, want), "rank %d: filtered tile differs" % rank

# ---- the one-shot collectives of the C ABI on per-rank grids: smap_allreduce, smap_reduce_scatter_rows
dm = sm.device_mapper
golden = case.sparse_map()
shard = frame_sharding.shard_range(len(frames), rank, world)


def integrate_shard():
    dm.clear()
    sm.integrate_frames(frames[i] for i in shard)


integrate_shard()
dm.allreduce()
assert np.array_equal(dm.map.cpu().numpy(), golden), "rank %d: smap_allreduce differs from the sequential grid" % rank
info = dm.comm_info()
x0, x1, y0, y1 = info["window"]
nz = np.nonzero(golden.sum(2))
assert x0 <= nz[0].min() and x1 >= nz[0].max() and y0 <= nz[1].min() and y1 >= nz[1].max(), info
assert info["pack"] == "u16" and info["bytes"] == (x1 - x0 + 1) * (((y1 - y0 + 1) * case.c + 1) // 2) * 4, info
integrate_shard()
tile, t0, t1, top, bottom = dm.reduce_scatter_rows()
assert (t0, t1) == (r0, r1)
assert np.array_equal(tile.cpu().numpy(), golden[r0 - top:r1 + bottom]), "rank %d: smap_reduce_scatter_rows differs" % rank
# a grid written from outside (non-integer values): the window is the whole grid and it travels as float64
integrate_shard()
dm.map.mul_(0.5)
dm.notify_map_modified()
dm.allreduce()
assert dm.comm_info()["pack"] == "f64" and dm.comm_info()["window"] == [0, case.mh - 1, 0, case.mw - 1]
assert np.array_equal(dm.map.cpu().numpy(), 0.5 * golden)

# ---- log-likelihood update: float64 exchange, summation order differs from the sequential reference
case_log = Case("cfg1_c5_log")
cfg = get_cfg_defaults()
cfg.OUTPUT_DIR = tempfile.mkdtemp()
path = os.path.join(cfg.OUTPUT_DIR, "cm.npy")
np.save(path, syn.synthetic_confusion_matrix(case_log.spec.get("cm_seed", 7)))
cfg.MAPPING.CONFUSION_MTX.LOAD_PATH = path
sm_log = SemanticMapping(cfg, device=local_rank)
frames_log = [syn.synthetic_frame(case_log.spec["seed"], f, case_log.spec["n_points"], blocky=(f in case_log.spec["blocky_frames"]))
              for f in range(case_log.spec["frames"])]
sm_log.EXCHANGE_EVERY, sm_log.FEED_BATCH = 1, 1
color_map = sm_log.mapping_replay(frames_log, "sharded_log", write_image=False)
assert sm_log.device_mapper.comm_info()["pack"] == "f64"
diff = color_map != case_log.arrays["rgb"]
assert diff.sum() == 0, "rank %d: %d rendered pixels differ in the log-likelihood case" % (rank, int(diff.any(axis=2).sum()))
dist.barrier()
print("rank %d/%d ok" % (rank, world))
dist.destroy_process_group()
