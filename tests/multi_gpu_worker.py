"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU): frame-sharded mapping_replay over
NCCL must reproduce the single-process golden result bit for bit (count grids are integer-valued)."""
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
color_map = sm.mapping_replay(frames, "sharded", write_image=(rank == 0))
assert np.array_equal(color_map, case.arrays["rgb"]), "rank %d: rendered map differs" % rank
filtered_full = sm.map
assert sha(filtered_full) == case.spec["filtered_sha"], "rank %d: filtered grid differs" % rank
# large-map variant: reduce-scatter by rows + halo exchange, every rank renders its own tile, image all-gathered
color_map = sm.mapping_replay(frames, "sharded_tiles", write_image=False, row_tiles=True)
assert np.array_equal(color_map, case.arrays["rgb"]), "rank %d: row-tiled rendered map differs" % rank
r0, r1 = frame_sharding.row_tile(case.mh, rank, world)
want = np.zeros_like(filtered_full)
want[r0:r1] = filtered_full[r0:r1]
assert np.array_equal(sm.map, want), "rank %d: filtered tile differs" % rank
dist.barrier()
print("rank %d/%d ok" % (rank, world))
dist.destroy_process_group()
