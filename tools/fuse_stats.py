"""Dev diagnostic: how k_fuse routes the points of the bench workload (needs a -DSMAP_FUSE_STATS build,
SMAP_LIB_PATH pointing at it): survivors of the float32 cull, points deferred to the float64 path, points
accepted by the float32 decisions."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_semantic_segmentation_b200 import synthetic as syn, _native
from vision_semantic_segmentation_b200.camera import camera_setup_1
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper
from vision_semantic_segmentation_b200.utils import transforms as tr

lib = _native.load()
for classes in (5, 19):
    labels, names, colors = syn.class_setup(classes == 19)
    dm = DeviceMapper(2000, 2000, colors, np.eye(len(labels)), [[100, 300], [800, 1000]], 0.1, 100.0, True,
                      names.index("lane"), cameras=[camera_setup_1()], device=0)
    out = (ctypes.c_ulonglong * 4)()
    lib.smap_debug_fuse_stats(out)
    n = 0
    for i in range(4):
        fr = syn.synthetic_frame(1000, i, 2000000, blocky=(i % 2 == 1), as_float64=False)
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        dm.integrate(dm.make_frame(torch.from_numpy(fr["points"]).cuda(), torch.from_numpy(fr["semantic_image"]).cuda(), T, 0))
        n += 2000000
    lib.smap_debug_fuse_stats(out)
    s, d, a = out[0], out[1], out[2]
    print("C=%d: points %d, cull survivors %d (%.2f %%), deferred %d (%.3f %% of survivors), float32-accepted %d (%.2f %%)"
          % (len(labels), n, s, 100.0 * s / n, d, 100.0 * d / max(s, 1), a, 100.0 * a / max(s, 1)))
    dm.close()
