"""Dev diagnostic (1 GPU, communicator of one rank): the phases of one streaming exchange with nothing to overlap."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_semantic_segmentation_b200 import _native, synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper
from vision_semantic_segmentation_b200.utils import transforms as tr

K = int(os.environ.get("K", "20"))
labels, names, colors = syn.class_setup(False)
dm = DeviceMapper(2000, 2000, colors, np.eye(5), [[100, 300], [800, 1000]], 0.1, 100.0, True, names.index("lane"),
                  cameras=[camera_setup_1()], device=0)
lib = _native.load()
ident = (ctypes.c_uint8 * 128)()
_native.check(lib.smap_comm_unique_id(ident))
_native.check(lib.smap_comm_init(dm._h, 1, 0, ident))
frames, keep = [], []
for i in range(16):
    fr = syn.synthetic_frame(1000, i, 2000000, blocky=(i % 2 == 1), as_float64=False)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    dp, di = torch.from_numpy(fr["points"]).cuda(), torch.from_numpy(fr["semantic_image"]).cuda()
    keep.append((dp, di))
    frames.append(dm.make_frame(dp, di, T, 0))
dm.set_streaming(True)
dm.clear()
for rep in range(4):
    dm.integrate_batch(frames[:K // 2]); dm.integrate_batch(frames[K // 2:K])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dm.exchange_async()
    t1 = time.perf_counter()
    dm.exchange_flush()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    info = dm.comm_info()
    print("alone: exchange_async %.0f us, flush (host) %.0f us, sync %.0f us; phases %s bytes %d window %s"
          % (1e6 * (t1 - t0), 1e6 * (t2 - t1), 1e6 * (t3 - t2), {k: round(v, 3) for k, v in info.items() if k.endswith("_ms")},
             info["bytes"], info["window"]))
