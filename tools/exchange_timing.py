"""Dev diagnostic (torchrun, >= 2 GPUs): where a block of K frames + one streaming exchange spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vision_semantic_segmentation_b200 import frame_sharding, synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper
from vision_semantic_segmentation_b200.utils import transforms as tr

rank, world, local = frame_sharding.init_from_env()
K = int(os.environ.get("K", "20"))
labels, names, colors = syn.class_setup(False)
dm = DeviceMapper(2000, 2000, colors, np.eye(5), [[100, 300], [800, 1000]], 0.1, 100.0, True, names.index("lane"),
                  cameras=[camera_setup_1()], device=local)
dm.init_comm()
frames, keep = [], []
for i in range(16):
    fr = syn.synthetic_frame(1000, rank * 100000 + i, 2000000, blocky=(i % 2 == 1), as_float64=False)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    dp, di = torch.from_numpy(fr["points"]).cuda(), torch.from_numpy(fr["semantic_image"]).cuda()
    keep.append((dp, di))
    frames.append(dm.make_frame(dp, di, T, 0))

def block():
    half = K // 2
    dm.integrate_batch((frames + frames)[:half]); dm.integrate_batch((frames + frames)[half:K])

def run(label, R, exchange, streaming):
    dm.set_streaming(streaming)
    dm.clear()
    for _ in range(3):
        block()
        if exchange: dm.exchange_async()
    if exchange: dm.exchange_flush()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_enq, host_x = 0.0, 0.0
    t_start = time.perf_counter()
    e0.record()
    for _ in range(R):
        t0 = time.perf_counter(); block(); t1 = time.perf_counter()
        if exchange: dm.exchange_async()
        t2 = time.perf_counter()
        host_enq += t1 - t0; host_x += t2 - t1
    if exchange: dm.exchange_flush()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_start
    info = dm.comm_info() if exchange else {}
    if rank == 0:
        print("%-34s gpu %.1f us/block  host enqueue %.1f us/block  host in exchange_async %.1f us  wall %.1f us/block  %s"
              % (label, 1e3 * e0.elapsed_time(e1) / R, 1e6 * host_enq / R, 1e6 * host_x / R, 1e6 * wall / R,
                 {k: round(v, 3) for k, v in info.items() if k.endswith("_ms")}), flush=True)
    dist.barrier()

run("no exchange, grid", 100, False, False)
run("no exchange, increments buffer", 100, False, True)
run("streaming exchange per block", 100, True, True)
os.environ["K2"] = "1"
K = 40
run("streaming exchange per 40 frames", 50, True, True)
K = 80
run("streaming exchange per 80 frames", 25, True, True)
dist.barrier()
dist.destroy_process_group()
