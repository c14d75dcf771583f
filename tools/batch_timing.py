"""Dev diagnostic: GPU time (CUDA events) and host time of one smap_integrate_batch call vs batch size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_semantic_segmentation_b200 import synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper
from vision_semantic_segmentation_b200.utils import transforms as tr

classes = int(os.environ.get("C", "5"))
labels, names, colors = syn.class_setup(classes == 19)
cam = camera_setup_1()
dm = DeviceMapper(2000, 2000, colors, np.eye(len(labels)), [[100, 300], [800, 1000]], 0.1, 100.0, True,
                  names.index("lane"), cameras=[cam], device=0)
frames, keep = [], []
for i in range(16):
    fr = syn.synthetic_frame(1000, i, 2000000, blocky=(i % 2 == 1), as_float64=False)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    dp, di = torch.from_numpy(fr["points"]).cuda(), torch.from_numpy(fr["semantic_image"]).cuda()
    keep.append((dp, di))
    frames.append(dm.make_frame(dp, di, T, 0))
same = [frames[0]] * 16
for name, fl in ((("ring", frames), ("same-frame", same)) if not os.environ.get("PROFILE") else ()):
    for b in (1, 2, 4, 8, 16):
        for _ in range(3):
            dm.integrate_batch(fl[:b])
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            dm.integrate_batch(fl[:b])
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print("%-10s batch %2d: gpu %.2f us/frame   host-enqueue %.2f us/frame" % (name, b, 1e3 * e0.elapsed_time(e1) / reps / b, 1e6 * t_host / reps / b))

if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            dm.integrate_batch(frames[:16])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
