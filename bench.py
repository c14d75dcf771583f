#!/usr/bin/env python
"""Benchmark of the semantic-mapping hot path (BASELINE.json metric: points fused/s, frames/s,
fraction of the HBM roofline, next to the CPU path timed on the same box).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU algorithm (numpy port)

A "step" is one frame of BASELINE.json configs[1]: a 2M-point local cloud + one 1920x1440
19-class label image, count-based update into the default 2000x2000 BEV grid.  Inputs are a ring of
distinct frames resident in HBM (ring >> L2), so every step streams its cloud and image from DRAM.
Frames are handed to the C ABI 16 at a time (smap_integrate_batch: one fused kernel per frame; inside a batch a launch
takes half of the resident block slots and the launches alternate over four internal streams, so two frames run side by
side; the count update needs no second kernel).  `roofline.kernel_ms` is the same kernel launched alone with the full
grid (the library's profiling mode serialises the launches on the caller's stream).
N > 1: one process per GPU (torchrun), frames sharded by rank (weak scaling: every rank integrates K
frames), one NCCL all-reduce of the grids at the end of the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from vision_semantic_segmentation_b200 import synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.utils import transforms as tr  # noqa: E402

SEED = 1000  # cfg2
BOUNDARY, RESOLUTION, RANGE_MAX = [[100, 300], [800, 1000]], 0.1, 100.0
MAP_H = MAP_W = 2000


def select_workload(args):
    """cfg2 (default, the configuration the metric is quoted on): count update, 2000 x 2000 grid at 0.1 m.
    cfg3 (BASELINE.json configs[2]): confusion-matrix log-likelihood update -- the ordered two-kernel path -- into a
    0.2 m, 2 km x 2 km grid (10^4 x 10^4 cells; 4 GB of float64 at 5 classes)."""
    global BOUNDARY, RESOLUTION, MAP_H, MAP_W
    if args.workload == "cfg3":
        BOUNDARY, RESOLUTION, MAP_H, MAP_W = [[0, 2000], [0, 2000]], 0.2, 10000, 10000


def update_matrix(args, labels):
    if args.workload != "cfg3":
        return np.eye(len(labels))
    # src/mapping_replay.py:104-109 / src/data/confusion_matrix.py:43-48,59-63 on a synthetic strictly positive matrix
    sub = syn.synthetic_confusion_matrix(11)[np.ix_(labels, labels)]
    return np.log(sub / np.sum(sub, axis=1)[:, np.newaxis])


def claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on stdout at
    the first collective unless told otherwise, and NCCL_DEBUG_FILE above is only a default the environment may
    override): hand file descriptor 1 over to stderr for the rest of the process and keep a private duplicate of the
    real stdout for the line."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=2000000)
    ap.add_argument("--classes", type=int, default=5, choices=[5, 19],
                    help="mapped classes: the reference's default 5 (LABELS=[2,1,8,10,3]) or all 19")
    ap.add_argument("--ring", type=int, default=16, help="distinct frames resident in HBM")
    ap.add_argument("--batch", type=int, default=16, help="frames handed to smap_integrate_batch per call")
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2: BASELINE.json configs[1], the configuration the metric is quoted on (default); "
                         "cfg3: configs[2], log-likelihood update into a 10^4 x 10^4 grid")
    ap.add_argument("--label-format", default="rgb", choices=["rgb", "ids"],
                    help="rgb: the (1440, 1920, 3) colour-coded label image the reference consumes (default, the "
                         "BASELINE.json workload); ids: the network's (1440, 1920) uint8 class-id plane "
                         "(SMAP_IMG_CLASS_IDS) for the device-resident legs as well")
    ap.add_argument("--ordered", action="store_true",
                    help="dev switch: force the ordered update (cell masks + k_apply) although the matrix is np.eye(C)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every millisecond
    (the timed region is only a few milliseconds long; nvidia-smi -lms cannot sample that fast), nvidia-smi as a
    fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop, self._thread, self._nvml = threading.Event(), None, None

    def _poll(self):
        nv, h = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def sample_now(self):
        """One sample from the calling thread (called right after the timed loop has been queued, while the GPU is
        still executing it): guarantees a sample under load even for a very short timed region."""
        if self._nvml is not None:
            nv, h = self._nvml
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: go through the PCI bus id of the torch device
            try:
                import torch
                pr = torch.cuda.get_device_properties(self.index)
                h = nv.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nvml = (nv, h)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            nv = self._nvml[0]
            names = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
            reasons = sorted(n for n, bit in names if self.reason_bits & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def make_ring(args, rank):
    """ring of distinct synthetic frames (host): points (N,4) f32, image (H,W,3) u8, T (4,4) f64"""
    frames = []
    for i in range(args.ring):
        f = rank * 100000 + i
        fr = syn.synthetic_frame(SEED, f, args.points, blocky=(i % 2 == 1), as_float64=False, with_ids=True)
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        frames.append((fr["points"], fr["semantic_image"], T, fr["semantic_ids"]))
    return frames


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the numpy restatement of the reference's CPU path (oracle/numpy_port.py)
# --------------------------------------------------------------------------------------------------
def cpu_frame_fn(args):
    from oracle import numpy_port  # the ONLY place bench.py touches oracle/: the CPU baseline being timed
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    labels, names, colors = syn.class_setup(args.classes == 19)
    cam, cm = camera_setup_1(), update_matrix(args, labels)
    grid = np.zeros((MAP_H, MAP_W, len(labels)))

    def run(points, image, T, ids=None):
        pcd = np.ascontiguousarray(points.T.astype(np.float64))
        masked, label, _, _ = numpy_port.project_pcd(pcd, T, cam.P, image, RANGE_MAX)
        numpy_port.update_map(grid, masked, label, colors, cm, BOUNDARY, RESOLUTION, True, names)
        return points.shape[0]
    return run


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run = cpu_frame_fn(args)
    # bounded sample: whole frames when a frame costs <~1 s, otherwise a fixed slice of each frame
    ring = make_ring(argparse.Namespace(ring=min(args.ring, 4), points=args.points), 0)
    t0 = time.perf_counter()
    run(*ring[0])
    t_frame = time.perf_counter() - t0
    budget = 150.0
    frac = min(1.0, budget / max(t_frame * (args.steps + args.warmup), 1e-9))
    n_sub = max(1000, int(args.points * frac))
    sample = [(p[:n_sub], im, T) for p, im, T, _ in ring]
    for i in range(args.warmup):
        run(*sample[i % len(sample)])
    t0 = time.perf_counter()
    pts = 0
    for i in range(args.steps):
        pts += run(*sample[i % len(sample)])
    dt = time.perf_counter() - t0
    value = pts / dt
    cores = blas_threads()
    desc = ("numpy port of project_pcd+update_map, %d points of each %d-point frame per step, %d steps; "
            "BLAS threads=%d of %d host cores (only the two dgemms are threaded, as in the reference)"
            % (n_sub, args.points, args.steps, cores, os.cpu_count() or 1))
    line = {
        "impl": "reference", "metric": "points_fused_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_sec": value / args.points,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)


def workload_config(args, world):
    ids = getattr(args, "label_format", "rgb") == "ids"
    cfg3 = getattr(args, "workload", "cfg2") == "cfg3"
    return {"workload": "mapping_replay %s: %d-point cloud + 1920x1440 19-class label image per frame%s, "
                        "%s update, %d mapped classes, grid %dx%d @ %.1f m"
                        % ("cfg3" if cfg3 else "cfg2", args.points, " (as uint8 class-id plane)" if ids else "",
                           "confusion-matrix log-likelihood" if cfg3 else "count", args.classes, MAP_H, MAP_W, RESOLUTION),
            "points_per_frame": args.points, "image": [1440, 1920] if ids else [1440, 1920, 3], "mapped_classes": args.classes,
            "grid": [MAP_H, MAP_W, args.classes], "update": "log-likelihood" if cfg3 else "count", "frames_per_rank": args.steps,
            "parallelism": "frames sharded over %d rank(s), all-reduce(sum) of the grid at the end" % world,
            "l2_policy": "inputs larger than L2: ring of %d distinct frames (%.0f MB) resident in HBM"
                         % (args.ring, args.ring * (args.points * 16 + 1440 * 1920 * (1 if ids else 3)) / 1e6)}


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from vision_semantic_segmentation_b200 import frame_sharding, _native
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    from vision_semantic_segmentation_b200.device_mapper import DeviceMapper

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    rank, world, local_rank = frame_sharding.init_from_env()
    dev = torch.device("cuda", local_rank)
    labels, names, colors = syn.class_setup(args.classes == 19)
    c = len(labels)
    cam, cm, lane = camera_setup_1(), update_matrix(args, labels), names.index("lane")
    dm = DeviceMapper(MAP_H, MAP_W, colors, cm, BOUNDARY, RESOLUTION, RANGE_MAX, True, lane, cameras=[cam], device=local_rank)

    if args.ordered:   # smap_clear re-arms the count update: undo that after every clear
        plain_clear = dm.clear

        def ordered_clear():
            plain_clear()
            dm.notify_map_modified()
        dm.clear = ordered_clear
        dm.notify_map_modified()

    ring_host = make_ring(args, rank)
    ring_dev, ring_pinned = [], []
    if args.label_format == "ids":
        dm.set_label_palette(syn.COLORS_19)
    for pts, img, T, ids in ring_host:
        dp, di = torch.from_numpy(pts).to(dev), torch.from_numpy(ids if args.label_format == "ids" else img).to(dev)
        ring_dev.append((dm.make_frame(dp, di, T, 0), dp, di))
    torch.cuda.synchronize()

    # ---- algorithmic bytes per frame: N, M, U measured on the device path itself (outside timed regions)
    n_pts = args.points
    m_list, u_list, k_list = [], [], []
    for frame, dp, di in ring_dev[: min(4, len(ring_dev))]:
        if args.label_format == "ids":   # the parity API takes RGB images: count the survivors with one
            masked, _ = dm.project(dm.make_frame(dp, torch.from_numpy(ring_host[len(m_list)][1]).to(dev),
                                                 ring_host[len(m_list)][2], 0))
        else:
            masked, _ = dm.project(frame)
        m_list.append(masked.shape[1])
        dm.clear()
        dm.integrate(frame)
        u_list.append(int(torch.count_nonzero(dm.map).item()))
        k_list.append(int(torch.count_nonzero(dm.map.sum(dim=2)).item()))
    M, U, Kc = float(np.mean(m_list)), float(np.mean(u_list)), float(np.mean(k_list))
    label_bytes = 1.0 if args.label_format == "ids" else 3.0
    bytes_per_frame = 16.0 * n_pts + label_bytes * M + 2.0 * 8.0 * U
    dm.clear()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ring_frames = [f for f, _, _ in ring_dev]

    def device_loop(steps, reduce_at_end):
        # frames go to the C ABI in batches (smap_integrate_batch: up to 16 frames per kernel launch)
        done = 0
        while done < steps:
            take = min(args.batch, steps - done, len(ring_frames))
            start = done % len(ring_frames)
            chunk = (ring_frames + ring_frames)[start:start + take]
            dm.integrate_batch(chunk)
            done += take
        if reduce_at_end and world > 1:
            frame_sharding.sum_grids(dm.map)

    # ---- device-resident throughput (`value`)
    device_loop(args.warmup, False)
    launches_before = dm.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    device_loop(args.steps, True)
    e1.record()
    if rank == 0:
        sampler.sample_now()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = dm.stats()["kernel_launches"] - launches_before
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * args.steps * n_pts / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_fuse: project + cull + lookup + update of one frame).
    # Its launch duration is measured live with CUDA events recorded by the library on the launching stream
    # (smap_set_profiling: the per-frame launches are then serialised on that stream); the algorithmic bytes are
    # SURVEY.md 8d's  16 N + 3 M + 2*8 U  per frame with N, M, U measured above on the device path.
    dm.clear()
    device_loop(min(16, args.warmup + 3), False)
    torch.cuda.synchronize()
    dm.set_profiling(True)
    device_loop(args.steps, False)
    st = dm.stats()
    dm.set_profiling(False)
    kernel_ms = st["stream_kernel_ms"] / max(st["profiled_frames"], 1)
    apply_ms = st["apply_kernel_ms"] / max(st["profiled_frames"], 1)
    peak, peak_src = peaks()
    achieved = bytes_per_frame / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")   # from the committed ncu --set full capture
    if args.label_format == "rgb" and args.workload == "cfg2" and not args.ordered and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("k_fuse_c%d" % args.classes)

    # ---- end to end through the host-buffer entry point: H2D of cloud + image and a D2H read every step
    e2e = None
    if not args.no_e2e:
        dm.clear()
        dm.set_label_palette(syn.COLORS_19)
        ring_pinned_ids = []
        for pts, img, T, ids in ring_host[: min(4, len(ring_host))]:
            hp, hi = torch.from_numpy(pts).pin_memory(), torch.from_numpy(img).pin_memory()
            ring_pinned.append((dm.make_frame(hp, hi, T, 0, host=True), hp, hi))
            hd = torch.from_numpy(ids).pin_memory()
            ring_pinned_ids.append((dm.make_frame(hp, hd, T, 0, host=True), hp, hd))
        h2d = ring_pinned[0][1].numel() * 4 + ring_pinned[0][2].numel()
        e2e_steps = min(args.steps, 50)
        result = torch.zeros(1, dtype=torch.float64, device=dev)
        host_result = torch.zeros(1, dtype=torch.float64).pin_memory()

        def e2e_loop(steps, ring=ring_pinned):
            for i in range(steps):
                dm.integrate_host(ring[i % len(ring)][0])
                # the step's "metric": evidence mass in the grid cell under the vehicle's first hit (8 bytes)
                host_result.copy_(dm.map.view(-1)[:1], non_blocking=False)
        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_steps * n_pts / float(tt.item()), "unit": "points/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8, "steps": e2e_steps,
               "note": "smap_integrate_host: pinned host cloud (float32 x,y,z,i) + RGB label image copied per step"}
        # the same loop fed with the network's class-id plane instead of the painted RGB image (SMAP_IMG_CLASS_IDS:
        # 1 byte per pixel over PCIe, same grid bit for bit -- tests/test_gpu_label_ids.py)
        grid_rgb = dm.map.clone()
        dm.clear()
        e2e_loop(2, ring_pinned_ids)
        barrier()
        dm.clear()
        e2e_loop(2, ring_pinned)   # same 2 warm-up frames as the RGB loop above saw before its timed region
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, ring_pinned_ids)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["class_ids"] = {"value": world * e2e_steps * n_pts / float(tt.item()), "unit": "points/s",
                            "h2d_bytes_per_step": int(ring_pinned_ids[0][1].numel() * 4 + ring_pinned_ids[0][2].numel()),
                            "d2h_bytes_per_step": 8, "steps": e2e_steps,
                            "same_grid_as_rgb": bool(torch.equal(grid_rgb, dm.map)),
                            "note": "label image handed over as the network's (1440, 1920) uint8 class-id plane"}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        run = cpu_frame_fn(args)
        run(*ring_host[0])  # warm-up
        t0 = time.perf_counter()
        pts = 0
        for i in range(args.cpu_frames):
            pts += run(*ring_host[(1 + i) % len(ring_host)])
        dt = time.perf_counter() - t0
        cpu = {"value": pts / dt, "unit": "points/s", "cores": blas_threads(), "kind": "port",
               "sample": "%d full frames of the same workload through oracle/numpy_port.py (numpy restatement of the "
                         "reference's project_pcd+update_map); %d host cores, BLAS threads only in the two dgemms"
                         % (args.cpu_frames, os.cpu_count() or 1)}

    if rank == 0:
        line = {
            "metric": "points_fused_per_sec", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "frames_per_sec": value / n_pts,
            "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("k_fuse<masks> (project+cull+lookup+RED.OR into the frame's cell masks, one frame per "
                                    "launch); the ordered k_apply replay is apply_kernel_ms_per_frame"
                                    if (args.workload == "cfg3" or args.ordered) else
                                    "k_fuse<count> (register-prefetched cloud, float32 certified project+cull+lookup+update, "
                                    "one frame per launch)"),
                         "kernel_ms": kernel_ms, "apply_kernel_ms_per_frame": apply_ms,
                         "step_frac": bytes_per_frame / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_frame": bytes_per_frame,
                         "N": n_pts, "M": M, "K_cells": Kc, "U_elements": U},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    args.out = claim_stdout()
    select_workload(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
