#!/usr/bin/env python
"""Benchmark of the semantic-mapping hot path (BASELINE.json metric: points fused/s, frames/s,
fraction of the HBM roofline, next to the CPU path timed on the same box).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's own CPU code (oracle/_ref)

A "step" is one frame of the workload (default cfg2 = BASELINE.json configs[1]: a 2M-point local cloud + one 1920x1440
19-class label image, count update into the default 2000x2000 BEV grid).  Inputs are a ring of distinct frames resident
in HBM (ring >> L2), so every step streams its cloud and image from DRAM.  Frames are handed to the C ABI in batches
(smap_integrate_batch: one fused kernel per frame; inside a batch a launch takes half of the resident block slots and
the launches alternate over internal streams, so two frames run side by side).

Timed region: R blocks of exactly K steps, back to back, R chosen so that the region lasts >= 50 ms on one GPU and
>= 300 ms on several (a 20-step block is 0.3 ms); barrier + synchronize on both sides, CUDA events, max over ranks.  `ms_per_step` = region / (R K); the per-block
figures (median / min / max) are printed next to it.
N > 1: one process per GPU (torchrun), frames sharded by rank (weak scaling: every rank integrates K frames per block).
Every --exchange-every (256) frames the ranks' increments are summed (smap_exchange_async: touched window only, counts
packed as uint16 pairs, NCCL all-reduce on an internal stream, overlapped with the next frames); the region ends when the
last exchange -- behind the last frame -- has been added to every rank's grid.  After the timed region the exchanged grid is compared with a plain
torch.distributed all-reduce of the per-rank grids (bit for bit in count mode).
Workloads: cfg2 (default), cfg3 (log-likelihood update, 10^4 x 10^4 grid), cfg4 (8000 frames in total, sharded: strong
scaling), cfg5 (cam1 + cam6 frames into a 10^4 x 10^4 grid, row-tiled filter + render + image gather inside the region).
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
    0.2 m, 2 km x 2 km grid (10^4 x 10^4 cells; 4 GB of float64 at 5 classes).
    cfg4 (configs[3]): cfg2's frames, --frames (8000) of them in total sharded over the ranks (strong scaling).
    cfg5 (configs[4]): frames alternate between the cam1 and cam6 calibrations, count update into the 10^4 x 10^4 grid,
    the map filtered + rendered by row tiles and the image gathered at the end of the region."""
    global BOUNDARY, RESOLUTION, MAP_H, MAP_W
    if args.workload in ("cfg3", "cfg5"):
        BOUNDARY, RESOLUTION, MAP_H, MAP_W = [[0, 2000], [0, 2000]], 0.2, 10000, 10000
    if args.workload == "cfg4":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.steps = max(1, args.frames // world)


def log_update(args):
    return args.workload == "cfg3"


def update_matrix(args, labels):
    if not log_update(args):
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
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=2000000)
    ap.add_argument("--classes", type=int, default=5, choices=[5, 19],
                    help="mapped classes: the reference's default 5 (LABELS=[2,1,8,10,3]) or all 19")
    ap.add_argument("--ring", type=int, default=16, help="distinct frames resident in HBM")
    ap.add_argument("--batch", type=int, default=32, help="frames handed to smap_integrate_batch per call (at most)")
    ap.add_argument("--repeats", type=int, default=0,
                    help="blocks of --steps steps in the timed region (0: as many as make the region last >= 50 ms)")
    ap.add_argument("--min-region-ms", type=float, default=0.0,
                    help="length the timed region is stretched to by repeating the block of --steps steps (0: 50 ms on "
                         "one GPU, 300 ms on several -- one late rank stalls every rank at the next exchange, and such "
                         "hiccups of 1 - 30 ms were seen once per run on the 8-GPU boxes whatever the launch path)")
    ap.add_argument("--exchange-every", type=int, default=256,
                    help="N > 1: frames a rank integrates between two exchanges of the streaming sum (as "
                         "SemanticMapping.EXCHANGE_EVERY); the last exchange ends the timed region")
    ap.add_argument("--frames", type=int, default=8000, help="cfg4: frames of the whole job, sharded over the ranks")
    ap.add_argument("--cpu-frames", type=int, default=6, help="frames of the cpu_baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=48, help="frames of the end-to-end legs")
    ap.add_argument("--clock-period-ms", type=float, default=1.0,
                    help="period of the NVML clock / throttle-reason polling thread during the timed region (0: only "
                         "the two direct samples taken while the region is queued)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg2: BASELINE.json configs[1], the configuration the metric is quoted on (default); "
                         "cfg3: configs[2], log-likelihood update into a 10^4 x 10^4 grid; cfg4: configs[3], 8000 frames "
                         "sharded (strong scaling); cfg5: configs[4], two cameras, large map, row-tiled render")
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

    def __init__(self, index, period_s=0.001):
        self.index, self.lines, self.proc, self.period_s = index, [], None, period_s
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
            time.sleep(self.period_s)

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

    def start(self, thread=True):
        """thread=False (ranks > 0): no polling thread, only the samples sample_now() takes"""
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
            if thread:
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
            if self._thread is not None:
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


def make_ring(args, rank, n=None):
    """ring of distinct synthetic frames (host): dicts with points (N,4) f32, semantic_image (H,W,3) u8, semantic_ids,
    pose, T (4,4) f64 world -> velodyne, camera slot (cfg5: cam1 / cam6 alternate)"""
    frames = []
    for i in range(args.ring if n is None else n):
        # every rank drives the SAME stretch of road (poses of frames 0 .. ring - 1: inside the map) with its own clouds and
        # label images, so the ranks do equal work.  (Up to round-2 commit 7b01586 rank r used frame indices 10^5 r + i,
        # i.e. poses hundreds of kilometres off the map: ranks 1 - 3 had nothing to update and ran 13 % faster than one
        # GPU alone, ranks 4 - 7 had coordinates beyond 2^20 m -- the exact-arithmetic path -- and ran 25 % slower.)
        fr = syn.synthetic_frame(SEED + 7919 * rank, i, args.points, blocky=(i % 2 == 1), as_float64=False, with_ids=True)
        fr["T"] = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        fr["camera"] = (i % 2) if args.workload == "cfg5" else 0
        fr["camera_id"] = 6 if fr["camera"] == 1 else 1
        frames.append(fr)
    return frames


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's OWN code (oracle/_ref: byte-compiled from /root/reference by
# oracle/build_ref.py, imported through oracle/ref_shim.py), the numpy restatement only if that tree is missing
# --------------------------------------------------------------------------------------------------
def make_cfg(cfg, args, tmp):
    labels, names, colors = syn.class_setup(args.classes == 19)
    cfg.OUTPUT_DIR = tmp
    cfg.LABELS, cfg.LABELS_NAMES, cfg.LABEL_COLORS = labels, names, colors
    cfg.MAPPING.BOUNDARY, cfg.MAPPING.RESOLUTION = BOUNDARY, RESOLUTION
    cfg.MAPPING.PCD.RANGE_MAX, cfg.MAPPING.PCD.USE_INTENSITY = RANGE_MAX, True
    if log_update(args):
        path = os.path.join(tmp, "cm.npy")
        np.save(path, syn.synthetic_confusion_matrix(11))
        cfg.MAPPING.CONFUSION_MTX.LOAD_PATH = path
    return cfg


def cpu_frame_fn(args):
    """-> (run(frame) -> points, render() -> seconds, kind): one frame through the CPU path -- project_pcd + update_map."""
    import tempfile
    from oracle import ref_shim   # bench.py touches oracle/ only here: the CPU baseline being timed
    labels, names, colors = syn.class_setup(args.classes == 19)
    if ref_shim.reference_available():
        ref = ref_shim.load_reference()
        sm = ref.SemanticMapping(make_cfg(ref.get_cfg_defaults(), args, tempfile.mkdtemp()))
        grid = np.zeros((sm.map_height, sm.map_width, sm.map_depth))
        cams = [sm.cam1, sm.cam6]

        def run(fr, n_sub=None):
            pcd = np.ascontiguousarray(fr["points"][:n_sub].T.astype(np.float64))   # the reference's frame format
            masked, label = sm.project_pcd(pcd, "world", fr["semantic_image"], fr["pose"], cams[fr["camera"]])
            sm.update_map(grid, masked, label)
            return pcd.shape[1]

        def render(crop=None):
            g = grid if crop is None else np.ascontiguousarray(grid[:crop, :crop])
            t0 = time.perf_counter()
            ref.render_bev_map(ref.apply_filter(g), sm.label_colors)
            return time.perf_counter() - t0
        return run, render, "reference"
    from oracle import numpy_port
    from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
    cams, cm = [camera_setup_1(), camera_setup_6()], update_matrix(args, labels)
    grid = np.zeros((MAP_H, MAP_W, len(labels)))

    def run(fr, n_sub=None):
        pcd = np.ascontiguousarray(fr["points"][:n_sub].T.astype(np.float64))
        masked, label, _, _ = numpy_port.project_pcd(pcd, fr["T"], cams[fr["camera"]].P, fr["semantic_image"], RANGE_MAX)
        numpy_port.update_map(grid, masked, label, colors, cm, BOUNDARY, RESOLUTION, True, names)
        return pcd.shape[1]
    return run, None, "port"


def set_blas_threads(n):
    """Pin the BLAS pool explicitly: torchrun exports OMP_NUM_THREADS=1, which would otherwise turn the N > 1 baseline
    into a single-threaded one."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n, user_api="blas")
    except Exception:
        return None


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def time_cpu(run, frames, n_frames, n_sub=None):
    run(frames[0], n_sub)  # warm-up
    t0 = time.perf_counter()
    pts = 0
    for i in range(n_frames):
        pts += run(frames[(1 + i) % len(frames)], n_sub)
    return pts / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    set_blas_threads(os.cpu_count() or 1)
    run, _, kind = cpu_frame_fn(args)
    # bounded sample: whole frames when a frame costs <~1 s, otherwise a fixed slice of each frame
    ring = make_ring(args, 0, n=min(args.ring, 4))
    t0 = time.perf_counter()
    run(ring[0])
    t_frame = time.perf_counter() - t0
    budget = 150.0
    frac = min(1.0, budget / max(t_frame * (args.steps + args.warmup), 1e-9))
    n_sub = max(1000, int(args.points * frac))
    for i in range(args.warmup):
        run(ring[i % len(ring)], n_sub)
    t0 = time.perf_counter()
    pts = 0
    for i in range(args.steps):
        pts += run(ring[i % len(ring)], n_sub)
    dt = time.perf_counter() - t0
    value = pts / dt
    cores = blas_threads()
    desc = ("%s project_pcd + update_map, %d points of each %d-point frame per step, %d steps; BLAS threads pinned to %d of %d "
            "host cores (only the two dgemms are threaded, everything else in the reference is single-threaded numpy)"
            % ("the reference's own src/mapping_replay.py (oracle/_ref, byte-compiled)" if kind == "reference" else
               "numpy port (oracle/numpy_port.py) of", n_sub, args.points, args.steps, cores, os.cpu_count() or 1))
    line = {
        "impl": "reference", "metric": "points_fused_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_sec": value / args.points,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)


def workload_config(args, world):
    ids = getattr(args, "label_format", "rgb") == "ids"
    wl = getattr(args, "workload", "cfg2")
    what = {"cfg2": "", "cfg3": "", "cfg4": ", %d frames in total sharded over the ranks" % args.frames,
            "cfg5": ", cam1 / cam6 frames alternating, row-tiled filter + render + image gather at the end"}[wl]
    return {"workload": "mapping_replay %s: %d-point cloud + 1920x1440 19-class label image per frame%s, "
                        "%s update, %d mapped classes, grid %dx%d @ %.1f m%s"
                        % (wl, args.points, " (as uint8 class-id plane)" if ids else "",
                           "confusion-matrix log-likelihood" if log_update(args) else "count", args.classes, MAP_H, MAP_W,
                           RESOLUTION, what),
            "points_per_frame": args.points, "image": [1440, 1920] if ids else [1440, 1920, 3], "mapped_classes": args.classes,
            "grid": [MAP_H, MAP_W, args.classes], "update": "log-likelihood" if log_update(args) else "count",
            "frames_per_rank_per_block": args.steps,
            "parallelism": ("single GPU" if world == 1 else
                            "frames sharded over %d ranks; the ranks' increments are summed every %d frames per rank (touched "
                            "window, packed counts, NCCL all-reduce on an internal stream, overlapped with the next frames)"
                            % (world, getattr(args, "exchange_every", 256))),
            "l2_policy": "inputs larger than L2: ring of %d distinct frames (%.0f MB) resident in HBM"
                         % (args.ring, args.ring * (args.points * 16 + 1440 * 1920 * (1 if ids else 3)) / 1e6)}


def balanced(steps, cap):
    """K steps in ceil(K / cap) batches of (almost) equal size: a 20-step block is 10 + 10, not 16 + 4."""
    nb = -(-steps // cap)
    base, extra = divmod(steps, nb)
    return [base + (1 if i < extra else 0) for i in range(nb)]


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import tempfile
    import torch
    import torch.distributed as dist
    from vision_semantic_segmentation_b200 import frame_sharding
    from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
    from vision_semantic_segmentation_b200.config.base_cfg import get_cfg_defaults
    from vision_semantic_segmentation_b200.device_mapper import DeviceMapper
    from vision_semantic_segmentation_b200.mapping_replay import SemanticMapping
    from vision_semantic_segmentation_b200.renderer import filter_and_render

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    rank, world, local_rank = frame_sharding.init_from_env()
    dev = torch.device("cuda", local_rank)
    labels, names, colors = syn.class_setup(args.classes == 19)
    c = len(labels)
    cm, lane = update_matrix(args, labels), names.index("lane")
    dm = DeviceMapper(MAP_H, MAP_W, colors, cm, BOUNDARY, RESOLUTION, RANGE_MAX, True, lane,
                      cameras=[camera_setup_1(), camera_setup_6()], device=local_rank)
    if world > 1:
        dm.init_comm()

    if args.ordered:   # smap_clear re-arms the count update: undo that after every clear
        plain_clear = dm.clear

        def ordered_clear():
            plain_clear()
            dm.notify_map_modified()
        dm.clear = ordered_clear
        dm.notify_map_modified()

    ring_host = make_ring(args, rank)
    ring_dev = []
    if args.label_format == "ids":
        dm.set_label_palette(syn.COLORS_19)
    for fr in ring_host:
        dp = torch.from_numpy(fr["points"]).to(dev)
        di = torch.from_numpy(fr["semantic_ids"] if args.label_format == "ids" else fr["semantic_image"]).to(dev)
        ring_dev.append((dm.make_frame(dp, di, fr["T"], fr["camera"]), dp, di))
    torch.cuda.synchronize()

    # ---- algorithmic bytes per frame: N, M, U measured on the device path itself (outside timed regions)
    n_pts = args.points
    m_list, u_list, k_list = [], [], []
    for i, (frame, dp, di) in enumerate(ring_dev[: min(4, len(ring_dev))]):
        if args.label_format == "ids":   # the parity API takes RGB images: count the survivors with one
            masked, _ = dm.project(dm.make_frame(dp, torch.from_numpy(ring_host[i]["semantic_image"]).to(dev),
                                                 ring_host[i]["T"], ring_host[i]["camera"]))
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
    pos = [0]   # position in the ring: blocks continue where the previous one stopped
    xch = {"since": 0, "count": 0, "on": False}   # N > 1, timed region: frames since the last exchange, exchanges so far

    def block(steps):
        """exactly `steps` frames through smap_integrate_batch, in balanced batches; inside the timed region of an
        N > 1 run the ranks' increments are handed over every --exchange-every frames"""
        for take in balanced(steps, args.batch):
            start = pos[0] % len(ring_frames)
            dm.integrate_batch([ring_frames[(start + j) % len(ring_frames)] for j in range(take)])
            pos[0] += take
            if xch["on"]:
                xch["since"] += take
                if xch["since"] >= args.exchange_every:
                    dm.exchange_async()
                    xch["since"] = 0
                    xch["count"] += 1

    def render_tiles():
        """cfg5: filter + render of this rank's row tile (one-row halos), image all-gathered"""
        r0, r1 = frame_sharding.row_tile(MAP_H, rank, world)
        top, bottom = (1 if r0 > 0 and r1 > r0 else 0), (1 if r1 < MAP_H and r1 > r0 else 0)
        rgb_tile = frame_sharding.render_row_tile(dm.map[r0 - top:r1 + bottom], top, bottom, colors)
        return frame_sharding.gather_rgb_rows(rgb_tile, MAP_H) if world > 1 else rgb_tile

    # ---- warm-up: kernels, and for N > 1 the collectives at the size they will have (NCCL sets its channels up lazily)
    if world > 1:
        dm.set_streaming(True)
    block(max(args.warmup, 3))
    if world > 1:
        for _ in range(3):
            block(min(args.exchange_every, 64))
            dm.exchange_async()
        dm.exchange_flush()
    if args.workload == "cfg5":
        render_tiles()
    torch.cuda.synchronize()
    # ---- how long is one block?  -> R blocks for a region of >= min_region_ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    block(args.steps)          # the first block of this length instantiates its launch graph
    e0.record()
    block(args.steps)
    e1.record()
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1), 1e-3)
    min_region_ms = args.min_region_ms if args.min_region_ms > 0 else (50.0 if world == 1 else 300.0)
    repeats = args.repeats if args.repeats > 0 else int(min(4000, max(1, np.ceil(min_region_ms / est))))
    if world > 1:   # every rank must run the same number of blocks
        t = torch.tensor([repeats], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        repeats = int(t.item())
        dm.exchange_async()
        dm.exchange_flush()
    dm.clear()

    # ---- device-resident throughput (`value`)
    launches_before = dm.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank, period_s=args.clock_period_ms * 1e-3)
    sampler.start(thread=(rank == 0 and args.clock_period_ms > 0))
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(repeats + 1)]
    e_end = torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()          # no collector pause on any rank inside the region (one late host stalls every rank's exchange)
    barrier()
    xch["on"] = world > 1
    t_host0 = time.perf_counter()
    for r in range(repeats):
        marks[r].record()
        block(args.steps)
        if rank != 0 and r == repeats // 2:
            sampler.sample_now()
    marks[repeats].record()
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host0)
    if world > 1:
        xch["on"] = False
        if xch["since"] > 0 or xch["count"] == 0:   # the frames behind the last regular exchange
            dm.exchange_async()
            xch["count"] += 1
        dm.exchange_flush()
    n_exchanges = xch["count"]
    if args.workload == "cfg5":
        rgb_full = render_tiles()
    e_end.record()
    sampler.sample_now()
    barrier()
    gc.enable()
    ms = marks[0].elapsed_time(e_end)
    block_ms = np.array([marks[r].elapsed_time(marks[r + 1]) for r in range(repeats)])
    clocks = sampler.stop()
    launches = dm.stats()["kernel_launches"] - launches_before
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own view of the region (the line's ms_per_step is the MAX): its time, its median block, how long
        # its host took to queue the region, its SM clock under load and throttle reasons
        mine = torch.tensor([ms, float(np.median(block_ms)), host_enqueue_ms, clocks.get("sm_mhz") or 0.0,
                             float(sampler.reason_bits), float(len(os.sched_getaffinity(0)))], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": i, "region_ms": float(v[0]), "median_block_us_per_step": 1e3 * float(v[1]) / args.steps,
                     "host_enqueue_ms": float(v[2]), "sm_mhz": float(v[3]), "clock_event_reason_bits": int(v[4]),
                     "cpus_bound_to": int(v[5])}
                    for i, v in enumerate(x.cpu().numpy() for x in allr)]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_steps = repeats * args.steps
    value = world * total_steps * n_pts / (ms * 1e-3)

    # ---- N > 1: the exchange on its own, and parity of the exchanged grid against torch.distributed's all-reduce
    collective = None
    if world > 1:
        # one block, then the exchange alone (nothing to overlap with): agreement + pack + all-reduce + unpack, host included
        n_x = min(args.exchange_every, max(args.steps, 16))   # frames behind the exchange that is timed and checked
        dm.clear()
        pos[0] = 0
        block(n_x)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dm.exchange_async()
        dm.exchange_flush()
        torch.cuda.synchronize()
        exch_ms = 1e3 * (time.perf_counter() - t0)
        got = dm.map.clone()
        phases = dm.comm_info()
        dm.set_streaming(False)
        dm.clear()
        pos[0] = 0
        block(n_x)
        want = dm.map.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        if log_update(args):
            err = float(((got - want).abs() / want.abs().clamp_min(1e-300)).max().item())
            parity = {"max_rel_err": err, "ok": bool(err <= 1e-5)}
        else:
            parity = {"bit_exact": bool(torch.equal(got, want)), "ok": bool(torch.equal(got, want))}
        parity["checksum"] = float(want.sum().item())
        tt = torch.tensor([exch_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        collective = {"exchange_alone_ms": float(tt.item()), "bytes_per_exchange": phases["bytes"], "grid_bytes": phases["grid_bytes"],
                      "window": phases["window"], "pack": phases["pack"], "exchanges_in_region": n_exchanges,
                      "frames_per_rank_between_exchanges": args.exchange_every, "frames_behind_the_timed_exchange": n_x,
                      "pack_ms": phases["pack_ms"], "reduce_ms": phases["reduce_ms"], "add_ms": phases["add_ms"],
                      "parity_vs_torch_all_reduce": parity,
                      "note": "exchange_alone_ms: one exchange with nothing to overlap (agreement + pack + NCCL all-reduce + "
                              "unpack-add, host time included); inside the region the exchanges run on an internal stream "
                              "under the next block's frames"}
        if not parity["ok"]:
            raise SystemExit("exchanged grid differs from torch.distributed all_reduce of the per-rank grids: %r" % (parity,))
        dm.clear()

    # ---- roofline of the dominant kernel (k_fuse: project + cull + lookup + update of one frame).
    # kernel_ms: the launch shape of the timed region -- a batch's launches on the internal streams, two half-grid
    # launches side by side -- timed with CUDA events around every smap_integrate_batch call of a pass (events recorded on
    # the launching stream; the batch's fork / join hang off it), per frame.  kernel_alone_ms: the same kernel launched
    # alone with the full grid (smap_set_profiling serialises the launches).  Algorithmic bytes: SURVEY.md 8d's
    # 16 N + 3 M + 2*8 U per frame with N, M, U measured above on the device path.
    pos[0] = 0
    block(16)
    torch.cuda.synchronize()
    bsz = balanced(args.steps, args.batch)[0]     # the batch size of the timed region
    n_b = max(4, min(64, total_steps // bsz))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_b + 1)]
    for b in range(n_b):
        evs[b].record()
        block(bsz)
    evs[n_b].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.median([evs[b].elapsed_time(evs[b + 1]) for b in range(n_b)])) / bsz
    dm.clear()
    dm.set_profiling(True)
    block(max(32, min(args.steps, 128)))
    st = dm.stats()
    dm.set_profiling(False)
    kernel_alone_ms = st["stream_kernel_ms"] / max(st["profiled_frames"], 1)
    apply_ms = st["apply_kernel_ms"] / max(st["profiled_frames"], 1)
    peak, peak_src = peaks()
    achieved = bytes_per_frame / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")   # from the committed ncu --set full capture
    if args.label_format == "rgb" and args.workload in ("cfg2", "cfg4") and not args.ordered and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("k_fuse_c%d" % args.classes)

    # ---- render leg: apply_filter + render_bev_map of the whole grid (once per replay), HBM-bound
    render = None
    if not args.no_render:
        pos[0] = 0
        block(min(64, max(args.steps, 16)))
        filter_and_render(dm.map, colors)
        torch.cuda.synchronize()
        revs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for i in range(5):
            revs[i].record()
            filter_and_render(dm.map, colors)
        revs[5].record()
        torch.cuda.synchronize()
        r_ms = float(np.median([revs[i].elapsed_time(revs[i + 1]) for i in range(5)]))
        r_bytes = float(MAP_H) * MAP_W * c * 8 + float(MAP_H) * MAP_W * 3
        render = {"kernel": "k_render<filter> (3x3 box filter + first-argmax colour, one pass)", "ms": r_ms,
                  "algorithmic_bytes": r_bytes, "achieved": r_bytes / (r_ms * 1e-3) / 1e9, "unit": "GB/s",
                  "frac": r_bytes / (r_ms * 1e-3) / 1e9 / peak}

    # ---- end to end through the reference-facing API: SemanticMapping.mapping_replay(input_list) with HOST frames in
    # pinned memory.  Inside the timed region: clear, H2D of every frame's cloud + label image (pinned, multi-buffered,
    # replay_feed), the fused kernels, filter + render, D2H of the rendered map.
    e2e = None
    if not args.no_e2e and args.workload in ("cfg2", "cfg3"):
        tmp = tempfile.mkdtemp()
        sm = SemanticMapping(make_cfg(get_cfg_defaults(), args, tmp), device=local_rank)
        sm.set_label_palette(syn.COLORS_19)
        k_e2e = max(8, min(args.e2e_steps, 4 * args.steps))
        # N > 1: the API shards the list over the ranks itself, so every rank hands it the SAME list
        src_frames = ring_host if world == 1 else make_ring(args, 0, n=min(8, args.ring))
        rgb_frames, ids_frames, pcd_frames, keep = [], [], [], []
        for fr in src_frames[: min(8, len(src_frames))]:
            hp = torch.from_numpy(fr["points"]).pin_memory()
            hi = torch.from_numpy(fr["semantic_image"]).pin_memory()
            hd = torch.from_numpy(fr["semantic_ids"]).pin_memory()
            h64 = torch.from_numpy(np.ascontiguousarray(fr["points"].T.astype(np.float64))).pin_memory()
            keep.append((hp, hi, hd, h64))
            base = {"pcd_frame_id": "world", "pose": fr["pose"], "camera_id": fr["camera_id"]}
            rgb_frames.append(dict(base, points=hp.numpy(), semantic_image=hi.numpy()))
            ids_frames.append(dict(base, points=hp.numpy(), semantic_ids=hd.numpy()))
            pcd_frames.append(dict(base, pcd=h64.numpy(), semantic_image=hi.numpy()))   # src/mapping.py:309-312

        def api_leg(which, frames):
            input_list = [frames[i % len(frames)] for i in range(k_e2e)]
            sm.mapping_replay(input_list, "warm", write_image=False)   # same shape: every slot set of the feed allocated
            barrier()
            t0 = time.perf_counter()
            rgb = sm.mapping_replay(input_list, "bench", write_image=False)
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            first = input_list[0]
            img = first["semantic_image"] if "semantic_image" in first else first["semantic_ids"]
            cloud = first["points"] if "points" in first else first["pcd"]
            return {"value": k_e2e * n_pts / float(tt.item()), "unit": "points/s",
                    "h2d_bytes_per_step": int(cloud.nbytes + img.nbytes), "d2h_bytes_per_step": int(rgb.nbytes // k_e2e),
                    "steps": k_e2e, "ms_per_step": 1e3 * float(tt.item()) / k_e2e,
                    "pcie_gbs": (cloud.nbytes + img.nbytes) * k_e2e / world / float(tt.item()) / 1e9,
                    "note": which}, rgb

        e2e, rgb_a = api_leg("SemanticMapping.mapping_replay(input_list): frames as host dictionaries in pinned memory "
                             "(points (N, 4) float32 + (1440, 1920, 3) RGB label image); timed: clear, pinned "
                             "multi-buffered H2D of every frame, fused kernels, filter + render, D2H of the rendered map",
                             rgb_frames)
        ids_leg, rgb_b = api_leg("same call, label image handed over as the network's (1440, 1920) uint8 class-id plane",
                                 ids_frames)
        ids_leg["same_render_as_rgb"] = bool(np.array_equal(rgb_a, rgb_b))
        e2e["class_ids"] = ids_leg
        ref_leg, rgb_c = api_leg("same call, cloud in the reference's own record layout: pcd (4, N) float64 "
                                 "(src/mapping.py:309-312), converted to the float4 layout on the device "
                                 "(smap_cloud_to_f32x4)", pcd_frames)
        ref_leg["same_render_as_float32"] = bool(np.array_equal(rgb_a, rgb_c))
        e2e["reference_record_format"] = ref_leg
        # what the link itself does: one large pinned H2D copy, timed alone
        big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dbig = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        dbig.copy_(big, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            dbig.copy_(big, non_blocking=True)
        torch.cuda.synchronize()
        e2e["pinned_h2d_peak_gbs"] = 4 * big.numel() / (time.perf_counter() - t0) / 1e9
        e2e["frac_of_pinned_h2d_peak"] = e2e["pcie_gbs"] / e2e["pinned_h2d_peak_gbs"]
        del big, dbig, sm, keep

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        run, cpu_render, kind = cpu_frame_fn(args)
        ncores = os.cpu_count() or 1
        set_blas_threads(ncores)
        v_all = time_cpu(run, ring_host, args.cpu_frames)
        cores = blas_threads()
        set_blas_threads(1)
        v_one = time_cpu(run, ring_host, max(2, args.cpu_frames // 2))
        set_blas_threads(ncores)
        cpu = {"value": v_all, "unit": "points/s", "cores": cores, "kind": kind, "single_thread_value": v_one,
               "sample": "%d full frames of the same workload through %s project_pcd + update_map; %d host cores, BLAS "
                         "threads pinned to %d (only the two dgemms are threaded); single_thread_value: BLAS pinned to 1"
                         % (args.cpu_frames, "the reference's own src/mapping_replay.py (oracle/_ref, byte-compiled from "
                            "/root/reference)" if kind == "reference" else "oracle/numpy_port.py (numpy restatement of)",
                            ncores, cores)}
        if render is not None and cpu_render is not None:
            crop = None if MAP_H * MAP_W <= 4000000 else 2000
            sec = cpu_render(crop)
            scale = 1.0 if crop is None else (MAP_H * MAP_W) / float(crop * crop)
            render["cpu_ms"] = 1e3 * sec * scale
            render["cpu_note"] = ("the reference's apply_filter + render_bev_map (src/renderer.py:32-59,175-189) on the host"
                                  + ("" if crop is None else ", timed on a %d x %d crop and scaled by area" % (crop, crop)))

    if rank == 0:
        line = {
            "metric": "points_fused_per_sec", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "repeats": repeats, "ms_per_step": ms / total_steps,
            "block_ms_per_step": {"median": float(np.median(block_ms)) / args.steps, "min": float(block_ms.min()) / args.steps,
                                  "max": float(block_ms.max()) / args.steps, "slowest_block": int(np.argmax(block_ms)),
                                  "blocks": int(len(block_ms))},
            "timed_region_ms": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "cfg4" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "frames_per_sec": value / n_pts,
            "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("k_fuse<masks> (project+cull+lookup+RED.OR into the frame's cell masks, one frame per "
                                    "launch); the ordered k_apply replay is apply_kernel_ms_per_frame"
                                    if (log_update(args) or args.ordered) else
                                    "k_fuse<count> (register-prefetched cloud, float32 certified project+cull+lookup+update, "
                                    "one frame per launch)"),
                         "kernel_ms": kernel_ms,
                         "kernel_ms_shape": "per frame in the launch shape of the timed region: batches of %d frames, one "
                                            "half-grid launch per frame, two side by side (events around each batch)" % bsz,
                         "kernel_alone_ms": kernel_alone_ms, "frac_alone": bytes_per_frame / (kernel_alone_ms * 1e-3) / 1e9 / peak,
                         "apply_kernel_ms_per_frame": apply_ms,
                         "step_frac": bytes_per_frame / (ms / total_steps * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_frame": bytes_per_frame,
                         "N": n_pts, "M": M, "K_cells": Kc, "U_elements": U},
            "render": render, "collective": collective, "per_rank": per_rank, "host_enqueue_ms": host_enqueue_ms,
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
