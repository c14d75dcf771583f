#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/ from the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference (imported unmodified through oracle/ref_shim.py) is executed on seeded
synthetic frames; what it returns is stored as

  tests/golden/manifest.json   one entry per case: generator parameters, sha256 of the
                               regenerated inputs (so a drifting RNG is detected, not
                               silently compared), sha256 of every reference output
  tests/golden/<case>.npz      compact copies of the outputs that make a failure
                               debuggable (kept-point indices, sparse map, rendered image)

The GPU box has no /root/reference: tests there regenerate the inputs from the manifest,
check the input hashes, and compare the oracle / the CUDA path with these files.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from vision_semantic_segmentation_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def sparse(a):
    flat = a.reshape(-1)
    idx = np.flatnonzero(flat)
    return idx.astype(np.int64), flat[idx]


def build_mapper(ref, full19, log_cm, cm_seed, boundary=None, resolution=None, range_max=None,
                 use_intensity=True):
    tmp = tempfile.mkdtemp()
    cfg = ref.get_cfg_defaults()
    cfg.OUTPUT_DIR = tmp
    cfg.LABELS, cfg.LABELS_NAMES, cfg.LABEL_COLORS = syn.class_setup(full19)
    if boundary is not None:
        cfg.MAPPING.BOUNDARY = boundary
    if resolution is not None:
        cfg.MAPPING.RESOLUTION = resolution
    if range_max is not None:
        cfg.MAPPING.PCD.RANGE_MAX = range_max
    cfg.MAPPING.PCD.USE_INTENSITY = use_intensity
    if log_cm:
        path = os.path.join(tmp, "cm.npy")
        np.save(path, syn.synthetic_confusion_matrix(cm_seed))
        cfg.MAPPING.CONFUSION_MTX.LOAD_PATH = path
    return ref.SemanticMapping(cfg)


def reference_uv(ref_mod, sm, pcd, frame_id, pose, cam):
    """The reference computes IXY inside project_pcd (src/mapping_replay.py:223-232) but does not
    return it; re-evaluate those statements with the reference's own helpers."""
    from src.utils.utils import homogenize, dehomogenize  # reference module (loaded by the shim)
    from src.utils.utils_ros import get_transform_from_pose
    if frame_id != "velodyne":
        T = np.linalg.inv(np.matmul(get_transform_from_pose(pose), sm.T_velodyne_to_basklink))
        velo = np.matmul(T, homogenize(pcd[0:3, :]))
    else:
        velo = homogenize(pcd[0:3, :])
    with np.errstate(all="ignore"):
        return dehomogenize(np.matmul(cam.P, velo)).astype(np.int32)


def run_case(ref, name, spec):
    sm = build_mapper(ref, spec["full19"], spec["log_cm"], spec.get("cm_seed", 7),
                      spec.get("boundary"), spec.get("resolution"), spec.get("range_max"),
                      spec.get("use_intensity", True))
    cam = sm.cam1 if spec.get("camera", 1) == 1 else sm.cam6
    grid = np.zeros((sm.map_height, sm.map_width, sm.map_depth))
    entry = dict(spec)
    entry["map_shape"] = [sm.map_height, sm.map_width, sm.map_depth]
    entry["frames_out"] = []
    arrays = {}
    for f in range(spec["frames"]):
        fr = syn.synthetic_frame(spec["seed"], f, spec["n_points"], height=spec["image_hw"][0],
                                 width=spec["image_hw"][1], blocky=(f in spec.get("blocky_frames", [])))
        pcd = fr["pcd"]
        frame_id = spec.get("pcd_frame_id", "world")
        if frame_id == "velodyne":
            # express the cloud in the velodyne frame and feed it as such (src/mapping_replay.py:229-230)
            from vision_semantic_segmentation_b200.utils import transforms as tr
            T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
            pcd = pcd.copy()
            pcd[0:3] = (T @ np.vstack((pcd[0:3], np.ones((1, pcd.shape[1])))))[0:3].astype(np.float32)
        masked, label = sm.project_pcd(pcd, frame_id, fr["semantic_image"], fr["pose"], cam)
        uv_all = reference_uv(ref, sm, pcd, frame_id, fr["pose"], cam)
        # indices of the kept points: recover from the compaction (order preserving)
        keep = np.zeros(pcd.shape[1], dtype=bool)
        h, w = fr["semantic_image"].shape[:2]
        # the reference's own mask statements (:235-240) on its own intermediates
        from src.utils.utils import homogenize
        from src.utils.utils_ros import get_transform_from_pose
        if frame_id != "velodyne":
            Tm = np.linalg.inv(np.matmul(get_transform_from_pose(fr["pose"]), sm.T_velodyne_to_basklink))
            vx = np.matmul(Tm, homogenize(pcd[0:3, :]))[0]
        else:
            vx = pcd[0]
        keep = (0 < vx) & (vx < sm.pcd_range_max) & (0 <= uv_all[0]) & (uv_all[0] < w) & \
               (0 <= uv_all[1]) & (uv_all[1] < h)
        assert keep.sum() == masked.shape[1] and np.array_equal(pcd[:, keep], masked)
        grid = sm.update_map(grid, masked, label)
        entry["frames_out"].append({
            "in_points_sha": sha(fr["points"]) if frame_id == "world" else sha(pcd),
            "in_image_sha": sha(fr["semantic_image"]),
            "M": int(masked.shape[1]),
            "masked_pcd_sha": sha(masked), "label_sha": sha(label), "uv_sha": sha(uv_all[:, keep]),
            "map_sha_after": sha(grid),
        })
        arrays["keep_idx_%d" % f] = np.flatnonzero(keep).astype(np.int32)
        if spec.get("store_full"):
            arrays["pcd_%d" % f] = pcd
            arrays["image_%d" % f] = fr["semantic_image"]
            arrays["pose_%d" % f] = fr["pose"].as_array()
            arrays["uv_%d" % f] = uv_all[:, keep]
            arrays["label_%d" % f] = label
    if spec.get("store_map", True):
        arrays["map_idx"], arrays["map_val"] = sparse(grid)
    filtered = ref.apply_filter(grid)
    rgb = ref.render_bev_map(filtered, sm.label_colors)
    rgb_raw = ref.render_bev_map(grid, sm.label_colors)
    c = sm.map_depth
    rng = np.random.default_rng(spec["seed"] + 99)
    priority = [int(v) for v in rng.permutation(c)]
    thresholds = [float(v) for v in rng.uniform(0.02, 0.4, c)]
    rgb_thr = ref.render_bev_map_with_thresholds(grid, sm.label_colors, priority=priority, thresholds=thresholds)
    entry.update({"map_sha": sha(grid), "filtered_sha": sha(filtered), "rgb_sha": sha(rgb),
                  "rgb_raw_sha": sha(rgb_raw), "rgb_thr_sha": sha(rgb_thr),
                  "priority": priority, "thresholds": thresholds,
                  "confusion_matrix_sha": sha(np.asarray(sm.confusion_matrix, dtype=np.float64))})
    arrays["confusion_matrix"] = np.asarray(sm.confusion_matrix, dtype=np.float64)
    arrays["rgb"] = rgb
    arrays["rgb_thr"] = rgb_thr
    if spec.get("store_full"):
        arrays["filtered_idx"], arrays["filtered_val"] = sparse(filtered)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    return entry


CASES = {
    # BASELINE.json configs[0]: 100k-point cloud + one 19-class label image (3 frames; frame 1 blocky)
    "cfg1_c5_count": dict(full19=False, log_cm=False, seed=0, frames=3, n_points=100000,
                          image_hw=[1440, 1920], blocky_frames=[1]),
    "cfg1_c5_log": dict(full19=False, log_cm=True, seed=0, frames=3, n_points=100000,
                        image_hw=[1440, 1920], blocky_frames=[1]),
    "cfg1_c19_count": dict(full19=True, log_cm=False, seed=0, frames=3, n_points=100000,
                           image_hw=[1440, 1920], blocky_frames=[1]),
    "cfg1_c19_log": dict(full19=True, log_cm=True, seed=0, frames=3, n_points=100000,
                         image_hw=[1440, 1920], blocky_frames=[1], store_map=False),  # 4 MB of doubles: hash only
    # camera 6 calibration, coarser grid, short range, intensity boost off
    "cam6_res02": dict(full19=True, log_cm=False, seed=40, frames=2, n_points=60000, image_hw=[1440, 1920],
                       camera=6, resolution=0.2, boundary=[[0, 600], [0, 1400]], range_max=60.0,
                       use_intensity=False),
    # small, fully stored (inputs + outputs): velodyne-frame cloud, small image, small grid
    "small_velodyne": dict(full19=True, log_cm=True, seed=11, frames=2, n_points=4000, image_hw=[1440, 1920],
                           pcd_frame_id="velodyne", resolution=0.5, boundary=[[1360, 1500], [500, 630]],
                           store_full=True, blocky_frames=[0, 1]),
}


def main():
    ref = ref_shim.load_reference()
    os.makedirs(OUT, exist_ok=True)
    manifest = {"generator": "oracle/make_golden.py", "numpy": np.__version__, "cases": {}}
    import cv2
    manifest["opencv"] = cv2.__version__
    for name, spec in CASES.items():
        manifest["cases"][name] = run_case(ref, name, spec)
        print(name, manifest["cases"][name]["map_shape"], [f["M"] for f in manifest["cases"][name]["frames_out"]])
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
