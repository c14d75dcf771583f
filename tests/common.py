"""Shared helpers of the test-suite: golden manifest access, frame regeneration, mapper set-up."""
import hashlib
import json
import os

import numpy as np

from vision_semantic_segmentation_b200 import synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
from vision_semantic_segmentation_b200.utils import transforms as tr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_arrays(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


class Case(object):
    """Everything needed to run one golden case through the oracle or the CUDA path."""

    def __init__(self, name):
        self.name = name
        self.spec = manifest()["cases"][name]
        s = self.spec
        self.labels, self.names, self.colors = syn.class_setup(s["full19"])
        self.boundary = s.get("boundary") or [[100, 300], [800, 1000]]
        self.resolution = s.get("resolution") or 0.1
        self.range_max = s.get("range_max") or 100.0
        self.use_intensity = s.get("use_intensity", True)
        self.mh, self.mw, self.c = s["map_shape"]
        self.lane = self.names.index("lane") if "lane" in self.names else -1
        self.cam = camera_setup_1() if s.get("camera", 1) == 1 else camera_setup_6()
        self.frame_id = s.get("pcd_frame_id", "world")
        self.arrays = golden_arrays(name)
        self.cm = np.ascontiguousarray(self.arrays["confusion_matrix"], dtype=np.float64)
        self.T_v2b = syn.velodyne_to_baselink()

    def frame(self, f):
        """(pcd (4,N) f64, points (N,4) f32 or None, image, T or None), input hashes verified."""
        s = self.spec
        fr = syn.synthetic_frame(s["seed"], f, s["n_points"], height=s["image_hw"][0], width=s["image_hw"][1],
                                 blocky=(f in s.get("blocky_frames", [])))
        out = s["frames_out"][f]
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ self.T_v2b)
        if self.frame_id == "velodyne":
            pcd = fr["pcd"].copy()
            pcd[0:3] = (T @ np.vstack((pcd[0:3], np.ones((1, pcd.shape[1])))))[0:3].astype(np.float32)
            assert sha(pcd) == out["in_points_sha"], "synthetic generator drifted (cloud)"
            points = np.ascontiguousarray(pcd.T.astype(np.float32))
            T = None
        else:
            pcd, points = fr["pcd"], fr["points"]
            assert sha(points) == out["in_points_sha"], "synthetic generator drifted (cloud)"
        assert sha(fr["semantic_image"]) == out["in_image_sha"], "synthetic generator drifted (image)"
        return pcd, points, fr["semantic_image"], T

    def sparse_map(self):
        m = np.zeros(self.mh * self.mw * self.c)
        m[self.arrays["map_idx"]] = self.arrays["map_val"]
        return m.reshape(self.mh, self.mw, self.c)


GOLDEN_CASES = ["cfg1_c5_count", "cfg1_c5_log", "cfg1_c19_count", "cfg1_c19_log", "cam6_res02", "small_velodyne"]
