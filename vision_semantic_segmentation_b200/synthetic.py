"""Seeded synthetic frames for parity tests and benchmarks.

There is no recorded drive (``input_list.hkl``) in the target environment, so tests
and ``bench.py`` feed the mapping path with frames of the shapes BASELINE.json names:
a local LiDAR cloud of N points plus one 1920x1440 colour-coded label image and the
vehicle pose.  A frame is the same dictionary the reference's live node records
(``src/mapping.py:309-312``): ``pcd`` (4, N) float64 rows x, y, z, intensity in the
world frame, ``pcd_frame_id``, ``semantic_image`` (H, W, 3) uint8 RGB, ``pose``.

Cloud values are float32-representable (PointCloud2 fields are FLOAT32), so the
``float4`` device copy and the float64 host copy hold identical numbers.
"""
import math

import numpy as np

from .utils.transforms import Pose, get_transform_from_pose, euler_matrix

# Colours of the 19 network classes (reference config/config_19.json, in label order).
COLORS_19 = np.array([
    [196, 196, 196], [140, 140, 200], [128, 64, 128], [244, 35, 232], [70, 70, 70],
    [220, 20, 60], [255, 0, 0], [255, 0, 100], [255, 255, 255], [70, 130, 180],
    [107, 142, 35], [100, 128, 160], [153, 153, 153], [220, 220, 0], [119, 11, 32],
    [0, 60, 100], [0, 0, 142], [0, 0, 230], [0, 0, 70]], dtype=np.uint8)

NAMES_19 = ["curb", "crosswalk", "road", "sidewalk", "building", "person", "bicyclist",
            "motorcyclist", "lane", "sky", "vegetation", "manhole", "pole", "traffic_sign",
            "bicycle", "bus", "car", "motorcycle", "truck"]

IMAGE_W, IMAGE_H = 1920, 1440


def velodyne_to_baselink():
    """Constant mounting transform (reference ``src/mapping_replay.py:140-144``)."""
    T = euler_matrix(0.0, 0.140, 0.0)
    T[0, 3], T[1, 3], T[2, 3] = 2.64, 0.0, 1.98
    return T


def synthetic_pose(frame_idx, origin=(-1200.0, 300.0), step=(3.0, 1.0), yaw0=0.3, dyaw=0.05):
    yaw = yaw0 + dyaw * frame_idx
    return Pose((origin[0] + step[0] * frame_idx, origin[1] + step[1] * frame_idx, 0.0),
                (0.0, 0.0, math.sin(0.5 * yaw), math.cos(0.5 * yaw)))


def synthetic_label_ids(rng, height=IMAGE_H, width=IMAGE_W, blocky=False, block=64, num_ids=len(COLORS_19)):
    """(H, W) class ids: what the segmentation network would output for the frame."""
    if blocky:
        gh, gw = -(-height // block), -(-width // block)
        ids = rng.integers(0, num_ids, (gh, gw))
        ids = np.repeat(np.repeat(ids, block, axis=0), block, axis=1)[:height, :width]
    else:
        ids = rng.integers(0, num_ids, (height, width))
    return ids


def synthetic_label_image(rng, height=IMAGE_H, width=IMAGE_W, blocky=False, block=64, colors=COLORS_19):
    return np.ascontiguousarray(colors[synthetic_label_ids(rng, height, width, blocky, block, len(colors))])


def synthetic_points(rng, n_points, pose):
    """(N, 4) float32 world-frame points x, y, z, intensity."""
    local = np.empty((4, n_points), dtype=np.float64)
    local[0] = rng.uniform(-5.0, 120.0, n_points)
    local[1] = rng.uniform(-60.0, 60.0, n_points)
    local[2] = rng.normal(-1.9, 0.5, n_points)
    local[3] = 1.0
    intensity = rng.uniform(0.0, 30.0, n_points).astype(np.float32)
    local[0:3] = local[0:3].astype(np.float32)
    T = get_transform_from_pose(pose) @ velodyne_to_baselink()
    world = (T @ local)[0:3].astype(np.float32)
    pts = np.empty((n_points, 4), dtype=np.float32)
    pts[:, 0:3] = world.T
    pts[:, 3] = intensity
    return pts


def synthetic_frame(seed, frame_idx, n_points, height=IMAGE_H, width=IMAGE_W, blocky=False,
                    pose=None, as_float64=True, with_ids=False):
    """One frame dictionary.  ``points`` (N,4) float32 is the device-friendly record;
    ``pcd`` (4,N) float64 is what the reference API takes (same values).  ``with_ids``: also ``semantic_ids``, the
    (H, W) uint8 class-id plane the label image was painted from (palette ``COLORS_19``)."""
    rng = np.random.default_rng(seed + frame_idx)
    pose = synthetic_pose(frame_idx) if pose is None else pose
    pts = synthetic_points(rng, n_points, pose)
    ids = synthetic_label_ids(rng, height, width, blocky)
    image = np.ascontiguousarray(COLORS_19[ids])
    frame = {"points": pts, "pcd_frame_id": "world", "semantic_image": image, "pose": pose}
    if with_ids:
        frame["semantic_ids"] = np.ascontiguousarray(ids.astype(np.uint8))
    if as_float64:
        frame["pcd"] = np.ascontiguousarray(pts.T.astype(np.float64))
    return frame


def synthetic_confusion_matrix(seed, num_class=19):
    """Strictly positive 19x19 count matrix with a strong diagonal (no -inf after log)."""
    rng = np.random.default_rng(seed)
    return (rng.integers(1, 50, (num_class, num_class)) + 1000 * np.eye(num_class, dtype=np.int64)).astype(np.float64)


def class_setup(full19=False):
    """(LABELS, LABELS_NAMES, LABEL_COLORS) for the default 5-class or the full 19-class run."""
    if full19:
        return list(range(19)), list(NAMES_19), COLORS_19.astype(int).tolist()
    labels = [2, 1, 8, 10, 3]
    return labels, ["road", "crosswalk", "lane", "vegetation", "sidewalk"], COLORS_19[labels].astype(int).tolist()
