"""Recorded-drive files: the replay input format either side of the mapping path.

The reference's live node appends one dictionary per camera frame to ``input_list``
(``src/mapping.py:309-313``: ``pcd`` (4, N) float64, ``pcd_frame_id``, ``semantic_image`` (H, W, 3) uint8,
``pose``) and dumps the list with hickle (``:324-326``); ``mapping_replay_dir`` loads it back
(``src/mapping_replay.py:146-159``).  hickle / HDF5 are not in the target image, so the same list is
stored as one ``.npz`` per drive.  Clouds are written as (N, 4) float32 -- the ``float4`` layout the kernels
read, and lossless for PointCloud2 data, whose fields are FLOAT32 -- unless a value is not
float32-representable, in which case the float64 (4, N) array is kept.

A frame may carry ``semantic_ids`` -- the network's (h, w) uint8 class-id plane -- instead of (or next to) the painted
``semantic_image`` (``label_image.py``; a third of the bytes at full resolution, a twelfth at ``IMAGE_SCALE`` 0.5),
with ``image_size`` (H, W) when the plane is smaller than the camera image.
"""
import numpy as np

from .utils.transforms import Pose

__all__ = ["save_input_list", "load_input_list"]


def _pose_array(pose):
    if isinstance(pose, Pose):
        return pose.as_array()
    p, o = pose.position, pose.orientation
    return np.array([p.x, p.y, p.z, o.x, o.y, o.z, o.w], dtype=np.float64)


def save_input_list(path, input_list, compressed=False):
    arrays = {"n_frames": np.array(len(input_list), dtype=np.int64)}
    for i, fr in enumerate(input_list):
        if "points" in fr:
            arrays["points_%d" % i] = np.ascontiguousarray(fr["points"], dtype=np.float32)
        else:
            pcd = np.asarray(fr["pcd"], dtype=np.float64)
            as32 = pcd.astype(np.float32)
            if pcd.shape[0] == 4 and np.array_equal(as32.astype(np.float64), pcd, equal_nan=True):
                arrays["points_%d" % i] = np.ascontiguousarray(as32.T)
            else:
                arrays["pcd_%d" % i] = pcd
        if fr.get("semantic_image") is None and fr.get("semantic_ids") is None:
            raise ValueError("frame %d has neither semantic_image nor semantic_ids" % i)
        if fr.get("semantic_image") is not None:
            arrays["image_%d" % i] = np.ascontiguousarray(fr["semantic_image"], dtype=np.uint8)
        if fr.get("semantic_ids") is not None:
            ids = np.ascontiguousarray(fr["semantic_ids"], dtype=np.uint8)
            if ids.ndim != 2:
                raise ValueError("semantic_ids must be a 2-D class-id plane")
            arrays["ids_%d" % i] = ids
            if fr.get("image_size") is not None:
                arrays["image_size_%d" % i] = np.array([int(v) for v in fr["image_size"]], dtype=np.int64)
        arrays["pose_%d" % i] = _pose_array(fr["pose"])
        arrays["frame_id_%d" % i] = np.array(str(fr["pcd_frame_id"]))
        arrays["camera_id_%d" % i] = np.array(int(fr.get("camera_id", 1)), dtype=np.int64)
    (np.savez_compressed if compressed else np.savez)(path, **arrays)


def load_input_list(path):
    if path.endswith(".hkl"):
        try:
            import hickle
        except ImportError:
            raise RuntimeError("reading %s needs hickle, which is not installed; re-record the drive as .npz "
                               "with replay_io.save_input_list" % path)
        with open(path, "rb") as f:
            return hickle.load(f)
    out = []
    with np.load(path, allow_pickle=False) as z:
        for i in range(int(z["n_frames"])):
            fr = {"pose": Pose.from_array(z["pose_%d" % i]),
                  "pcd_frame_id": str(z["frame_id_%d" % i]), "camera_id": int(z["camera_id_%d" % i])}
            if "image_%d" % i in z.files:
                fr["semantic_image"] = z["image_%d" % i]
            if "ids_%d" % i in z.files:
                fr["semantic_ids"] = z["ids_%d" % i]
                if "image_size_%d" % i in z.files:
                    fr["image_size"] = tuple(int(v) for v in z["image_size_%d" % i])
            if "points_%d" % i in z.files:
                fr["points"] = z["points_%d" % i]
            else:
                fr["pcd"] = z["pcd_%d" % i]
            out.append(fr)
    return out
