"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's CPU mapping path.

Purpose: (1) the CPU baseline that ``bench.py`` times on the GPU box's host cores
(``cpu_baseline.kind = "port"``; the reference itself is Python + numpy, so a numpy port
has the same cost structure: two BLAS dgemms, a dozen full-length float64 temporaries,
one fancy-index read-modify-write per class); (2) a second, independently written
checker beside ``smap_oracle.c``.

Bit-exactness note: the dgemm rounding (which fused chain OpenBLAS uses) depends on the
host CPU's kernel; the scalar C oracle fixes that chain explicitly and is THE bit-exact
spec.  This port agrees with it wherever OpenBLAS uses the same chain (it does on the
build container, see tests/test_oracle.py) and otherwise differs only for points that
sit within one ulp of a pixel boundary.

Follows: src/mapping_replay.py:214-301 (project_pcd, update_map), src/renderer.py:32-59,
131-172, 175-189, src/data/confusion_matrix.py:25-63, src/camera.py:21-35.
"""
import numpy as np

PCD_ORIGIN_OFFSET = (1369.0496826171875, 562.84814453125)  # src/mapping_replay.py:261


def world_to_velodyne(T_base_to_origin, T_velodyne_to_baselink):
    """src/mapping_replay.py:225-226"""
    return np.linalg.inv(np.matmul(T_base_to_origin, T_velodyne_to_baselink))


def project_pcd(pcd, T, P, image, range_max):
    """src/mapping_replay.py:223-246.  T is None for pcd_frame_id == 'velodyne'."""
    n = pcd.shape[1]
    homo = np.vstack((pcd[0:3, :], np.ones((1, n))))
    velo = np.matmul(T, homo) if T is not None else homo
    q = np.matmul(P, velo)
    with np.errstate(all="ignore"):
        ixy = (q[:-1] / q[-1]).astype(np.int32)
    in_front = (0 < velo[0]) & (velo[0] < range_max)
    h, w = image.shape[0], image.shape[1]
    on_image = (0 <= ixy[0]) & (ixy[0] < w) & (0 <= ixy[1]) & (ixy[1] < h)
    keep = on_image & in_front
    masked = pcd[:, keep]
    idx = ixy[:, keep]
    label = image[idx[1], idx[0]].T
    return masked, label, idx, keep


def update_map(map_, pcd, label, colors, cm, boundary, resolution, use_intensity, label_names,
               offset=PCD_ORIGIN_OFFSET):
    """src/mapping_replay.py:248-301; mutates and returns map_."""
    mh, mw = map_.shape[0], map_.shape[1]
    local_xy = pcd[0:2] + np.array([[offset[0]], [offset[1]]])
    lo = np.array([[boundary[0][0]], [boundary[1][0]]])
    with np.errstate(all="ignore"):
        cell = ((local_xy - lo) / resolution).astype(np.int32)
    on_grid = (0 <= cell[0]) & (cell[0] < mh) & (0 <= cell[1]) & (cell[1] < mw)
    colors = np.asarray(colors)
    for i, name in enumerate(label_names):
        # the reference's logical_and(*rows) takes row 3 as out=, so only R and G are compared
        hit = (label[0] == colors[i][0]) & (label[1] == colors[i][1]) & on_grid
        map_[cell[0, hit], cell[1, hit], :] += cm[:, i].reshape(1, -1)
        if use_intensity and name == "lane":
            strong = ((pcd[3] < 2) | (pcd[3] > 14)) & hit
            map_[cell[0, strong], cell[1, strong], i] += 2
    return map_


def apply_filter(src):
    """src/renderer.py:175-189"""
    import cv2
    k = np.ones((3, 3), dtype=np.float32)
    k /= 9
    return cv2.filter2D(src, -1, k)


def render_bev_map(map_, colors):
    """src/renderer.py:32-59"""
    colors = np.asarray(colors)
    if map_.shape[2] != len(colors):
        raise ValueError("Each channel should have a color!")
    out = np.zeros(map_.shape[:2] + (3,), dtype=np.uint8)
    best = np.argmax(map_, axis=2)
    for i in range(map_.shape[2]):
        out[best == i] = colors[i]
    out[np.sum(map_, axis=2) == 0] = 0
    return out


def render_bev_map_with_thresholds(map_, colors, priority, thresholds):
    """src/renderer.py:131-172"""
    colors = np.asarray(colors)
    c = map_.shape[2]
    priority = np.arange(c) if priority is None else np.asarray(priority)
    total = np.sum(map_, axis=2, keepdims=True)
    prob = np.divide(map_, total, out=np.zeros_like(map_), where=(total != 0))[:, :, priority]
    known = np.sum(map_, axis=2) != 0
    out = np.zeros(map_.shape[:2] + (3,), dtype=np.uint8)
    for i in range(c):
        out[(prob[:, :, i] >= thresholds[i]) & known] = colors[priority[i]]
    return out


def confusion_submatrix_log(cm_full, indices):
    """src/data/confusion_matrix.py:43-48,59-63 with to_probability=True, use_log=True"""
    sub = cm_full[np.ix_(indices, indices)]
    sub = sub / np.sum(sub, axis=1)[:, np.newaxis]
    with np.errstate(divide="ignore"):
        return np.log(sub)


def camera_P(K, Rt):
    """src/camera.py:21-35,102-117 : R = Rt[:3,:3].T, t = -R Rt[:3,3], P = K [R t]"""
    R = Rt[0:3, 0:3].T
    t = -np.matmul(R, Rt[0:3, 3:4])
    return np.matmul(K, np.concatenate([R, t], axis=1))
