"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/smap_oracle.c.

The scalar C restatement is the bit-exact checker for the CUDA path.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsmap_oracle.so")
_lib = None

_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "smap_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsmap_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.smo_project.restype = ctypes.c_int64
        L.smo_project.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64, _f64p, _f64p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_double, _u8p, _i32p, _i32p]
        L.smo_gather.restype = None
        L.smo_gather.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64, _u8p, _i32p, _i32p, _u8p, ctypes.c_int,
                                 ctypes.c_int64, _f64p, _u8p, _i32p]
        L.smo_update.restype = ctypes.c_int
        L.smo_update.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, ctypes.c_int64, _u8p,
                                 ctypes.c_int64, ctypes.c_int64, _u8p, _f64p, ctypes.c_double, ctypes.c_double,
                                 ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, _i64p]
        L.smo_filter.restype = None
        L.smo_filter.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p]
        L.smo_render.restype = None
        L.smo_render.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p, _u8p]
        L.smo_render_thresholds.restype = None
        L.smo_render_thresholds.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p, _i32p, _f64p, _u8p]
        L.smo_np_sum.restype = ctypes.c_double
        L.smo_np_sum.argtypes = [_f64p, ctypes.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


PCD_ORIGIN_OFFSET = (1369.0496826171875, 562.84814453125)  # src/mapping_replay.py:261


def project_pcd(pcd, T, P, image, range_max):
    """Restates SemanticMapping.project_pcd (src/mapping_replay.py:214-246) given the host-side
    4x4 ``T`` (None for the velodyne frame).  Returns masked_pcd (4,M) f64, label (3,M) u8,
    image_idx (2,M) i32, keep (N,) bool."""
    pcd = np.ascontiguousarray(pcd, dtype=np.float64)
    image = np.ascontiguousarray(image, dtype=np.uint8)
    n = pcd.shape[1]
    P = np.ascontiguousarray(P, dtype=np.float64)
    Tp = None
    if T is not None:
        T = np.ascontiguousarray(T, dtype=np.float64)
        Tp = _p(T, _f64p)
    keep = np.zeros(n, dtype=np.uint8)
    iu = np.zeros(n, dtype=np.int32)
    iv = np.zeros(n, dtype=np.int32)
    m = lib().smo_project(_p(pcd, _f64p), n, n, Tp, _p(P, _f64p), image.shape[1], image.shape[0],
                          float(range_max), _p(keep, _u8p), _p(iu, _i32p), _p(iv, _i32p))
    out_pcd = np.empty((pcd.shape[0], m), dtype=np.float64)
    out_label = np.empty((3, m), dtype=np.uint8)
    out_uv = np.empty((2, m), dtype=np.int32)
    lib().smo_gather(_p(pcd, _f64p), n, n, _p(keep, _u8p), _p(iu, _i32p), _p(iv, _i32p), _p(image, _u8p),
                     image.shape[1], m, _p(out_pcd, _f64p), _p(out_label, _u8p), _p(out_uv, _i32p))
    return out_pcd, out_label, out_uv, keep.astype(bool)


def update_map(map_, pcd, label, colors, cm, boundary, resolution, use_intensity, lane_index,
               offset=PCD_ORIGIN_OFFSET):
    """Restates SemanticMapping.update_map (src/mapping_replay.py:248-301); in place.  Returns stats
    [points with a class on the grid, touched cells K, touched (cell,class) pairs, boosted cells]."""
    assert map_.dtype == np.float64 and map_.flags.c_contiguous
    pcd = np.ascontiguousarray(pcd, dtype=np.float64)
    label = np.ascontiguousarray(label, dtype=np.uint8)
    colors = np.ascontiguousarray(np.asarray(colors).astype(np.uint8))
    cm = np.ascontiguousarray(cm, dtype=np.float64)
    mh, mw, c = map_.shape
    m = pcd.shape[1]
    stats = np.zeros(4, dtype=np.int64)
    rc = lib().smo_update(_p(map_, _f64p), mh, mw, c, _p(pcd, _f64p), m, _p(label, _u8p), m, m, _p(colors, _u8p),
                          _p(cm, _f64p), float(boundary[0][0]), float(boundary[1][0]), float(resolution),
                          float(offset[0]), float(offset[1]), int(bool(use_intensity)), int(lane_index),
                          _p(stats, _i64p))
    if rc != 0:
        raise RuntimeError("smo_update failed: %d" % rc)
    return stats


def apply_filter(src):
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty_like(src)
    mh, mw, c = src.shape
    lib().smo_filter(_p(src, _f64p), mh, mw, c, _p(dst, _f64p))
    return dst


def render_bev_map(map_, colors):
    map_ = np.ascontiguousarray(map_, dtype=np.float64)
    colors = np.ascontiguousarray(np.asarray(colors).astype(np.uint8))
    mh, mw, c = map_.shape
    rgb = np.empty((mh, mw, 3), dtype=np.uint8)
    lib().smo_render(_p(map_, _f64p), mh, mw, c, _p(colors, _u8p), _p(rgb, _u8p))
    return rgb


def render_bev_map_with_thresholds(map_, colors, priority, thresholds):
    map_ = np.ascontiguousarray(map_, dtype=np.float64)
    colors = np.ascontiguousarray(np.asarray(colors).astype(np.uint8))
    mh, mw, c = map_.shape
    priority = np.ascontiguousarray(np.arange(c) if priority is None else priority, dtype=np.int32)
    thresholds = np.ascontiguousarray(thresholds, dtype=np.float64)
    assert len(priority) == c and len(thresholds) >= c
    rgb = np.empty((mh, mw, 3), dtype=np.uint8)
    lib().smo_render_thresholds(_p(map_, _f64p), mh, mw, c, _p(colors, _u8p), _p(priority, _i32p),
                                _p(thresholds, _f64p), _p(rgb, _u8p))
    return rgb


def np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().smo_np_sum(_p(a, _f64p), a.shape[0])
