"""ctypes binding of the C ABI in ``include/smap.h`` (``csrc/libsmap_b200.so``).

This is the only door between the Python host code and the CUDA kernels.  There is no
CPU fallback: if the library is missing, or no CUDA device is present when a device
call is made, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMAP_LIB_PATH") or os.path.join(_HERE, "csrc", "libsmap_b200.so")  # env: kernel-variant sweeps

ABI_VERSION = 5   # SMAP_ABI_VERSION of include/smap.h this binding was written against
SMAP_PTS_F32X4 = 0
SMAP_PTS_F64_SOA = 1
SMAP_IMG_RGB = 0
SMAP_IMG_CLASS_IDS = 1
SMAP_MAX_CLASSES = 31
SMAP_MAX_CAMERAS = 8

EXPORTS = [
    "smap_abi_version", "smap_last_error", "smap_device_count", "smap_device_info", "smap_create",
    "smap_destroy", "smap_set_camera", "smap_set_classes", "smap_set_label_palette", "smap_project", "smap_update", "smap_integrate",
    "smap_integrate_batch", "smap_integrate_host", "smap_apply_filter", "smap_render", "smap_filter_render",
    "smap_render_thresholds", "smap_eval_counts", "smap_map_ptr", "smap_clear", "smap_notify_map_modified", "smap_download", "smap_upload", "smap_get_stats", "smap_set_profiling", "smap_debug_set_frame_tag", "smap_debug_fast32",
    "smap_debug_nearest_map", "smap_cloud_to_f32x4",
    "smap_comm_unique_id", "smap_comm_init", "smap_comm_attach", "smap_comm_destroy", "smap_allreduce",
    "smap_reduce_scatter_rows", "smap_comm_get_info", "smap_comm_streaming", "smap_exchange_async", "smap_exchange_flush",
    "smap_clamp_negative", "smap_warp_perspective", "smap_hull_components", "smap_hull_row_extremes", "smap_debug_bounds",
]
SMAP_COMM_ID_BYTES = 128


class SmapConfig(ctypes.Structure):
    _fields_ = [
        ("map_height", ctypes.c_int32), ("map_width", ctypes.c_int32), ("num_classes", ctypes.c_int32),
        ("use_intensity", ctypes.c_int32), ("lane_index", ctypes.c_int32), ("device", ctypes.c_int32),
        ("boundary_x_min", ctypes.c_double), ("boundary_y_min", ctypes.c_double), ("resolution", ctypes.c_double),
        ("origin_offset_x", ctypes.c_double), ("origin_offset_y", ctypes.c_double), ("range_max", ctypes.c_double),
        ("map_dev", ctypes.c_void_p), ("map_is_zero", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class SmapFrame(ctypes.Structure):
    _fields_ = [
        ("points_dev", ctypes.c_void_p), ("n_points", ctypes.c_int64), ("ld", ctypes.c_int64),
        ("layout", ctypes.c_int32), ("camera", ctypes.c_int32), ("image_dev", ctypes.c_void_p),
        ("image_width", ctypes.c_int32), ("image_height", ctypes.c_int32), ("has_transform", ctypes.c_int32),
        ("image_format", ctypes.c_int32), ("world_to_velodyne", ctypes.c_double * 16),
        ("ids_width", ctypes.c_int32), ("ids_height", ctypes.c_int32),
    ]


class SmapStats(ctypes.Structure):
    _fields_ = [("frames", ctypes.c_int64), ("points", ctypes.c_int64), ("touched_cells", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64), ("profiled_frames", ctypes.c_int64),
                ("stream_kernel_ms", ctypes.c_double), ("apply_kernel_ms", ctypes.c_double)]


class SmapCommInfo(ctypes.Structure):
    _fields_ = [("n_ranks", ctypes.c_int32), ("rank", ctypes.c_int32), ("window", ctypes.c_int32 * 4),
                ("pack", ctypes.c_int32), ("reserved", ctypes.c_int32), ("bytes", ctypes.c_int64),
                ("grid_bytes", ctypes.c_int64), ("exchanges", ctypes.c_int64), ("pack_ms", ctypes.c_double),
                ("reduce_ms", ctypes.c_double), ("add_ms", ctypes.c_double), ("host_wait_ms", ctypes.c_double)]


class SmapError(RuntimeError):
    def __init__(self, code, message):
        super(SmapError, self).__init__("smap error %d: %s" % (code, message))
        self.code = code


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA extension %s not found; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vision_semantic_segmentation_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.smap_abi_version.restype = ctypes.c_int
    if L.smap_abi_version() != ABI_VERSION:
        raise RuntimeError("%s has C-ABI version %d, this package binds version %d: rebuild it (no CPU fallback)"
                           % (LIB_PATH, L.smap_abi_version(), ABI_VERSION))
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.smap_abi_version.restype = i32
    L.smap_abi_version.argtypes = []
    L.smap_last_error.restype = ctypes.c_char_p
    L.smap_last_error.argtypes = []
    L.smap_device_count.restype = i32
    L.smap_device_count.argtypes = []
    L.smap_device_info.restype = i32
    L.smap_device_info.argtypes = [i32, ctypes.c_char_p, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.smap_create.restype = i32
    L.smap_create.argtypes = [ctypes.POINTER(SmapConfig), ctypes.POINTER(vp)]
    L.smap_destroy.restype = i32
    L.smap_destroy.argtypes = [vp]
    L.smap_set_camera.restype = i32
    L.smap_set_camera.argtypes = [vp, i32, ctypes.POINTER(dbl)]
    L.smap_set_classes.restype = i32
    L.smap_set_classes.argtypes = [vp, vp, vp]
    L.smap_set_label_palette.restype = i32
    L.smap_set_label_palette.argtypes = [vp, vp, i32]
    L.smap_project.restype = i32
    L.smap_project.argtypes = [vp, ctypes.POINTER(SmapFrame), vp, vp, vp, vp, i64, ctypes.POINTER(i64), vp]
    L.smap_update.restype = i32
    L.smap_update.argtypes = [vp, vp, vp, i64, vp, i64, i64, vp]
    L.smap_integrate.restype = i32
    L.smap_integrate.argtypes = [vp, ctypes.POINTER(SmapFrame), vp]
    L.smap_integrate_batch.restype = i32
    L.smap_integrate_batch.argtypes = [vp, ctypes.POINTER(SmapFrame), i32, vp]
    L.smap_integrate_host.restype = i32
    L.smap_integrate_host.argtypes = [vp, ctypes.POINTER(SmapFrame), vp]
    L.smap_apply_filter.restype = i32
    L.smap_apply_filter.argtypes = [vp, i32, i32, i32, vp, i32, vp]
    L.smap_render.restype = i32
    L.smap_render.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp]
    L.smap_filter_render.restype = i32
    L.smap_filter_render.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp]
    L.smap_render_thresholds.restype = i32
    L.smap_render_thresholds.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    L.smap_eval_counts.restype = i32
    L.smap_eval_counts.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, vp]
    L.smap_map_ptr.restype = i32
    L.smap_map_ptr.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    L.smap_clear.restype = i32
    L.smap_clear.argtypes = [vp, vp]
    L.smap_set_profiling.restype = i32
    L.smap_set_profiling.argtypes = [vp, i32]
    L.smap_notify_map_modified.restype = i32
    L.smap_notify_map_modified.argtypes = [vp]
    L.smap_debug_set_frame_tag.restype = i32
    L.smap_debug_set_frame_tag.argtypes = [vp, ctypes.c_uint32]
    L.smap_debug_fast32.restype = i32
    L.smap_debug_fast32.argtypes = [ctypes.POINTER(SmapConfig), ctypes.POINTER(SmapFrame), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
    L.smap_debug_nearest_map.restype = i32
    L.smap_debug_nearest_map.argtypes = [i32, i32, vp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    L.smap_download.restype = i32
    L.smap_download.argtypes = [vp, vp]
    L.smap_upload.restype = i32
    L.smap_upload.argtypes = [vp, vp]
    L.smap_get_stats.restype = i32
    L.smap_get_stats.argtypes = [vp, ctypes.POINTER(SmapStats)]
    L.smap_cloud_to_f32x4.restype = i32
    L.smap_cloud_to_f32x4.argtypes = [vp, i64, i64, vp, vp, i32, vp]
    L.smap_comm_unique_id.restype = i32
    L.smap_comm_unique_id.argtypes = [vp]
    L.smap_comm_init.restype = i32
    L.smap_comm_init.argtypes = [vp, i32, i32, vp]
    L.smap_comm_attach.restype = i32
    L.smap_comm_attach.argtypes = [vp, vp]
    L.smap_comm_destroy.restype = i32
    L.smap_comm_destroy.argtypes = [vp]
    L.smap_allreduce.restype = i32
    L.smap_allreduce.argtypes = [vp, vp]
    L.smap_reduce_scatter_rows.restype = i32
    L.smap_reduce_scatter_rows.argtypes = [vp, vp, i64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                           ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp]
    L.smap_comm_streaming.restype = i32
    L.smap_comm_streaming.argtypes = [vp, i32, vp]
    L.smap_exchange_async.restype = i32
    L.smap_exchange_async.argtypes = [vp, vp]
    L.smap_exchange_flush.restype = i32
    L.smap_exchange_flush.argtypes = [vp, vp]
    L.smap_clamp_negative.restype = i32
    L.smap_clamp_negative.argtypes = [vp, i64, i32, vp]
    L.smap_hull_components.restype = i32
    L.smap_hull_components.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp]
    L.smap_hull_row_extremes.restype = i32
    L.smap_hull_row_extremes.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp]
    L.smap_debug_bounds.restype = i32
    L.smap_debug_bounds.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    L.smap_warp_perspective.restype = i32
    L.smap_warp_perspective.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_double), vp, i32, i32, i32, vp]
    L.smap_comm_get_info.restype = i32
    L.smap_comm_get_info.argtypes = [vp, ctypes.POINTER(SmapCommInfo)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SmapError(rc, (load().smap_last_error() or b"").decode("utf-8", "replace"))
    return rc


def require_cuda():
    """torch supplies device memory and streams; the kernels need a CUDA device. No fallback."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("vision_semantic_segmentation_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback for the mapping path.")
    return torch


def current_stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
