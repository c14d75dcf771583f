"""Handle on the device-resident BEV grid: thin object layer over the C ABI (``include/smap.h``).

``DeviceMapper`` owns one ``smap_handle`` on one GPU.  The grid itself is a ``torch`` CUDA tensor
(torch is the allocator and the stream provider; ``torch.distributed`` reduces this tensor across
ranks) whose pointer is lent to the handle.  All arithmetic happens in the CUDA kernels.
"""
import ctypes

import numpy as np

from . import _native
from ._native import SmapConfig, SmapFrame, SmapStats, SMAP_PTS_F32X4, SMAP_PTS_F64_SOA

PCD_ORIGIN_OFFSET = (1369.0496826171875, 562.84814453125)  # reference src/mapping_replay.py:261

_IDENTITY16 = (ctypes.c_double * 16)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)


class DeviceMapper(object):
    def __init__(self, map_height, map_width, label_colors, update_matrix, boundary, resolution, range_max,
                 use_intensity, lane_index, cameras, device=None, origin_offset=PCD_ORIGIN_OFFSET):
        torch = _native.require_cuda()
        self._lib = _native.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        colors = np.ascontiguousarray(np.asarray(label_colors).astype(np.uint8))
        cm = np.ascontiguousarray(update_matrix, dtype=np.float64)
        c = colors.shape[0]
        if colors.shape != (c, 3) or cm.shape != (c, c):
            raise ValueError("label_colors must be (C,3) and the update matrix (C,C)")
        if not 1 <= c <= _native.SMAP_MAX_CLASSES:
            raise ValueError("between 1 and %d classes are supported" % _native.SMAP_MAX_CLASSES)
        self.map_height, self.map_width, self.num_classes = int(map_height), int(map_width), c
        with torch.cuda.device(self.device):
            self.map = torch.zeros((self.map_height, self.map_width, c), dtype=torch.float64, device=self.device)
        cfg = SmapConfig()
        cfg.map_height, cfg.map_width, cfg.num_classes = self.map_height, self.map_width, c
        cfg.use_intensity = int(bool(use_intensity))
        cfg.lane_index = int(lane_index)
        cfg.device = self.device.index
        cfg.boundary_x_min, cfg.boundary_y_min = float(boundary[0][0]), float(boundary[1][0])
        cfg.resolution = float(resolution)
        cfg.origin_offset_x, cfg.origin_offset_y = float(origin_offset[0]), float(origin_offset[1])
        cfg.range_max = float(range_max)
        cfg.map_dev = self.map.data_ptr()
        cfg.map_is_zero = 1   # torch.zeros above: the handle may use the atomic count update (include/smap.h)
        handle = ctypes.c_void_p()
        _native.check(self._lib.smap_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._h = handle
        self._colors, self._cm = colors, cm
        _native.check(self._lib.smap_set_classes(self._h, colors.ctypes.data_as(ctypes.c_void_p),
                                                 cm.ctypes.data_as(ctypes.c_void_p)))
        self._camera_slots = {}
        for slot, cam in enumerate(cameras):
            P = np.ascontiguousarray(cam.P if hasattr(cam, "P") else cam, dtype=np.float64)
            if P.shape != (3, 4):
                raise ValueError("camera projection must be 3x4")
            _native.check(self._lib.smap_set_camera(self._h, slot, P.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
            self._camera_slots[id(cam)] = slot
        self._keepalive = []

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.smap_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return _native.current_stream_ptr(self.device)

    def camera_slot(self, camera):
        if isinstance(camera, int):
            return camera
        try:
            return self._camera_slots[id(camera)]
        except KeyError:
            raise ValueError("unknown camera calibration object; pass it in `cameras` at construction")

    def set_label_palette(self, labels):
        """Palette of the segmentation network (``labels`` of the dataset config, or an (n, 3) colour array): needed
        before frames carry class-id planes instead of RGB label images (``label_image.py``, SMAP_IMG_CLASS_IDS)."""
        from .label_image import palette_of
        pal = palette_of(labels)
        _native.check(self._lib.smap_set_label_palette(self._h, pal.ctypes.data_as(ctypes.c_void_p), pal.shape[0]))
        self._palette = pal

    def make_frame(self, points, image, world_to_velodyne, camera=0, host=False, image_size=None):
        """Describe one frame for the C ABI.  ``points``: torch tensor (N,4) float32 [float4 layout] or
        (4,N) float64 [the reference's pcd layout]; ``image``: (H,W,3) uint8 RGB label image, or an (h,w) uint8
        class-id plane (the network's output; ``image_size=(H, W)`` is then the camera resolution it would be
        upscaled to, default its own shape; float32 clouds only); ``world_to_velodyne``: 4x4
        float64 numpy or None for a cloud already in the velodyne frame.  host=True: tensors live in
        (pinned) host memory and are meant for ``integrate_host``."""
        torch = _native.require_cuda()
        f = SmapFrame()
        if points.dtype == torch.float32:
            if points.dim() != 2 or points.shape[1] != 4 or not points.is_contiguous():
                raise ValueError("float32 clouds must be contiguous (N, 4)")
            f.layout, f.n_points, f.ld = SMAP_PTS_F32X4, points.shape[0], 0
        elif points.dtype == torch.float64:
            if points.dim() != 2 or points.shape[0] < 4 or points.stride(1) != 1:
                raise ValueError("float64 clouds must be (4, N) with unit column stride")
            f.layout, f.n_points, f.ld = SMAP_PTS_F64_SOA, points.shape[1], points.stride(0)
            if points.shape[1] == 0:
                f.ld = 0
        else:
            raise TypeError("cloud must be float32 (N,4) or float64 (4,N)")
        if image.dtype == torch.uint8 and image.dim() == 2 and image.is_contiguous():
            if getattr(self, "_palette", None) is None:
                raise ValueError("class-id planes need the network's palette: call set_label_palette first")
            if f.layout != SMAP_PTS_F32X4:
                raise ValueError("class-id planes go with float32 (N, 4) clouds")
            f.image_format = _native.SMAP_IMG_CLASS_IDS
            f.ids_height, f.ids_width = image.shape[0], image.shape[1]
            f.image_height, f.image_width = (image.shape[0], image.shape[1]) if image_size is None else \
                (int(image_size[0]), int(image_size[1]))
        elif image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3 or not image.is_contiguous():
            raise ValueError("label image must be contiguous (H, W, 3) uint8, or an (h, w) uint8 class-id plane")
        else:
            f.image_format = _native.SMAP_IMG_RGB
            f.image_height, f.image_width = image.shape[0], image.shape[1]
        if not host and (not points.is_cuda or not image.is_cuda):
            raise ValueError("device frames need CUDA tensors")
        f.points_dev = points.data_ptr()
        f.image_dev = image.data_ptr()
        f.camera = self.camera_slot(camera)
        if world_to_velodyne is None:
            f.has_transform = 0
            ctypes.memmove(f.world_to_velodyne, _IDENTITY16, 128)
        else:
            T = np.ascontiguousarray(world_to_velodyne, dtype=np.float64)
            if T.shape != (4, 4):
                raise ValueError("world_to_velodyne must be 4x4")
            f.has_transform = 1
            ctypes.memmove(f.world_to_velodyne, T.ctypes.data, 128)
        f._tensors = (points, image)   # the struct holds raw pointers: keep the tensors alive as long as it lives
        return f

    # ------------------------------------------------------------------ the path
    def integrate(self, frame):
        """project_pcd + update_map of one frame, fused (bit-exact for any update matrix)."""
        _native.check(self._lib.smap_integrate(self._h, ctypes.byref(frame), self._stream()))

    def integrate_host(self, frame):
        """``smap_integrate_host``: the frame's tensors live in (pinned) host memory and the copies are asynchronous
        (torch's caching host allocator knows nothing about copies issued by the library), so the frame stays
        referenced here until an event recorded behind its kernels has completed."""
        torch = _native.require_cuda()
        _native.check(self._lib.smap_integrate_host(self._h, ctypes.byref(frame), self._stream()))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._keepalive.append((ev, frame))
        while len(self._keepalive) > 2 and self._keepalive[0][0].query():
            del self._keepalive[0]

    def integrate_batch(self, frames):
        arr = (SmapFrame * len(frames))(*frames)
        _native.check(self._lib.smap_integrate_batch(self._h, arr, len(frames), self._stream()))

    def project(self, frame, want_uv=False, want_keep=False):
        """Parity API of project_pcd: returns (masked_pcd (4,M) f64, label (3,M) u8[, uv (2,M) i32][, keep (N,) bool])."""
        torch = _native.require_cuda()
        n = int(frame.n_points)
        with torch.cuda.device(self.device):
            out_pcd = torch.empty((4, n), dtype=torch.float64, device=self.device)
            out_label = torch.empty((3, n), dtype=torch.uint8, device=self.device)
            out_uv = torch.empty((2, n), dtype=torch.int32, device=self.device) if want_uv else None
            out_keep = torch.empty((n,), dtype=torch.uint8, device=self.device) if want_keep else None
            m = ctypes.c_int64(0)
            _native.check(self._lib.smap_project(
                self._h, ctypes.byref(frame), out_pcd.data_ptr(), out_label.data_ptr(),
                out_uv.data_ptr() if want_uv else None, out_keep.data_ptr() if want_keep else None,
                n, ctypes.byref(m), self._stream()))
        m = m.value
        res = [out_pcd[:, :m].contiguous(), out_label[:, :m].contiguous()]
        if want_uv:
            res.append(out_uv[:, :m].contiguous())
        if want_keep:
            res.append(out_keep.bool())
        return tuple(res)

    def update(self, pcd, label, map_tensor=None):
        """Parity API of update_map on CUDA tensors: pcd (4,M) float64, label (3,M) uint8."""
        torch = _native.require_cuda()
        if pcd.dtype != torch.float64 or label.dtype != torch.uint8 or not pcd.is_cuda or not label.is_cuda:
            raise TypeError("update needs CUDA float64 pcd and uint8 label")
        if pcd.shape[0] < 4 or label.shape[0] != 3 or pcd.shape[1] != label.shape[1]:
            raise ValueError("pcd must be (4,M) and label (3,M)")
        m = pcd.shape[1]
        if m and (pcd.stride(1) != 1 or label.stride(1) != 1):
            pcd, label = pcd.contiguous(), label.contiguous()
        target = self.map if map_tensor is None else map_tensor
        if target.shape != self.map.shape or target.dtype != torch.float64 or not target.is_contiguous():
            raise ValueError("map must be a contiguous float64 tensor of shape %s" % (tuple(self.map.shape),))
        _native.check(self._lib.smap_update(self._h, target.data_ptr(), pcd.data_ptr(), pcd.stride(0) if m else 0,
                                            label.data_ptr(), label.stride(0) if m else 0, m, self._stream()))
        return target

    # ------------------------------------------------------------------ multi-GPU
    def init_comm(self, group=None):
        """Collective over a ``torch.distributed`` process group: gives the handle its own NCCL communicator.  Rank 0
        draws the unique id (``smap_comm_unique_id``), ``torch.distributed`` carries its 128 bytes to the other ranks
        (plumbing), every rank calls ``smap_comm_init``.  After this ``allreduce`` / ``reduce_scatter_rows`` run
        entirely under the C ABI (pack kernels + NCCL)."""
        import torch.distributed as dist
        if getattr(self, "_comm_ready", False):
            return
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = (ctypes.c_uint8 * _native.SMAP_COMM_ID_BYTES)()
        if rank == 0:
            _native.check(self._lib.smap_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (ctypes.c_uint8 * _native.SMAP_COMM_ID_BYTES).from_buffer_copy(box[0])
        _native.check(self._lib.smap_comm_init(self._h, world, rank, ident))
        self._comm_ready = True

    def allreduce(self):
        """Every rank's grid becomes the sum over the ranks (``smap_allreduce``: touched window only, counts packed)."""
        _native.check(self._lib.smap_allreduce(self._h, self._stream()))

    def reduce_scatter_rows(self):
        """``smap_reduce_scatter_rows``: returns ``(tile, r0, r1, top, bottom)``; ``tile`` holds the summed rows
        ``[r0 - top, r1 + bottom)`` of the map as a (rows, MW, C) float64 CUDA tensor."""
        torch = _native.require_cuda()
        info = self.comm_info()
        per = -(-self.map_height // max(info["n_ranks"], 1))
        with torch.cuda.device(self.device):
            tile = torch.empty((per + 2, self.map_width, self.num_classes), dtype=torch.float64, device=self.device)
        r0, r1, top, bottom = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32())
        _native.check(self._lib.smap_reduce_scatter_rows(self._h, tile.data_ptr(), per + 2, ctypes.byref(r0), ctypes.byref(r1),
                                                         ctypes.byref(top), ctypes.byref(bottom), self._stream()))
        rows = r1.value - r0.value + top.value + bottom.value
        return tile[:rows], r0.value, r1.value, top.value, bottom.value

    def set_streaming(self, on=True):
        """``smap_comm_streaming``: integrate into buffers of local increments that ``exchange_async`` sums into every
        rank's grid while the next frames are integrated (``include/smap.h``)."""
        _native.check(self._lib.smap_comm_streaming(self._h, int(bool(on)), self._stream()))

    def exchange_async(self):
        """Collective (same call sequence on every rank): hand the increments integrated since the previous call to
        the other ranks; returns without waiting for the collective (``smap_exchange_async``)."""
        _native.check(self._lib.smap_exchange_async(self._h, self._stream()))

    def exchange_flush(self):
        """The current stream waits until every exchange queued so far has been added to the grid."""
        _native.check(self._lib.smap_exchange_flush(self._h, self._stream()))

    def comm_info(self):
        info = _native.SmapCommInfo()
        _native.check(self._lib.smap_comm_get_info(self._h, ctypes.byref(info)))
        return {"n_ranks": info.n_ranks, "rank": info.rank, "window": list(info.window), "pack": ("u16", "u32", "f64")[info.pack],
                "bytes": info.bytes, "grid_bytes": info.grid_bytes, "exchanges": info.exchanges,
                "pack_ms": info.pack_ms, "reduce_ms": info.reduce_ms, "add_ms": info.add_ms, "host_wait_ms": info.host_wait_ms}

    def cloud_to_f32x4(self, pcd, out, flag):
        """(4, N) float64 CUDA cloud -> (N, 4) float32 ``out``; ``flag`` (int32 CUDA tensor of one element, zeroed by
        the caller) is raised when a value is not float32-representable (``smap_cloud_to_f32x4``)."""
        _native.check(self._lib.smap_cloud_to_f32x4(pcd.data_ptr(), pcd.stride(0) if pcd.shape[1] else 0, pcd.shape[1],
                                                    out.data_ptr(), flag.data_ptr(), self.device.index, self._stream()))

    def notify_map_modified(self):
        """Call after writing into ``self.map`` from outside (``copy_``, arithmetic, ...): the grid may no longer hold
        integer-valued counts, so the handle switches from the atomic count update to the ordered update."""
        _native.check(self._lib.smap_notify_map_modified(self._h))

    def clear(self):
        _native.check(self._lib.smap_clear(self._h, self._stream()))

    def stats(self):
        s = SmapStats()
        _native.check(self._lib.smap_get_stats(self._h, ctypes.byref(s)))
        return {"frames": s.frames, "points": s.points, "touched_cells": s.touched_cells,
                "kernel_launches": s.kernel_launches, "profiled_frames": s.profiled_frames,
                "stream_kernel_ms": s.stream_kernel_ms, "apply_kernel_ms": s.apply_kernel_ms}

    def set_profiling(self, on=True):
        """Time the streaming / apply kernels with CUDA events (see include/smap.h); totals come back in stats()."""
        _native.check(self._lib.smap_set_profiling(self._h, int(bool(on))))
