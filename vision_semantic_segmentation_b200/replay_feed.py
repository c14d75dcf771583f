"""Host -> device feed of recorded frames (the transfer half of the replay record format).

The reference's replay loop hands one frame dictionary at a time to ``project_pcd`` / ``update_map``
(``src/mapping_replay.py:175-196``); every array is a pageable numpy array.  Here the same list is streamed to the GPU
so that the PCIe link, not the host, sets the pace:

* frames travel in batches of ``batch`` frames through ``depth`` sets of device slots (cloud + label image per slot);
  the copies of batch k + 1 are queued on a copy stream while batch k is integrated (``smap_integrate_batch``) on
  the caller's stream -- no per-frame synchronisation, the host only waits when every slot set is in flight;
* arrays that already live in pinned memory (``torch`` pinned tensors, numpy views of them, ``cudaHostRegister``-ed
  buffers) are copied from where they are; pageable arrays are first copied -- by a few threads -- into pinned staging
  buffers of the slot (a pageable ``cudaMemcpy`` would do the same on one thread and block);
* the reference's own cloud layout, (4, N) float64 (``src/mapping.py:178-180,309-312``), is copied as it is and
  converted to the fast kernel's float4 layout ON THE DEVICE (``smap_cloud_to_f32x4``), which also checks that every
  value is float32-representable (PointCloud2 fields are FLOAT32, so recorded clouds are); a cloud that is not keeps the
  float64 layout and the float64 kernel, so results never change.  The flags of a batch come back with one 4 x batch
  byte copy that the host reads when the batch's copies have landed -- by then the previous batch is still running;
* CUDA tensors are used where they are.

torch is the allocator and the stream / event provider here; all arithmetic is in the CUDA kernels behind the C ABI.
"""
import collections

import numpy as np

from . import _native

__all__ = ["FrameFeeder", "FeedItem"]

FeedItem = collections.namedtuple("FeedItem", "cloud image world_to_velodyne camera image_size")
FeedItem.__doc__ = """One frame for the feeder: ``cloud`` (N, 4) float32 or (4, N) float64 (numpy / torch, host or CUDA),
``image`` (H, W, 3) uint8 RGB label image or (h, w) uint8 class-id plane, ``world_to_velodyne`` 4x4 float64 or None,
``camera`` calibration object or slot, ``image_size`` (H, W) for class-id planes smaller than the camera image."""


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Slot(object):
    """Device (and, when needed, pinned host) buffers of one frame in flight."""
    __slots__ = ("pts32", "pts64", "img", "h_pts", "h_img", "refs")

    def __init__(self):
        self.pts32 = self.pts64 = self.img = self.h_pts = self.h_img = None
        self.refs = []


class _SlotSet(object):
    def __init__(self, torch, device, batch):
        self.slots = [_Slot() for _ in range(batch)]
        self.copied = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.in_flight = False
        with torch.cuda.device(device):
            self.flags_dev = torch.zeros(batch, dtype=torch.int32, device=device)
        self.flags_host = torch.zeros(batch, dtype=torch.int32).pin_memory()
        self.frames = []      # (slot index, FeedItem, kind, n_points) of the staged batch
        self.any64 = False


class FrameFeeder(object):
    def __init__(self, device_mapper, batch=8, depth=3, copy_threads=4):
        torch = _native.require_cuda()
        self.torch, self.dm = torch, device_mapper
        self.device = device_mapper.device
        self.batch, self.depth = int(batch), int(depth)
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream(device=self.device)
        self.sets = [_SlotSet(torch, self.device, self.batch) for _ in range(self.depth)]
        self._pool = None
        self._copy_threads = int(copy_threads)
        self.stats = {"frames": 0, "h2d_bytes": 0, "staged_bytes": 0, "converted_clouds": 0, "float64_clouds": 0}

    # ------------------------------------------------------------------ buffers
    def _dev_buffer(self, cur, nbytes):
        torch = self.torch
        if cur is None or cur.numel() < nbytes:
            with torch.cuda.device(self.device):
                cur = torch.empty(int(nbytes + nbytes // 8 + 256), dtype=torch.uint8, device=self.device)
        return cur

    def _pinned_buffer(self, cur, nbytes):
        torch = self.torch
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(int(nbytes + nbytes // 8 + 256), dtype=torch.uint8).pin_memory()
        return cur

    def _host_copy(self, dst_u8, src_np):
        """Pageable numpy array -> pinned staging bytes, split over a few threads (numpy releases the GIL)."""
        src = src_np.reshape(-1).view(np.uint8)
        dst = dst_u8.numpy()[:src.size]
        n = src.size
        if n < (4 << 20) or self._copy_threads <= 1:
            np.copyto(dst, src)
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(self._copy_threads)
        step = -(-n // self._copy_threads)
        step = (step + 4095) // 4096 * 4096
        futs = [self._pool.submit(np.copyto, dst[o:o + step], src[o:o + step]) for o in range(0, n, step)]
        for f in futs:
            f.result()

    def _to_device(self, slot, which, arr, nbytes):
        """Queue the H2D copy of a contiguous host array (numpy or CPU tensor) into the slot's device buffer `which`
        on the copy stream; returns the uint8 device view of exactly nbytes."""
        torch = self.torch
        dev = self._dev_buffer(getattr(slot, which), nbytes)
        setattr(slot, which, dev)
        if nbytes == 0:
            return dev[:0]
        host = arr if _is_torch(arr) else torch.from_numpy(arr)
        host = host.reshape(-1).view(torch.uint8)
        if not host.is_pinned():
            hname = "h_pts" if which in ("pts32", "pts64") else "h_img"
            stage = self._pinned_buffer(getattr(slot, hname), nbytes)
            setattr(slot, hname, stage)
            self._host_copy(stage, host.numpy())
            host = stage[:nbytes]
            self.stats["staged_bytes"] += nbytes
        else:
            slot.refs.append(arr)     # the caller's pinned memory must outlive the asynchronous copy
        dev[:nbytes].copy_(host, non_blocking=True)
        self.stats["h2d_bytes"] += nbytes
        return dev[:nbytes]

    # ------------------------------------------------------------------ staging
    def _stage(self, sset, items):
        """Queue the copies (and conversions) of one batch on the copy stream."""
        torch = self.torch
        if sset.in_flight:
            sset.done.synchronize()       # the batch that last used these slots has been integrated
            sset.in_flight = False
        sset.frames, sset.any64 = [], False
        with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
            for i, it in enumerate(items):
                slot = sset.slots[i]
                slot.refs = []
                cloud, image = it.cloud, it.image
                # ---- label image
                if _is_torch(image) and image.is_cuda:
                    img = image.to(self.device).contiguous()
                    if img.dtype != torch.uint8:
                        raise TypeError("label image must be uint8")
                    slot.refs.append(img)
                else:
                    if not _is_torch(image):
                        image = np.ascontiguousarray(image, dtype=np.uint8)
                    elif image.dtype != torch.uint8 or not image.is_contiguous():
                        image = image.to(torch.uint8).contiguous()
                    img = self._to_device(slot, "img", image, int(np.prod(image.shape))).view(tuple(image.shape))
                # ---- cloud
                if _is_torch(cloud) and cloud.is_cuda:
                    cloud = cloud.to(self.device)
                    if cloud.dtype == torch.float32:
                        kind, pts = 32, cloud.contiguous()
                    else:
                        pts = cloud.to(torch.float64)
                        if pts.dim() != 2 or pts.shape[0] < 4:
                            raise ValueError("float64 clouds must be (4, N)")
                        if pts.shape[1] and pts.stride(1) != 1:
                            pts = pts.contiguous()
                        kind = 64
                    slot.refs.append(pts)
                else:
                    if not _is_torch(cloud):
                        cloud = np.asarray(cloud)
                        is32 = cloud.dtype == np.float32 and cloud.ndim == 2 and cloud.shape[1] == 4
                        cloud = np.ascontiguousarray(cloud) if is32 else np.ascontiguousarray(cloud, dtype=np.float64)
                    else:
                        is32 = cloud.dtype == torch.float32 and cloud.dim() == 2 and cloud.shape[1] == 4
                        cloud = cloud.contiguous() if is32 else cloud.to(torch.float64).contiguous()
                    if is32:
                        kind = 32
                        n = cloud.shape[0]
                        pts = self._to_device(slot, "pts32", cloud, n * 16).view(torch.float32).view(n, 4)
                    else:
                        if len(cloud.shape) != 2 or cloud.shape[0] < 4:
                            raise ValueError("float64 clouds must be (4, N)")
                        kind = 64
                        rows, n = cloud.shape[0], cloud.shape[1]
                        pts = self._to_device(slot, "pts64", cloud, rows * n * 8).view(torch.float64).view(rows, n)
                if kind == 64:
                    # the reference's layout: convert on the device, keep the float64 cloud for the (rare) cloud that
                    # does not survive the round trip
                    n = pts.shape[1]
                    slot.pts32 = self._dev_buffer(slot.pts32, n * 16)
                    p32 = slot.pts32[:n * 16].view(torch.float32).view(n, 4)
                    sset.flags_dev[i:i + 1].zero_()
                    self.dm.cloud_to_f32x4(pts, p32, sset.flags_dev[i:i + 1])   # on the copy stream (current here)
                    sset.any64 = True
                    sset.frames.append((i, it, 64, (p32, pts), img))
                else:
                    sset.frames.append((i, it, 32, (pts, None), img))
            if sset.any64:
                sset.flags_host.copy_(sset.flags_dev, non_blocking=True)
            sset.copied.record(self.copy_stream)

    # ------------------------------------------------------------------ integration
    def _submit(self, sset):
        torch = self.torch
        dm = self.dm
        stream = torch.cuda.current_stream(self.device)
        if sset.any64:
            sset.copied.synchronize()     # the flags: the batch before this one is normally still being integrated
            flags = sset.flags_host.tolist()
        else:
            flags = None
        stream.wait_event(sset.copied)
        runs, cur_layout = [], None
        for i, it, kind, (p32, p64), img in sset.frames:
            use64 = kind == 64 and flags[i] != 0
            if use64 and img.dim() == 2:
                raise ValueError("class-id planes need a cloud of float32-representable values")
            if kind == 64:
                self.stats["float64_clouds" if use64 else "converted_clouds"] += 1
            pts = p64 if use64 else p32
            fr = dm.make_frame(pts, img, it.world_to_velodyne, it.camera, image_size=it.image_size)
            layout = 64 if use64 else 32
            if layout != cur_layout:
                runs.append([])
                cur_layout = layout
            runs[-1].append(fr)
        for run in runs:                  # frames of one smap_integrate_batch call share a point layout; order is kept
            dm.integrate_batch(run)
        sset.done.record(stream)
        sset.in_flight = True
        self.stats["frames"] += len(sset.frames)

    def run(self, items, after_batch=None):
        """Integrate every FeedItem of the iterable ``items`` into the mapper's grid, in order.  ``after_batch(n)`` is
        called with the number of frames integrated so far after every batch has been queued."""
        torch = self.torch
        staged = collections.deque()
        batch, k, total = [], 0, 0
        # copies may read tensors the caller produced on the current stream
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))

        def flush_batch():
            nonlocal k
            sset = self.sets[k % self.depth]
            k += 1
            self._stage(sset, batch)
            staged.append(sset)

        def submit_one():
            nonlocal total
            sset = staged.popleft()
            self._submit(sset)
            total += len(sset.frames)
            if after_batch is not None:
                after_batch(total)

        for it in items:
            batch.append(it)
            if len(batch) == self.batch:
                flush_batch()
                batch = []
                if len(staged) == self.depth - 1:   # keep depth - 1 batches of copies queued behind the running one
                    submit_one()
        if batch:
            flush_batch()
        while staged:
            submit_one()
        return total

    def drain(self):
        """Wait until nothing of the caller's (pinned) memory is being read any more."""
        for sset in self.sets:
            if sset.in_flight:
                sset.done.synchronize()
                sset.in_flight = False
            for s in sset.slots:
                s.refs = []
