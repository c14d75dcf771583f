"""Convex hulls of the largest connected regions of one class: the reference's ``src/semantic_convex_hull.py`` on the GPU.

``generate_convex_hull(img_src, vis=False, index_care_about=1, index_to_vitualize=None, top_number=1,
area_threshold=30)`` keeps the reference's signature, defaults and result (``src/semantic_convex_hull.py:17-91``): a list
of ``(2, V + 1)`` integer arrays, one closed polygon (x row, y row) per selected component.

What touches every pixel runs as CUDA kernels behind the C ABI (``smap_hull_components``: class mask, 3 x 3 erosion,
8-connected component labelling, component areas; ``smap_hull_row_extremes``: the first and last column of a component
in every image row).  The host keeps what is tiny: the ranking of the components (``collections.Counter.most_common``
order: largest area first, ties in raster order of the first pixel) and ``cv2.convexHull`` of the at most 2 H row
extremes of a component -- the same polygon OpenCV returns for all of its pixels, vertex order included (checked on the
reference itself: ``oracle/make_golden_hull.py``, ``tests/test_convex_hull.py``).  OpenCV is the reference's own
dependency; there is no CPU fallback for the per-pixel part.
"""
import ctypes

import numpy as np

from . import _native


def _components(img_dev, index):
    torch = _native.require_cuda()
    h, w = int(img_dev.shape[0]), int(img_dev.shape[1])
    scratch = torch.empty((h, w), dtype=torch.uint8, device=img_dev.device)
    labels = torch.empty((h, w), dtype=torch.int32, device=img_dev.device)
    areas = torch.empty((h, w), dtype=torch.int32, device=img_dev.device)
    with torch.cuda.device(img_dev.device):
        _native.check(_native.load().smap_hull_components(
            ctypes.c_void_p(img_dev.data_ptr()), h, w, int(index), ctypes.c_void_p(scratch.data_ptr()),
            ctypes.c_void_p(labels.data_ptr()), ctypes.c_void_p(areas.data_ptr()), img_dev.device.index,
            _native.current_stream_ptr(img_dev.device)))
    return labels, areas


def label_components(img_src, index_care_about=1):
    """The eroded class mask's 8-connected components, numbered 1.. in raster order of their first pixel as
    ``skimage.measure.label`` numbers them (0 = background); returns (labels (h, w) int32 numpy, areas per label)."""
    torch = _native.require_cuda()
    dev = _to_device_u8(img_src)
    labels, areas = _components(dev, index_care_about)
    flat = labels.view(-1)
    roots = torch.nonzero(flat == torch.arange(flat.numel(), device=flat.device, dtype=torch.int32)).view(-1)
    number = torch.zeros(flat.numel() + 1, dtype=torch.int32, device=flat.device)
    number[roots] = torch.arange(1, roots.numel() + 1, dtype=torch.int32, device=flat.device)
    out = number[flat.long()]            # index -1 (background) -> the spare last entry, 0
    return out.view(labels.shape).cpu().numpy(), areas.view(-1)[roots].cpu().numpy()


def _raster_ordered_extremes(lo, hi):
    """(x, y) of the row extremes in RASTER order (row by row, smallest column first, a row's single pixel once).
    The order matters: cv2.convexHull finishes by rotating its output so that the hull's indices into the INPUT array
    ascend, and the reference hands it the component's pixels in raster order -- with the extremes in raster order any
    two hull vertices compare as they do there, so the polygon starts at the same vertex."""
    ys = np.nonzero(hi >= 0)[0]
    both = np.stack([np.stack([lo[ys], ys], 1), np.stack([hi[ys], ys], 1)], 1).reshape(-1, 2)
    keep = np.ones(len(both), bool)
    keep[1::2] = lo[ys] != hi[ys]
    return both[keep].astype(np.int32)


def _to_device_u8(img_src):
    torch = _native.require_cuda()
    if isinstance(img_src, np.ndarray):
        if img_src.dtype != np.uint8:
            # the reference hands the array to cv2.erode, which rejects int32 / int64 label planes with cv2.error
            raise ValueError("generate_convex_hull takes a uint8 label image (cv2.erode rejects %s in the reference)" % img_src.dtype)
        return torch.from_numpy(np.ascontiguousarray(img_src)).cuda()
    if not img_src.is_cuda or img_src.dtype != torch.uint8:
        raise ValueError("img_src must be a uint8 numpy array or a uint8 CUDA tensor")
    return img_src.contiguous()


def generate_convex_hull(img_src, vis=False, index_care_about=1, index_to_vitualize=None, top_number=1, area_threshold=30):
    """
        Generate the convex hull
        Args:
            img: input img (h, w) uint8 label image (numpy array or CUDA tensor)
            vis: the reference drew matplotlib figures; not available here (raises)
            index_care_about: index that will be used to generate the convex hull
            index_to_vitualize: component numbers to use (raster order of the first pixel, from 1) instead of the largest
            top_number: the number most common label decided to choose
            area_threshold: only consider the connected component which contains the points greater than this area_threshold
        Returns:
            vertices: extracted vertices; list of numpy arrays; array shape- -- [2, number of vertices]
    """
    try:
        import cv2
    except ImportError as e:
        raise ImportError("generate_convex_hull needs OpenCV on the host for cv2.convexHull (opencv-python)") from e
    torch = _native.require_cuda()
    if vis:
        raise NotImplementedError("vis=True draws matplotlib figures in the reference; not available here")
    if len(img_src.shape) != 2:
        raise ValueError("not enough values to unpack" if len(img_src.shape) < 2 else "too many values to unpack (expected 2)")
    rows, cols = int(img_src.shape[0]), int(img_src.shape[1])
    if index_care_about == 0:
        raise SystemExit(0)   # the reference logs an error and calls exit(0) (:33-35)
    dev = _to_device_u8(img_src)
    labels, areas = _components(dev, index_care_about)
    flat = labels.view(-1)
    roots = torch.nonzero(flat == torch.arange(flat.numel(), device=flat.device, dtype=torch.int32)).view(-1)
    if roots.numel() == 0:
        return []
    roots_h = roots.cpu().numpy()
    areas_h = areas.view(-1)[roots].cpu().numpy()
    if index_to_vitualize is None:
        # Counter(...).most_common(top_number): by count, ties in order of first appearance = raster order of the first pixel
        order = sorted(range(len(roots_h)), key=lambda k: -int(areas_h[k]))[:top_number]
        chosen = [k for k in order if areas_h[k] > area_threshold]
    else:
        chosen = [int(k) - 1 for k in index_to_vitualize]
    vertices = []
    rowmin = torch.empty(rows, dtype=torch.int32, device=dev.device)
    rowmax = torch.empty(rows, dtype=torch.int32, device=dev.device)
    lib = _native.load()
    for k in chosen:
        if k < 0 or k >= len(roots_h):
            # np.concatenate of an empty list in the reference (:67)
            raise ValueError("need at least one array to concatenate")
        with torch.cuda.device(dev.device):
            _native.check(lib.smap_hull_row_extremes(
                ctypes.c_void_p(labels.data_ptr()), rows, cols, int(roots_h[k]), ctypes.c_void_p(rowmin.data_ptr()),
                ctypes.c_void_p(rowmax.data_ptr()), dev.device.index, _native.current_stream_ptr(dev.device)))
        pts = _raster_ordered_extremes(rowmin.cpu().numpy(), rowmax.cpu().numpy())
        hull = cv2.convexHull(pts)                                  # a component of one pixel: an empty array, cv2.error, as in the reference
        nodes = np.concatenate([np.squeeze(hull), hull[0, :, :].reshape(1, -1)], axis=0).T     # (:74-75)
        vertices.append(nodes)
    return vertices
