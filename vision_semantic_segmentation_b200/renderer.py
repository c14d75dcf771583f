"""BEV map rendering on the GPU behind the reference's ``src/renderer.py`` functions.

``render_bev_map`` (``src/renderer.py:32-59``), ``render_bev_map_with_thresholds`` (``:131-172``) and
``apply_filter`` (``:175-189``) keep their names, argument meaning and error behaviour.  Arrays may be
numpy (host; uploaded, result returned as numpy) or CUDA ``torch`` tensors (result stays on the device).
``filter_and_render`` is the fused form of ``apply_filter`` followed by ``render_bev_map``
(``src/mapping_replay.py:198-200``) that never writes the filtered grid unless asked to.
"""
import ctypes

import numpy as np

from . import _native

__all__ = ["render_bev_map", "render_bev_map_with_thresholds", "apply_filter", "filter_and_render"]


def _as_device_map(map_):
    """-> (float64 contiguous CUDA tensor of shape (H, W, C), was_numpy)."""
    torch = _native.require_cuda()
    if isinstance(map_, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(map_, dtype=np.float64)).cuda()
        return t, True
    if not isinstance(map_, torch.Tensor):
        raise TypeError("map must be a numpy array or a torch tensor")
    if not map_.is_cuda:
        return map_.to(dtype=torch.float64).contiguous().cuda(), False
    return map_.to(dtype=torch.float64).contiguous(), False


def _colors_u8(label_colors):
    for c in label_colors:
        if len(c) != 3:
            raise ValueError("Color should be an RGB value.")
    return np.ascontiguousarray(np.asarray(label_colors).astype(np.uint8))


def _u8p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def render_bev_map(map, label_colors):
    """Colour every cell by its arg-max class; all-zero cells stay black."""
    assert len(map.shape) == 3
    colors = _colors_u8(label_colors)
    if map.shape[2] != len(colors):
        raise ValueError("Each channel should have a color!")
    dev, was_numpy = _as_device_map(map)
    torch = _native.require_cuda()
    mh, mw, c = dev.shape
    rgb = torch.empty((mh, mw, 3), dtype=torch.uint8, device=dev.device)
    with torch.cuda.device(dev.device):
        _native.check(_native.load().smap_render(dev.data_ptr(), mh, mw, c, _u8p(colors), rgb.data_ptr(),
                                                 dev.device.index, _native.current_stream_ptr(dev.device)))
    return rgb.cpu().numpy() if was_numpy else rgb


def render_bev_map_with_thresholds(map, label_colors, priority=None, thresholds=[0.01, 0.01, 0.01, 0.01, 0.01]):
    """Paint class ``priority[i]`` wherever its normalised evidence reaches ``thresholds[i]``;
    later entries of ``priority`` overwrite earlier ones."""
    assert len(map.shape) == 3
    colors = _colors_u8(label_colors)
    num_channels = map.shape[2]
    if num_channels != len(colors):
        raise ValueError("Each channel should have a color.")
    if priority is not None and num_channels != len(priority):
        raise ValueError("Each channel should have a priority.")
    if priority is None:
        priority = np.arange(num_channels)
    if len(thresholds) < num_channels:
        # the reference indexes thresholds[i] for every channel and dies with IndexError
        raise IndexError("list index out of range")
    priority = np.ascontiguousarray(priority, dtype=np.int32)
    thresholds = np.ascontiguousarray(np.asarray(thresholds, dtype=np.float64)[:num_channels])
    dev, was_numpy = _as_device_map(map)
    torch = _native.require_cuda()
    mh, mw, c = dev.shape
    rgb = torch.empty((mh, mw, 3), dtype=torch.uint8, device=dev.device)
    with torch.cuda.device(dev.device):
        _native.check(_native.load().smap_render_thresholds(
            dev.data_ptr(), mh, mw, c, _u8p(colors), _u8p(priority), _u8p(thresholds), rgb.data_ptr(),
            dev.device.index, _native.current_stream_ptr(dev.device)))
    return rgb.cpu().numpy() if was_numpy else rgb


def apply_filter(src):
    """3x3 box filter per class channel (``cv2.filter2D`` semantics, reflect-101 border)."""
    assert len(src.shape) == 3
    dev, was_numpy = _as_device_map(src)
    torch = _native.require_cuda()
    mh, mw, c = dev.shape
    dst = torch.empty_like(dev)
    with torch.cuda.device(dev.device):
        _native.check(_native.load().smap_apply_filter(dev.data_ptr(), mh, mw, c, dst.data_ptr(), dev.device.index,
                                                       _native.current_stream_ptr(dev.device)))
    return dst.cpu().numpy() if was_numpy else dst


def filter_and_render(map, label_colors, return_filtered=False):
    """``render_bev_map(apply_filter(map), label_colors)`` in one kernel."""
    assert len(map.shape) == 3
    colors = _colors_u8(label_colors)
    if map.shape[2] != len(colors):
        raise ValueError("Each channel should have a color!")
    dev, was_numpy = _as_device_map(map)
    torch = _native.require_cuda()
    mh, mw, c = dev.shape
    rgb = torch.empty((mh, mw, 3), dtype=torch.uint8, device=dev.device)
    filtered = torch.empty_like(dev) if return_filtered else None
    with torch.cuda.device(dev.device):
        _native.check(_native.load().smap_filter_render(
            dev.data_ptr(), mh, mw, c, _u8p(colors), rgb.data_ptr(),
            filtered.data_ptr() if filtered is not None else None, dev.device.index,
            _native.current_stream_ptr(dev.device)))
    if was_numpy:
        rgb = rgb.cpu().numpy()
        filtered = filtered.cpu().numpy() if filtered is not None else None
    return (rgb, filtered) if return_filtered else rgb
