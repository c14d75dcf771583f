"""Planar projection of the semantic image onto the map plane: the reference's ``src/homography.py`` on the GPU.

``generate_homography(im_src, pts_src, pts_dst, vis=False, out_size=None, return_h=False)`` keeps the reference's
signature and behaviour (``src/homography.py:22-76``): the 3 x 3 homography comes from ``cv2.findHomography`` on the host
(four anchor correspondences, host-side 3 x 3 work like the 4 x 4 pose matrices of the point path; OpenCV is a declared
dependency of the reference), the warp -- ``cv2.warpPerspective(im_src, h, (W, H))``, the part that touches every pixel of
the (MH, MW) map image -- runs as a CUDA kernel (``smap_warp_perspective``) that reproduces OpenCV's 8-bit INTER_LINEAR
result bit for bit (``oracle/warp_port.py`` is the restatement it is tested against; there is no CPU fallback).

Caller: ``update_map_planar`` (``src/mapping.py:446-488``).  In the reference the warped image is compared with the label
NAMES and therefore never updates a cell; :meth:`mapping.SemanticMapping.update_map_planar` keeps that
observable behaviour, and this module provides the projection step for whoever wants the warped image itself.
"""
import ctypes

import numpy as np

from . import _native


def find_homography(pts_src, pts_dst):
    """``cv2.findHomography(pts_src, pts_dst)[0]`` (host, 3 x 3 float64)."""
    try:
        import cv2
    except ImportError as e:   # the reference's own dependency (requirements.txt: opencv-python)
        raise ImportError("generate_homography needs OpenCV on the host for cv2.findHomography (opencv-python)") from e
    h, _ = cv2.findHomography(np.asarray(pts_src), np.asarray(pts_dst))
    return h


def warp_perspective(image, h, dsize):
    """``cv2.warpPerspective(image, h, dsize)`` (INTER_LINEAR, constant border 0) on the GPU.

    image: (H, W) or (H, W, C <= 4) uint8, numpy array or CUDA tensor; dsize = (width, height) as in OpenCV.
    Returns the same kind of array it was given."""
    torch = _native.require_cuda()
    was_numpy = isinstance(image, np.ndarray)
    if was_numpy:
        if image.dtype != np.uint8:
            raise ValueError("warp_perspective takes 8-bit images (the semantic label image)")
        dev = torch.from_numpy(np.ascontiguousarray(image)).cuda()
    else:
        dev = image
        if not dev.is_cuda or dev.dtype != torch.uint8:
            raise ValueError("image must be a uint8 numpy array or a uint8 CUDA tensor")
        dev = dev.contiguous()
    if dev.dim() not in (2, 3) or (dev.dim() == 3 and not 1 <= dev.shape[2] <= 4):
        raise ValueError("image must be (H, W) or (H, W, C) with 1 <= C <= 4")
    cn = 1 if dev.dim() == 2 else int(dev.shape[2])
    width, height = int(dsize[0]), int(dsize[1])
    shape = (height, width) if dev.dim() == 2 else (height, width, cn)
    out = torch.empty(shape, dtype=torch.uint8, device=dev.device)
    hm = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(3, 3))
    with torch.cuda.device(dev.device):
        _native.check(_native.load().smap_warp_perspective(
            ctypes.c_void_p(dev.data_ptr()), int(dev.shape[0]), int(dev.shape[1]), cn,
            hm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_void_p(out.data_ptr()), height, width,
            dev.device.index, _native.current_stream_ptr(dev.device)))
    return out.cpu().numpy() if was_numpy else out


def generate_homography(im_src, pts_src, pts_dst, vis=False, out_size=None, return_h=False):
    """Homography-transformed image from point correspondences (``src/homography.py:22-76``).

    pts_src, pts_dst: n x 2 arrays; out_size: [width, height] of the output (default: the size of ``im_src``).
    ``vis=True`` opened OpenCV windows in the reference; there is no display on this path, it raises."""
    assert (len(pts_src[1]) == 2)
    assert (len(pts_dst[1]) == 2)
    if vis:
        raise NotImplementedError("vis=True draws into cv2.imshow windows in the reference; not available here")
    h = find_homography(pts_src, pts_dst)
    if out_size is None:
        im_dst = warp_perspective(im_src, h, (im_src.shape[1], im_src.shape[0]))
    else:
        im_dst = warp_perspective(im_src, h, (out_size[0], out_size[1]))
    if return_h:
        return im_dst, h
    return im_dst
