"""TEST INFRASTRUCTURE (oracle): numpy restatement of the planar-projection warp of the reference,
``generate_homography`` -> ``cv2.warpPerspective(im_src, h, (W, H))`` (/root/reference/src/homography.py:39,53-55,
called from ``update_map_planar``, src/mapping.py:465-466).

The arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV (``opencv-python``, unpinned in
requirements.txt:3; 4.13.0 in this image).  Its published algorithm for 8-bit images, INTER_LINEAR, BORDER_CONSTANT(0):

* ``M = inv(h)`` by the closed 3 x 3 adjugate formula (``cv::invert``, double);
* destination blocks of ``bw0 x bh0`` pixels (64 x 16 for images at least that large); per row of a block
  ``X0 = M0 bx + M1 y + M2`` (``bx`` = first column of the block) and per pixel ``W = W0 + M6 x1``,
  ``W = W ? 32 / W : 0``, ``fX = max(INT_MIN, min(INT_MAX, (X0 + M0 x1) W))``, ``X = cvRound(fX)`` -- source
  coordinates in fixed point with 5 fractional bits; integer part saturated to int16;
* bilinear taps with the fixed-point table ``(32 - ax)(32 - ay) 32`` ... (sums to 2^15 exactly, so OpenCV's table fix-up
  never fires), taps outside the source read the border value 0, result ``(acc + 2^14) >> 15``.

Pinned: ``tests/test_warp.py`` holds this file to ``cv2.warpPerspective`` itself wherever cv2 imports, and to the
committed vectors of ``oracle/make_golden_warp.py`` (generated with the cv2 of this image).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import anything under oracle/.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


def invert3(h):
    """cv::invert of a 3 x 3 double matrix (direct formula; checked bit for bit against cv2.invert)."""
    s = np.asarray(h, dtype=np.float64).reshape(3, 3)
    d = (s[0, 0] * (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) - s[0, 1] * (s[1, 0] * s[2, 2] - s[1, 2] * s[2, 0])
         + s[0, 2] * (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]))
    if d == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty(9)
    t[0] = (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) * d
    t[1] = (s[0, 2] * s[2, 1] - s[0, 1] * s[2, 2]) * d
    t[2] = (s[0, 1] * s[1, 2] - s[0, 2] * s[1, 1]) * d
    t[3] = (s[1, 2] * s[2, 0] - s[1, 0] * s[2, 2]) * d
    t[4] = (s[0, 0] * s[2, 2] - s[0, 2] * s[2, 0]) * d
    t[5] = (s[0, 2] * s[1, 0] - s[0, 0] * s[1, 2]) * d
    t[6] = (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]) * d
    t[7] = (s[0, 1] * s[2, 0] - s[0, 0] * s[2, 1]) * d
    t[8] = (s[0, 0] * s[1, 1] - s[0, 1] * s[1, 0]) * d
    return t.reshape(3, 3)


def block_size(width, height):
    """WarpPerspectiveInvoker's block: BLOCK_SZ = 32, bh0 = min(16, h), bw0 = min(1024 / bh0, w), bh0 = min(1024 / bw0, h)."""
    bh0 = min(16, height)
    bw0 = min(1024 // bh0, width)
    bh0 = min(1024 // bw0, height)
    return bw0, bh0


def _std_min_max(v):
    """std::max((double)INT_MIN, std::min((double)INT_MAX, v)) with the C++ comparison semantics (NaN -> INT_MAX)."""
    t = np.where(v < 2147483647.0, v, 2147483647.0)
    return np.where(-2147483648.0 < t, t, -2147483648.0)


def warp_perspective(img, h, dsize):
    """cv2.warpPerspective(img, h, dsize) for uint8 images of 1..4 channels (INTER_LINEAR, BORDER_CONSTANT 0)."""
    width, height = int(dsize[0]), int(dsize[1])
    m = invert3(h).ravel()
    sh, sw = img.shape[:2]
    cn = img.shape[2] if img.ndim == 3 else 1
    src = np.ascontiguousarray(img).reshape(sh, sw, cn).astype(np.int64)
    bw0, _ = block_size(width, height)
    ys, xs = np.mgrid[0:height, 0:width]
    bx = (xs // bw0) * bw0
    x1 = (xs - bx).astype(np.float64)
    bx = bx.astype(np.float64)
    ys = ys.astype(np.float64)
    x0 = m[0] * bx + m[1] * ys + m[2]
    y0 = m[3] * bx + m[4] * ys + m[5]
    w0 = m[6] * bx + m[7] * ys + m[8]
    with np.errstate(all="ignore"):
        w = w0 + m[6] * x1
        w = np.where(w != 0, INTER_TAB_SIZE / w, 0.0)
        fx = _std_min_max((x0 + m[0] * x1) * w)
        fy = _std_min_max((y0 + m[3] * x1) * w)
    xi = np.rint(fx).astype(np.int64)
    yi = np.rint(fy).astype(np.int64)
    sx = np.clip(xi >> INTER_BITS, -32768, 32767)
    sy = np.clip(yi >> INTER_BITS, -32768, 32767)
    ax = xi & (INTER_TAB_SIZE - 1)
    ay = yi & (INTER_TAB_SIZE - 1)
    w00, w01 = (32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32
    w10, w11 = (32 - ax) * ay * 32, ax * ay * 32

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)]
        return np.where(ok[..., None], v, 0)

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None]
           + tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None])
    out = ((acc + (1 << 14)) >> 15).astype(np.uint8)
    return out.reshape(height, width, cn) if img.ndim == 3 else out.reshape(height, width)
