"""PNG output of the rendered map.

The reference saves the colour map with ``cv2.imwrite`` (``src/mapping_replay.py:204-206``), i.e. the
RGB array is interpreted as BGR and the file holds the channels swapped.  The same bytes are produced
here: through OpenCV when it is importable, otherwise by a minimal zlib PNG encoder.
"""
import struct
import zlib

import numpy as np


def _png_bytes(rgb):
    h, w, _ = rgb.shape
    raw = np.empty((h, 1 + 3 * w), dtype=np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = rgb.reshape(h, 3 * w)

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + chunk(b"IEND", b""))


def imwrite(path, color_map):
    color_map = np.ascontiguousarray(color_map, dtype=np.uint8)
    try:
        import cv2
        if cv2.imwrite(path, color_map):
            return True
    except ImportError:
        pass
    with open(path, "wb") as f:
        f.write(_png_bytes(color_map[:, :, ::-1]))  # cv2 would store arr[..., 2] as red
    return True
