"""The label-image producer boundary: what happens between the segmentation
network's output and the image ``project_pcd`` indexes.

In the reference the segmentation node takes the network's class-id plane, upscales it to the camera
resolution with ``cv2.resize(..., interpolation=cv2.INTER_NEAREST)``
(``src/vision_semantic_segmentation_node.py:109-110``), paints it with the dataset palette
(``apply_color_map``, ``src/network/deeplab_v3_plus/data/utils/mapillary_visualization.py:70-89``) and
publishes the RGB image; the mapper then compares R and G of every hit pixel with ``cfg.LABEL_COLORS``
(``src/mapping_replay.py:276``).  The CUDA path can take the class-id plane itself
(``SMAP_IMG_CLASS_IDS`` in ``include/smap.h``): 1 byte per network pixel instead of 3 per camera pixel over
PCIe, no colour compare in the kernel.  The functions below are the HOST restatement of the two producer
steps -- they define what the kernel must reproduce and let a caller build the RGB image the reference
would have seen; the mapping path itself never calls them.
"""
import math

import numpy as np


def get_labels(config_file):
    """``labels`` list of a dataset config such as ``config/config_19.json``
    (``mapillary_visualization.py:10-19``)."""
    import json
    with open(config_file) as f:
        return json.load(f)["labels"]


def palette_of(labels):
    """(n, 3) uint8 palette from a ``labels`` list (dicts with a ``color`` entry) or an array of colours."""
    if len(labels) and isinstance(labels[0], dict):
        labels = [lab["color"] for lab in labels]
    pal = np.asarray(labels)
    if pal.size == 0:
        return np.zeros((0, 3), dtype=np.uint8)
    if pal.ndim != 2 or pal.shape[1] != 3:
        raise ValueError("a palette is a list of RGB triples")
    if pal.shape[0] > 256:
        raise ValueError("a uint8 class-id plane addresses at most 256 colours")
    return np.ascontiguousarray(pal.astype(np.uint8))


def apply_color_map(label_array, labels):
    """Paint a class-id array, (H, W) or (B, H, W), with the palette; ids without a palette entry stay black
    (the canvas is ``np.zeros``).  Same results as ``mapillary_visualization.py:70-89``."""
    label_array = np.asarray(label_array)
    if label_array.ndim not in (2, 3):
        raise NotImplementedError
    pal = palette_of(labels)
    table = np.zeros((max(256, int(label_array.max(initial=0)) + 1), 3), dtype=np.uint8)
    table[:pal.shape[0]] = pal
    idx = label_array.astype(np.int64)
    out = table[np.clip(idx, 0, table.shape[0] - 1)]
    out[(idx < 0)] = 0
    return out


def nearest_index_map(dst, src):
    """Source index of every destination index along one axis of ``cv2.resize(..., INTER_NEAREST)``:
    ``min(floor(x * ifx), src - 1)`` with ``ifx = 1.0 / (dst / src)`` evaluated in double, as OpenCV does."""
    if dst <= 0 or src <= 0:
        raise ValueError("sizes must be positive")
    ifx = 1.0 / (float(dst) / float(src))
    return np.array([min(int(math.floor(x * ifx)), src - 1) for x in range(dst)], dtype=np.int64)


def upscale_nearest(plane, width, height):
    """``cv2.resize(plane, (width, height), interpolation=cv2.INTER_NEAREST)`` for a 2-D plane or an (h, w, c) image."""
    plane = np.asarray(plane)
    iy = nearest_index_map(height, plane.shape[0])
    ix = nearest_index_map(width, plane.shape[1])
    return np.ascontiguousarray(plane[iy][:, ix])


def paint_class_ids(ids, labels, width=None, height=None):
    """The RGB image the reference's node publishes for a network output ``ids`` (h, w) uint8:
    nearest-neighbour upscale to (height, width), then ``apply_color_map``."""
    ids = np.asarray(ids)
    if ids.ndim != 2:
        raise ValueError("a class-id plane is 2-D")
    width = ids.shape[1] if width is None else width
    height = ids.shape[0] if height is None else height
    return np.ascontiguousarray(apply_color_map(upscale_nearest(ids, width, height), labels))


def id_class_bits(labels, label_colors):
    """256-entry table: bit i of entry ``id`` is set when the palette colour of ``id`` equals ``label_colors[i]`` in R
    and G -- the reference's colour compare (``src/mapping_replay.py:276``) folded over the palette, collisions and the
    black of unknown ids included.  The library builds the same table on its own (``smap_set_label_palette``); this
    copy documents it and serves the tests."""
    pal = np.zeros((256, 3), dtype=np.uint8)
    p = palette_of(labels)
    pal[:p.shape[0]] = p
    colors = np.asarray(label_colors).astype(np.uint8)
    bits = np.zeros(256, dtype=np.uint32)
    for i in range(colors.shape[0]):
        hit = (pal[:, 0] == colors[i, 0]) & (pal[:, 1] == colors[i, 1])
        bits[hit] |= np.uint32(1 << i)
    return bits
