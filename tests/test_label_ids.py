"""CPU: the label-image producer boundary (class-id planes).  The host restatement in
``label_image.py`` against golden vectors made by the real reference (``oracle/make_golden_ids.py``: the reference's
own ``apply_color_map`` after ``cv2.resize(..., INTER_NEAREST)``), and the index map / multiply-shift the library
hands to the kernel (host-only entry point, no device needed) against the same restatement."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests.common import GOLDEN, sha
from vision_semantic_segmentation_b200 import _native, label_image, synthetic as syn


def golden_ids():
    with open(os.path.join(GOLDEN, "label_ids.json")) as f:
        return json.load(f), np.load(os.path.join(GOLDEN, "label_ids.npz"))


def make_ids(spec):
    rng = np.random.default_rng(spec["seed"])
    return rng.integers(0, spec["max_id"], tuple(spec["ids_hw"])).astype(np.uint8)


def test_palette_is_the_reference_config():
    manifest, arrays = golden_ids()
    assert np.array_equal(arrays["palette"], syn.COLORS_19)


@pytest.mark.parametrize("name", ["full_res", "half_res", "scale_03", "awkward", "tiny", "tiny_same"])
def test_paint_class_ids_matches_reference_golden(name):
    manifest, arrays = golden_ids()
    spec = manifest["cases"][name]
    ids = make_ids(spec)
    assert sha(ids) == spec["ids_sha"], "generator drifted"
    h, w = spec["out_hw"]
    assert sha(label_image.upscale_nearest(ids, w, h)) == spec["upscaled_sha"]
    painted = label_image.paint_class_ids(ids, arrays["palette"], w, h)
    assert painted.dtype == np.uint8 and painted.shape == (h, w, 3)
    assert sha(painted) == spec["colored_sha"]
    if "colored_" + name in arrays:
        assert np.array_equal(painted, arrays["colored_" + name])


def test_apply_color_map_shapes_and_unknown_ids():
    labels = [{"color": [1, 2, 3]}, {"color": [4, 5, 6]}]
    out = label_image.apply_color_map(np.array([[0, 1], [2, 255]], dtype=np.uint8), labels)
    assert out.tolist() == [[[1, 2, 3], [4, 5, 6]], [[0, 0, 0], [0, 0, 0]]]
    assert label_image.apply_color_map(np.zeros((2, 3, 4), dtype=np.int64), labels).shape == (2, 3, 4, 3)
    with pytest.raises(NotImplementedError):
        label_image.apply_color_map(np.zeros((4,), dtype=np.uint8), labels)
    with pytest.raises(ValueError):
        label_image.palette_of([[1, 2], [3, 4]])


def test_nearest_index_map_against_opencv():
    cv2 = pytest.importorskip("cv2")
    for dst, src in [(1920, 960), (1920, 576), (1920, 130), (1440, 185), (1000, 333), (40, 5), (640, 640), (100, 250)]:
        ramp = np.arange(src, dtype=np.float32).reshape(1, src)
        want = cv2.resize(ramp, (dst, 1), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
        assert np.array_equal(label_image.nearest_index_map(dst, src), want), (dst, src)


def library_map(dst, src):
    lib = _native.load()
    tab = np.zeros(dst, dtype=np.uint16)
    mul, shift = ctypes.c_uint32(0), ctypes.c_uint32(0)
    _native.check(lib.smap_debug_nearest_map(dst, src, tab.ctypes.data_as(ctypes.c_void_p), ctypes.byref(mul),
                                             ctypes.byref(shift)))
    return tab.astype(np.int64), mul.value, shift.value


def test_library_index_map_and_multiply_shift():
    """The table the library tabulates equals the restatement for every size pair tried; whenever it reports a
    multiply-shift, 32-bit (x * mul) >> shift reproduces the whole table; and the awkward sizes really exercise the
    table path of the kernel."""
    n_tab = 0
    pairs = [(1920, s) for s in list(range(1, 1921, 13)) + [960, 576, 480, 130, 1920]]
    pairs += [(1440, s) for s in list(range(1, 1441, 11)) + [720, 432, 185, 1440]]
    pairs += [(65535, 65535), (65535, 1), (65535, 40000), (1, 1), (2, 1), (3, 2)]
    for dst, src in pairs:
        tab, mul, shift = library_map(dst, src)
        assert np.array_equal(tab, label_image.nearest_index_map(dst, src)), (dst, src)
        if shift:
            x = np.arange(dst, dtype=np.uint64)
            assert int((x * np.uint64(mul)).max()) < 2 ** 32
            assert np.array_equal(((x * np.uint64(mul)) >> np.uint64(shift)).astype(np.int64), tab), (dst, src)
        else:
            n_tab += 1
    assert library_map(1920, 960)[2] != 0 and library_map(1440, 720)[2] != 0 and library_map(1920, 1920)[2] != 0
    assert library_map(1920, 130)[2] == 0, "the 'awkward' golden case is meant to take the table path"
    assert 0 < n_tab < len(pairs) // 2
    # downscaling (src > dst) always goes through the table
    tab, mul, shift = library_map(100, 250)
    assert shift == 0 and np.array_equal(tab, label_image.nearest_index_map(100, 250))
    lib = _native.load()
    assert lib.smap_debug_nearest_map(0, 5, tab.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ctypes.c_uint32()),
                                      ctypes.byref(ctypes.c_uint32())) == -1


def test_id_class_bits_reproduces_the_rg_compare():
    """Folding palette and LABEL_COLORS into id -> class bits gives, for every id, the bits the reference's R,G compare
    (src/mapping_replay.py:276) finds in the painted pixel -- including ids without a palette entry (black) matching the
    classes whose colour has R == G == 0 (car, motorcycle, truck of the 19-class palette)."""
    for full19 in (False, True):
        labels, names, colors = syn.class_setup(full19)
        bits = label_image.id_class_bits(syn.COLORS_19, colors)
        painted = label_image.apply_color_map(np.arange(256, dtype=np.uint8).reshape(1, 256), syn.COLORS_19)[0]
        for i, col in enumerate(colors):
            want = (painted[:, 0] == col[0]) & (painted[:, 1] == col[1])
            assert np.array_equal((bits >> i) & 1, want.astype(np.uint32))
        if full19:
            assert bin(int(bits[200])).count("1") == 3 and bits[16] == bits[17] == bits[18] == bits[200]
        else:
            assert bits[200] == 0
