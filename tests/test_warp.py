"""The planar-projection warp (generate_homography -> cv2.warpPerspective, src/homography.py:22-76).

CPU: oracle/warp_port.py against cv2 itself (when cv2 imports) and against the committed vectors of
oracle/make_golden_warp.py.  GPU: smap_warp_perspective (through homography.warp_perspective / generate_homography)
against the same vectors and the port, bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import warp_port
from oracle.make_golden_warp import CASES, case_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "warp.json")) as f:
    GOLDEN = json.load(f)["cases"]
SMALL = [n for n in CASES if n != "label_image_to_map"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(name):
    """the case's inputs with the homography taken from the committed vector (cv2.findHomography need not be rerun)"""
    img, _, dsize = case_inputs(name) if "h" in CASES[name] else _inputs_without_cv2(name)
    assert sha(img) == GOLDEN[name]["image_sha"]
    return img, np.array(GOLDEN[name]["h"]).reshape(3, 3), dsize


def _inputs_without_cv2(name):
    spec = dict(CASES[name])
    spec["h"] = np.array(GOLDEN[name]["h"]).reshape(3, 3).tolist()
    saved = CASES[name]
    CASES[name] = spec
    try:
        return case_inputs(name)
    finally:
        CASES[name] = saved


@pytest.mark.parametrize("name", SMALL)
def test_port_matches_committed_opencv_output(name):
    img, h, dsize = inputs(name)
    got = warp_port.warp_perspective(img, h, dsize)
    assert list(got.shape) == GOLDEN[name]["out_shape"]
    assert sha(got) == GOLDEN[name]["out_sha"]
    if "out" in GOLDEN[name]:
        assert np.array_equal(got.ravel(), np.array(GOLDEN[name]["out"], np.uint8))


def test_port_matches_opencv_on_random_homographies():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for t in range(25):
        sh, sw = int(rng.integers(8, 160)), int(rng.integers(8, 220))
        cn = int(rng.integers(1, 5))
        img = rng.integers(0, 256, (sh, sw) if cn == 1 else (sh, sw, cn), dtype=np.uint8)
        w, hh = int(rng.integers(3, 230)), int(rng.integers(3, 170))
        src = np.array([[0, 0], [sw, 0], [sw, sh], [0, sh]], np.float64) + rng.normal(0, 4, (4, 2))
        dst = np.array([[0, 0], [w, 0], [w, hh], [0, hh]], np.float64) + rng.normal(0, 0.25 * min(w, hh), (4, 2))
        h, _ = cv2.findHomography(src, dst)
        if h is None:
            continue
        assert np.array_equal(warp_port.invert3(h), cv2.invert(h)[1])
        assert np.array_equal(warp_port.warp_perspective(img, h, (w, hh)), cv2.warpPerspective(img, h, (w, hh))), t


def test_block_size_rule():
    assert warp_port.block_size(2000, 2000) == (64, 16)
    assert warp_port.block_size(37, 11) == (37, 11)
    assert warp_port.block_size(200, 5) == (200, 5)
    assert warp_port.block_size(300, 3) == (300, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_warp_matches_opencv_vectors(name):
    torch = pytest.importorskip("torch")
    from vision_semantic_segmentation_b200 import homography
    img, h, dsize = inputs(name)
    got = homography.warp_perspective(img, h, dsize)
    assert list(got.shape) == GOLDEN[name]["out_shape"]
    assert sha(got) == GOLDEN[name]["out_sha"]
    dev = homography.warp_perspective(torch.from_numpy(img).cuda(), h, dsize)     # device in, device out
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)
    if name != "label_image_to_map":
        assert np.array_equal(got, warp_port.warp_perspective(img, h, dsize))


@pytest.mark.gpu
def test_gpu_generate_homography_like_the_reference_call():
    """update_map_planar's call (src/mapping.py:465-466): image, projected anchors, map anchors, out_size = [MW, MH]."""
    cv2 = pytest.importorskip("cv2")
    from vision_semantic_segmentation_b200 import homography
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (144, 192, 3), dtype=np.uint8)
    pts_image = np.array([[60.0, 90.0], [130.0, 90.0], [180.0, 140.0], [10.0, 140.0]])
    anchors = np.array([[60.0, 40.0], [140.0, 40.0], [140.0, 180.0], [60.0, 180.0]])
    im_dst, h = homography.generate_homography(img, pts_image, anchors, out_size=[200, 220], return_h=True)
    h_ref, _ = cv2.findHomography(pts_image, anchors)
    assert np.array_equal(h, h_ref)
    assert np.array_equal(im_dst, cv2.warpPerspective(img, h_ref, (200, 220)))
    assert np.array_equal(homography.generate_homography(img, pts_image, anchors),
                          cv2.warpPerspective(img, h_ref, (192, 144)))
    with pytest.raises(ValueError):
        homography.warp_perspective(img.astype(np.float32), h, (10, 10))
