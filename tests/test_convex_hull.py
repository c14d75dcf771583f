"""generate_convex_hull (src/semantic_convex_hull.py:17-91).

The golden vertex lists come from the reference's own function (oracle/make_golden_hull.py).  CPU: the host restatement
oracle/hull_port.py against them.  GPU: the CUDA path (class mask + erosion + connected components + row extremes
through the C ABI, hull of the row extremes on the host) against them, and its labelling against scipy / OpenCV."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import hull_port
from oracle.make_golden_hull import CASES, case_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "convex_hull.json")) as f:
    GOLDEN = json.load(f)
SMALL = [n for n in CASES if n != "full_resolution"]


def image(name):
    img = case_image(name)
    assert hashlib.sha256(img.tobytes()).hexdigest() == GOLDEN[name]["image_sha"]
    return img


def same(vertices, name):
    want = GOLDEN[name]["vertices"]
    assert len(vertices) == len(want)
    for got, ref in zip(vertices, want):
        assert np.array_equal(np.asarray(got), np.asarray(ref))


@pytest.mark.parametrize("name", SMALL)
def test_port_matches_reference_vertices(name):
    pytest.importorskip("cv2")
    pytest.importorskip("scipy")
    spec = CASES[name]
    same(hull_port.generate_convex_hull(image(name), index_care_about=spec["index"], **spec["kw"]), name)


def _reduced(pts):
    """row extremes of a point set in raster order, as semantic_convex_hull._raster_ordered_extremes lists them"""
    from vision_semantic_segmentation_b200.semantic_convex_hull import _raster_ordered_extremes
    h = int(pts[:, 1].max()) + 1
    lo, hi = np.full(h, 0x7fffffff, np.int32), np.full(h, -1, np.int32)
    np.minimum.at(lo, pts[:, 1], pts[:, 0])
    np.maximum.at(hi, pts[:, 1], pts[:, 0])
    return _raster_ordered_extremes(lo, hi)


def test_row_extremes_carry_the_whole_hull():
    """what the CUDA path relies on: cv2.convexHull of a component's row extremes (in raster order) is cv2.convexHull
    of all its pixels, starting vertex included -- blobs, and thin / diagonal / collinear components, where OpenCV's
    final rotation by input index decides the start"""
    cv2 = pytest.importorskip("cv2")
    pytest.importorskip("scipy")
    checked = 0
    for name in ("three_largest", "touching_borders", "other_class", "named_components"):
        labels = hull_port.erode_and_label(image(name), CASES[name]["index"])
        for sel in range(1, labels.max() + 1):
            ys, xs = np.where(labels == sel)
            pts = np.stack([xs, ys], 1)[1:].astype(np.int32)
            if len(pts) == 0:
                continue
            assert np.array_equal(cv2.convexHull(pts), cv2.convexHull(_reduced(pts)))
            checked += 1
    rng = np.random.default_rng(9)
    for t in range(600):
        h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        m = np.zeros((h, w), bool)
        kind = t % 4
        yy, xx = np.ogrid[:h, :w]
        if kind == 0:      # a diagonal band
            m[np.abs((yy - rng.integers(0, h)) * int(rng.choice([-1, 1])) - (xx - rng.integers(0, w))) <= rng.integers(0, 2)] = True
        elif kind == 1:    # a single row / column / a few collinear pixels
            if rng.random() < 0.5:
                m[rng.integers(0, h), rng.integers(0, w // 2):] = True
            else:
                m[rng.integers(0, h // 2):, rng.integers(0, w)] = True
        elif kind == 2:    # sparse random pixels (any raster-ordered point set)
            m[rng.random((h, w)) < 0.08] = True
        else:              # a rectangle with a notch
            m[h // 4: 3 * h // 4 + 1, w // 4: 3 * w // 4 + 1] = True
            m[h // 4, w // 4] = False
        ys, xs = np.where(m)
        pts = np.stack([xs, ys], 1)[1:].astype(np.int32)
        if len(pts) == 0:
            continue
        assert np.array_equal(cv2.convexHull(pts), cv2.convexHull(_reduced(pts))), (t, pts.tolist())
        checked += 1
    assert checked > 500


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_convex_hull_matches_reference(name):
    pytest.importorskip("torch")
    pytest.importorskip("cv2")
    from vision_semantic_segmentation_b200 import semantic_convex_hull as sch
    spec = CASES[name]
    same(sch.generate_convex_hull(image(name), index_care_about=spec["index"], **spec["kw"]), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["three_largest", "touching_borders", "named_components", "full_resolution"])
def test_gpu_components_match_scipy_labelling(name):
    torch = pytest.importorskip("torch")
    pytest.importorskip("cv2")
    pytest.importorskip("scipy")
    from vision_semantic_segmentation_b200 import semantic_convex_hull as sch
    img = image(name)
    want = hull_port.erode_and_label(img, CASES[name]["index"])
    got, areas = sch.label_components(torch.from_numpy(img).cuda(), CASES[name]["index"])
    assert np.array_equal(got, want)
    assert np.array_equal(areas, np.bincount(want.ravel())[1:])


@pytest.mark.gpu
def test_gpu_convex_hull_argument_errors():
    pytest.importorskip("torch")
    from vision_semantic_segmentation_b200 import semantic_convex_hull as sch
    img = image("named_components")
    with pytest.raises(ValueError):
        sch.generate_convex_hull(img.astype(np.int64))           # cv2.erode rejects it in the reference
    with pytest.raises(ValueError):
        sch.generate_convex_hull(img, index_to_vitualize=[999])  # np.concatenate of nothing in the reference
    with pytest.raises(SystemExit):
        sch.generate_convex_hull(img, index_care_about=0)
    assert sch.generate_convex_hull(np.zeros((30, 30), np.uint8)) == []
