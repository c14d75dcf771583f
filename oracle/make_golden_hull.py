"""Generates tests/golden/convex_hull.json by running the reference's own generate_convex_hull
(/root/reference/src/semantic_convex_hull.py:17-91, unmodified, imported through oracle/ref_shim.py) on seeded label
images.  The images are regenerated from the seeds by the tests (sha256-checked).

    python oracle/make_golden_hull.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "crosswalk_like": {"seed": 1, "shape": [360, 480], "blobs": 6, "index": 1, "kw": {}},
    "three_largest": {"seed": 2, "shape": [200, 260], "blobs": 9, "index": 1, "kw": {"top_number": 3}},
    "other_class": {"seed": 3, "shape": [150, 150], "blobs": 5, "index": 7, "kw": {"top_number": 2, "area_threshold": 5}},
    "touching_borders": {"seed": 4, "shape": [64, 96], "blobs": 12, "index": 1, "kw": {"top_number": 4, "area_threshold": 0}},
    "named_components": {"seed": 5, "shape": [90, 120], "blobs": 7, "index": 1, "kw": {"index_to_vitualize": [2, 1]}},
    "nothing_left": {"seed": 6, "shape": [40, 40], "blobs": 0, "index": 1, "kw": {}},
    "below_threshold": {"seed": 7, "shape": [40, 40], "blobs": 1, "index": 1, "kw": {"area_threshold": 100000}},
    "full_resolution": {"seed": 8, "shape": [1440, 1920], "blobs": 14, "index": 1, "kw": {"top_number": 3}},
}


def case_image(name):
    spec = CASES[name]
    rng = np.random.default_rng(spec["seed"])
    h, w = spec["shape"]
    img = rng.integers(2, 6, (h, w)).astype(np.uint8)          # other classes everywhere
    img[img == spec["index"]] = 0
    yy, xx = np.ogrid[:h, :w]
    for _ in range(spec["blobs"]):
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        ry, rx = int(rng.integers(3, max(4, h // 4))), int(rng.integers(3, max(4, w // 4)))
        kind = rng.random()
        if kind < 0.4:
            img[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = spec["index"]
        elif kind < 0.8:
            img[max(0, cy - ry):cy + ry, max(0, cx - rx):cx + rx] = spec["index"]
        else:   # a slanted band: diagonal (8-connected only) contacts
            img[np.abs((yy - cy) - (xx - cx)) <= 2] = spec["index"]
    holes = rng.random((h, w)) < 0.01
    img[holes & (img == spec["index"])] = 0
    return img


def main():
    from oracle import ref_shim
    ref = ref_shim.load_reference_convex_hull()
    out = {}
    for name, spec in CASES.items():
        img = case_image(name)
        verts = ref(img, index_care_about=spec["index"], **spec["kw"])
        out[name] = {"image_sha": hashlib.sha256(img.tobytes()).hexdigest(),
                     "vertices": [np.asarray(v).tolist() for v in verts]}
        print(name, [np.asarray(v).shape for v in verts])
    with open(os.path.join(ROOT, "tests", "golden", "convex_hull.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
