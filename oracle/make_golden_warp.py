"""Generates tests/golden/warp.json: sha256 of cv2.warpPerspective outputs (the call generate_homography makes,
/root/reference/src/homography.py:53-55) on seeded inputs, with the cv2 of this image (4.13.0).  The inputs are
regenerated from the seeds by the tests; small cases are stored in full.

    python oracle/make_golden_warp.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case_inputs(name):
    """(image, h, (W, H)) of a named case -- shared with tests/test_warp.py"""
    spec = CASES[name]
    rng = np.random.default_rng(spec["seed"])
    sh, sw, cn = spec["src"]
    shape = (sh, sw) if cn == 1 else (sh, sw, cn)
    if spec.get("labels"):
        from vision_semantic_segmentation_b200 import synthetic as syn
        tiles = rng.integers(0, 19, (-(-sh // 48), -(-sw // 48)))
        ids = np.kron(tiles, np.ones((48, 48), np.int64))[:sh, :sw]
        img = np.ascontiguousarray(syn.COLORS_19[ids].astype(np.uint8))
    else:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
    w, hh = spec["dst"]
    if "h" in spec:
        h = np.array(spec["h"], np.float64)
    else:
        q = np.array(spec["quad"], np.float64) * [sw, sh] + rng.normal(0, spec.get("jitter", 3.0), (4, 2))
        d = np.array(spec["dquad"], np.float64) * [w, hh]
        import cv2
        h, _ = cv2.findHomography(q, d)
    return img, h, (w, hh)


CASES = {
    # the reference's use: the 1920 x 1440 label image onto the 2000 x 2000 map image, anchors = a road trapezoid
    "label_image_to_map": {"seed": 1, "src": [1440, 1920, 3], "dst": [2000, 2000], "labels": True,
                           "quad": [[0.31, 0.62], [0.68, 0.62], [0.94, 0.97], [0.05, 0.97]],
                           "dquad": [[0.3, 0.2], [0.7, 0.2], [0.7, 0.9], [0.3, 0.9]]},
    "random_rgb": {"seed": 2, "src": [173, 251, 3], "dst": [320, 200],
                   "quad": [[0.1, 0.1], [0.9, 0.05], [0.95, 0.9], [0.02, 0.95]],
                   "dquad": [[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]]},
    "gray_small_blocks": {"seed": 3, "src": [40, 50, 1], "dst": [37, 11],      # fewer than 16 rows, fewer than 64 columns
                          "h": [[1.1, 0.05, -3.0], [0.02, 0.9, 4.0], [1e-4, -2e-4, 1.0]]},
    "horizon_inside": {"seed": 4, "src": [120, 160, 3], "dst": [200, 150],     # W changes sign inside the output
                       "h": [[0.828992242252, -0.0816999587753, -2.3235768092], [-0.0434733725593, 0.908443578308, -1.68646703894],
                             [0.00318554885133, 0.00239853089982, 0.985646291646]]},
    "magnify_4ch": {"seed": 5, "src": [9, 7, 4], "dst": [130, 70], "h": [[17.0, 0.5, 3.0], [-0.7, 8.0, 2.0], [0.0, 0.0, 1.0]]},
    "far_outside": {"seed": 6, "src": [30, 30, 2], "dst": [65, 17], "h": [[1e-3, 0.0, 40000.0], [0.0, 1e-3, -70000.0], [0.0, 0.0, 1.0]]},
    "singular": {"seed": 7, "src": [20, 20, 3], "dst": [16, 16], "h": [[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.0, 0.0, 1.0]]},
}


def main():
    import cv2
    out = {"cv2_version": cv2.__version__, "cases": {}}
    for name in CASES:
        img, h, dsize = case_inputs(name)
        ref = cv2.warpPerspective(img, h, dsize)
        rec = {"image_sha": hashlib.sha256(img.tobytes()).hexdigest(), "h": [float(v) for v in np.asarray(h).ravel()],
               "out_sha": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest(), "out_shape": list(ref.shape),
               "nonzero": int(np.count_nonzero(ref))}
        if ref.size <= 4096:
            rec["out"] = ref.ravel().tolist()
        out["cases"][name] = rec
        print(name, rec["out_shape"], rec["nonzero"])
    with open(os.path.join(ROOT, "tests", "golden", "warp.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
