#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors of the label-image producer step from the REAL
reference.

Run in the build container (needs /root/reference and cv2):   python -m oracle.make_golden_ids
For seeded class-id planes of several shapes it executes what the reference's segmentation node does between the
network and the publisher (src/vision_semantic_segmentation_node.py:109-113):

    image_out = cv2.resize(ids, (W, H), interpolation=cv2.INTER_NEAREST)
    colored   = apply_color_map(image_out, labels)       # the reference's own function, imported unmodified,
                                                         # labels = get_labels(config/config_19.json)
and stores  tests/golden/label_ids.json  (generator parameters + sha256 of ids and of the painted image) and
tests/golden/label_ids.npz  (the palette, and the painted image of the small cases in full).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (ids height, ids width) -> (H, W); ids drawn from [0, max_id): ids >= 19 have no palette entry
CASES = {
    "full_res": dict(ids_hw=[1440, 1920], out_hw=[1440, 1920], max_id=19, seed=500),
    "half_res": dict(ids_hw=[720, 960], out_hw=[1440, 1920], max_id=24, seed=501),          # IMAGE_SCALE 0.5
    "scale_03": dict(ids_hw=[432, 576], out_hw=[1440, 1920], max_id=19, seed=502),          # IMAGE_SCALE 0.3
    "awkward": dict(ids_hw=[185, 130], out_hw=[1440, 1920], max_id=256, seed=503),          # no multiply-shift reproduces it
    "tiny": dict(ids_hw=[7, 5], out_hw=[33, 40], max_id=22, seed=504, store=True),
    "tiny_same": dict(ids_hw=[9, 11], out_hw=[9, 11], max_id=30, seed=505, store=True),
}


def make_ids(spec):
    rng = np.random.default_rng(spec["seed"])
    return rng.integers(0, spec["max_id"], tuple(spec["ids_hw"])).astype(np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    ref = ref_shim.load_reference_color_map()
    labels = ref.get_labels(ref.config_19)
    arrays = {"palette": np.array([lab["color"] for lab in labels], dtype=np.uint8)}
    manifest = {"generator": "oracle/make_golden_ids.py", "opencv": cv2.__version__, "numpy": np.__version__,
                "palette_source": "config/config_19.json", "cases": {}}
    for name, spec in CASES.items():
        ids = make_ids(spec)
        h, w = spec["out_hw"]
        image_out = cv2.resize(ids, (w, h), interpolation=cv2.INTER_NEAREST)
        colored = np.squeeze(ref.apply_color_map(image_out, labels)).astype(np.uint8)
        entry = dict(spec)
        entry.update({"ids_sha": sha(ids), "upscaled_sha": sha(image_out), "colored_sha": sha(colored)})
        manifest["cases"][name] = entry
        if spec.get("store"):
            arrays["colored_" + name] = colored
        print(name, ids.shape, "->", colored.shape)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "label_ids.npz"), **arrays)
    with open(os.path.join(OUT, "label_ids.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
