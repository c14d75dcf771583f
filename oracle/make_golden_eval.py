#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors of the evaluation step from the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_eval

``/root/reference/test/test_semantic_mapping.py`` does not parse as a whole (a second ``else:`` at line 70), but the
two pieces the mapping path calls are syntactically complete on their own:

    lines   6 - 18    def convert_labels(gmap, mask=None)
    lines 117 - 161   Test.test_single_map  and  Test.iou

This script slices exactly those line ranges out of the unmodified file, ``exec``s them (``convert_labels`` at module
level, the two methods as the body of a bare ``class Test``) and runs them on seeded colour maps / ground-truth label
maps -- among them the golden render of ``cfg1_c5_count`` (the reference's own ``render_bev_map`` output).  It stores
in ``tests/golden/eval.json`` the generator parameters and what the reference returned and printed: the IoU list and
the missing rate (return values of ``Test.iou``), the accuracy list and mean accuracy (parsed from the line the
reference prints), and, for the degenerate cases, the exception it raises (``ZeroDivisionError`` on an empty union:
``iou = intersection / union`` divides two Python floats, test/test_semantic_mapping.py:140-141).
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REF_TEST = os.path.join(os.environ.get("SMAP_REFERENCE_DIR", "/root/reference"), "test", "test_semantic_mapping.py")
OUT = os.path.join(ROOT, "tests", "golden", "eval.json")

# the colours convert_labels knows (1..5), black, and three colours it must ignore
PALETTE = np.array([[0, 0, 0], [128, 64, 128], [140, 140, 200], [255, 255, 255], [244, 35, 232], [107, 142, 35],
                    [128, 64, 129], [255, 255, 0], [0, 60, 100]], dtype=np.uint8)

CASES = {
    "grid_2000": dict(shape=[2000, 2000], truth_shape=[2100, 2300], shift=[37, 91], mask=False, seed=700),
    "grid_2000_mask": dict(shape=[2000, 2000], truth_shape=[2000, 2000], shift=[0, 0], mask=True, seed=701),
    "small": dict(shape=[3, 5], truth_shape=[4, 9], shift=[1, 4], mask=False, seed=702),
    "odd": dict(shape=[333, 1027], truth_shape=[400, 1100], shift=[67, 0], mask=True, seed=703),
    # degenerate: no crosswalk anywhere -> empty union; no known ground truth at all
    "no_crosswalk": dict(shape=[16, 16], truth_shape=[16, 16], shift=[0, 0], mask=False, seed=704, drop=[2]),
    "no_truth": dict(shape=[8, 8], truth_shape=[8, 8], shift=[0, 0], mask=False, seed=705, truth_max=1),
    # the reference's own render of the first golden mapping case, scored against a seeded ground truth
    "golden_render": dict(render_of="cfg1_c5_count", truth_shape=[2000, 2000], shift=[0, 0], mask=False, seed=706),
}


def load_reference_pieces():
    with open(REF_TEST) as f:
        lines = f.read().split("\n")
    ns = {"np": np}
    exec("\n".join(lines[5:18]), ns)                       # lines 6-18: convert_labels
    exec("class Test:\n" + "\n".join(lines[116:161]), ns)  # lines 117-161: test_single_map, iou (method bodies)
    return ns["convert_labels"], ns["Test"]


def make_inputs(name, spec):
    """(rgb (H, W, 3) uint8, truth (h, w) float64, mask or None) -- also used by tests/ to regenerate the inputs"""
    rng = np.random.default_rng(spec["seed"])
    if "render_of" in spec:
        rgb = np.load(os.path.join(ROOT, "tests", "golden", spec["render_of"] + ".npz"))["rgb"]
    else:
        ids = rng.integers(0, len(PALETTE), tuple(spec["shape"]))
        for d in spec.get("drop", []):
            ids[ids == d] = 0
        rgb = PALETTE[ids]
    truth = rng.integers(0, spec.get("truth_max", 4), tuple(spec["truth_shape"])).astype(np.float64)
    for d in spec.get("drop", []):
        truth[truth == d] = 0
    mask = None
    if spec["mask"]:
        mask = (rng.uniform(size=(rgb.shape[0] + 3, rgb.shape[1] + 2)) > 0.3).astype(np.float64)
    return rgb, truth, mask


def parse_printed(text):
    """the numbers of the reference's two verbose lines (test/test_semantic_mapping.py:147-155)"""
    out = {}
    for line in text.splitlines():
        if line.startswith("Accuracy for"):
            parts = line.replace("\t", " ").replace(":", " ").split()
            # Accuracy for road X crosswalk Y lane Z mean Accuracy W
            out["acc"] = [float(parts[3]), float(parts[5]), float(parts[7])]
            out["accuracy"] = float(parts[10])
        elif line.startswith("Overall Missing rate"):
            out["miss_printed"] = float(line.split(":")[1])
    return out


def enc(v):
    v = float(v)
    return "nan" if v != v else ("inf" if v == float("inf") else ("-inf" if v == float("-inf") else v))


def main():
    convert_labels, Test = load_reference_pieces()
    manifest = {"generator": "oracle/make_golden_eval.py", "numpy": np.__version__,
                "reference_lines": "test/test_semantic_mapping.py:6-18,117-161", "cases": {}}
    for name, spec in CASES.items():
        rgb, truth, mask = make_inputs(name, spec)
        t = Test.__new__(Test)
        t.class_lists, t.d = [1, 2, 3], {0: "road", 1: "crosswalk", 2: "lane"}
        t.shift_w, t.shift_h = spec["shift"]
        t.ground_truth_mask = truth
        entry = dict(spec)
        generated = convert_labels(rgb, mask)
        entry["labels_hist"] = [int(np.sum(generated == k)) for k in range(6)]
        gmap = truth[t.shift_w:generated.shape[0] + t.shift_w, t.shift_h:generated.shape[1] + t.shift_h]
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ious, miss = t.iou(gmap, generated, verbose=True)
                if mask is None:   # test_single_map = convert_labels(global_map) without a mask + iou(verbose)
                    buf2 = io.StringIO()
                    with contextlib.redirect_stdout(buf2):
                        t.test_single_map(rgb)
                    assert buf2.getvalue() == buf.getvalue(), "test_single_map and iou disagree"
            entry["iou"] = [enc(v) for v in ious]
            entry["miss"] = enc(miss)
            entry.update({k: ([enc(x) for x in v] if isinstance(v, list) else enc(v))
                          for k, v in parse_printed(buf.getvalue()).items()})
            entry["printed"] = buf.getvalue().splitlines()
        except ZeroDivisionError:
            entry["raises"] = "ZeroDivisionError"
        manifest["cases"][name] = entry
        print(name, {k: entry.get(k) for k in ("iou", "miss", "acc", "accuracy", "raises")})
    with open(OUT, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
