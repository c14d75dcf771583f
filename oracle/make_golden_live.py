#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors of the live node's host logic (src/mapping.py) from the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_live
The reference's live-node class is imported unmodified through oracle/ref_shim.py (rospy, tf, cv_bridge, hickle ...
are inert placeholders) and two things are recorded in tests/golden/live_node.json:

* queue synchronisation: ``update_pcd`` / ``update_pose`` (src/mapping.py:185-259) executed on seeded stamp
  sequences -- which queue entry is picked for a target stamp and how much of the queue is kept;
* ``mapping()`` (src/mapping.py:292-355) driven frame by frame over the frames of the golden case cfg1_c5_count,
  with ``save_map_to_file`` raised for the last one: the generator asserts that the map it leaves and the image it
  writes are the ones ``mapping_replay`` produced for the same frames (tests/golden/manifest.json), so the existing
  golden vectors pin the live entry point as well; the fact and the recorded frame count are stored.
"""
import hashlib
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from vision_semantic_segmentation_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_live():
    ref_shim.load_reference()
    saved = list(sys.path)
    saved_test = sys.modules.get('test'), sys.modules.get('test.test_semantic_mapping')
    sys.path[:0] = [ref_shim.REF, os.path.join(ref_shim.REF, 'src')]
    test_pkg = types.ModuleType('test')
    test_pkg.__path__ = []
    test_mod = types.ModuleType('test.test_semantic_mapping')
    test_mod.Test = object
    sys.modules['test'], sys.modules['test.test_semantic_mapping'] = test_pkg, test_mod
    try:
        return importlib.import_module('src.mapping')
    finally:
        sys.path[:] = saved
        for key, val in zip(('test', 'test.test_semantic_mapping'), saved_test):
            if val is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = val


class Stamp(float):
    """Stand-in for rospy.Time: ordered, subtractable, with the two fields the reference's log lines read."""
    secs = property(lambda self: int(self))
    nsecs = property(lambda self: int((self - int(self)) * 1e9))

    def __sub__(self, other):
        return Stamp(float(self) - float(other))


def queue_cases(live):
    rng = np.random.default_rng(77)
    seqs = [[1.0], [1.0, 2.0], [1.0, 2.0, 3.0, 4.0, 5.0], [0.5, 0.75, 2.0, 2.0, 3.5, 9.0]]
    for _ in range(12):
        seqs.append(sorted(float(v) for v in np.round(rng.uniform(0, 10, int(rng.integers(1, 9))), 2)))
    cases = []
    for stamps in seqs:
        targets = [stamps[0] - 1.0, stamps[-1] + 1.0] + list(stamps) + \
                  [float(v) for v in np.round(rng.uniform(stamps[0] - 0.5, stamps[-1] + 0.5, 6), 3)]
        mids = [(a + b) / 2 for a, b in zip(stamps, stamps[1:])]
        for target in targets + mids:
            # point clouds
            me = types.SimpleNamespace(pcd_header_queue=[types.SimpleNamespace(stamp=Stamp(s)) for s in stamps],
                                       pcd_queue=list(range(len(stamps))))
            pcd, stamp = live.SemanticMapping.update_pcd(me, Stamp(target))
            # poses
            me2 = types.SimpleNamespace(pose_queue=[types.SimpleNamespace(header=types.SimpleNamespace(stamp=Stamp(s)), pose=i)
                                                    for i, s in enumerate(stamps)])
            pose, stamp2 = live.SemanticMapping.update_pose(me2, Stamp(target))
            cases.append({"stamps": stamps, "target": target,
                          "pcd_pick": int(pcd), "pcd_stamp": float(stamp), "pcd_left": [int(v) for v in me.pcd_queue],
                          "pose_pick": int(pose), "pose_stamp": float(stamp2),
                          "pose_left": [int(m.pose) for m in me2.pose_queue]})
    return cases


def live_mapping_equals_replay(live):
    with open(os.path.join(OUT, "manifest.json")) as f:
        spec = json.load(f)["cases"]["cfg1_c5_count"]
    golden = np.load(os.path.join(OUT, "cfg1_c5_count.npz"))
    tmp = tempfile.mkdtemp()
    base_cfg = importlib.import_module('src.config.base_cfg')
    cfg = base_cfg.get_cfg_defaults()
    cfg.OUTPUT_DIR = tmp
    cfg.MAPPING.INPUT_DIR = tmp
    cfg.LABELS, cfg.LABELS_NAMES, cfg.LABEL_COLORS = syn.class_setup(False)
    # the labelled point cloud is only PUBLISHED (src/mapping.py:317-318); building the ROS message needs real
    # sensor_msgs constants, so the publisher helper is replaced by a no-op -- nothing on the map path reads it
    live.create_point_cloud = lambda *a, **k: None
    sm = live.SemanticMapping(cfg)
    assert sm.depth_method in ('points_map', 'points_raw')
    for f in range(spec["frames"]):
        fr = syn.synthetic_frame(spec["seed"], f, spec["n_points"], blocky=(f in spec["blocky_frames"]))
        sm.pcd, sm.pcd_frame_id = fr["pcd"], "world"
        sm.save_map_to_file = f == spec["frames"] - 1
        sm.mapping(fr["semantic_image"], fr["pose"], sm.cam1)
    import cv2
    written = cv2.imread(os.path.join(sm.output_dir, "global_map.png"))
    assert sha(sm.map) == spec["filtered_sha"], "live mapping() and mapping_replay() disagree on the map"
    assert np.array_equal(written, golden["rgb"]), "live mapping() and mapping_replay() disagree on the image"
    return {"case": "cfg1_c5_count", "frames_recorded": len(sm.input_list),
            "recorded_keys": sorted(sm.input_list[0].keys()),
            "map_after_save_is": "filtered_sha", "image_is": "rgb (as written by cv2.imwrite and read back)"}


def main():
    live = load_live()
    out = {"generator": "oracle/make_golden_live.py", "queue_cases": queue_cases(live),
           "mapping": live_mapping_equals_replay(live)}
    with open(os.path.join(OUT, "live_node.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(len(out["queue_cases"]), "queue cases;", out["mapping"])


if __name__ == "__main__":
    main()
