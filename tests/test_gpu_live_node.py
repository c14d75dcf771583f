"""GPU: the live node's entry point (``mapping.SemanticMapping``: callbacks + ``mapping()``) driven with messages, against
the golden vectors of the real reference.  ``oracle/make_golden_live.py`` ran the reference's own live ``mapping()``
over these frames and found the map and the written image identical to its ``mapping_replay`` outputs, so the committed
cfg1_c5_count vectors are the expected results here."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from tests.common import Case, sha  # noqa: E402
from tests.test_gpu_api import make_cfg  # noqa: E402
from vision_semantic_segmentation_b200 import replay_io, synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.mapping import SemanticMapping  # noqa: E402


def msg(stamp, frame_id="world", **payload):
    return types.SimpleNamespace(header=types.SimpleNamespace(stamp=stamp, frame_id=frame_id), **payload)


def drive(sm, case, cloud_of, image_of):
    """Feed the case's frames as the three message streams of the live node; the last pose trips the clock."""
    n = case.spec["frames"]
    frames = [syn.synthetic_frame(case.spec["seed"], f, case.spec["n_points"], with_ids=True,
                                  blocky=(f in case.spec.get("blocky_frames", []))) for f in range(n)]
    sm.test_cut_time = 10 * (n - 1)
    for f, fr in enumerate(frames):
        t = 10.0 * f
        sm.pcd_callback(msg(t - 0.02, "world", points=cloud_of(fr)))
        sm.pcd_callback(msg(t + 4.0, "world", points=np.zeros((1, 4), np.float32)))       # a later cloud: not picked
        sm.pose_callback(msg(t + 0.01, pose=fr["pose"]))
        assert not sm.done
        sm.image_callback(msg(t, "camera1", image=image_of(fr)))
    return frames


@pytest.mark.parametrize("variant", ["float4", "reference_pcd", "class_ids", "with_point_cloud_consumer"])
def test_live_node_reproduces_reference_golden(tmp_path, variant):
    case = Case("cfg1_c5_count")
    cfg = make_cfg(tmp_path, case, False)
    cfg.MAPPING.INPUT_DIR = str(tmp_path / "recorded")
    sm = SemanticMapping(cfg)
    shown, clouds = [], []
    sm.on_semantic_local_map = shown.append
    cloud_of, image_of = (lambda fr: fr["points"]), (lambda fr: fr["semantic_image"])
    if variant == "reference_pcd":
        cloud_of = lambda fr: fr["pcd"]                                                  # noqa: E731
    if variant == "class_ids":
        sm.set_label_palette(syn.COLORS_19)
        image_of = lambda fr: fr["semantic_ids"]                                         # noqa: E731
    if variant == "with_point_cloud_consumer":
        sm.on_semantic_point_cloud = lambda pcd, label, frame_id: clouds.append((pcd, label, frame_id))
    assert sm.map is None
    drive(sm, case, cloud_of, image_of)
    assert sm.done and sm.save_map_to_file
    assert sha(sm.map) == case.spec["filtered_sha"]                  # self.map = apply_filter(self.map)
    assert np.array_equal(sm.color_map, case.arrays["rgb"]) and len(shown) == 1 and shown[0] is sm.color_map
    assert os.path.exists(os.path.join(sm.output_dir, "global_map.png"))
    if variant == "with_point_cloud_consumer":
        assert len(clouds) == case.spec["frames"]
        for (pcd, label, frame_id), out in zip(clouds, case.spec["frames_out"]):
            assert pcd.is_cuda and frame_id == "world" and pcd.shape[1] == out["M"]
            assert sha(pcd.cpu().numpy()) == out["masked_pcd_sha"] and sha(label.cpu().numpy()) == out["label_sha"]
    # the recorded drive replays to the same map
    back = replay_io.load_input_list(os.path.join(sm.input_dir, "input_list.npz"))
    assert len(back) == case.spec["frames"]
    assert ("semantic_ids" in back[0]) == (variant == "class_ids")
    again = SemanticMapping(make_cfg(tmp_path / "again", case, False))
    if variant == "class_ids":
        again.set_label_palette(syn.COLORS_19)
    assert np.array_equal(again.mapping_replay(back, "again", write_image=False), case.arrays["rgb"])


def test_planar_update_is_the_references_clamp(tmp_path):
    """cfg.MAPPING.DEPTH_METHOD other than points_map / points_raw: the reference's update_map_planar
    (src/mapping.py:446-488) compares the warped uint8 image with label NAMES, so -- probed on the unmodified reference,
    SURVEY.md 8a A7 -- no cell is ever incremented and the only effect is map_local[map_local < 0] = 0."""
    case = Case("cfg1_c5_count")
    sm = SemanticMapping(make_cfg(tmp_path, case, False))
    rng = np.random.default_rng(4)
    grid = rng.normal(size=(sm.map_height, sm.map_width, sm.map_depth))
    grid[0, 0, 0], grid[0, 0, 1], grid[0, 1, 0] = -0.0, np.nan, -np.inf
    want = grid.copy()
    want[want < 0] = 0                                   # the statement of the reference, on the same array
    image = syn.synthetic_frame(1, 0, 10)["semantic_image"]
    got = grid.copy()
    ret = sm.update_map_planar(got, image, sm.cam1)
    assert ret is got and np.array_equal(got, want, equal_nan=True) and np.signbit(got[0, 0, 0])
    dev = torch.from_numpy(grid).cuda()
    assert sm.update_map_planar(dev, image, sm.cam1) is dev
    assert np.array_equal(dev.cpu().numpy(), want, equal_nan=True)
    # the live entry point: a planar frame changes nothing in a fresh map and is not recorded
    sm.depth_method = "planar"
    sm.mapping(image, syn.synthetic_pose(0), sm.cam1)
    assert sm.input_list == [] and not np.any(sm.map)
