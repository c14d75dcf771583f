"""GPU: the reference-facing Python API (SemanticMapping, renderer functions) with numpy in / numpy out,
checked against the golden vectors of the real reference."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from tests.common import Case, sha  # noqa: E402
from vision_semantic_segmentation_b200 import replay_io, synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.config.base_cfg import get_cfg_defaults  # noqa: E402
from vision_semantic_segmentation_b200.mapping_replay import SemanticMapping, main  # noqa: E402
from vision_semantic_segmentation_b200 import renderer  # noqa: E402


def make_cfg(tmp_path, case, log_cm):
    cfg = get_cfg_defaults()
    cfg.OUTPUT_DIR = str(tmp_path / "out")
    cfg.LABELS, cfg.LABELS_NAMES, cfg.LABEL_COLORS = syn.class_setup(case.spec["full19"])
    if log_cm:
        path = str(tmp_path / "cm.npy")
        np.save(path, syn.synthetic_confusion_matrix(case.spec.get("cm_seed", 7)))
        cfg.MAPPING.CONFUSION_MTX.LOAD_PATH = path
    return cfg


def frame_dicts(case, reference_style):
    out = []
    for f in range(case.spec["frames"]):
        fr = syn.synthetic_frame(case.spec["seed"], f, case.spec["n_points"], blocky=(f in case.spec.get("blocky_frames", [])))
        if reference_style:
            fr.pop("points")   # exactly what the reference's live node records: float64 (4, N) pcd
        out.append(fr)
    return out


@pytest.mark.parametrize("name,log_cm", [("cfg1_c5_count", False), ("cfg1_c19_log", True)])
def test_project_pcd_and_update_map_numpy_roundtrip(tmp_path, name, log_cm):
    case = Case(name)
    sm = SemanticMapping(make_cfg(tmp_path, case, log_cm))
    assert (sm.map_height, sm.map_width, sm.map_depth) == tuple(case.spec["map_shape"])
    assert np.array_equal(np.asarray(sm.confusion_matrix, dtype=np.float64), case.cm)
    assert sm.map is None
    grid = np.zeros((sm.map_height, sm.map_width, sm.map_depth))
    for f, (fr, out) in enumerate(zip(frame_dicts(case, True), case.spec["frames_out"])):
        masked, label = sm.project_pcd(fr["pcd"], fr["pcd_frame_id"], fr["semantic_image"], fr["pose"], sm.cam1)
        assert isinstance(masked, np.ndarray) and masked.dtype == np.float64 and label.dtype == np.uint8
        assert sha(masked) == out["masked_pcd_sha"] and sha(label) == out["label_sha"]
        ret = sm.update_map(grid, masked, label)
        assert ret is grid and sha(grid) == out["map_sha_after"]
    assert sm.project_pcd(None, "world", None, None, sm.cam1) is None


@pytest.mark.parametrize("name,log_cm,reference_style", [("cfg1_c5_count", False, True), ("cfg1_c5_count", False, False),
                                                         ("cfg1_c19_log", True, False)])
def test_mapping_replay_end_to_end(tmp_path, name, log_cm, reference_style):
    case = Case(name)
    sm = SemanticMapping(make_cfg(tmp_path, case, log_cm))
    color_map = sm.mapping_replay(frame_dicts(case, reference_style), "golden")
    assert np.array_equal(color_map, case.arrays["rgb"])
    assert sha(sm.map) == case.spec["filtered_sha"]            # self.map = apply_filter(self.map), as the reference
    png = os.path.join(sm.output_dir, "global_map_golden.png")
    assert os.path.exists(png)
    import cv2
    assert np.array_equal(cv2.imread(png), color_map)          # cv2 round trip (BGR in, BGR out)
    # a second replay starts from a fresh grid
    assert np.array_equal(sm.mapping_replay(frame_dicts(case, reference_style), "again", write_image=False), case.arrays["rgb"])


def test_replay_dir_reads_npz_records(tmp_path):
    case = Case("cfg1_c5_count")
    cfg = make_cfg(tmp_path, case, False)
    rec = tmp_path / "records"
    rec.mkdir()
    replay_io.save_input_list(str(rec / "input_list_0.npz"), frame_dicts(case, True))
    cfg.MAPPING.INPUT_DIR = str(rec)
    sm = SemanticMapping(cfg)
    sm.mapping_replay_dir()
    assert sha(sm.map) == case.spec["filtered_sha"]
    assert os.path.exists(os.path.join(sm.output_dir, "global_map_input_list_0.png"))
    # the command line entry point: --cfg FILE
    ypath = tmp_path / "run.yaml"
    ypath.write_text("OUTPUT_DIR: %s\nMAPPING:\n  INPUT_DIR: %s\n" % (tmp_path / "cli_out", rec))
    main(["--cfg", str(ypath)])
    assert os.path.exists(str(tmp_path / "cli_out" / "version_0" / "global_map_input_list_0.png"))


def test_map_setter_switches_to_ordered_update(tmp_path):
    """Assigning a non-integer grid must not break exactness of a following count update."""
    case = Case("cfg1_c5_count")
    sm = SemanticMapping(make_cfg(tmp_path, case, False))
    rng = np.random.default_rng(0)
    start = rng.uniform(0, 1, (sm.map_height, sm.map_width, sm.map_depth)) * 1e-3 + 0.1
    sm.map = start.copy()
    fr = frame_dicts(case, False)[0]
    sm.integrate_frame(fr)
    from oracle import c_oracle
    want = start.copy()
    pcd, _, image, T = case.frame(0)
    mp, lab, _, _ = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
    c_oracle.update_map(want, mp, lab, case.colors, case.cm, case.boundary, case.resolution, True, case.lane)
    assert np.array_equal(sm.map, want)


def test_renderer_functions_numpy_and_errors():
    rng = np.random.default_rng(1)
    m = rng.integers(0, 4, (50, 60, 5)).astype(np.float64)
    colors = syn.COLORS_19[:5]
    from oracle import c_oracle
    assert np.array_equal(renderer.render_bev_map(m, colors), c_oracle.render_bev_map(m, colors))
    assert np.array_equal(renderer.apply_filter(m), c_oracle.apply_filter(m))
    pr, th = [3, 4, 0, 2, 1], [0.1, 0.1, 0.5, 0.2, 0.05]
    assert np.array_equal(renderer.render_bev_map_with_thresholds(m, colors, priority=pr, thresholds=th),
                          c_oracle.render_bev_map_with_thresholds(m, colors, pr, th))
    with pytest.raises(ValueError):
        renderer.render_bev_map(m, colors[:4])
    with pytest.raises(ValueError):
        renderer.render_bev_map_with_thresholds(m, colors, priority=[0, 1])
    with pytest.raises(IndexError):
        renderer.render_bev_map_with_thresholds(np.zeros((4, 4, 7)), syn.COLORS_19[:7])
