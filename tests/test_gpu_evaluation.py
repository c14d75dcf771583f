"""GPU: the evaluation step -- smap_eval_counts against what the REFERENCE's own ``convert_labels`` +
``Test.iou`` / ``test_single_map`` (test/test_semantic_mapping.py:6-18,117-161) returned, printed and raised on the
same seeded inputs (tests/golden/eval.json, written by oracle/make_golden_eval.py from the unmodified file), and
against the host-array form of the same functions.  Integer sums, so every score must be the same float, not merely
close."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from tests.common import Case  # noqa: E402
from vision_semantic_segmentation_b200 import evaluation as ev  # noqa: E402

CLASS_COLORS = np.array([[0, 0, 0], [128, 64, 128], [140, 140, 200], [255, 255, 255], [244, 35, 232], [107, 142, 35],
                         [128, 64, 129], [255, 255, 0], [0, 60, 100]], dtype=np.uint8)   # 6..8: not evaluation classes


def host_scores(rgb, truth, shift_w, shift_h, mask=None):
    t = ev.Test.__new__(ev.Test)
    t.class_lists, t.d, t.logger = [1, 2, 3], {0: "road", 1: "crosswalk", 2: "lane"}, None
    generated = ev.convert_labels(rgb, mask)
    gmap = truth[shift_w:generated.shape[0] + shift_w, shift_h:generated.shape[1] + shift_h]
    return t.iou(gmap, generated)


@pytest.mark.parametrize("shape,truth_shape,shift,with_mask", [
    ((2000, 2000), (2100, 2300), (37, 91), False), ((2000, 2000), (2000, 2000), (0, 0), True),
    ((3, 5), (4, 9), (1, 4), False), ((1, 1), (1, 1), (0, 0), False), ((333, 1027), (400, 1100), (67, 0), True)])
def test_device_counts_match_numpy_iou(shape, truth_shape, shift, with_mask):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    rgb = CLASS_COLORS[rng.integers(0, len(CLASS_COLORS), shape)]
    truth = rng.integers(0, 4, truth_shape).astype(np.float64)   # truth.npy is a float64 label map in the reference
    mask = None
    if with_mask:
        mask = (rng.uniform(size=(shape[0] + 3, shape[1] + 2)) > 0.3).astype(np.float64)   # larger than the map, as mask.npy may be
    for color_map in (rgb, torch.from_numpy(rgb).cuda()):
        counts = ev.device_counts(color_map, truth, shift[0], shift[1], mask=mask)
        try:
            want_ious, want_miss = host_scores(rgb, truth, shift[0], shift[1], mask)
        except ZeroDivisionError:
            # a class absent from both maps (the 1 x 1 case): the reference's Test.iou raises, and so does the device path
            with pytest.raises(ZeroDivisionError):
                ev.scores_from_counts(counts)
            continue
        ious, accs, accuracy, miss = ev.scores_from_counts(counts)
        assert all(_same(a, b) for a, b in zip(ious, want_ious))
        assert miss == want_miss
    generated = ev.convert_labels(rgb, mask)
    gmap = truth[shift[0]:shape[0] + shift[0], shift[1]:shape[1] + shift[1]]
    assert counts[9] == int(np.sum(gmap > 0)) and counts[11] == int(np.sum((gmap == generated)[gmap > 0]))
    assert counts[6:9] == [int(np.sum(generated == k)) for k in (1, 2, 3)]


def _dec(v):
    return float(v) if isinstance(v, str) else v


def _same(a, b):
    return a == b or (a != a and b != b)


@pytest.mark.parametrize("name", ["grid_2000", "grid_2000_mask", "small", "odd", "no_crosswalk", "no_truth",
                                  "golden_render"])
def test_device_counts_match_the_reference(name):
    """smap_eval_counts -> scores == what the reference's own functions produced (tests/golden/eval.json)."""
    import json
    import os
    from oracle.make_golden_eval import make_inputs
    from tests.common import GOLDEN
    with open(os.path.join(GOLDEN, "eval.json")) as f:
        spec = json.load(f)["cases"][name]
    rgb, truth, mask = make_inputs(name, spec)
    counts = ev.device_counts(torch.from_numpy(rgb).cuda(), truth, spec["shift"][0], spec["shift"][1], mask=mask)
    # convert_labels: the label histogram of the converted map (classes 1..3 are counted by the kernel)
    assert counts[6:9] == spec["labels_hist"][1:4]
    if spec.get("raises") == "ZeroDivisionError":
        with pytest.raises(ZeroDivisionError):
            ev.scores_from_counts(counts)
        return
    ious, accs, accuracy, miss = ev.scores_from_counts(counts)
    assert all(_same(a, _dec(b)) for a, b in zip(ious, spec["iou"])), (ious, spec["iou"])
    assert _same(miss, _dec(spec["miss"]))
    # the accuracies only exist in the line the reference prints ('{}'.format of the float: repr, round-trips)
    assert all(_same(a, _dec(b)) for a, b in zip(accs, spec["acc"])), (accs, spec["acc"])
    assert _same(accuracy, _dec(spec["accuracy"]))
    if mask is None:
        # Test.test_single_map end to end: the very lines the reference printed
        t = ev.Test.__new__(ev.Test)
        t.class_lists, t.d, t.logger = [1, 2, 3], {0: "road", 1: "crosswalk", 2: "lane"}, None
        t.shift_w, t.shift_h = spec["shift"]
        t.ground_truth_mask, t.mask = truth, None
        said = []
        t._say = said.append
        t.test_single_map(torch.from_numpy(rgb).cuda())
        assert said == spec["printed"]


def test_device_counts_errors():
    rgb = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(ValueError, match="inside"):
        ev.device_counts(rgb, np.zeros((4, 4)), 1, 0)
    with pytest.raises(ValueError, match="integers"):
        ev.device_counts(rgb, np.full((4, 4), 0.5))
    with pytest.raises(ValueError):
        ev.device_counts(np.zeros((4, 4), np.uint8), np.zeros((4, 4)))


def test_mapping_replay_scores_the_rendered_map(tmp_path):
    """cfg.GROUND_TRUTH_DIR set: mapping_replay scores the map it rendered (src/mapping_replay.py:208-210); the logged
    numbers are those of the numpy path on the golden render of the real reference."""
    from tests.test_gpu_api import frame_dicts, make_cfg
    from vision_semantic_segmentation_b200.mapping_replay import SemanticMapping
    case = Case("cfg1_c5_count")
    rng = np.random.default_rng(3)
    truth = rng.integers(0, 4, (case.mh, case.mw)).astype(np.float64)
    gt_dir = tmp_path / "gt"
    gt_dir.mkdir()
    np.save(str(gt_dir / "truth.npy"), truth)
    cfg = make_cfg(tmp_path, case, False)
    cfg.GROUND_TRUTH_DIR = str(gt_dir)
    sm = SemanticMapping(cfg)
    said = []
    sm.logger.log = lambda msg: said.append(msg)
    color_map = sm.mapping_replay(frame_dicts(case, False), "scored", write_image=False)
    assert np.array_equal(color_map, case.arrays["rgb"])
    want_ious, want_miss = host_scores(case.arrays["rgb"], truth, 0, 0)
    assert any(line.startswith("IOU for road: %s" % want_ious[0]) for line in said), said
    assert "Overall Missing rate: %s" % want_miss in said
