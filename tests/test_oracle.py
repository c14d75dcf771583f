"""CPU: the oracle (C restatement) against the golden vectors produced by the real reference,
and against numpy for the pieces whose order of operations matters."""
import numpy as np
import pytest

from oracle import c_oracle, numpy_port
from tests.common import Case, GOLDEN_CASES, sha


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    case = Case(name)
    grid = np.zeros((case.mh, case.mw, case.c))
    for f, out in enumerate(case.spec["frames_out"]):
        pcd, _, image, T = case.frame(f)
        masked, label, uv, keep = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
        assert masked.shape[1] == out["M"]
        assert np.array_equal(np.flatnonzero(keep), case.arrays["keep_idx_%d" % f])
        assert sha(masked) == out["masked_pcd_sha"]
        assert sha(label) == out["label_sha"]
        assert sha(uv) == out["uv_sha"]
        c_oracle.update_map(grid, masked, label, case.colors, case.cm, case.boundary, case.resolution,
                            case.use_intensity, case.lane)
        assert sha(grid) == out["map_sha_after"], "frame %d" % f
    assert sha(grid) == case.spec["map_sha"]
    if "map_idx" in case.arrays.files:
        assert np.array_equal(grid, case.sparse_map())
    filtered = c_oracle.apply_filter(grid)
    assert sha(filtered) == case.spec["filtered_sha"]
    rgb = c_oracle.render_bev_map(filtered, case.colors)
    assert np.array_equal(rgb, case.arrays["rgb"]) and sha(rgb) == case.spec["rgb_sha"]
    assert sha(c_oracle.render_bev_map(grid, case.colors)) == case.spec["rgb_raw_sha"]
    thr = c_oracle.render_bev_map_with_thresholds(grid, case.colors, case.spec["priority"], case.spec["thresholds"])
    assert np.array_equal(thr, case.arrays["rgb_thr"]) and sha(thr) == case.spec["rgb_thr_sha"]


def test_numpy_port_matches_c_oracle_small():
    case = Case("small_velodyne")
    g1 = np.zeros((case.mh, case.mw, case.c))
    g2 = g1.copy()
    for f in range(case.spec["frames"]):
        pcd, _, image, T = case.frame(f)
        a = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
        b = numpy_port.project_pcd(pcd, T, case.cam.P, image, case.range_max)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        c_oracle.update_map(g1, a[0], a[1], case.colors, case.cm, case.boundary, case.resolution,
                            case.use_intensity, case.lane)
        numpy_port.update_map(g2, b[0], b[1], case.colors, case.cm, case.boundary, case.resolution,
                              case.use_intensity, case.names)
    assert np.array_equal(g1, g2)
    assert np.array_equal(c_oracle.apply_filter(g1), numpy_port.apply_filter(g2))
    assert np.array_equal(c_oracle.render_bev_map(g1, case.colors), numpy_port.render_bev_map(g2, case.colors))


@pytest.mark.parametrize("c", [1, 2, 5, 8, 9, 10, 16, 17, 19, 25, 31, 32])
def test_class_axis_sum_and_argmax_follow_numpy(c):
    rng = np.random.default_rng(c)
    # values whose sum depends on the order of additions, with exact cancellations and NaN / inf sprinkled in
    m = rng.choice([1e16, -1e16, 1.0, -1.0, 3.0, 0.1, -0.1, 0.0, 2.5e-7], size=(40, 50, c))
    m[0, 0, :] = 0.0
    m[1, 1, c // 2] = np.nan
    m[2, 2, c - 1] = np.inf
    m[3, 3, 0] = -np.inf
    colors = rng.integers(0, 256, (c, 3))
    with np.errstate(all="ignore"):
        want_sum = np.sum(m, axis=2)
        for i in range(40):
            for j in range(0, 50, 7):
                got = c_oracle.np_sum(m[i, j])
                assert got == want_sum[i, j] or (np.isnan(got) and np.isnan(want_sum[i, j]))
        assert np.array_equal(c_oracle.render_bev_map(m, colors), numpy_port.render_bev_map(m, colors))
        pr = rng.permutation(c)
        th = rng.uniform(-0.5, 0.5, c)
        assert np.array_equal(c_oracle.render_bev_map_with_thresholds(m, colors, pr, th),
                              numpy_port.render_bev_map_with_thresholds(m, colors, pr, th))


def test_filter_matches_opencv_on_general_data():
    rng = np.random.default_rng(5)
    for shape in [(1, 1, 3), (1, 7, 2), (6, 1, 4), (2, 2, 2), (2, 3, 7), (33, 65, 5), (40, 31, 19)]:
        src = rng.normal(0, 1e3, shape) * rng.choice([1.0, 1e-9, 1e9], shape)
        assert np.array_equal(c_oracle.apply_filter(src), numpy_port.apply_filter(src)), shape


def test_empty_cloud_and_out_of_grid():
    case = Case("cfg1_c5_count")
    image = np.zeros((1440, 1920, 3), np.uint8)
    empty = np.zeros((4, 0))
    masked, label, uv, keep = c_oracle.project_pcd(empty, np.eye(4), case.cam.P, image, 100.0)
    assert masked.shape == (4, 0) and label.shape == (3, 0)
    grid = np.zeros((case.mh, case.mw, case.c))
    st = c_oracle.update_map(grid, masked, label, case.colors, case.cm, case.boundary, 0.1, True, case.lane)
    assert not grid.any() and st[1] == 0
    # truncation toward zero: a point 0.05 m below the boundary still lands in cell 0
    pcd = np.array([[100 - 1369.0496826171875 - 0.05], [800 - 562.84814453125 + 0.05], [0.0], [0.0]])
    lab = np.array([[128], [64], [128]], np.uint8)
    c_oracle.update_map(grid, pcd, lab, case.colors, case.cm, case.boundary, 0.1, True, case.lane)
    assert grid[0, 0, 0] == 1.0 and grid.sum() == 1.0
    # non-finite coordinates are dropped
    bad = np.array([[np.nan, np.inf, -np.inf], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    _, _, _, keep = c_oracle.project_pcd(bad, np.eye(4), case.cam.P, image, 100.0)
    assert not keep.any()
