"""GPU: the CUDA path (through the C ABI) against the golden vectors of the real reference and
against the C oracle on the same seeded inputs.  Bit-exact everywhere (integer indices, labels,
count grids, log-likelihood grids, rendered images)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import c_oracle  # noqa: E402
from tests.common import Case, GOLDEN_CASES, sha  # noqa: E402
from vision_semantic_segmentation_b200 import synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper  # noqa: E402
from vision_semantic_segmentation_b200 import renderer  # noqa: E402
from vision_semantic_segmentation_b200.utils import transforms as tr  # noqa: E402


def make_mapper(case):
    return DeviceMapper(case.mh, case.mw, case.colors, case.cm, case.boundary, case.resolution, case.range_max,
                        case.use_intensity, case.lane, cameras=[case.cam], device=0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("layout,ordered", [("f32x4", False), ("f64soa", False), ("f32x4", True)])
def test_fused_path_matches_reference_golden(name, layout, ordered):
    """ordered=True forces the two-kernel ordered update also for the count update (which otherwise adds its
    increments with float64 atomics when the grid holds integer counts)."""
    case = Case(name)
    dm = make_mapper(case)
    if ordered:
        dm.notify_map_modified()
    for f, out in enumerate(case.spec["frames_out"]):
        pcd, points, image, T = case.frame(f)
        cloud = dev(points) if layout == "f32x4" else dev(pcd)
        dm.integrate(dm.make_frame(cloud, dev(image), T, 0))
        assert sha(dm.map.cpu().numpy()) == out["map_sha_after"], "frame %d" % f
    grid = dm.map.cpu().numpy()
    assert sha(grid) == case.spec["map_sha"]
    # fused filter + render, and the separate kernels
    rgb, filtered = renderer.filter_and_render(dm.map, case.colors, return_filtered=True)
    assert sha(filtered.cpu().numpy()) == case.spec["filtered_sha"]
    assert np.array_equal(rgb.cpu().numpy(), case.arrays["rgb"])
    assert sha(renderer.apply_filter(dm.map).cpu().numpy()) == case.spec["filtered_sha"]
    assert sha(renderer.render_bev_map(filtered, case.colors).cpu().numpy()) == case.spec["rgb_sha"]
    assert sha(renderer.render_bev_map(dm.map, case.colors).cpu().numpy()) == case.spec["rgb_raw_sha"]
    thr = renderer.render_bev_map_with_thresholds(dm.map, case.colors, case.spec["priority"], case.spec["thresholds"])
    assert np.array_equal(thr.cpu().numpy(), case.arrays["rgb_thr"])
    dm.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_parity_api_project_and_update(name):
    case = Case(name)
    dm = make_mapper(case)
    for f, out in enumerate(case.spec["frames_out"]):
        pcd, points, image, T = case.frame(f)
        for cloud in (dev(points), dev(pcd)):
            masked, label, uv, keep = dm.project(dm.make_frame(cloud, dev(image), T, 0), want_uv=True, want_keep=True)
            assert masked.shape[1] == out["M"]
            assert np.array_equal(np.flatnonzero(keep.cpu().numpy()), case.arrays["keep_idx_%d" % f])
            assert sha(masked.cpu().numpy()) == out["masked_pcd_sha"]
            assert sha(label.cpu().numpy()) == out["label_sha"]
            assert sha(uv.cpu().numpy()) == out["uv_sha"]
        dm.update(masked, label)
        assert sha(dm.map.cpu().numpy()) == out["map_sha_after"]
    dm.close()


def test_edge_cases_empty_nonfinite_truncation():
    case = Case("cfg1_c5_count")
    dm = make_mapper(case)
    image = dev(np.zeros((1440, 1920, 3), np.uint8))
    # empty cloud, both layouts
    for cloud in (torch.empty((0, 4), dtype=torch.float32, device="cuda"),
                  torch.empty((4, 0), dtype=torch.float64, device="cuda")):
        fr = dm.make_frame(cloud, image, np.eye(4), 0)
        dm.integrate(fr)
        masked, label = dm.project(fr)
        assert masked.shape == (4, 0) and label.shape == (3, 0)
        dm.update(masked, label)
    assert not dm.map.any().item()
    # truncation toward zero keeps a point 0.05 m below the boundary in cell 0; non-finite points vanish
    pcd = np.array([[100 - 1369.0496826171875 - 0.05, np.nan, np.inf], [800 - 562.84814453125 + 0.05, 0.0, 0.0],
                    [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    lab = np.array([[128, 128, 128], [64, 64, 64], [128, 128, 128]], np.uint8)
    dm.update(dev(pcd), dev(lab))
    got = dm.map.cpu().numpy()
    assert got[0, 0, 0] == 1.0 and got.sum() == 1.0
    bad = np.array([[np.nan, np.inf, -np.inf, 1e300], [0.0, 0.0, 0.0, 1e300], [0.0, 0.0, 0.0, -1e300], [0.0] * 4])
    _, _, keep = dm.project(dm.make_frame(dev(bad), image, np.eye(4), 0), want_keep=True)
    assert not keep.any().item()
    dm.close()


@pytest.mark.parametrize("c", [2, 5, 8, 9, 16, 19, 31, 32])
def test_render_kernels_follow_numpy_order_on_general_data(c):
    rng = np.random.default_rng(100 + c)
    m = rng.choice([1e16, -1e16, 1.0, -1.0, 3.0, 0.1, -0.1, 0.0, 2.5e-7], size=(37, 71, c))
    m[0, 0, :] = 0.0
    m[1, 1, c // 2] = np.nan
    m[2, 2, c - 1] = np.inf
    m[3, 3, 0] = -np.inf
    colors = rng.integers(0, 256, (c, 3))
    pr, th = rng.permutation(c), rng.uniform(-0.5, 0.5, c)
    md = dev(m)
    with np.errstate(all="ignore"):
        assert np.array_equal(renderer.render_bev_map(md, colors).cpu().numpy(), c_oracle.render_bev_map(m, colors))
        assert np.array_equal(renderer.render_bev_map_with_thresholds(md, colors, pr, th).cpu().numpy(),
                              c_oracle.render_bev_map_with_thresholds(m, colors, pr, th))
        smooth = rng.normal(0, 1e3, (37, 71, c)) * rng.choice([1.0, 1e-9, 1e9], (37, 71, c))
        want = c_oracle.apply_filter(smooth)
        assert np.array_equal(renderer.apply_filter(dev(smooth)).cpu().numpy(), want)
        rgb, filt = renderer.filter_and_render(dev(smooth), colors, return_filtered=True)
        assert np.array_equal(filt.cpu().numpy(), want)
        assert np.array_equal(rgb.cpu().numpy(), c_oracle.render_bev_map(want, colors))


@pytest.mark.parametrize("shape", [(5, 2, 5), (40, 34, 5), (70, 100, 5), (33, 64, 19), (67, 130, 7), (35, 96, 31),
                                   (130, 66, 19), (3, 32, 1)])
def test_render_bulk_staged_tiles(shape):
    """Grids the bulk-copy staged render kernel takes (odd number of classes, even number of columns): several tiles
    in both directions, right-edge tiles of 2 / 4 / 32 columns, bottom tiles of a few rows, NaN / inf / signed zeros."""
    rng = np.random.default_rng(sum(shape))
    c = shape[2]
    m = rng.normal(0, 1e3, shape) * rng.choice([1.0, 1e-9, 1e9, 0.0], shape)
    m[rng.random(shape) < 0.02] = -0.0
    m[0, 0, :] = 0.0
    m[shape[0] // 2, shape[1] // 2, c // 2] = np.nan
    m[shape[0] - 1, shape[1] - 1, c - 1] = np.inf
    m[0, shape[1] - 1, 0] = -np.inf
    colors = rng.integers(0, 256, (c, 3))
    with np.errstate(all="ignore"):
        want = c_oracle.apply_filter(m)
        rgb, filt = renderer.filter_and_render(dev(m), colors, return_filtered=True)
        assert np.array_equal(filt.cpu().numpy(), want, equal_nan=True)
        assert np.array_equal(np.signbit(filt.cpu().numpy()), np.signbit(want))
        assert np.array_equal(rgb.cpu().numpy(), c_oracle.render_bev_map(want, colors))
        assert np.array_equal(renderer.filter_and_render(dev(m), colors).cpu().numpy(), c_oracle.render_bev_map(want, colors))
        assert np.array_equal(renderer.render_bev_map(dev(m), colors).cpu().numpy(), c_oracle.render_bev_map(m, colors))
        # a view that starts 8 bytes into an allocation: not 16-byte aligned, the register-staged kernel takes it
        flat = torch.empty(m.size + 1, dtype=torch.float64, device="cuda")
        view = flat[1:].view(shape)
        view.copy_(dev(m))
        assert np.array_equal(renderer.filter_and_render(view, colors).cpu().numpy(), c_oracle.render_bev_map(want, colors))


@pytest.mark.parametrize("shape", [(1, 1, 3), (1, 7, 2), (6, 1, 4), (2, 2, 2), (8, 32, 5), (9, 33, 5), (64, 100, 19)])
def test_filter_borders(shape):
    rng = np.random.default_rng(7)
    src = rng.normal(0, 10, shape)
    assert np.array_equal(renderer.apply_filter(dev(src)).cpu().numpy(), c_oracle.apply_filter(src))


@pytest.mark.parametrize("full19,log_cm,ordered", [(False, False, False), (True, False, False), (True, False, True),
                                                  (True, True, True)])
def test_full_size_frame_against_oracle(full19, log_cm, ordered):
    """BASELINE.json configs[1] shape: 2M-point cloud + 1920x1440 frame, checked in full against the oracle,
    plus size-independent properties (linearity of the count grid in the number of replays)."""
    labels, names, colors = syn.class_setup(full19)
    c = len(labels)
    cm = np.eye(c)
    if log_cm:
        from oracle import numpy_port
        cm = numpy_port.confusion_submatrix_log(syn.synthetic_confusion_matrix(7), labels)
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    cam = camera_setup_1()
    boundary, res, mh, mw = [[100, 300], [800, 1000]], 0.1, 2000, 2000
    lane = names.index("lane")
    dm = DeviceMapper(mh, mw, colors, cm, boundary, res, 100.0, True, lane, cameras=[cam], device=0)
    if ordered:
        dm.notify_map_modified()
    ref = np.zeros((mh, mw, c))
    for f in range(2):
        fr = syn.synthetic_frame(1000, f, 2000000, blocky=(f == 1))
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        dm.integrate(dm.make_frame(dev(fr["points"]), dev(fr["semantic_image"]), T, 0))
        mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
        st = c_oracle.update_map(ref, mp, lab, colors, cm, boundary, res, True, lane)
        # K is counted by the ordered apply; the tagged count update of a float4 cloud replays nothing
        assert dm.stats()["touched_cells"] == (st[1] if ordered else -1)
        assert np.array_equal(dm.map.cpu().numpy(), ref), "frame %d" % f
    if not log_cm:
        # replaying the same two frames again doubles every count exactly
        for f in range(2):
            fr = syn.synthetic_frame(1000, f, 2000000, blocky=(f == 1))
            T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
            dm.integrate(dm.make_frame(dev(fr["points"]), dev(fr["semantic_image"]), T, 0))
        assert np.array_equal(dm.map.cpu().numpy(), 2.0 * ref)
    rgb = renderer.filter_and_render(dm.map, colors)
    want = c_oracle.render_bev_map(c_oracle.apply_filter(dm.map.cpu().numpy()), colors)
    assert np.array_equal(rgb.cpu().numpy(), want)
    dm.close()


def test_host_staged_integrate_matches_device_path():
    case = Case("cfg1_c19_count")
    a, b = make_mapper(case), make_mapper(case)
    keep = []
    for f in range(3):
        pcd, points, image, T = case.frame(f)
        a.integrate(a.make_frame(dev(points), dev(image), T, 0))
        hp, hi = torch.from_numpy(points).pin_memory(), torch.from_numpy(image).pin_memory()
        keep.append((hp, hi))
        b.integrate_host(b.make_frame(hp, hi, T, 0, host=True))
    torch.cuda.synchronize()
    assert sha(b.map.cpu().numpy()) == case.spec["map_sha"]
    assert torch.equal(a.map, b.map)
    a.close()
    b.close()


def _oracle_grid(case, frames):
    ref = np.zeros((case.mh, case.mw, case.c))
    for pcd, _, image, T in frames:
        mp, lab, _, _ = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
        c_oracle.update_map(ref, mp, lab, case.colors, case.cm, case.boundary, case.resolution,
                            case.use_intensity, case.lane)
    return ref


@pytest.mark.parametrize("name,ordered", [("cfg1_c5_count", False), ("cfg1_c19_count", False), ("cfg1_c19_count", True),
                                          ("cfg1_c19_log", True)])
def test_batched_launch_equals_sequential_frames(name, ordered):
    """Up to 16 frames share one launch (one mask slot each); repeating a frame inside a batch must count it
    twice, and 37 frames exercise full batches plus a ragged tail."""
    case = Case(name)
    dm = make_mapper(case)
    if ordered:
        dm.notify_map_modified()
    base = [case.frame(f) for f in range(3)]
    order = [0, 1, 2, 2, 0, 1, 1] * 5 + [2, 0]
    frames_dev, keep = [], []
    for i in order:
        pcd, points, image, T = base[i]
        dp, di = dev(points), dev(image)
        keep.append((dp, di))
        frames_dev.append(dm.make_frame(dp, di, T, 0))
    # an empty frame in the middle of a batch is legal
    empty = torch.empty((0, 4), dtype=torch.float32, device="cuda")
    frames_dev.insert(5, dm.make_frame(empty, keep[0][1], np.eye(4), 0))
    dm.integrate_batch(frames_dev)
    got = dm.map.cpu().numpy()
    want = _oracle_grid(case, [base[i] for i in order])
    assert np.array_equal(got, want)
    dm.close()


@pytest.mark.parametrize("name,ordered", [("cfg1_c5_count", False), ("cfg1_c19_count", False), ("cfg1_c5_count", True)])
def test_graph_and_plain_launches_agree(name, ordered, monkeypatch):
    """A chunk's k_fuse launches go out as one CUDA graph launch (re-parameterised per chunk) or, with
    SMAP_FUSE_GRAPH=0, as per-frame launches on the internal streams: same grids, chunk after chunk (the second and
    third call reuse and update the instantiated graphs; 5 frames and 21 frames need different ones)."""
    case = Case(name)
    base = [case.frame(f) for f in range(3)]
    grids = []
    for use_graph in ("1", "0"):
        monkeypatch.setenv("SMAP_FUSE_GRAPH", use_graph)
        dm = make_mapper(case)
        if ordered:
            dm.notify_map_modified()
        keep = []
        got = []
        for order in ([0, 1, 2, 1, 0], [2, 2, 1, 0, 1, 2, 0] * 3, [1, 0, 2, 2, 1]):
            frames_dev = []
            for i in order:
                pcd, points, image, T = base[i]
                dp, di = dev(points), dev(image)
                keep.append((dp, di))
                frames_dev.append(dm.make_frame(dp, di, T, 0))
            dm.integrate_batch(frames_dev)
            got.append(dm.map.cpu().numpy().copy())
        grids.append(got)
        dm.close()
    want = _oracle_grid(case, [base[i] for i in [0, 1, 2, 1, 0] + [2, 2, 1, 0, 1, 2, 0] * 3 + [1, 0, 2, 2, 1]])
    for a, b in zip(*grids):
        assert np.array_equal(a, b)
    assert np.array_equal(grids[0][-1], want)


def test_many_classes_and_mask_slots_stay_clean():
    """28 classes (4 register chunks in the apply kernel, boost bit 28); 20 frames through the same mask slot:
    any word left uncleared by k_apply would leak into a later frame."""
    rng = np.random.default_rng(3)
    colors = np.zeros((28, 3), np.int64)
    colors[:19] = syn.COLORS_19
    colors[19:, 0] = np.arange(1, 10)  # colours that never occur: classes 19..27 stay empty
    colors[19:, 1] = 7
    names = list(syn.NAMES_19) + ["x%d" % i for i in range(9)]
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    cam = camera_setup_1()
    boundary, res, mh, mw = [[100, 300], [800, 1000]], 0.5, 400, 400
    lane = names.index("lane")
    for cm in (np.eye(28), -rng.uniform(0.1, 5.0, (28, 28))):
        dm = DeviceMapper(mh, mw, colors, cm, boundary, res, 100.0, True, lane, cameras=[cam], device=0)
        ref = np.zeros((mh, mw, 28))
        for f in range(20):
            fr = syn.synthetic_frame(77, f % 4, 20000, blocky=True)
            T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
            dm.integrate(dm.make_frame(dev(fr["points"]), dev(fr["semantic_image"]), T, 0))
            mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
            c_oracle.update_map(ref, mp, lab, colors, cm, boundary, res, True, lane)
        assert np.array_equal(dm.map.cpu().numpy(), ref)
        dm.close()


def test_precull_is_conservative_near_the_frustum_and_range_borders():
    """Points placed within float32 noise of every culling boundary must get the exact decision."""
    case = Case("cfg1_c19_count")
    dm = make_mapper(case)
    pcd, points, image, T = case.frame(0)
    masked, label, uv, keep = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
    # take kept points, move them along the viewing ray to just around range_max / just around x = 0, and
    # sideways to just around the image borders, by solving in the velodyne frame
    Tinv = np.linalg.inv(T)
    rng = np.random.default_rng(1)
    velo = (T @ np.vstack((pcd[0:3, keep][:, :20000], np.ones((1, min(20000, keep.sum()))))))
    out = []
    for target in (100.0, 1e-4, 99.9999, 0.0):
        v = velo.copy()
        scale = (target + rng.normal(0, 2e-5, v.shape[1])) / v[0]
        v[0:3] *= scale
        out.append(v)
    # lateral sweep: pixel columns within +-0.01 px of 0 and W
    P = case.cam.P
    for ucol in (0.0, -1.0, 1920.0, 1919.999):
        v = velo.copy()
        depth = v[0].copy()
        # choose y so that u == ucol (+ noise) given x, z:  u = (P0 . v) / (P2 . v)
        a = P[0, 1] - ucol * P[2, 1]
        b = (P[0, 0] - ucol * P[2, 0]) * v[0] + (P[0, 2] - ucol * P[2, 2]) * v[2] + (P[0, 3] - ucol * P[2, 3])
        v[1] = -b / a + rng.normal(0, 1e-5, v.shape[1]) * depth / 1800.0
        out.append(v)
    velo_all = np.hstack(out)
    world = (Tinv @ velo_all)[0:3].astype(np.float32)
    pts = np.ascontiguousarray(np.vstack((world, rng.uniform(0, 30, (1, world.shape[1])).astype(np.float32))).T)
    pcd2 = np.ascontiguousarray(pts.T.astype(np.float64))
    mp, lab, _, keep2 = c_oracle.project_pcd(pcd2, T, case.cam.P, image, case.range_max)
    assert 0.05 < keep2.mean() < 0.95  # the set really straddles the boundaries
    ref = np.zeros((case.mh, case.mw, case.c))
    c_oracle.update_map(ref, mp, lab, case.colors, case.cm, case.boundary, case.resolution, True, case.lane)
    for cloud in (dev(pts), dev(pcd2)):
        dm.clear()
        dm.integrate(dm.make_frame(cloud, dev(image), T, 0))
        assert np.array_equal(dm.map.cpu().numpy(), ref)
    dm.close()


@pytest.mark.parametrize("offset,hw", [(1, (1440, 1920)), (4, (1440, 1920)), (0, (1439, 1917)), (3, (1439, 1917))])
def test_label_image_alignment_and_odd_sizes(offset, hw):
    """The fused kernel reads R and G with one aligned 8-byte load when the image allows it and with byte loads
    otherwise (base not 8-byte aligned, size not a multiple of 8); both must give the reference's labels, for the
    tagged (5 classes) and the masked (19 classes) count update."""
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    cam = camera_setup_1()
    boundary, res, mh, mw = [[100, 300], [800, 1000]], 0.1, 2000, 2000
    for full19 in (False, True):
        labels, names, colors = syn.class_setup(full19)
        c = len(labels)
        lane = names.index("lane")
        dm = DeviceMapper(mh, mw, colors, np.eye(c), boundary, res, 100.0, True, lane, cameras=[cam], device=0)
        ref = np.zeros((mh, mw, c))
        for f in range(2):
            fr = syn.synthetic_frame(55, f, 150000, height=hw[0], width=hw[1], blocky=(f == 1))
            T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
            image = fr["semantic_image"]
            buf = torch.zeros(image.size + 16, dtype=torch.uint8, device="cuda")
            view = buf[offset:offset + image.size].view(hw[0], hw[1], 3)
            view.copy_(torch.from_numpy(image))
            assert view.data_ptr() % 8 == offset % 8
            dm.integrate(dm.make_frame(dev(fr["points"]), view, T, 0))
            mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, image, 100.0)
            c_oracle.update_map(ref, mp, lab, colors, np.eye(c), boundary, res, True, lane)
            assert np.array_equal(dm.map.cpu().numpy(), ref), "frame %d" % f
        dm.close()


def test_points_on_cell_boundaries():
    """Kept points snapped to within float32 noise of BEV cell edges -- including the grid's own edges and the
    (-1, 0) strip that truncates to cell 0 -- must land in the reference's cells: the float32 cell decision has to
    hand exactly these to the float64 / exact path."""
    case = Case("cfg1_c19_count")
    dm = make_mapper(case)
    pcd, points, image, T = case.frame(0)
    _, _, _, keep = c_oracle.project_pcd(pcd, T, case.cam.P, image, case.range_max)
    rng = np.random.default_rng(5)
    base = pcd[:, keep][:, :30000].copy()
    off = np.array([1369.0496826171875, 562.84814453125])
    b0 = np.array([case.boundary[0][0], case.boundary[1][0]], dtype=np.float64)
    clouds = []
    for axis in (0, 1):
        v = base.copy()
        g = ((v[axis] + off[axis]) - b0[axis]) / case.resolution
        k = np.rint(g)
        # a third of the points go to the grid's edges: -1, 0 and the last cell's upper edge
        edge = rng.choice([-1.0, 0.0, float(case.mh if axis == 0 else case.mw)], size=k.shape)
        k = np.where(rng.random(k.shape) < 0.33, edge, k)
        x = k * case.resolution + b0[axis] - off[axis]
        x32 = x.astype(np.float32)
        # the float32 value next to the edge, one ulp below, one ulp above
        step = rng.integers(-1, 2, size=x32.shape)
        x32 = np.where(step < 0, np.nextafter(x32, np.float32(-np.inf)), np.where(step > 0, np.nextafter(x32, np.float32(np.inf)), x32))
        v[axis] = x32.astype(np.float64)
        clouds.append(v)
    pcd2 = np.ascontiguousarray(np.hstack(clouds))
    pts = np.ascontiguousarray(pcd2.T.astype(np.float32))
    assert np.array_equal(pts.T.astype(np.float64), pcd2)
    mp, lab, _, keep2 = c_oracle.project_pcd(pcd2, T, case.cam.P, image, case.range_max)
    assert keep2.mean() > 0.2
    ref = np.zeros((case.mh, case.mw, case.c))
    c_oracle.update_map(ref, mp, lab, case.colors, case.cm, case.boundary, case.resolution, True, case.lane)
    for ordered in (False, True):
        dm.clear()
        if ordered:
            dm.notify_map_modified()
        dm.integrate(dm.make_frame(dev(pts), dev(image), T, 0))
        assert np.array_equal(dm.map.cpu().numpy(), ref), "ordered=%s" % ordered
    dm.close()


def test_cfg3_shape_log_likelihood_on_a_2km_grid():
    """BASELINE.json configs[2] shape: confusion-matrix log-likelihood update into a 0.2 m, 2 km x 2 km grid
    (10^4 x 10^4 x 5 float64 = 4 GB), then the count update on the same grid size (tag planes of 2.4 GB each);
    whole grids compared with the oracle."""
    from oracle import numpy_port
    from vision_semantic_segmentation_b200.camera import camera_setup_1
    labels, names, colors = syn.class_setup(False)
    c = len(labels)
    cam = camera_setup_1()
    boundary, res, mh, mw = [[0, 2000], [0, 2000]], 0.2, 10000, 10000
    lane = names.index("lane")
    log_cm = numpy_port.confusion_submatrix_log(syn.synthetic_confusion_matrix(11), labels)
    frames = []
    for f in range(3):
        fr = syn.synthetic_frame(2000, 7 * f, 400000, blocky=(f == 1))   # poses 21 m apart
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        frames.append((fr, T))
    for cm in (log_cm, np.eye(c)):
        dm = DeviceMapper(mh, mw, colors, cm, boundary, res, 100.0, True, lane, cameras=[cam], device=0)
        ref = np.zeros((mh, mw, c))
        keep = [(dev(fr["points"]), dev(fr["semantic_image"])) for fr, _ in frames]   # frames only hold raw pointers
        dm.integrate_batch([dm.make_frame(dp, di, T, 0) for (dp, di), (_, T) in zip(keep, frames)])
        for fr, T in frames:
            mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
            c_oracle.update_map(ref, mp, lab, colors, cm, boundary, res, True, lane)
        got = dm.map.cpu().numpy()
        assert np.count_nonzero(ref) > 50000
        assert np.array_equal(got, ref)
        del got
        dm.close()
        del dm
        torch.cuda.empty_cache()


@pytest.mark.parametrize("log_cm", [False, True])
def test_cfg5_shape_two_cameras_in_one_batch(log_cm):
    """BASELINE.json configs[4] shape: frames of cam1 and cam6 (each with its own projection matrix) fused into one
    grid, alternating inside a batch."""
    from oracle import numpy_port
    from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
    labels, names, colors = syn.class_setup(True)
    c = len(labels)
    cams = [camera_setup_1(), camera_setup_6()]
    boundary, res, mh, mw = [[0, 600], [0, 1400]], 0.2, 3000, 7000
    lane = names.index("lane")
    cm = numpy_port.confusion_submatrix_log(syn.synthetic_confusion_matrix(5), labels) if log_cm else np.eye(c)
    dm = DeviceMapper(mh, mw, colors, cm, boundary, res, 100.0, True, lane, cameras=cams, device=0)
    ref = np.zeros((mh, mw, c))
    batch, keep = [], []
    for f in range(6):
        cam = cams[f % 2]
        fr = syn.synthetic_frame(4000, f, 120000, blocky=(f % 3 == 0))
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        dp, di = dev(fr["points"]), dev(fr["semantic_image"])
        keep.append((dp, di))
        batch.append(dm.make_frame(dp, di, T, cam))
        mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
        c_oracle.update_map(ref, mp, lab, colors, cm, boundary, res, True, lane)
    dm.integrate_batch(batch)
    assert np.array_equal(dm.map.cpu().numpy(), ref)
    rgb = renderer.filter_and_render(dm.map, colors)
    assert np.array_equal(rgb.cpu().numpy(), c_oracle.render_bev_map(c_oracle.apply_filter(ref), colors))
    dm.close()


@pytest.mark.parametrize("world", [2, 3, 7])
def test_row_tiled_render_equals_whole_map_render(world):
    """frame_sharding.render_row_tile on row tiles with one-row halos (what sum_grid_row_tile hands every rank)
    reproduces the whole-map filter + render, seams and true borders included."""
    from vision_semantic_segmentation_b200 import frame_sharding
    rng = np.random.default_rng(world)
    mh, mw, c = 61, 45, 5
    grid = rng.integers(0, 4, (mh, mw, c)).astype(np.float64) * (rng.random((mh, mw, 1)) < 0.4)
    colors = rng.integers(0, 256, (c, 3))
    whole_rgb, whole_f = renderer.filter_and_render(dev(grid), colors, return_filtered=True)
    rows_rgb, rows_f = [], []
    for rank in range(world):
        r0, r1 = frame_sharding.row_tile(mh, rank, world)
        top = 1 if r0 > 0 and r1 > r0 else 0
        bottom = 1 if r1 < mh and r1 > r0 else 0
        tile = dev(grid[r0 - top:r1 + bottom])
        rgb, filt = frame_sharding.render_row_tile(tile, top, bottom, colors, return_filtered=True)
        assert rgb.shape[0] == r1 - r0
        rows_rgb.append(rgb)
        rows_f.append(filt)
    assert torch.equal(torch.cat(rows_rgb), whole_rgb)
    assert torch.equal(torch.cat(rows_f), whole_f)


def test_frame_tag_space_wraps_cleanly():
    """The count update de-duplicates with uint32 frame tags that only grow; when the counter is about to overflow
    the tag planes are re-zeroed and it restarts.  A test hook moves the counter next to 2^32 so that the wrap
    happens in the middle of 40 frames (three batches); every count must still be right."""
    from vision_semantic_segmentation_b200 import _native
    case = Case("cfg1_c5_count")
    dm = make_mapper(case)
    lib = _native.load()
    frames = [case.frame(f % 3) for f in range(3)]
    keep = [(dev(points), dev(image)) for _, points, image, _ in frames]
    batch = [dm.make_frame(keep[f % 3][0], keep[f % 3][1], frames[f % 3][3], 0) for f in range(40)]
    dm.integrate_batch(batch[:5])                      # allocates the tag planes, tags 1..5
    _native.check(lib.smap_debug_set_frame_tag(dm._h, 0xffffffff - 25))
    dm.integrate_batch(batch[5:])                      # 16 + 16 + 3 frames: the second batch trips the wrap
    # frames 0, 1, 2 were integrated 14, 13 and 13 times; counts are additive
    per = [_oracle_grid(case, [frames[k]]) for k in range(3)]
    ref = 14.0 * per[0] + 13.0 * per[1] + 13.0 * per[2]
    assert np.array_equal(dm.map.cpu().numpy(), ref)
    assert lib.smap_debug_set_frame_tag(dm._h, 1) != 0   # only forward
    dm.close()
