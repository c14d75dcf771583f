"""GPU: class-id planes (SMAP_IMG_CLASS_IDS) through the fused kernel.  Integrating the network's
id plane must give, bit for bit, the grid the reference computes from the RGB image its node would have published
(nearest-neighbour upscale + apply_color_map): checked against the golden vectors of the real reference (full
resolution) and against the C oracle run on the painted image (reduced resolutions, ids without a palette entry,
multiply-shift and table index maps, all three update modes, batches, the host-buffer entry point, the Python API)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import c_oracle  # noqa: E402
from tests.common import Case, sha  # noqa: E402
from vision_semantic_segmentation_b200 import _native, label_image, synthetic as syn  # noqa: E402
from vision_semantic_segmentation_b200.camera import camera_setup_1  # noqa: E402
from vision_semantic_segmentation_b200.device_mapper import DeviceMapper  # noqa: E402
from vision_semantic_segmentation_b200.utils import transforms as tr  # noqa: E402

W, H = syn.IMAGE_W, syn.IMAGE_H
BOUNDARY, RES, MH, MW = [[100, 300], [800, 1000]], 0.1, 2000, 2000


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def frame_with_ids(case, f):
    """The golden case's frame f plus the id plane its label image was painted from."""
    s = case.spec
    fr = syn.synthetic_frame(s["seed"], f, s["n_points"], height=s["image_hw"][0], width=s["image_hw"][1],
                             blocky=(f in s.get("blocky_frames", [])), with_ids=True)
    assert sha(fr["semantic_image"]) == s["frames_out"][f]["in_image_sha"]
    assert np.array_equal(label_image.paint_class_ids(fr["semantic_ids"], syn.COLORS_19), fr["semantic_image"])
    return fr


@pytest.mark.parametrize("name,ordered", [("cfg1_c5_count", False), ("cfg1_c5_count", True), ("cfg1_c19_count", False),
                                          ("cfg1_c5_log", False), ("cfg1_c19_log", False), ("cam6_res02", False)])
def test_full_resolution_ids_match_reference_golden(name, ordered):
    """tags (5 classes), masks + float64 REDs (19 classes), ordered update (log-likelihood / forced)."""
    case = Case(name)
    dm = DeviceMapper(case.mh, case.mw, case.colors, case.cm, case.boundary, case.resolution, case.range_max,
                      case.use_intensity, case.lane, cameras=[case.cam], device=0)
    dm.set_label_palette(syn.COLORS_19)
    if ordered:
        dm.notify_map_modified()
    for f, out in enumerate(case.spec["frames_out"]):
        fr = frame_with_ids(case, f)
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ case.T_v2b)
        dm.integrate(dm.make_frame(dev(fr["points"]), dev(fr["semantic_ids"]), T, 0))
        assert sha(dm.map.cpu().numpy()) == out["map_sha_after"], "frame %d" % f
    assert sha(dm.map.cpu().numpy()) == case.spec["map_sha"]
    dm.close()


def oracle_grid(frames, colors, cm, lane, c):
    ref = np.zeros((MH, MW, c))
    cam = camera_setup_1()
    for fr, image, T in frames:
        mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, image, 100.0)
        c_oracle.update_map(ref, mp, lab, colors, cm, BOUNDARY, RES, True, lane)
    return ref


@pytest.mark.parametrize("ids_hw,max_id", [((720, 960), 24), ((432, 576), 19), ((185, 130), 256), ((1440, 1920), 40)])
@pytest.mark.parametrize("full19,log_cm", [(False, False), (True, False), (False, True)])
def test_reduced_resolution_ids_against_oracle(ids_hw, max_id, full19, log_cm):
    """Network output at a reduced resolution (IMAGE_SCALE < 1), ids beyond the palette (painted black: with the 19
    classes they hit car + motorcycle + truck at once), index map as multiply-shift or as table."""
    labels, names, colors = syn.class_setup(full19)
    c, lane = len(labels), names.index("lane")
    cm = np.eye(c)
    if log_cm:
        cm = np.log(np.random.default_rng(5).uniform(0.01, 1.0, (c, c)))
    dm = DeviceMapper(MH, MW, colors, cm, BOUNDARY, RES, 100.0, True, lane, cameras=[camera_setup_1()], device=0)
    dm.set_label_palette(syn.COLORS_19)
    rgb = DeviceMapper(MH, MW, colors, cm, BOUNDARY, RES, 100.0, True, lane, cameras=[camera_setup_1()], device=0)
    frames, keep = [], []
    rng = np.random.default_rng(ids_hw[0] + max_id)
    for f in range(2):
        fr = syn.synthetic_frame(700, f, 150000)
        ids = rng.integers(0, max_id, ids_hw).astype(np.uint8)
        if f == 1:   # spatially coherent: many points per (cell, class)
            ids = np.repeat(np.repeat(ids[::8, ::8], 8, axis=0), 8, axis=1)[:ids_hw[0], :ids_hw[1]].copy()
        image = label_image.paint_class_ids(ids, syn.COLORS_19, W, H)
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        frames.append((fr, image, T))
        pts, dids, dimg = dev(fr["points"]), dev(ids), dev(image)
        keep.append((pts, dids, dimg))
        dm.integrate(dm.make_frame(pts, dids, T, 0, image_size=(H, W)))
        rgb.integrate(rgb.make_frame(pts, dimg, T, 0))
    want = oracle_grid(frames, colors, cm, lane, c)
    got = dm.map.cpu().numpy()
    assert np.count_nonzero(want) > 1000
    assert np.array_equal(got, want)
    assert np.array_equal(rgb.map.cpu().numpy(), want)
    dm.close()
    rgb.close()


def test_ids_in_batches_and_through_host_buffers():
    """A 16-frame batch mixing RGB images and id planes of two resolutions (the index map changes inside the batch),
    and smap_integrate_host with a pinned id plane."""
    labels, names, colors = syn.class_setup(False)
    c, lane = len(labels), names.index("lane")
    cm = np.eye(c)
    dm = DeviceMapper(MH, MW, colors, cm, BOUNDARY, RES, 100.0, True, lane, cameras=[camera_setup_1()], device=0)
    dm.set_label_palette(syn.COLORS_19)
    host = DeviceMapper(MH, MW, colors, cm, BOUNDARY, RES, 100.0, True, lane, cameras=[camera_setup_1()], device=0)
    host.set_label_palette(syn.COLORS_19)
    rng = np.random.default_rng(77)
    shapes = [(1440, 1920), (720, 960), None, (185, 130)]
    frames, batch, keep = [], [], []
    for f in range(16):
        fr = syn.synthetic_frame(900, f, 40000)
        T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
        shape = shapes[f % 4]
        pts = dev(fr["points"])
        if shape is None:
            image = fr["semantic_image"]
            dimg = dev(image)
            batch.append(dm.make_frame(pts, dimg, T, 0))
            hp, hi = torch.from_numpy(fr["points"]).pin_memory(), torch.from_numpy(image).pin_memory()
            host.integrate_host(host.make_frame(hp, hi, T, 0, host=True))
        else:
            ids = rng.integers(0, 21, shape).astype(np.uint8)
            image = label_image.paint_class_ids(ids, syn.COLORS_19, W, H)
            dimg = dev(ids)
            batch.append(dm.make_frame(pts, dimg, T, 0, image_size=(H, W)))
            hp, hi = torch.from_numpy(fr["points"]).pin_memory(), torch.from_numpy(ids).pin_memory()
            host.integrate_host(host.make_frame(hp, hi, T, 0, host=True, image_size=(H, W)))
        keep.append((pts, dimg, hp, hi))
        frames.append((fr, image, T))
    dm.integrate_batch(batch)
    want = oracle_grid(frames, colors, cm, lane, c)
    assert np.array_equal(dm.map.cpu().numpy(), want)
    assert np.array_equal(host.map.cpu().numpy(), want)
    dm.close()
    host.close()


def test_ids_error_paths():
    labels, names, colors = syn.class_setup(False)
    dm = DeviceMapper(MH, MW, colors, np.eye(5), BOUNDARY, RES, 100.0, True, 2, cameras=[camera_setup_1()], device=0)
    fr = syn.synthetic_frame(1, 0, 1000, with_ids=True)
    pts, ids = dev(fr["points"]), dev(fr["semantic_ids"])
    with pytest.raises(ValueError, match="palette"):
        dm.make_frame(pts, ids, None, 0)
    dm.set_label_palette(syn.COLORS_19)
    with pytest.raises(ValueError, match="float32"):
        dm.make_frame(dev(fr["pcd"]), ids, None, 0)
    frame = dm.make_frame(pts, ids, None, 0)
    with pytest.raises(_native.SmapError, match="RGB"):
        dm.project(frame)
    frame.image_format = 7
    with pytest.raises(_native.SmapError, match="image format"):
        dm.integrate(frame)
    # a fresh handle without a palette refuses id planes at the C ABI as well
    raw = DeviceMapper(MH, MW, colors, np.eye(5), BOUNDARY, RES, 100.0, True, 2, cameras=[camera_setup_1()], device=0)
    frame.image_format = _native.SMAP_IMG_CLASS_IDS
    with pytest.raises(_native.SmapError, match="palette"):
        raw.integrate(frame)
    dm.close()
    raw.close()


def test_mapping_replay_with_semantic_ids_matches_reference_golden(tmp_path):
    """The reference-facing API fed with the network's id planes reproduces the real reference's map and render."""
    from vision_semantic_segmentation_b200.config.base_cfg import get_cfg_defaults
    from vision_semantic_segmentation_b200.mapping_replay import SemanticMapping
    case = Case("cfg1_c5_count")
    cfg = get_cfg_defaults()
    cfg.OUTPUT_DIR = str(tmp_path)
    cfg.LABELS, cfg.LABELS_NAMES, cfg.LABEL_COLORS = case.labels, case.names, case.colors
    sm = SemanticMapping(cfg)
    sm.set_label_palette([{"color": [int(v) for v in col]} for col in syn.COLORS_19])
    frames = []
    for f in range(len(case.spec["frames_out"])):
        fr = frame_with_ids(case, f)
        fr.pop("semantic_image")
        if f == 1:
            fr.pop("points")   # reference-style record: (4, N) float64 pcd only
        frames.append(fr)
    color_map = sm.mapping_replay(frames, "ids", write_image=False)
    assert sha(sm.map) == case.spec["filtered_sha"]
    assert np.array_equal(color_map, case.arrays["rgb"])
