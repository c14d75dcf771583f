"""CPU: host-side mirror of the reference interface, the C-ABI surface, frame sharding over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import numpy_port
from vision_semantic_segmentation_b200 import _native, frame_sharding, replay_io, synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
from vision_semantic_segmentation_b200.config.base_cfg import get_cfg_defaults
from vision_semantic_segmentation_b200.data.confusion_matrix import ConfusionMatrix
from vision_semantic_segmentation_b200.utils import transforms as tr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "smap.h")).read()
    declared = sorted(set(re.findall(r"SMAP_API\s+(?:const\s+char\s*\*|int)\s*(smap_\w+)\s*\(", header)))
    assert len(declared) >= 20
    assert sorted(_native.EXPORTS) == declared
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smap_abi_version() == _native.ABI_VERSION == 5
    # struct layouts agree with the header (sizes as computed by the C compiler)
    probe = r'''
    #include <stdio.h>
    #include "smap.h"
    int main(void){printf("%zu %zu %zu\n", sizeof(smap_config), sizeof(smap_frame), sizeof(smap_stats));return 0;}
    '''
    exe = os.path.join(ROOT, "tests", "_abi_probe")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe],
                   input=probe.encode(), check=True)
    try:
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    finally:
        os.remove(exe)
    assert sizes == [ctypes.sizeof(_native.SmapConfig), ctypes.sizeof(_native.SmapFrame),
                     ctypes.sizeof(_native.SmapStats)]


def test_bad_arguments_are_reported_not_thrown():
    lib = _native.load()
    h = ctypes.c_void_p()
    cfg = _native.SmapConfig()
    cfg.map_height, cfg.map_width, cfg.num_classes, cfg.resolution = 0, 10, 5, 0.1
    assert lib.smap_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"empty grid" in lib.smap_last_error()
    cfg.map_height, cfg.num_classes = 10, 32
    assert lib.smap_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    with pytest.raises(_native.SmapError):
        _native.check(lib.smap_render(None, 4, 4, 3, None, None, 0, None))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vision_semantic_segmentation_b200 import renderer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        renderer.render_bev_map(np.zeros((4, 4, 2)), [[1, 2, 3], [4, 5, 6]])
    with pytest.raises(ValueError):
        renderer.render_bev_map(np.zeros((4, 4, 2)), [[1, 2, 3]])
    with pytest.raises(ValueError):
        renderer.render_bev_map(np.zeros((4, 4, 2)), [[1, 2], [4, 5]])
    with pytest.raises(ValueError):
        renderer.render_bev_map_with_thresholds(np.zeros((4, 4, 2)), [[1, 2, 3], [4, 5, 6]], priority=[0])


def test_config_defaults_and_merge(tmp_path):
    cfg = get_cfg_defaults()
    assert cfg.LABELS == [2, 1, 8, 10, 3] and cfg.MAPPING.RESOLUTION == 0.1
    assert cfg.MAPPING.BOUNDARY == [[100, 300], [800, 1000]] and cfg.MAPPING.PCD.RANGE_MAX == 100.0
    y = tmp_path / "c.yaml"
    y.write_text("TASK_NAME: t\nMAPPING:\n  RESOLUTION: 0.2\n  BOUNDARY: [[0, 600], [0, 1400]]\n  PCD:\n    RANGE_MAX: 60\n"
                 "VISION_SEM_SEG:\n  IMAGE_SCALE: 0.5\n  SEM_SEG_NETWORK:\n    MODEL:\n      BACKBONE: x\n")
    cfg.merge_from_file(str(y))
    assert cfg.MAPPING.RESOLUTION == 0.2 and cfg.MAPPING.PCD.RANGE_MAX == 60.0 and cfg.TASK_NAME == "t"
    assert get_cfg_defaults().MAPPING.RESOLUTION == 0.1  # defaults untouched
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MAPPING.NOPE", 1])
    cfg.merge_from_list(["MAPPING.PCD.USE_INTENSITY", False])
    assert cfg.MAPPING.PCD.USE_INTENSITY is False
    bad = tmp_path / "bad.yaml"
    bad.write_text("UNKNOWN_KEY: 1\n")
    with pytest.raises(KeyError):
        cfg.merge_from_file(str(bad))


def test_cameras_and_transforms():
    for cam, K, Rt in ((camera_setup_1(), None, None), (camera_setup_6(), None, None)):
        assert cam.P.shape == (3, 4) and cam.imSize == [1920, 1440]
        assert np.allclose(cam.R @ cam.R.T, np.eye(3), atol=1e-12)
        assert np.array_equal(cam.P, np.matmul(cam.K, np.concatenate([cam.R, cam.t], axis=1)))
    pose = syn.synthetic_pose(3)
    T = tr.get_transform_from_pose(pose)
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
    assert np.array_equal(T[:3, 3], [pose.position.x, pose.position.y, pose.position.z])
    E = tr.euler_matrix(0.0, 0.140, 0.0)
    assert np.isclose(E[0, 2], np.sin(0.140)) and np.isclose(E[2, 0], -np.sin(0.140)) and E[1, 1] == 1.0
    x = np.arange(6.0).reshape(2, 3)
    assert np.array_equal(tr.dehomogenize(tr.homogenize(x) * 2.0), x)


def test_confusion_matrix(tmp_path):
    full = syn.synthetic_confusion_matrix(7)
    p = tmp_path / "cm.npy"
    np.save(p, full)
    cm = ConfusionMatrix(str(p))
    idx = [2, 1, 8, 10, 3]
    assert np.array_equal(cm.get_submatrix(idx, True, True), numpy_port.confusion_submatrix_log(full, idx))
    assert np.allclose(np.exp(cm.get_submatrix(idx, True, True)).sum(1), 1.0)
    with pytest.raises(ValueError):
        cm.get_submatrix([0, 19])
    assert cm.get_submatrix([]) == [] and len(cm) == 19


def test_replay_record_roundtrip(tmp_path):
    frames = [syn.synthetic_frame(3, f, 500, height=48, width=64) for f in range(2)]
    frames[1].pop("points")  # reference-style frame: only the float64 pcd
    frames[1]["camera_id"] = 6
    path = str(tmp_path / "input_list_0.npz")
    replay_io.save_input_list(path, frames)
    back = replay_io.load_input_list(path)
    assert len(back) == 2
    for a, b in zip(frames, back):
        want = a["points"] if "points" in a else a["pcd"].T.astype(np.float32)
        assert np.array_equal(b["points"], want)
        assert np.array_equal(a["semantic_image"], b["semantic_image"])
        assert np.array_equal(a["pose"].as_array(), b["pose"].as_array())
        assert b["pcd_frame_id"] == "world"
    assert back[1]["camera_id"] == 6 and back[0]["camera_id"] == 1
    odd = dict(frames[0])
    odd.pop("points")
    odd["pcd"] = odd["pcd"] + 1e-9  # not float32-representable -> kept as float64
    replay_io.save_input_list(path, [odd])
    assert np.array_equal(replay_io.load_input_list(path)[0]["pcd"], odd["pcd"])


def test_replay_record_with_class_id_planes(tmp_path):
    """Frames that carry the network's class-id plane instead of (or next to) the painted image."""
    frames = [syn.synthetic_frame(4, f, 300, height=48, width=64, with_ids=True) for f in range(3)]
    frames[0].pop("semantic_image")                                   # ids only, full resolution
    frames[1]["semantic_ids"] = frames[1]["semantic_ids"][::2, ::2].copy()   # ids only, half resolution
    frames[1].pop("semantic_image")
    frames[1]["image_size"] = (48, 64)
    path = str(tmp_path / "input_list_0.npz")
    replay_io.save_input_list(path, frames, compressed=True)
    back = replay_io.load_input_list(path)
    assert "semantic_image" not in back[0] and "semantic_image" not in back[1] and "semantic_image" in back[2]
    for a, b in zip(frames, back):
        assert b["semantic_ids"].dtype == np.uint8 and np.array_equal(a["semantic_ids"], b["semantic_ids"])
    assert back[1]["image_size"] == (48, 64) and "image_size" not in back[0]
    with pytest.raises(ValueError):
        replay_io.save_input_list(path, [{k: v for k, v in frames[0].items() if k != "semantic_ids"}])
    with pytest.raises(ValueError):
        replay_io.save_input_list(path, [dict(frames[0], semantic_ids=np.zeros((2, 2, 3), np.uint8))])


def test_shard_ranges_cover_every_frame_once():
    for n in (0, 1, 7, 100, 8000):
        for world in (1, 2, 3, 4, 8):
            seen = [i for r in range(world) for i in frame_sharding.shard_range(n, r, world)]
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        frame_sharding.shard_range(10, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import c_oracle
from vision_semantic_segmentation_b200 import frame_sharding, synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1
from vision_semantic_segmentation_b200.utils import transforms as tr
rank, world, _ = frame_sharding.init_from_env(backend="gloo")
assert frame_sharding.rank_and_world() == (rank, world)
labels, names, colors = syn.class_setup(False)
cam = camera_setup_1(); cm = np.eye(5); B = [[100, 300], [800, 1000]]
def add(grid, f):
    fr = syn.synthetic_frame(5, f, 3000, blocky=True)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    mp, lab, _, _ = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
    c_oracle.update_map(grid, mp, lab, colors, cm, B, 0.5, True, 2)
n_frames = 5
mine = np.zeros((400, 400, 5))
for f in frame_sharding.shard_range(n_frames, rank, world):
    add(mine, f)
partial = mine.copy()
t = torch.from_numpy(mine)
frame_sharding.sum_grids(t)
full = np.zeros((400, 400, 5))
for f in range(n_frames):
    add(full, f)
assert np.array_equal(t.numpy(), full), "sharded sum differs from the sequential grid"
# row tiles: reduce-scatter by rows + one-row halos, filter + render per tile, image all-gathered
tile, r0, r1, top, bottom = frame_sharding.sum_grid_row_tile(torch.from_numpy(partial))
assert (r0, r1) == frame_sharding.row_tile(400, rank, world)
assert np.array_equal(tile.numpy(), full[r0 - top:r1 + bottom]), "row tile / halo differs"
f_tile = c_oracle.apply_filter(tile.numpy())   # the oracle stands in for the CUDA renderer on CPU
rgb_tile = c_oracle.render_bev_map(f_tile, colors)[top:tile.shape[0] - bottom]
rgb = frame_sharding.gather_rgb_rows(torch.from_numpy(np.ascontiguousarray(rgb_tile)), 400)
assert np.array_equal(rgb.numpy(), c_oracle.render_bev_map(c_oracle.apply_filter(full), colors)), "tiled render differs"
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank %%d ok\n" %% rank)   # one write per rank: the two ranks share a pipe
sys.stdout.flush()
'''


def test_frame_sharding_world_size_2_gloo(tmp_path):
    """Host-side N>1 logic on CPU: per-rank partial grids (built here with the oracle standing in for the
    kernels) all-reduced over gloo equal the sequential grid exactly."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_evaluation_host_functions_known_answers():
    """convert_labels / Test.iou (test/test_semantic_mapping.py:6-18,127-161) on a hand-made map, and the same scores
    from the twelve sums the device kernel returns (scores_from_counts): no GPU needed for either."""
    from vision_semantic_segmentation_b200 import evaluation as ev
    colors = {1: (128, 64, 128), 2: (140, 140, 200), 3: (255, 255, 255), 4: (244, 35, 232), 5: (107, 142, 35)}
    want = np.array([[1, 1, 2, 0], [3, 3, 4, 5], [0, 2, 2, 1]])
    rgb = np.zeros(want.shape + (3,), dtype=np.uint8)
    for k, col in colors.items():
        rgb[want == k] = col
    rgb[0, 3] = (128, 64, 129)            # one channel off: not a class
    assert np.array_equal(ev.convert_labels(rgb), want)
    mask = np.ones((5, 6))
    mask[0, 0] = 0
    masked = ev.convert_labels(rgb, mask)
    assert masked[0, 0] == 0 and np.array_equal(masked.ravel()[1:], want.ravel()[1:])
    truth = np.array([[1, 2, 2, 0], [3, 0, 0, 3], [1, 2, 3, 1]], dtype=np.float64)
    t = ev.Test.__new__(ev.Test)
    t.class_lists, t.d, t.logger = [1, 2, 3], {0: "road", 1: "crosswalk", 2: "lane"}, None
    ious, miss = t.iou(truth, want.astype(np.float64))
    assert ious == [2 / 4, 2 / 4, 1 / 4]                   # road 2/(3+3-2), crosswalk 2/(3+3-2), lane 1/(3+2-1)
    assert miss == 1 - 8 / 9
    counts = [2, 2, 1, 3, 3, 3, 3, 3, 2, 9, 8, 5]
    ious2, accs, accuracy, miss2 = ev.scores_from_counts(counts)
    assert ious2 == ious and miss2 == miss and accs == [2 / 3, 2 / 3, 1 / 3] and accuracy == 5 / 9
    with pytest.raises(ZeroDivisionError):   # an empty union: the reference divides two Python floats (:140-141)
        ev.scores_from_counts([0] * 12)


# ---------------------------------------------------------------------------------------------- evaluation (N3)
@pytest.mark.parametrize("name", ["grid_2000_mask", "small", "odd", "no_crosswalk", "no_truth", "golden_render"])
def test_evaluation_host_functions_match_the_reference(name):
    """convert_labels / Test.iou on host arrays == the reference's own functions on the same seeded inputs
    (tests/golden/eval.json, from test/test_semantic_mapping.py:6-18,117-161 run unmodified)."""
    import json
    import os
    from oracle.make_golden_eval import make_inputs
    from tests.common import GOLDEN
    from vision_semantic_segmentation_b200 import evaluation as ev
    with open(os.path.join(GOLDEN, "eval.json")) as f:
        spec = json.load(f)["cases"][name]
    rgb, truth, mask = make_inputs(name, spec)
    generated = ev.convert_labels(rgb, mask)
    assert [int(np.sum(generated == k)) for k in range(6)] == spec["labels_hist"]
    t = ev.Test.__new__(ev.Test)
    t.class_lists, t.d, t.logger = [1, 2, 3], {0: "road", 1: "crosswalk", 2: "lane"}, None
    said = []
    t._say = said.append
    gmap = truth[spec["shift"][0]:generated.shape[0] + spec["shift"][0], spec["shift"][1]:generated.shape[1] + spec["shift"][1]]
    if spec.get("raises") == "ZeroDivisionError":
        with pytest.raises(ZeroDivisionError):
            t.iou(gmap, generated)
        return
    ious, miss = t.iou(gmap, generated, verbose=True)
    dec = lambda v: float(v) if isinstance(v, str) else v
    same = lambda a, b: a == b or (a != a and b != b)
    assert all(same(a, dec(b)) for a, b in zip(ious, spec["iou"]))
    assert same(miss, dec(spec["miss"]))
    assert said == spec["printed"]
