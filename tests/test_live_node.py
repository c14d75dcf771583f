"""CPU: host logic of the live node's entry point (``mapping.py``, mirror of the reference's ``src/mapping.py``):
queue synchronisation against golden vectors recorded from the real reference (``oracle/make_golden_live.py``), message
handling, recording -- with a stand-in for the device so that no GPU is needed."""
import json
import os
import types

import numpy as np
import pytest

from tests.common import GOLDEN
from vision_semantic_segmentation_b200 import mapping as live
from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
from vision_semantic_segmentation_b200.utils import transforms as tr


def bare_node(**attrs):
    """A SemanticMapping without its constructor (which needs an output directory and, later, a device)."""
    sm = live.SemanticMapping.__new__(live.SemanticMapping)
    sm.pcd_queue, sm.pcd_header_queue, sm.pose_queue = [], [], []
    sm.pcd = sm.pcd_frame_id = sm.pose = None
    sm.cam1, sm.cam6 = camera_setup_1(), camera_setup_6()
    sm.depth_method, sm.test_cut_time, sm.save_map_to_file = "points_map", 100, False
    sm.input_list, sm.record_inputs, sm.done = [], True, False
    sm.on_semantic_point_cloud = sm.on_semantic_local_map = None
    for k, v in attrs.items():
        setattr(sm, k, v)
    return sm


def header(stamp, frame_id="world"):
    return types.SimpleNamespace(stamp=stamp, frame_id=frame_id)


def test_queue_synchronisation_matches_the_reference():
    with open(os.path.join(GOLDEN, "live_node.json")) as f:
        cases = json.load(f)["queue_cases"]
    assert len(cases) > 200
    for c in cases:
        sm = bare_node()
        sm.pcd_header_queue = [header(s) for s in c["stamps"]]
        sm.pcd_queue = list(range(len(c["stamps"])))
        pcd, stamp = sm.update_pcd(c["target"])
        assert (pcd, stamp, sm.pcd_queue) == (c["pcd_pick"], c["pcd_stamp"], c["pcd_left"]), c
        assert len(sm.pcd_header_queue) == len(sm.pcd_queue)
        sm.pose_queue = [types.SimpleNamespace(header=header(s), pose=i) for i, s in enumerate(c["stamps"])]
        pose, stamp = sm.update_pose(c["target"])
        assert (pose, stamp, [m.pose for m in sm.pose_queue]) == (c["pose_pick"], c["pose_stamp"], c["pose_left"]), c


def test_callbacks_queue_and_clock():
    sm = bare_node()
    cloud = np.zeros((5, 4), np.float32)
    sm.pcd_callback(types.SimpleNamespace(header=header(1.0, "velodyne"), points=cloud))
    assert sm.pcd_queue[0] is cloud and sm.pcd_frame_id == "velodyne" and sm.pcd_header_queue[0].stamp == 1.0
    sm.pose_callback(types.SimpleNamespace(header=header(99.0), pose="p"))
    assert not sm.save_map_to_file
    sm.pose_callback(types.SimpleNamespace(header=header(types.SimpleNamespace(secs=100, nsecs=5)), pose="q"))
    assert sm.save_map_to_file and len(sm.pose_queue) == 2
    with pytest.raises(TypeError, match="points"):
        sm.pcd_callback(types.SimpleNamespace(header=header(2.0), width=3))
    with pytest.raises(TypeError, match="image"):
        sm.image_callback(types.SimpleNamespace(header=header(2.0, "camera1")))
    with pytest.raises(ValueError, match="camera"):
        sm.image_callback(types.SimpleNamespace(header=header(2.0, "camera9"), image=np.zeros((2, 2, 3), np.uint8)))
    origin = sm.set_global_map_pose()
    assert (origin.position.x, origin.position.y) == (-1369.0496826171875, -562.84814453125)


class FakeDevice(object):
    """Records what mapping() asks of the device."""

    def __init__(self):
        self.device, self.calls = "fake", []

    def clear(self):
        self.calls.append("clear")

    def integrate(self, frame):
        self.calls.append(("integrate", frame))


class NodeOnFakeDevice(live.SemanticMapping):
    fake = None
    device_mapper = property(lambda self: self.fake)

    def _frame_for(self, pcd, pcd_frame_id, image, pose, camera_calibration, image_size=None):
        return (pcd_frame_id, np.asarray(image).shape, camera_calibration, image_size), None


def test_image_callback_synchronises_records_and_integrates():
    sm = NodeOnFakeDevice.__new__(NodeOnFakeDevice)
    sm.__dict__.update(bare_node().__dict__)
    sm.fake = FakeDevice()
    image = np.zeros((4, 6, 3), np.uint8)
    # nothing queued yet: the image is dropped, as in the reference (src/mapping.py:281-285)
    sm.image_callback(types.SimpleNamespace(header=header(1.0, "camera1"), image=image))
    assert sm.fake.calls == [] and sm.input_list == []
    clouds = [np.full((3, 4), k, np.float32) for k in range(3)]
    for k, c in enumerate(clouds):
        sm.pcd_callback(types.SimpleNamespace(header=header(float(k)), points=c))
    sm.image_callback(types.SimpleNamespace(header=header(1.2, "camera1"), image=image))
    assert sm.fake.calls == [], "no pose yet"
    poses = [tr.Pose((k, 0, 0), (0, 0, 0, 1)) for k in range(3)]
    for k, p in enumerate(poses):
        sm.pose_callback(types.SimpleNamespace(header=header(float(k) + 0.5), pose=p))
    sm.image_callback(types.SimpleNamespace(header=header(1.2, "camera6"), image=image))
    assert sm.fake.calls[0] == "clear" and sm.fake.calls[1][0] == "integrate" and len(sm.fake.calls) == 2
    assert sm.fake.calls[1][1] == ("world", (4, 6, 3), sm.cam6, None)
    assert np.array_equal(sm.pcd, clouds[1]) and sm.pose is poses[1]          # 1.0 is closer to 1.2 than 2.0; 1.5 than 0.5
    rec = sm.input_list[0]
    assert sorted(rec) == ["camera_id", "pcd_frame_id", "points", "pose", "semantic_image"] and rec["camera_id"] == 6
    assert rec["points"] is not clouds[1] and np.array_equal(rec["points"], clouds[1])   # a copy, as np.array(self.pcd)
    # a class-id plane is recorded as such, with the camera resolution it stands for; the map is not cleared again
    sm.image_callback(types.SimpleNamespace(header=header(2.4, "camera1"), image=np.zeros((2, 3), np.uint8)))
    assert sm.fake.calls[2][1] == ("world", (2, 3), sm.cam1, (1440, 1920)) and "clear" not in sm.fake.calls[2:]
    assert "semantic_ids" in sm.input_list[1] and sm.input_list[1]["image_size"] == (1440, 1920)
    # the reference's own cloud layout is recorded under the reference's key
    sm.pcd_callback(types.SimpleNamespace(header=header(2.45), points=np.zeros((4, 7))))
    sm.image_callback(types.SimpleNamespace(header=header(2.45, "camera1"), image=image))
    assert "pcd" in sm.input_list[2] and "points" not in sm.input_list[2]
    sm.input_list.pop()
    sm.fake.calls.pop()
    sm.record_inputs = False
    sm.image_callback(types.SimpleNamespace(header=header(2.5, "camera1"), image=image))
    assert len(sm.input_list) == 2 and len(sm.fake.calls) == 4


def test_planar_method_and_extrinsics():
    # the planar method: nothing is recorded or integrated (src/mapping.py:319-320), only update_map_planar runs -- and that
    # needs the device (no CPU fallback): tests/test_gpu_live_node.py::test_planar_update_is_the_references_clamp
    sm = NodeOnFakeDevice.__new__(NodeOnFakeDevice)
    sm.__dict__.update(bare_node(depth_method="planar").__dict__)
    sm.fake, sm._map_valid, planar_calls = FakeDevice(), False, []
    sm.update_map_planar = lambda m, image, cam: planar_calls.append((m, image.shape, cam))
    sm.mapping(np.zeros((2, 2, 3), np.uint8), None, sm.cam1)
    assert sm.fake.calls == ["clear"] and planar_calls == [(None, (2, 2, 3), sm.cam1)] and sm.input_list == []
    sm = bare_node()
    T_v2b = tr.euler_matrix(0.0, 0.140, 0.0)
    T_v2b[0:3, 3] = [2.64, 0, 1.98]
    sm.T_cam1_to_base, sm.T_cam6_to_base = T_v2b @ sm.cam1.T, T_v2b @ sm.cam6.T
    pose = tr.Pose((10.0, -3.0, 0.5), (0.0, 0.0, np.sin(0.2), np.cos(0.2)))
    want = np.linalg.inv(tr.get_transform_from_pose(pose) @ sm.T_cam6_to_base)[0:3]
    assert np.array_equal(sm.get_extrinsics(pose, "camera6"), want)
    with pytest.raises(ValueError):
        sm.get_extrinsics(pose, "camera7")
    with pytest.raises(RuntimeError, match="ROS"):
        live.main([])
