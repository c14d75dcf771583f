#!/usr/bin/env python
"""Live semantic mapping -- the reference's ``src/mapping.py`` entry point on a B200, without a ROS dependency.

The reference's live node (``src/mapping.py:39-355``) is a rospy class: three subscribers feed ``pcd_callback``,
``pose_callback`` and ``image_callback``; every segmented camera image picks the point cloud and the pose closest in
time (``update_pcd`` / ``update_pose``) and calls ``mapping(semantic_image, pose, camera_calibration)``, which records
the frame in ``input_list``, runs ``project_pcd`` + ``update_map`` and -- once ``save_map_to_file`` is raised by the
clock -- dumps the recorded drive, smooths, renders, writes ``global_map.png``, evaluates and shuts the node down.

``SemanticMapping`` below keeps those methods, attributes and their order of effects; the per-frame arithmetic is the
fused CUDA path of ``mapping_replay.SemanticMapping`` (same kernels, same C ABI), so a frame costs microseconds
instead of a good fraction of a second.  What is ROS in the reference is a plain Python hook here:

* messages are duck-typed: anything with ``.header.stamp`` (ordered, subtractable; ``rospy.Time`` or a float) and
  ``.header.frame_id``.  A point-cloud message carries its cloud as ``.points`` ((N, 4) float32 x, y, z, intensity,
  or the reference's (4, N) float64) -- or is a real ``sensor_msgs/PointCloud2`` when ROS is installed; an image
  message carries ``.image`` ((H, W, 3) uint8 RGB label image, or the network's (h, w) uint8 class-id plane) -- or is
  a real ``sensor_msgs/Image`` when ``cv_bridge`` is installed; a pose message carries ``.pose``;
* the two publishers (``/semantic_point_cloud``, ``/semantic_local_map``) are the optional callables
  ``on_semantic_point_cloud(pcd_in_range, pcd_label, frame_id)`` and ``on_semantic_local_map(color_map)``.  Without a
  point-cloud consumer the labelled cloud is never materialised (fused kernel); with one, ``project_pcd`` +
  ``update_map`` run as two calls exactly as in the reference;
* ``rospy.signal_shutdown`` becomes ``self.done = True``;
* the recorded drive is written with ``replay_io`` (``input_list.npz``; hickle is not part of the target image);
* ``main()`` wires the callbacks to rospy subscribers when rospy is importable and says so when it is not.

``update_map_planar`` (``:446-488``) does what the reference's does to the map: in the reference the warped uint8 image is
compared with label *names*, never matches, and so never updates a cell; only the clamp
``map[map < 0] = 0`` runs, and that is what runs here.  Not carried over: ``add_car_to_map`` (``:490-526``, "not
tested, may have bug", no caller).
"""
import os
import os.path as osp

import numpy as np

from . import mapping_replay, replay_io
from .config.base_cfg import get_cfg_defaults
from .renderer import filter_and_render
from .utils.transforms import Pose, get_transform_from_pose

__all__ = ["SemanticMapping", "main"]

_POINTS_METHODS = ("points_map", "points_raw")


def _closest_in_queue(stamps, target_stamp):
    """The queue-synchronisation rule of ``update_pcd`` / ``update_pose`` (``src/mapping.py:185-259``).

    Walk the queue in arrival order; at the first neighbouring pair that brackets the target strictly
    (``stamp[i] < target < stamp[i + 1]``) pick the closer of the two (the earlier one on a tie) and drop everything
    before ``i``.  If no pair brackets it -- the target is older or newer than the whole queue, or coincides with a
    stamp -- pick the newest entry and keep only that one.  Returns (index picked, index of the first entry kept)."""
    for i in range(len(stamps) - 1):
        if stamps[i + 1] > target_stamp:
            if stamps[i] < target_stamp:
                diff_2 = stamps[i + 1] - target_stamp
                diff_1 = target_stamp - stamps[i]
                return (i + 1 if diff_1 > diff_2 else i), i
    last = len(stamps) - 1
    return last, last


class SemanticMapping(mapping_replay.SemanticMapping):
    """The live node's class (``src/mapping.py:39``): callbacks + ``mapping()`` on top of the device-resident grid."""

    def __init__(self, cfg, device=None):
        super(SemanticMapping, self).__init__(cfg, device=device)
        self.depth_method = cfg.MAPPING.DEPTH_METHOD
        self.input_list = []
        self.unique_input_dict = {}
        self.record_inputs = True            # the reference always records; switch off for drives that outgrow memory
        self.done = False                    # rospy.signal_shutdown('Done with the mapping')
        self.on_semantic_point_cloud = None  # pub_pcd.publish(create_point_cloud(...))
        self.on_semantic_local_map = None    # pub_semantic_local_map.publish(...)
        self.color_map = None                # the rendered map once it has been saved

    # ------------------------------------------------------------------ callbacks (src/mapping.py:172-290)
    def pcd_callback(self, msg):
        """Queue a point cloud (``:172-183``)."""
        self.pcd_queue.append(self._decode_cloud(msg))
        self.pcd_header_queue.append(msg.header)
        self.pcd_frame_id = msg.header.frame_id

    def update_pcd(self, target_stamp):
        """The queued cloud closest to ``target_stamp`` and its stamp; older entries are dropped (``:185-219``)."""
        pick, keep = _closest_in_queue([h.stamp for h in self.pcd_header_queue], target_stamp)
        header, pcd = self.pcd_header_queue[pick], self.pcd_queue[pick]
        self.pcd_header_queue = self.pcd_header_queue[keep::]
        self.pcd_queue = self.pcd_queue[keep::]
        return pcd, header.stamp

    def pose_callback(self, msg):
        """Queue a pose; the clock passing ``cfg.TEST_END_TIME`` asks for the map to be saved (``:221-226``)."""
        self.pose_queue.append(msg)
        if getattr(msg.header.stamp, "secs", msg.header.stamp) >= self.test_cut_time:
            self.save_map_to_file = True

    def set_global_map_pose(self):
        """The reference broadcasts the map origin as a TF frame here (``:228-236``): the minimum x, y of the point
        map, so that map coordinates are positive.  There is no TF tree to tell; the pose is returned instead (the
        same constants enter the cell index, ``device_mapper.PCD_ORIGIN_OFFSET``)."""
        return Pose((-1369.0496826171875, -562.84814453125, 0.0), (0.0, 0.0, 0.0, 1.0))

    def update_pose(self, target_stamp):
        """The queued pose closest to ``target_stamp`` and its stamp (``:238-259``; same rule as ``update_pcd``)."""
        pick, keep = _closest_in_queue([m.header.stamp for m in self.pose_queue], target_stamp)
        msg = self.pose_queue[pick]
        self.pose_queue = self.pose_queue[keep::]
        return msg.pose, msg.header.stamp

    def image_callback(self, msg):
        """A segmented camera image arrived: synchronise cloud and pose to it and map it (``:261-290``)."""
        image_in = self._decode_image(msg)
        if msg.header.frame_id == "camera1":
            camera_calibration = self.cam1
        elif msg.header.frame_id == "camera6":
            camera_calibration = self.cam6
        else:
            # the reference logs a warning and then fails on the unbound name; say what is wrong instead
            raise ValueError("cannot find camera for frame_id %s" % msg.header.frame_id)
        if self.depth_method in _POINTS_METHODS:
            if len(self.pcd_header_queue) == 0:
                return
            self.pcd, self.pcd_time = self.update_pcd(msg.header.stamp)
        if len(self.pose_queue) == 0:
            return
        self.pose, self.pose_time = self.update_pose(msg.header.stamp)
        self.map_pose = self.set_global_map_pose()
        self.mapping(image_in, self.pose, camera_calibration)

    # ------------------------------------------------------------------ one frame (src/mapping.py:292-355)
    def mapping(self, semantic_image, pose, camera_calibration):
        """Integrate the current cloud (``self.pcd`` / ``self.pcd_frame_id``) seen through ``semantic_image`` at ``pose``
        into the map; when ``save_map_to_file`` is set, also finish the drive: dump the recorded inputs, smooth, render,
        write ``global_map.png``, evaluate, hand the image to ``on_semantic_local_map`` and set ``done``."""
        dm = self.device_mapper
        if not self._map_valid:               # self.map = np.zeros(...) on the first frame
            dm.clear()
            self._map_valid = True
        if self.depth_method not in _POINTS_METHODS:
            # src/mapping.py:319-320: nothing is recorded, nothing is projected; see update_map_planar
            self.update_map_planar(None, semantic_image, camera_calibration)
            self._finish_frame(dm)
            return
        ids = getattr(semantic_image, "ndim", 3) == 2 or (hasattr(semantic_image, "dim") and semantic_image.dim() == 2)
        if self.record_inputs:
            # the reference's record (:309-312); a float4 cloud goes under "points", the key replay gives that layout
            frame_input_dict = {"points" if _is_float4(self.pcd) else "pcd": _copy(self.pcd),
                                "pcd_frame_id": self.pcd_frame_id, "pose": pose,
                                "semantic_ids" if ids else "semantic_image": _copy(semantic_image)}
            if ids:
                frame_input_dict["image_size"] = (int(camera_calibration.imSize[1]), int(camera_calibration.imSize[0]))
            if camera_calibration is self.cam6:
                frame_input_dict["camera_id"] = 6
            self.input_list.append(frame_input_dict)
        if self.on_semantic_point_cloud is not None and not ids:
            # somebody wants the labelled cloud: the two calls of the reference, results stay on the device
            pcd_in_range, pcd_label = self.project_pcd(_to_device(self.pcd, dm.device), self.pcd_frame_id,
                                                       _to_device(semantic_image, dm.device), pose, camera_calibration)
            self.on_semantic_point_cloud(pcd_in_range, pcd_label, self.pcd_frame_id)
            self.update_map(dm.map, pcd_in_range, pcd_label)
        else:
            size = (int(camera_calibration.imSize[1]), int(camera_calibration.imSize[0])) if ids else None
            frame, keep = self._frame_for(self.pcd, self.pcd_frame_id, semantic_image, pose, camera_calibration,
                                          image_size=size)
            dm.integrate(frame)

        self._finish_frame(dm)

    def _finish_frame(self, dm):
        """``src/mapping.py:322-355``: when ``save_map_to_file`` is set, finish the drive."""
        if self.save_map_to_file:
            if self.record_inputs:
                os.makedirs(self.input_dir, exist_ok=True)
                print("writing input_list ...")
                replay_io.save_input_list(osp.join(self.input_dir, "input_list.npz"), self.input_list)
            os.makedirs(self.output_dir, exist_ok=True)
            color_map, filtered = filter_and_render(dm.map, self.label_colors, return_filtered=True)
            dm.map.copy_(filtered)            # self.map = apply_filter(self.map)
            dm.notify_map_modified()
            self.color_map = color_map.cpu().numpy()
            output_file = osp.join(self.output_dir, "global_map.png")
            print("Saving image to", output_file)
            from .utils.image_io import imwrite
            imwrite(output_file, self.color_map)
            if self.ground_truth_dir != "":
                from .evaluation import Test
                Test(ground_truth_dir=self.ground_truth_dir, logger=self.logger).test_single_map(color_map)
            if self.on_semantic_local_map is not None:
                self.on_semantic_local_map(self.color_map)
            self.done = True

    def update_map_planar(self, map_local, image, cam):
        """The planar (homography) update of the reference (``src/mapping.py:446-488``), i.e. what it does to the map:
        it warps ``image`` onto the map plane (``generate_homography``, needs a live TF tree) and then tests
        ``image_on_map[:, :, 0] == self.label_names[i]`` -- a uint8 array against a *string*, which numpy evaluates to
        ``False`` -- so no cell is ever incremented (probed on the unmodified reference); the one
        statement with an effect is ``map_local[map_local < 0] = 0`` (``:481``).  Exactly that runs here
        (``smap_clamp_negative``); the warp, whose result the reference discards, is not computed.
        ``map_local``: numpy array (clamped in place and returned, as the reference does), CUDA tensor, or None for
        the device grid."""
        import ctypes
        from . import _native
        torch = _native.require_cuda()
        lib = _native.load()
        if map_local is None:
            dm = self.device_mapper
            target = dm.map
        elif isinstance(map_local, np.ndarray):
            target = torch.from_numpy(np.ascontiguousarray(map_local, dtype=np.float64)).to(self.device_mapper.device)
        else:
            target = map_local
            if not target.is_cuda or target.dtype != torch.float64 or not target.is_contiguous():
                raise ValueError("map must be a contiguous float64 CUDA tensor or a numpy array")
        with torch.cuda.device(target.device):
            _native.check(lib.smap_clamp_negative(ctypes.c_void_p(target.data_ptr()), target.numel(), target.device.index,
                                                  _native.current_stream_ptr(target.device)))
        if map_local is None:
            return None     # the device grid keeps only zeros and counts >= 0 or log-probabilities <= 0: bookkeeping unchanged
        if isinstance(map_local, np.ndarray):
            map_local[...] = target.cpu().numpy()
        return map_local

    def get_extrinsics(self, pose, camera_id):
        """World -> camera 3x4 extrinsics for ``camera1`` / ``camera6`` at ``pose`` (``:528-541``)."""
        T_base_to_origin = get_transform_from_pose(pose)
        if camera_id == "camera1":
            T_cam_to_origin = np.matmul(T_base_to_origin, self.T_cam1_to_base)
        elif camera_id == "camera6":
            T_cam_to_origin = np.matmul(T_base_to_origin, self.T_cam6_to_base)
        else:
            raise ValueError("unable to find camera to base for camera_id %s" % camera_id)
        return np.linalg.inv(T_cam_to_origin)[0:3]

    # ------------------------------------------------------------------ message decoding
    @staticmethod
    def _decode_cloud(msg):
        points = getattr(msg, "points", None)
        if points is not None:
            return points
        try:
            from sensor_msgs import point_cloud2 as pc2
        except ImportError:
            raise TypeError("point-cloud messages need a .points array ((N, 4) float32 or (4, N) float64) when ROS "
                            "(sensor_msgs) is not installed")
        pcd = np.empty((4, msg.width))      # src/mapping.py:178-180
        for i, el in enumerate(pc2.read_points(msg, field_names=("x", "y", "z", "intensity"), skip_nans=True)):
            pcd[:, i] = el
        return pcd

    @staticmethod
    def _decode_image(msg):
        image = getattr(msg, "image", None)
        if image is not None:
            return image
        try:
            from cv_bridge import CvBridge
        except ImportError:
            raise TypeError("image messages need an .image array ((H, W, 3) uint8 label image or (h, w) uint8 class ids) "
                            "when ROS (cv_bridge) is not installed")
        return CvBridge().imgmsg_to_cv2(msg, desired_encoding="passthrough")


def _is_float4(cloud):
    """(N, 4) float32 x, y, z, intensity -- as opposed to the reference's (4, N) float64."""
    shape, dtype = getattr(cloud, "shape", ()), str(getattr(cloud, "dtype", ""))
    return len(shape) == 2 and shape[1] == 4 and dtype.endswith("float32")


def _copy(a):
    """np.array(x) of the reference's recording (``:309-312``): the queue entry may be overwritten later."""
    if a is None:
        return None
    if hasattr(a, "clone"):
        return a.clone()
    return np.array(a)


def _to_device(a, device):
    """CUDA tensor of a host array (so that project_pcd returns device tensors); tensors pass through."""
    if hasattr(a, "is_cuda"):
        return a.to(device)
    torch = mapping_replay._native.require_cuda()
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.uint8):
        a = np.ascontiguousarray(a, dtype=np.float64)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def main(argv=None):
    """``rosrun``-style entry point (``src/mapping.py:543-575``): needs rospy for the subscriptions."""
    try:
        import rospy
        from geometry_msgs.msg import PoseStamped
        from sensor_msgs.msg import Image, PointCloud2
    except ImportError:
        raise RuntimeError("the live node needs ROS (rospy, sensor_msgs, geometry_msgs); without it, build a "
                           "mapping.SemanticMapping yourself and feed pcd_callback / pose_callback / image_callback, "
                           "or replay a recorded drive with mapping_replay")
    import sys
    rospy.init_node("semantic_mapping")
    cfg = get_cfg_defaults()
    # the last two arguments belong to roslaunch (src/mapping.py:556-558)
    args = mapping_replay.parse_args(sys.argv[1:-2] if argv is None else argv)
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    sm = SemanticMapping(cfg)
    rospy.Subscriber("/current_pose", PoseStamped, sm.pose_callback)
    rospy.Subscriber("/camera1/semantic", Image, sm.image_callback)
    rospy.Subscriber("/camera6/semantic", Image, sm.image_callback)
    if sm.depth_method == "points_map":
        rospy.Subscriber("/reduced_map", PointCloud2, sm.pcd_callback)
    elif sm.depth_method == "points_raw":
        rospy.Subscriber("/points_raw", PointCloud2, sm.pcd_callback)
    rate = rospy.Rate(20)
    while not rospy.is_shutdown() and not sm.done:
        rate.sleep()


if __name__ == "__main__":
    main()
