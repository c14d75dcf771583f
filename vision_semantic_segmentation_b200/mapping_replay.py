#!/usr/bin/env python
"""Offline semantic mapping -- the reference's ``src/mapping_replay.py`` entry point on a B200.

``SemanticMapping`` keeps the constructor options, attributes and methods of the reference class
(``src/mapping_replay.py:38-301``): ``project_pcd``, ``update_map``, ``mapping_replay``,
``mapping_replay_dir`` / ``mapping_replay_file`` and the ``--cfg FILE`` command line.  The per-frame
work (transform, projection, culling, label lookup, Bayesian cell update) and the final filter /
arg-max rendering run in the CUDA kernels of ``csrc/`` through the C ABI; the host keeps only what the
reference also does once per frame on the host (pose -> 4x4 matrix, its inverse).

Differences a caller can see, all additive:
* arrays may be CUDA ``torch`` tensors as well as numpy arrays (results then stay on the device);
* a frame dictionary may carry ``"points"`` -- the cloud as (N, 4) float32 x, y, z, intensity -- instead
  of / next to the reference's ``"pcd"`` (4, N) float64, and an optional ``"camera_id"`` (1 or 6;
  the reference's replay hard-codes camera 1, ``src/mapping_replay.py:182``);
* recorded drives are read from ``.npz`` records (``replay_io``) because hickle / HDF5 are not part of
  the target image; ``.hkl`` files are still read when hickle is importable;
* with ``torch.distributed`` initialised, ``mapping_replay`` shards the frames over the ranks and sums
  the per-rank grids with one all-reduce (``frame_sharding``).
"""
import argparse
import os
import os.path as osp
import sys

import numpy as np

from . import _native
from .camera import camera_setup_1, camera_setup_6
from .config.base_cfg import get_cfg_defaults
from .data.confusion_matrix import ConfusionMatrix
from .device_mapper import DeviceMapper
from .renderer import render_bev_map, render_bev_map_with_thresholds, apply_filter, filter_and_render  # noqa: F401
from .utils.logger import MyLogger
from .utils.transforms import euler_matrix, get_transform_from_pose

__all__ = ["SemanticMapping", "main"]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class SemanticMapping(object):
    """BEV semantic grid built from LiDAR clouds and colour-coded segmentation images."""

    def __init__(self, cfg, device=None):
        assert len(cfg.LABELS) == len(cfg.LABELS_NAMES) == len(cfg.LABEL_COLORS)

        output_dir = cfg.OUTPUT_DIR
        if "@" in output_dir:
            # '@' is the project root; a TASK_NAME sub-folder is created below it
            output_dir = output_dir.replace("@", osp.join(osp.dirname(osp.abspath(__file__)), "../"))
            output_dir = osp.abspath(osp.join(output_dir, cfg.TASK_NAME))
        self.logger = MyLogger("mapping", save_dir=output_dir, use_timestamp=False)
        self.output_dir = self.logger.save_dir

        self.pose = None
        self.pose_queue = []
        self.pose_time = None
        self.cam1 = camera_setup_1()
        self.cam6 = camera_setup_6()

        self.pcd = None
        self.pcd_frame_id = None
        self.pcd_queue = []
        self.pcd_header_queue = []
        self.pcd_time = None
        self.pcd_range_max = cfg.MAPPING.PCD.RANGE_MAX
        self.use_pcd_intensity = cfg.MAPPING.PCD.USE_INTENSITY

        self.map_pose = None
        self.save_map_to_file = False
        self.map_boundary = cfg.MAPPING.BOUNDARY
        self.resolution = cfg.MAPPING.RESOLUTION
        self.label_names = cfg.LABELS_NAMES
        self.label_colors = np.array(cfg.LABEL_COLORS)

        self.map_height = int((self.map_boundary[0][1] - self.map_boundary[0][0]) / self.resolution)
        self.map_width = int((self.map_boundary[1][1] - self.map_boundary[1][0]) / self.resolution)
        self.map_depth = len(self.label_names)

        self.position_rel = np.array([[0, 0, 0]]).T
        self.yaw_rel = 0
        self.preprocessing()
        self.test_cut_time = cfg.TEST_END_TIME

        if cfg.MAPPING.CONFUSION_MTX.LOAD_PATH != "":
            cm = ConfusionMatrix(load_path=cfg.MAPPING.CONFUSION_MTX.LOAD_PATH)
            self.confusion_matrix = cm.get_submatrix(cfg.LABELS, to_probability=True, use_log=True)
        else:
            self.confusion_matrix = np.eye(len(self.label_names))

        self.logger.log("Running with configuration:\n" + str(cfg))
        self.ground_truth_dir = cfg.GROUND_TRUTH_DIR
        self.input_dir = cfg.MAPPING.INPUT_DIR

        lanes = [i for i, name in enumerate(self.label_names) if name == "lane"]
        if len(lanes) > 1:
            raise NotImplementedError("more than one class named 'lane' is not supported by the cell-mask layout")
        self._lane_index = lanes[0] if lanes else -1
        self._device_arg = device
        self._label_palette = None   # the network's palette, for frames that carry class-id planes (set_label_palette)
        self._dev = None        # DeviceMapper, created on first device use
        self._host_map = None   # numpy grid assigned by the caller before the device exists

    # ------------------------------------------------------------------ constants
    def preprocessing(self):
        """Constant matrices (``src/mapping_replay.py:117-138``)."""
        self.T_velodyne_to_basklink = self.set_velodyne_to_baselink()
        self.T_cam1_to_base = np.matmul(self.T_velodyne_to_basklink, self.cam1.T)
        self.T_cam6_to_base = np.matmul(self.T_velodyne_to_basklink, self.cam6.T)
        self.discretize_matrix_inv = np.array([
            [self.resolution, 0, self.map_boundary[0][0]],
            [0, self.resolution, self.map_boundary[1][1]],
            [0, 0, 1]], dtype=np.float64)
        self.discretize_matrix = np.linalg.inv(self.discretize_matrix_inv)
        w, h = self.map_width, self.map_height
        self.anchor_points = np.array([[w, w / 3, w, w / 3], [h / 4, h / 4, h * 3 / 4, h * 3 / 4]])
        self.anchor_points_2 = np.array([[w, w / 2, w / 2, w], [h / 4, h / 4, h * 3 / 4, h * 3 / 4]])

    def set_velodyne_to_baselink(self):
        T = euler_matrix(0.0, 0.140, 0.0)
        T[0:3, -1::] = np.array([[2.64, 0, 1.98]]).T
        return T

    # ------------------------------------------------------------------ device state
    @property
    def device_mapper(self):
        if self._dev is None:
            self._dev = DeviceMapper(
                self.map_height, self.map_width, self.label_colors, self.confusion_matrix, self.map_boundary,
                self.resolution, self.pcd_range_max, self.use_pcd_intensity, self._lane_index,
                cameras=[self.cam1, self.cam6], device=self._device_arg)
            if self._label_palette is not None:
                self._dev.set_label_palette(self._label_palette)
            if self._host_map is not None:
                self._dev.map.copy_(_native.require_cuda().from_numpy(self._host_map))
                self._dev.notify_map_modified()
                self._host_map = None
        return self._dev

    def set_label_palette(self, labels):
        """Palette of the segmentation network: the ``labels`` list of its dataset config
        (``cfg.VISION_SEM_SEG.SEM_SEG_NETWORK.DATASET_CONFIG``, read with ``label_image.get_labels``) or an (n, 3)
        colour array.  Frames may then carry ``"semantic_ids"`` -- the network's (h, w) uint8 class-id plane -- instead
        of ``"semantic_image"``; the result is what the reference computes from the image its node would have painted
        (``label_image.paint_class_ids``)."""
        from .label_image import palette_of
        self._label_palette = palette_of(labels)
        if self._dev is not None:
            self._dev.set_label_palette(self._label_palette)

    @property
    def map_device(self):
        """The grid as a CUDA tensor (MH, MW, C) float64 -- no copy."""
        return self.device_mapper.map

    @property
    def map(self):
        """The grid as numpy (a fresh copy), or None before the first frame -- as in the reference."""
        if self._dev is None:
            return self._host_map
        if not self._map_valid:
            return None
        return self._dev.map.cpu().numpy()

    @map.setter
    def map(self, value):
        if value is None:
            self._map_valid = False
            self._host_map = None
            return
        if _is_torch(value):
            self.device_mapper.map.copy_(value)
            self._dev.notify_map_modified()
        elif self._dev is None:
            self._host_map = np.ascontiguousarray(value, dtype=np.float64)
        else:
            self._dev.map.copy_(_native.require_cuda().from_numpy(np.ascontiguousarray(value, dtype=np.float64)))
            self._dev.notify_map_modified()
        self._map_valid = True

    _map_valid = False

    # ------------------------------------------------------------------ per-frame host math
    def world_to_velodyne(self, pose):
        """``inv(T_base_to_origin @ T_velodyne_to_baselink)`` (``src/mapping_replay.py:225-226``)."""
        return np.linalg.inv(np.matmul(get_transform_from_pose(pose), self.T_velodyne_to_basklink))

    def _frame_for(self, pcd, pcd_frame_id, image, pose, camera_calibration, image_size=None):
        """Move one frame's inputs to the device (if needed) and describe it for the C ABI."""
        torch = _native.require_cuda()
        dm = self.device_mapper
        dev = dm.device
        if _is_torch(pcd):
            pts = pcd.to(dev)
            if pts.dtype != torch.float32:
                pts = pts.to(torch.float64)
                if pts.stride(1) != 1:
                    pts = pts.contiguous()
        else:
            pcd = np.asarray(pcd)
            if pcd.dtype == np.float32 and pcd.ndim == 2 and pcd.shape[1] == 4:
                pts = torch.from_numpy(np.ascontiguousarray(pcd)).to(dev)
            else:
                pts = torch.from_numpy(np.ascontiguousarray(pcd, dtype=np.float64)).to(dev)
        if _is_torch(image):
            img = image.to(dev).contiguous()
        else:
            img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(dev)
        T = self.world_to_velodyne(pose) if pcd_frame_id != "velodyne" else None
        if img.dim() == 2 and pts.dtype != torch.float32:
            # class-id planes ride the float4 kernel; recorded clouds hold float32-representable values
            # (PointCloud2 fields are FLOAT32, src/mapping.py:178-180), anything else would change results
            p32 = pts[0:4].to(torch.float32)
            if not torch.equal(p32.to(torch.float64), pts[0:4]):
                raise ValueError("class-id planes need a cloud of float32-representable values")
            pts = p32.t().contiguous()
        return dm.make_frame(pts, img, T, camera_calibration, image_size=image_size), (pts, img)

    # ------------------------------------------------------------------ reference API
    def project_pcd(self, pcd, pcd_frame_id, image, pose, camera_calibration):
        """Points visible in the image and their RGB labels (``src/mapping_replay.py:214-246``).

        pcd (4, N) float64 [or (N, 4) float32]; returns ``masked_pcd`` (4, M) float64 and ``label`` (3, M) uint8,
        numpy for numpy inputs, CUDA tensors for tensor inputs."""
        if pcd is None:
            return
        frame, keep = self._frame_for(pcd, pcd_frame_id, image, pose, camera_calibration)
        masked, label = self.device_mapper.project(frame)
        if _is_torch(pcd):
            return masked, label
        return masked.cpu().numpy(), label.cpu().numpy()

    def update_map(self, map, pcd, label):
        """Bayesian per-cell update (``src/mapping_replay.py:248-301``); mutates and returns ``map``."""
        torch = _native.require_cuda()
        dm = self.device_mapper
        as_dev = lambda a, dt: (a.to(dm.device) if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(a)).to(dm.device)).to(dt)
        pcd_d, label_d = as_dev(pcd, torch.float64), as_dev(label, torch.uint8)
        if _is_torch(map):
            if map.is_cuda and map.dtype == torch.float64 and map.is_contiguous():
                dm.update(pcd_d, label_d, map)
                return map
            tmp = map.to(device=dm.device, dtype=torch.float64).contiguous()
            dm.update(pcd_d, label_d, tmp)
            map.copy_(tmp)
            return map
        if map.shape != (self.map_height, self.map_width, self.map_depth):
            raise ValueError("map has shape %s, expected %s" % (map.shape, (self.map_height, self.map_width, self.map_depth)))
        tmp = torch.from_numpy(np.ascontiguousarray(map, dtype=np.float64)).to(dm.device)
        dm.update(pcd_d, label_d, tmp)
        map[...] = tmp.cpu().numpy()
        return map

    def _feed_item(self, frame_input_dict):
        """One recorded frame (``src/mapping.py:309-312``; optional ``points`` / ``camera_id`` / ``semantic_ids`` /
        ``image_size`` keys, see the module docstring) as a ``replay_feed.FeedItem``; None when it has no cloud."""
        from .replay_feed import FeedItem
        pcd = frame_input_dict["points"] if "points" in frame_input_dict else frame_input_dict["pcd"]
        if pcd is None:
            return None
        cam = self.cam1
        if frame_input_dict.get("camera_id", 1) == 6:
            cam = self.cam6
        if frame_input_dict.get("semantic_ids") is not None:
            # the network's class-id plane; camera resolution from "image_size" (H, W), default the 1920x1440 of
            # the calibrations (src/camera.py:102-135)
            image = frame_input_dict["semantic_ids"]
            size = frame_input_dict.get("image_size")
            if size is None:
                size = (int(cam.imSize[1]), int(cam.imSize[0]))
        else:
            image, size = frame_input_dict["semantic_image"], None
        T = self.world_to_velodyne(frame_input_dict["pose"]) if frame_input_dict["pcd_frame_id"] != "velodyne" else None
        return FeedItem(pcd, image, T, cam, size)

    def integrate_frame(self, frame_input_dict):
        """Fused project_pcd + update_map of one recorded frame into the device grid."""
        self.integrate_frames([frame_input_dict])

    def integrate_frames(self, frame_input_dicts, after_batch=None):
        """Fused project_pcd + update_map of a sequence of recorded frames, in order: streamed to the GPU through the
        pinned, multi-buffered feed of ``replay_feed`` and integrated ``FEED_BATCH`` frames per ``smap_integrate_batch``
        call.  Returns the number of frames integrated."""
        from .replay_feed import FrameFeeder
        dm = self.device_mapper
        if self._feeder is None:
            self._feeder = FrameFeeder(dm, batch=self.FEED_BATCH, depth=self.FEED_DEPTH)
        items = (it for it in (self._feed_item(fr) for fr in frame_input_dicts) if it is not None)
        return self._feeder.run(items, after_batch=after_batch)

    FEED_BATCH = 8     # frames per smap_integrate_batch call of the host-fed replay (PCIe-bound: larger buys nothing)
    FEED_DEPTH = 3     # slot sets: one being integrated, two with copies queued
    EXCHANGE_EVERY = 256  # multi-GPU: frames a rank integrates between two exchanges of the streaming sum
    _feeder = None

    def mapping_replay(self, input_list, file_name, write_image=True, row_tiles=False):
        """Map all frames of ``input_list`` into a fresh grid, smooth, render, save
        ``global_map_<file_name>.png`` (``src/mapping_replay.py:175-211``).  Returns the colour map.
        Under torch.distributed the frames are sharded over the ranks.  With NCCL the per-rank
        increments are summed into every rank's grid by the library's streaming exchange (``smap_exchange_async``:
        touched window only, counts packed, overlapped with the integration of the next frames); other backends
        (gloo in CPU tests) all-reduce the grids through torch.distributed at the end.
        ``row_tiles=False``: every rank filters and renders the whole map (``self.map`` = the filtered map, as in the
        reference).  ``row_tiles=True`` (large maps): every rank filters and renders its own row tile (one-row halos
        from the neighbouring tiles), the image is all-gathered; ``self.map`` then holds the filtered rows of this
        rank's tile only, zeros elsewhere."""
        from . import frame_sharding
        dm = self.device_mapper
        rank, world = frame_sharding.rank_and_world()
        native_comm = world > 1 and frame_sharding.backend_is_nccl()
        if native_comm:
            dm.init_comm()
            dm.set_streaming(True)
        dm.clear()
        self._map_valid = True
        shard = frame_sharding.shard_range(len(input_list), rank, world)
        if native_comm:
            # the same number of exchanges on every rank: one per EXCHANGE_EVERY frames of the longest shard
            longest = -(-len(input_list) // world)
            n_exchanges = max(1, -(-longest // self.EXCHANGE_EVERY))
            state = {"done": 0}

            def after_batch(n_frames):
                while state["done"] < n_exchanges - 1 and n_frames >= (state["done"] + 1) * self.EXCHANGE_EVERY:
                    dm.exchange_async()
                    state["done"] += 1
            self.integrate_frames((input_list[idx] for idx in shard), after_batch=after_batch)
            while state["done"] < n_exchanges:
                dm.exchange_async()
                state["done"] += 1
            dm.exchange_flush()
            dm.set_streaming(False)
        else:
            self.integrate_frames(input_list[idx] for idx in shard)
        if world > 1 and row_tiles:
            if native_comm:     # every rank holds the summed grid: its tile and the halo rows are slices of it
                r0, r1 = frame_sharding.row_tile(dm.map.shape[0], rank, world)
                top, bottom = (1 if r0 > 0 and r1 > r0 else 0), (1 if r1 < dm.map.shape[0] and r1 > r0 else 0)
                tile = dm.map[r0 - top:r1 + bottom]
            else:
                tile, r0, r1, top, bottom = frame_sharding.sum_grid_row_tile(dm.map)
            rgb_tile, filtered = frame_sharding.render_row_tile(tile, top, bottom, self.label_colors, return_filtered=True)
            color_map = frame_sharding.gather_rgb_rows(rgb_tile, dm.map.shape[0])
            filtered = filtered.clone()
            dm.map.zero_()
            dm.map[r0:r1].copy_(filtered)
        else:
            if world > 1 and not native_comm:
                frame_sharding.sum_grids(dm.map)
            color_map, filtered = filter_and_render(dm.map, self.label_colors, return_filtered=True)
            dm.map.copy_(filtered)  # self.map = apply_filter(self.map)
        dm.notify_map_modified()
        color_map_dev = color_map
        color_map = color_map.cpu().numpy()

        if write_image and rank == 0:
            os.makedirs(self.output_dir, exist_ok=True)
            output_file = osp.join(self.output_dir, "global_map_" + file_name + ".png")
            print("Saving image to", output_file)
            from .utils.image_io import imwrite
            imwrite(output_file, color_map)
        if self.ground_truth_dir != "" and rank == 0:
            from .evaluation import Test
            # scored where it was rendered (smap_eval_counts): src/mapping_replay.py:208-210
            Test(ground_truth_dir=self.ground_truth_dir, logger=self.logger).test_single_map(color_map_dev)
        return color_map

    def mapping_replay_dir(self):
        """Replay every recorded drive in ``cfg.MAPPING.INPUT_DIR`` (``src/mapping_replay.py:146-159``)."""
        from . import replay_io
        if os.path.exists(self.input_dir):
            for file_name in sorted(os.listdir(self.input_dir)):
                if file_name.endswith((".hkl", ".npz")):
                    path = os.path.join(self.input_dir, file_name)
                    print("Loading input file " + path)
                    input_list = replay_io.load_input_list(path)
                    print("Input file loaded!")
                    self.mapping_replay(input_list, file_name[0:-4])

    def mapping_replay_file(self, file_name=None):
        """Replay ``input_list_0`` of ``cfg.MAPPING.INPUT_DIR`` (``src/mapping_replay.py:161-172``)."""
        from . import replay_io
        if file_name is None:
            file_name = "input_list_0.npz" if os.path.exists(os.path.join(self.input_dir, "input_list_0.npz")) \
                else "input_list_0.hkl"
        path = os.path.join(self.input_dir, file_name)
        print("Loading input file " + path)
        input_list = replay_io.load_input_list(path)
        print("Input file loaded!")
        return self.mapping_replay(input_list, file_name[0:-4])


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="B200 semantic mapping replay")
    parser.add_argument("--cfg", dest="config_file", default="", metavar="FILE", help="path to config file", type=str)
    return parser.parse_args(sys.argv[1:] if argv is None else argv)


def main(argv=None):
    cfg = get_cfg_defaults()
    args = parse_args(argv)
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    sm = SemanticMapping(cfg)
    sm.mapping_replay_dir()


if __name__ == "__main__":
    main()
