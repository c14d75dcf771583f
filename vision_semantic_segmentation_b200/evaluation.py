"""Map evaluation against a ground-truth BEV label map: the step after rendering.

Follows the reference's ``test/test_semantic_mapping.py`` (``convert_labels`` ``:6-18``, ``Test.iou``
``:127-161``, ``Test.test_single_map`` ``:117-125``): the rendered colour map is converted back to integer
labels and IoU / accuracy / missing rate are reported for road, crosswalk and lane.  The reference file
does not parse (a second ``else:`` at ``:70``), but ``convert_labels`` and the two ``Test`` methods are complete on
their own: ``oracle/make_golden_eval.py`` slices those lines out of the unmodified file, runs them, and
``tests/golden/eval.json`` pins this module (and ``smap_eval_counts``) to what they return, print and raise.

``Test.test_single_map`` -- what ``mapping_replay`` calls on the rendered map -- counts on the device
(``smap_eval_counts``: colour -> label and every sum of ``iou`` in one pass over the image, integer counts, exact), so
the rendered map does not have to leave the GPU to be scored; ``convert_labels`` and ``Test.iou`` keep the reference's
numpy signatures (label arrays in, lists out) for callers that hold host arrays.
"""
import os

import numpy as np

__all__ = ["convert_labels", "device_counts", "scores_from_counts", "Test"]

_EVAL_COLORS = (((128, 64, 128), 1), ((140, 140, 200), 2), ((255, 255, 255), 3), ((244, 35, 232), 4),
                ((107, 142, 35), 5))


def convert_labels(gmap, mask=None):
    """(H, W, 3) colour map -> (H, W) labels 1 road, 2 crosswalk, 3 lane, 4 sidewalk, 5 vegetation."""
    if mask is None:
        mask = np.ones(gmap.shape[:2])
    else:
        mask = mask[:gmap.shape[0], :gmap.shape[1]]
    out = np.zeros(gmap.shape[:2])
    for color, value in _EVAL_COLORS:
        out[np.logical_and(np.all(gmap == np.array(color), axis=-1), mask)] = value
    return out


def _as_device_u8(a, device, what):
    """uint8 CUDA tensor of a label map / mask / image given as numpy or torch; values must survive the cast."""
    from . import _native
    torch = _native.require_cuda()
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(a))
    a = a.to(device)
    if a.dtype != torch.uint8:
        b = a.to(torch.uint8)
        if not torch.equal(b.to(a.dtype), a):
            raise ValueError("%s holds values that are not integers in 0..255" % what)
        a = b
    return a.contiguous()


def device_counts(color_map, truth, shift_w=0, shift_h=0, mask=None):
    """The twelve integer sums behind ``Test.iou`` for a rendered (H, W, 3) colour map against the ground-truth label
    map slice ``truth[shift_w : H + shift_w, shift_h : W + shift_h]`` (``test_single_map``), computed by
    ``smap_eval_counts`` on the GPU.  Returns a list of 12 ints (layout: ``include/smap.h``)."""
    from . import _native
    torch = _native.require_cuda()
    lib = _native.load()
    if isinstance(color_map, torch.Tensor) and color_map.is_cuda:
        dev = color_map.device
    elif isinstance(truth, torch.Tensor) and truth.is_cuda:
        dev = truth.device
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
    rgb = _as_device_u8(color_map, dev, "the colour map")
    if rgb.dim() != 3 or rgb.shape[2] != 3:
        raise ValueError("the colour map must be (H, W, 3)")
    gt = _as_device_u8(truth, dev, "the ground-truth label map")
    if gt.dim() != 2:
        raise ValueError("the ground-truth label map must be 2-D")
    mh, mw = int(rgb.shape[0]), int(rgb.shape[1])
    if shift_w < 0 or shift_h < 0 or shift_w + mh > gt.shape[0] or shift_h + mw > gt.shape[1]:
        # numpy would silently clip the slice and then fail to broadcast it against the map
        raise ValueError("the map (%d x %d at %d, %d) does not lie inside the ground truth %s"
                         % (mh, mw, shift_w, shift_h, tuple(gt.shape)))
    mk = None if mask is None else _as_device_u8((mask != 0) if not isinstance(mask, torch.Tensor) else (mask != 0),
                                                   dev, "the mask")
    with torch.cuda.device(dev):
        counts = torch.empty(12, dtype=torch.int64, device=dev)
        _native.check(lib.smap_eval_counts(
            rgb.data_ptr(), mh, mw, gt.data_ptr(), int(gt.shape[0]), int(gt.shape[1]), int(shift_w), int(shift_h),
            None if mk is None else mk.data_ptr(), 0 if mk is None else int(mk.shape[0]),
            0 if mk is None else int(mk.shape[1]), counts.data_ptr(), dev.index,
            _native.current_stream_ptr(dev)))
        return [int(v) for v in counts.cpu().tolist()]


def scores_from_counts(counts):
    """(ious, accs, accuracy, miss) from the sums, with the reference's formulas AND its division behaviour
    (``Test.iou``, test/test_semantic_mapping.py:134-145): ``iou = intersection / union`` divides two Python floats, so
    a class absent from both maps raises ``ZeroDivisionError``; the other quotients are numpy divisions (an empty
    denominator gives nan, silently here, with numpy's RuntimeWarning in the reference).  Pinned to the reference's own
    output by tests/golden/eval.json (oracle/make_golden_eval.py)."""
    ious, accs = [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(3):
            inter = float(counts[k])
            g, m = np.int64(counts[3 + k]), np.int64(counts[6 + k])
            union = float(g + m - inter)
            ious.append(inter / union)          # ZeroDivisionError on an empty union, as the reference
            accs.append(inter / g)
        known = np.int64(counts[9])
        miss = 1 - np.int64(counts[10]) / known
        accuracy = np.int64(counts[11]) / known
    return ious, accs, accuracy, miss


class Test(object):
    def __init__(self, ground_truth_dir="./", shift_h=0, shift_w=0, logger=None):
        truth = os.path.join(ground_truth_dir, "truth.npy")
        if not os.path.exists(truth):
            raise FileNotFoundError("%s not found: pre-process the ground-truth BEV images into truth.npy / mask.npy"
                                    % truth)
        self.ground_truth_mask = np.load(truth)
        mask = os.path.join(ground_truth_dir, "mask.npy")
        self.mask = np.load(mask) if os.path.exists(mask) else None
        self.d = {0: "road", 1: "crosswalk", 2: "lane"}
        self.class_lists = [1, 2, 3]
        self.shift_w, self.shift_h = shift_w, shift_h
        self.logger = logger

    def _say(self, msg):
        if self.logger is not None:
            self.logger.log(msg)
        else:
            print(msg)

    def test_single_map(self, global_map):
        """Score a rendered map (numpy or CUDA tensor, (H, W, 3)) against the ground truth; counts on the device."""
        dev = _cuda_device_of(global_map)
        cache = self.__dict__.setdefault("_truth_dev", {})   # one copy of the ground truth per GPU
        if dev not in cache:
            cache[dev] = _as_device_u8(self.ground_truth_mask, dev, "the ground-truth label map")
        counts = device_counts(global_map, cache[dev], self.shift_w, self.shift_h)
        ious, accs, accuracy, miss = scores_from_counts(counts)
        self._report(ious, accs, accuracy, miss)
        return ious, miss

    def iou(self, gmap, generate_map, latex_mode=False, verbose=False):
        """Host-array form of the reference's ``Test.iou`` (label maps in, ``(iou_lists, miss)`` out): the same twelve
        sums, taken with numpy, through the same ``scores_from_counts``."""
        counts = [0] * 12
        for k, cls in enumerate(self.class_lists):
            g, m = gmap == cls, generate_map == cls
            counts[k], counts[3 + k], counts[6 + k] = int(np.sum(g & m)), int(np.sum(g)), int(np.sum(m))
        known = gmap > 0
        counts[9] = int(np.sum(known))
        counts[10] = int(np.sum(known & (generate_map > 0)))
        counts[11] = int(np.sum((gmap == generate_map)[known]))
        ious, accs, accuracy, miss = scores_from_counts(counts)
        if verbose:
            self._report(ious, accs, accuracy, miss)
        return ious, miss

    def _report(self, ious, accs, accuracy, miss):
        self._say("IOU for {}: {}\t{}: {}\t{}:{}\tmIOU: {}".format(
            self.d[0], ious[0], self.d[1], ious[1], self.d[2], ious[2], np.mean(ious)))
        self._say("Accuracy for {}: {}\t{}: {}\t{}:{}\tmean Accuracy: {}".format(
            self.d[0], accs[0], self.d[1], accs[1], self.d[2], accs[2], accuracy))
        self._say("Overall Missing rate: {}".format(miss))


def _cuda_device_of(a):
    from . import _native
    torch = _native.require_cuda()
    if isinstance(a, torch.Tensor) and a.is_cuda:
        return a.device
    return torch.device("cuda", torch.cuda.current_device())
