"""Map evaluation against a ground-truth BEV label map: the step after rendering.

Follows the reference's ``test/test_semantic_mapping.py`` (``convert_labels`` ``:6-18``, ``Test.iou``
``:127-161``, ``Test.test_single_map`` ``:117-125``): the rendered colour map is converted back to integer
labels and IoU / accuracy / missing rate are reported for road, crosswalk and lane.  The reference file
does not parse (a second ``else:`` at ``:70``); the logic below is the intended one.  Host-side numpy --
it runs once per replay and is not on the hot path.
"""
import os

import numpy as np

__all__ = ["convert_labels", "Test"]

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
        generated = convert_labels(global_map)
        truth = self.ground_truth_mask[self.shift_w:generated.shape[0] + self.shift_w,
                                       self.shift_h:generated.shape[1] + self.shift_h]
        return self.iou(truth, generated, verbose=True)

    def iou(self, gmap, generate_map, latex_mode=False, verbose=False):
        ious, accs = [], []
        for cls in self.class_lists:
            g, m = gmap == cls, generate_map == cls
            inter = float(np.sum(g & m))
            union = float(np.sum(g) + np.sum(m) - inter)
            ious.append(inter / union if union else float("nan"))
            accs.append(inter / np.sum(g) if np.sum(g) else float("nan"))
        known = gmap > 0
        miss = 1 - np.sum(known & (generate_map > 0)) / max(np.sum(known), 1)
        accuracy = np.sum((gmap == generate_map)[known]) / max(np.sum(known), 1)
        if verbose:
            self._say("IOU for {}: {}\t{}: {}\t{}:{}\tmIOU: {}".format(
                self.d[0], ious[0], self.d[1], ious[1], self.d[2], ious[2], np.mean(ious)))
            self._say("Accuracy for {}: {}\t{}: {}\t{}:{}\tmean Accuracy: {}".format(
                self.d[0], accs[0], self.d[1], accs[1], self.d[2], accs[2], accuracy))
            self._say("Overall Missing rate: {}".format(miss))
        return ious, miss
