"""TEST INFRASTRUCTURE (oracle): host restatement of ``generate_convex_hull`` (/root/reference/src/semantic_convex_hull.py:
17-91) -- class mask, ``cv2.erode`` 3 x 3, 8-connected labelling, ``Counter.most_common`` ranking, ``cv2.convexHull`` of
the component's pixels without the first one.

Third-party dependency absent from this image: scikit-image (``skimage.measure.label``, ``scikit-image>=0.11.2`` in the
reference's requirements).  Restated with ``scipy.ndimage.label`` (8-connectivity; both number the components in raster
order of their first pixel, and the result does not depend on the numbering unless the caller names component indices).
Pinned: ``oracle/make_golden_hull.py`` runs the unmodified reference function (through the same stand-in for the missing
import) and commits its vertex lists; ``tests/test_convex_hull.py`` holds this port and the CUDA path to them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import anything under oracle/.
"""
from collections import Counter

import numpy as np


def generate_convex_hull(img_src, index_care_about=1, index_to_vitualize=None, top_number=1, area_threshold=30):
    import cv2
    from scipy import ndimage
    img = np.array(img_src)
    mask = (img == index_care_about).astype(np.uint8)                         # :37-38
    eroded = cv2.erode(mask, np.ones((3, 3), np.uint8), iterations=1)          # :45
    labels, _ = ndimage.label(eroded, structure=np.ones((3, 3)))               # :51
    if np.all(labels == 0):
        return []
    if index_to_vitualize is None:
        count = Counter(labels[labels != 0].reshape(-1)).most_common(top_number)   # :59
        index_to_vitualize = [x[0] for x in count if x[1] > area_threshold]
    vertices = []
    for sel in index_to_vitualize:
        ys, xs = np.where(labels == sel)
        pts = np.stack([xs, ys], 1)[1:].astype(np.int32)                       # first point dropped, (x, y) order :70-71
        hull = cv2.convexHull(pts)
        vertices.append(np.concatenate([np.squeeze(hull), hull[0, :, :].reshape(1, -1)], axis=0).T)
    return vertices


def erode_and_label(img_src, index_care_about=1):
    import cv2
    from scipy import ndimage
    mask = (np.array(img_src) == index_care_about).astype(np.uint8)
    eroded = cv2.erode(mask, np.ones((3, 3), np.uint8), iterations=1)
    return ndimage.label(eroded, structure=np.ones((3, 3)))[0]
