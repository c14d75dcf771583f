"""Host-side 4x4 rigid transforms (float64, numpy).

The reference obtains these from ROS ``tf`` (``tf.transformations.euler_matrix`` at
``src/mapping_replay.py:141``; ``TransformerROS.fromTranslationRotation`` at
``src/utils/utils_ros.py:104-111``).  ROS is not part of the target environment, so
the published formulas (quaternion order x, y, z, w; static-xyz Euler angles) are
written out here.  They run once per frame on the host; the GPU consumes the
resulting matrix bit-for-bit, so device parity does not depend on them.
"""
import math

import numpy as np

_EPS = np.finfo(np.float64).eps * 4.0


class _XYZ(object):
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class _XYZW(object):
    __slots__ = ("x", "y", "z", "w")

    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)


class Pose(object):
    """Duck-type of ``geometry_msgs/Pose``: ``.position.{x,y,z}``, ``.orientation.{x,y,z,w}``."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=(0.0, 0.0, 0.0, 1.0)):
        self.position = _XYZ(*position)
        self.orientation = _XYZW(*orientation)

    def as_array(self):
        p, o = self.position, self.orientation
        return np.array([p.x, p.y, p.z, o.x, o.y, o.z, o.w], dtype=np.float64)

    @staticmethod
    def from_array(a):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        return Pose(a[0:3], a[3:7])


def translation_matrix(t):
    m = np.identity(4)
    m[0, 3], m[1, 3], m[2, 3] = t[0], t[1], t[2]
    return m


def quaternion_matrix(q_xyzw):
    """Homogeneous rotation matrix of a quaternion given as (x, y, z, w)."""
    q = np.array(q_xyzw[:4], dtype=np.float64)
    n = float(np.dot(q, q))
    if n < _EPS:
        return np.identity(4)
    q = q * math.sqrt(2.0 / n)
    o = np.outer(q, q)
    m = np.identity(4)
    m[0, 0] = 1.0 - o[1, 1] - o[2, 2]
    m[0, 1] = o[0, 1] - o[2, 3]
    m[0, 2] = o[0, 2] + o[1, 3]
    m[1, 0] = o[0, 1] + o[2, 3]
    m[1, 1] = 1.0 - o[0, 0] - o[2, 2]
    m[1, 2] = o[1, 2] - o[0, 3]
    m[2, 0] = o[0, 2] - o[1, 3]
    m[2, 1] = o[1, 2] + o[0, 3]
    m[2, 2] = 1.0 - o[0, 0] - o[1, 1]
    return m


def euler_matrix(roll, pitch, yaw):
    """Static-frame x-y-z ('sxyz') Euler angles -> homogeneous rotation matrix."""
    sr, sp, sy = math.sin(roll), math.sin(pitch), math.sin(yaw)
    cr, cp, cy = math.cos(roll), math.cos(pitch), math.cos(yaw)
    m = np.identity(4)
    m[0, 0] = cp * cy
    m[0, 1] = sp * (sr * cy) - cr * sy
    m[0, 2] = sp * (cr * cy) + sr * sy
    m[1, 0] = cp * sy
    m[1, 1] = sp * (sr * sy) + cr * cy
    m[1, 2] = sp * (cr * sy) - sr * cy
    m[2, 0] = -sp
    m[2, 1] = cp * sr
    m[2, 2] = cp * cr
    return m


def get_transform_from_pose(pose):
    """pose -> 4x4 'base_link to origin' (reference ``src/utils/utils_ros.py:104-111``)."""
    p, o = pose.position, pose.orientation
    return np.dot(translation_matrix((p.x, p.y, p.z)), quaternion_matrix((o.x, o.y, o.z, o.w)))


def homogenize(x):
    """(d, n) -> (d+1, n) with a row of ones (reference ``src/utils/utils.py:68-70``)."""
    return np.vstack((x, np.ones((1, x.shape[1]))))


def dehomogenize(x):
    """(d+1, n) -> (d, n) dividing by the last row (reference ``src/utils/utils.py:73-75``)."""
    return x[:-1] / x[-1]
