"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the UNMODIFIED reference (``/root/reference/src/mapping_replay.py`` and
friends) inside THIS container by registering stub modules for the packages the
reference imports but never needs on the mapping path (rospy, tf, cv_bridge,
hickle, yacs, matplotlib, ...).  It exists for exactly two jobs:

* ``oracle/make_golden.py`` runs the real reference on seeded synthetic inputs
  and stores its outputs under ``tests/golden/``;
* ``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is
  absent, i.e. on the GPU box) cross-checks the C / numpy restatements in this
  directory against the real thing.

Nothing here can travel to the GPU box (``/root/reference`` does not exist
there); nothing in ``vision_semantic_segmentation_b200`` may import this module.

Stubs (none of them restates reference code; they replace *third-party*
packages that are absent from this image):

* ``yacs.config.CfgNode``  -- dict with attribute access, ``clone``, ``merge_from_file``.
* ``tf`` / ``tf.transformations`` -- ``euler_matrix``, ``quaternion_matrix``,
  ``translation_matrix`` and ``TransformerROS.fromTranslationRotation`` restated
  from the published algorithm of ROS ``geometry/tf`` (Gohlke's
  ``transformations.py``; quaternion order x, y, z, w; static-xyz Euler angles).
  The reference calls them at ``src/mapping_replay.py:141`` and
  ``src/utils/utils_ros.py:104-111``.
* everything else (rospy, cv_bridge, *_msgs, hickle, matplotlib, skimage,
  ``test.test_semantic_mapping`` -- which has a SyntaxError upstream) is an inert
  placeholder.
"""
import copy
import math
import os
import sys
import types

import numpy as np

# oracle/build_ref.py: the reference's modules byte-compiled into one archive (imported with zipimport)
_BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_bytecode.zip")


def _pick_reference():
    env = os.environ.get("SMAP_REFERENCE_DIR")
    if env:
        return env
    if os.path.isfile("/root/reference/src/mapping_replay.py"):
        return "/root/reference"
    return _BUILT   # the GPU box: the byte-compiled copy that travelled with the repo snapshot


REF = _pick_reference()


def reference_available():
    if REF.endswith(".zip"):
        if not os.path.isfile(REF):
            return False
        import zipfile
        with zipfile.ZipFile(REF) as z:
            return "src/mapping_replay.pyc" in z.namelist()
    return any(os.path.isfile(os.path.join(REF, "src", "mapping_replay" + ext)) for ext in (".py", ".pyc"))


def reference_kind():
    """'source' (the tree under /root/reference) or 'bytecode' (oracle/_ref, byte-compiled from it)."""
    return "source" if os.path.isfile(os.path.join(REF, "src", "mapping_replay.py")) else "bytecode"


# --------------------------------------------------------------------------- #
# third-party restatements needed by the stubs
# --------------------------------------------------------------------------- #
_EPS = np.finfo(float).eps * 4.0


def translation_matrix(direction):
    M = np.identity(4)
    M[:3, 3] = direction[:3]
    return M


def quaternion_matrix(quaternion):
    q = np.array(quaternion[:4], dtype=np.float64, copy=True)
    nq = np.dot(q, q)
    if nq < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
        (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)


def euler_matrix(ai, aj, ak, axes='sxyz'):
    if axes != 'sxyz':
        raise NotImplementedError(axes)
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    M = np.identity(4)
    M[0, 0] = cj * ck
    M[0, 1] = sj * sc - cs
    M[0, 2] = sj * cc + ss
    M[1, 0] = cj * sk
    M[1, 1] = sj * ss + cc
    M[1, 2] = sj * cs - sc
    M[2, 0] = -sj
    M[2, 1] = cj * si
    M[2, 2] = cj * ci
    return M


class _Vec(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Pose(object):
    """geometry_msgs/Pose look-alike."""

    def __init__(self, position=(0, 0, 0), orientation=(0, 0, 0, 1)):
        self.position = _Vec(x=float(position[0]), y=float(position[1]), z=float(position[2]))
        self.orientation = _Vec(x=float(orientation[0]), y=float(orientation[1]),
                                z=float(orientation[2]), w=float(orientation[3]))


class _CfgNode(dict):
    """Minimal stand-in for yacs.config.CfgNode (third-party, absent here)."""

    def __init__(self, init=None):
        super(_CfgNode, self).__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split('.')
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v


class _Anything(object):
    """Inert object: any attribute / call returns another inert object."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()


def _inert_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k: _Anything  # PEP 562: any missing name is an inert class
    return m


_loaded = None


def load_reference():
    """Import the reference's mapping modules; returns a namespace with
    SemanticMapping, render_bev_map, render_bev_map_with_thresholds, apply_filter,
    ConfusionMatrix, get_cfg_defaults, camera_setup_1/6, Pose."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF)

    # numpy-1 aliases the reference still uses (src/mapping_replay.py:127)
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'int'):
        np.int = int

    class TransformerROS(object):
        def __init__(self, *a, **k):
            pass

        def fromTranslationRotation(self, translation, rotation):
            return np.dot(translation_matrix(translation), quaternion_matrix(rotation))

    tf_trans = _inert_module('tf.transformations', euler_matrix=euler_matrix,
                             quaternion_matrix=quaternion_matrix,
                             translation_matrix=translation_matrix)
    tf_mod = _inert_module('tf', TransformerROS=TransformerROS, transformations=tf_trans)
    yacs_cfg = types.ModuleType('yacs.config')
    yacs_cfg.CfgNode = _CfgNode
    yacs = types.ModuleType('yacs')
    yacs.config = yacs_cfg
    geometry_msgs_msg = _inert_module('geometry_msgs.msg', Pose=Pose)

    class _DummyTest(object):
        def __init__(self, *a, **k):
            pass

        def test_single_map(self, *a, **k):
            return None

    stubs = {
        'tf': tf_mod, 'tf.transformations': tf_trans,
        'yacs': yacs, 'yacs.config': yacs_cfg,
        'rospy': _inert_module('rospy'),
        'cv_bridge': _inert_module('cv_bridge'),
        'geometry_msgs': _inert_module('geometry_msgs', msg=geometry_msgs_msg),
        'geometry_msgs.msg': geometry_msgs_msg,
        'sensor_msgs': _inert_module('sensor_msgs'),
        'sensor_msgs.msg': _inert_module('sensor_msgs.msg'),
        'sensor_msgs.point_cloud2': _inert_module('sensor_msgs.point_cloud2'),
        'std_msgs': _inert_module('std_msgs'),
        'std_msgs.msg': _inert_module('std_msgs.msg'),
        'tf_conversions': _inert_module('tf_conversions'),
        'hickle': _inert_module('hickle'),
        'matplotlib': _inert_module('matplotlib'),
        'matplotlib.pyplot': _inert_module('matplotlib.pyplot'),
        'mpl_toolkits': _inert_module('mpl_toolkits'),
        'mpl_toolkits.mplot3d': _inert_module('mpl_toolkits.mplot3d'),
        'skimage': _inert_module('skimage'),
        'skimage.measure': _inert_module('skimage.measure'),
    }
    for name, mod in stubs.items():
        sys.modules.setdefault(name, mod)

    saved_path = list(sys.path)
    saved_test = sys.modules.get('test'), sys.modules.get('test.test_semantic_mapping')
    sys.path[:0] = [REF, os.path.join(REF, 'src')]
    test_pkg = types.ModuleType('test')
    test_pkg.__path__ = []
    test_mod = types.ModuleType('test.test_semantic_mapping')
    test_mod.Test = _DummyTest
    sys.modules['test'] = test_pkg
    sys.modules['test.test_semantic_mapping'] = test_mod
    try:
        import importlib
        mr = importlib.import_module('src.mapping_replay')
        rd = importlib.import_module('src.renderer')
        cam = importlib.import_module('src.camera')
        cm = importlib.import_module('src.data.confusion_matrix')
        cfg = importlib.import_module('src.config.base_cfg')
    finally:
        sys.path[:] = saved_path
        for key, val in zip(('test', 'test.test_semantic_mapping'), saved_test):
            if val is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = val

    ns = types.SimpleNamespace(
        SemanticMapping=mr.SemanticMapping,
        render_bev_map=rd.render_bev_map,
        render_bev_map_with_thresholds=rd.render_bev_map_with_thresholds,
        apply_filter=rd.apply_filter,
        ConfusionMatrix=cm.ConfusionMatrix,
        get_cfg_defaults=cfg.get_cfg_defaults,
        camera_setup_1=cam.camera_setup_1,
        camera_setup_6=cam.camera_setup_6,
        Camera=cam.Camera,
        Pose=Pose,
        quaternion_matrix=quaternion_matrix,
        euler_matrix=euler_matrix,
        translation_matrix=translation_matrix,
    )
    _loaded = ns
    return ns


def load_reference_color_map():
    """The reference's label-image producer functions, unmodified:
    ``apply_color_map`` and ``get_labels`` of
    ``src/network/deeplab_v3_plus/data/utils/mapillary_visualization.py`` (the dataset class and matplotlib that the
    module imports at the top are inert placeholders: neither function touches them)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF)
    import importlib.util
    stubs = {
        'deeplab_v3_plus': _inert_module('deeplab_v3_plus'),
        'deeplab_v3_plus.data': _inert_module('deeplab_v3_plus.data'),
        'deeplab_v3_plus.data.dataset': _inert_module('deeplab_v3_plus.data.dataset'),
        'deeplab_v3_plus.data.dataset.mapillary': _inert_module('deeplab_v3_plus.data.dataset.mapillary',
                                                               MapillaryVistas=_Anything),
        'matplotlib': _inert_module('matplotlib'),
        'matplotlib.pyplot': _inert_module('matplotlib.pyplot'),
    }
    added = [name for name in stubs if name not in sys.modules]
    for name in added:
        sys.modules[name] = stubs[name]
    try:
        path = os.path.join(REF, 'src', 'network', 'deeplab_v3_plus', 'data', 'utils', 'mapillary_visualization.py')
        spec = importlib.util.spec_from_file_location('_ref_mapillary_visualization', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in added:
            if name.startswith('deeplab_v3_plus'):
                sys.modules.pop(name, None)
    return types.SimpleNamespace(apply_color_map=mod.apply_color_map, get_labels=mod.get_labels,
                                 config_19=os.path.join(REF, 'config', 'config_19.json'))


def load_reference_convex_hull():
    """The reference's ``generate_convex_hull`` (``src/semantic_convex_hull.py:17-91``), unmodified.  Its one absent
    dependency on the path is ``skimage.measure.label`` (scikit-image is not in this image): the stand-in below labels
    the 8-connected components with ``scipy.ndimage.label`` -- numbered in raster order of their first pixel, as
    scikit-image numbers them -- and the function's result does not depend on the numbering anyway unless the caller
    names component indices itself (``index_to_vitualize``)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF)
    import importlib.util
    from scipy import ndimage

    def label(img, connectivity=None, **kw):
        nd = img.ndim
        conn = nd if connectivity is None else connectivity
        structure = ndimage.generate_binary_structure(nd, conn)
        out, _ = ndimage.label(img, structure=structure)
        return out.astype(np.int64)

    measure = _inert_module('skimage.measure', label=label)
    stubs = {'skimage': _inert_module('skimage', measure=measure), 'skimage.measure': measure,
             'matplotlib': _inert_module('matplotlib'), 'matplotlib.pyplot': _inert_module('matplotlib.pyplot'),
             'rospy': _inert_module('rospy')}
    saved = {name: sys.modules.get(name) for name in stubs}
    sys.modules.update(stubs)
    try:
        path = os.path.join(REF, 'src', 'semantic_convex_hull.py')
        spec = importlib.util.spec_from_file_location('_ref_semantic_convex_hull', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name, val in saved.items():
            if val is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = val
    return mod.generate_convex_hull
