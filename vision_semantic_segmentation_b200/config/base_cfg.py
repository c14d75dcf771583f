"""Configuration tree of the mapping path.

Mirrors the option names and defaults of the reference's ``src/config/base_cfg.py:12-83``
(``get_cfg_defaults()`` -> node with ``clone`` / ``merge_from_file`` / ``merge_from_list``).
yacs is not available in the target image, so ``CfgNode`` below is a small
attribute-access dictionary with the subset of the yacs interface the reference
entry points use (``src/mapping_replay.py:321-325``).

The ``VISION_SEM_SEG.SEM_SEG_NETWORK`` subtree (network weights, backbone) belongs to
the segmentation node, which is outside the mapping path; only the keys a mapping
config file may legally mention are kept so that the reference's YAML files merge.
"""
import copy

__all__ = ["CfgNode", "get_cfg_defaults"]


class CfgNode(dict):
    """Attribute-access dict; unknown keys are rejected on merge (as yacs does)."""

    def __init__(self, init=None):
        super(CfgNode, self).__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if (isinstance(v, dict) and not isinstance(v, CfgNode)) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def _merge_dict(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("Non-existent config key: {}".format(full))
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError("Config key {} expects a mapping".format(full))
                self[k]._merge_dict(v, path + [k])
            else:
                self[k] = _coerce(v, self[k], full)

    def merge_from_file(self, cfg_filename):
        import yaml
        with open(cfg_filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self._merge_dict(loaded, [])

    def merge_from_other_cfg(self, other):
        self._merge_dict(other, [])

    def merge_from_list(self, cfg_list):
        if len(cfg_list) % 2 != 0:
            raise ValueError("Override list has odd length: {}".format(cfg_list))
        for full, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            parts = full.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError("Non-existent config key: {}".format(full))
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: {}".format(full))
            if isinstance(v, str):
                import yaml
                try:
                    v = yaml.safe_load(v)
                except Exception:
                    pass
            node[parts[-1]] = _coerce(v, node[parts[-1]], full)

    def dump(self):
        import yaml
        return yaml.safe_dump(_to_plain(self), default_flow_style=None)

    def __str__(self):
        return self.dump()


def _to_plain(node):
    if isinstance(node, dict):
        return {k: _to_plain(v) for k, v in node.items()}
    if isinstance(node, tuple):
        return list(node)
    return node


def _coerce(new, old, full):
    """Same permissive casts yacs allows: int<->float, list<->tuple; else types must agree."""
    if old is None or new is None or type(new) is type(old):
        return new
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, (list, tuple)) and isinstance(new, (list, tuple)):
        return type(old)(new)
    if isinstance(old, int) and not isinstance(old, bool) and isinstance(new, float):
        return new
    raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(old), type(new), full))


_C = CfgNode()

# ---- general (reference src/config/base_cfg.py:30-57) ----
_C.TASK_NAME = "cfn_mtx_with_intensity"
_C.OUTPUT_DIR = "@/outputs"          # '@' = project root
_C.TEST_END_TIME = 1581541450
_C.GROUND_TRUTH_DIR = ""
_C.RNG_SEED = -1
_C.LABELS = [2, 1, 8, 10, 3]
_C.LABELS_NAMES = ["road", "crosswalk", "lane", "vegetation", "sidewalk"]
_C.LABEL_COLORS = [
    [128, 64, 128],
    [140, 140, 200],
    [255, 255, 255],
    [107, 142, 35],
    [244, 35, 232],
]

# ---- mapping (reference src/config/base_cfg.py:62-83) ----
_C.MAPPING = CfgNode()
_C.MAPPING.RESOLUTION = 0.1
_C.MAPPING.BOUNDARY = [[100, 300], [800, 1000]]
_C.MAPPING.DEPTH_METHOD = "points_map"
_C.MAPPING.PCD = CfgNode()
_C.MAPPING.PCD.USE_INTENSITY = True
_C.MAPPING.PCD.RANGE_MAX = 100.0
_C.MAPPING.CONFUSION_MTX = CfgNode()
_C.MAPPING.CONFUSION_MTX.LOAD_PATH = ""
_C.MAPPING.INPUT_DIR = ""

# ---- segmentation node options a mapping YAML may carry (reference :88-112) ----
_C.VISION_SEM_SEG = CfgNode()
_C.VISION_SEM_SEG.IMAGE_SCALE = 1.0
_C.VISION_SEM_SEG.SEM_SEG_NETWORK = CfgNode({
    "OUTPUT_DIR": "@", "OUTPUT_NAME": "",
    "TRAIN_DATASET": "Mapillary",
    "DATASET_CONFIG": "/mnt/avl_shared/qinru/iros2020/resnext50_os8/config.json",
    "DATASET": {"NAME": "AVL", "IN_CHANNELS": 3, "NUM_CLASSES": 19, "ROOT_DIR": ""},
    "MODEL": {
        "TYPE": "DeepLabv3+",
        "WEIGHT": "/mnt/avl_shared/qinru/iros2020/resnext50_os8/run1/model_best.pth",
        "SYNC_BN": False,
        "BACKBONE": "resnext50_32x4d",
        "OUTPUT_STRIDE": 8,
        "ASPP": {"OUT_CHANNELS": 256, "DROPOUT": 0.5},
        "DECODER": {"LOW_LEVEL_OUT_CHANNELS": 256},
    },
})


def get_cfg_defaults():
    """Fresh copy of the defaults (reference ``src/config/base_cfg.py:15-19``)."""
    return _C.clone()
