"""ORACLE tooling (authoring container only): makes the UNMODIFIED reference importable.

The reference (`/root/reference`, package `vidgen`) needs three pure-Python packages that are
not installed here and cannot be fetched (no network): fvcore, yacs (via fvcore) and termcolor.
None of them does arithmetic.  `install()` registers small in-memory stand-ins for exactly the
names the reference imports and puts the reference on sys.path; nothing is written to disk and
nothing of the reference is copied.  Used by tests/golden/make_golden.py and
tests/test_oracle_vs_reference.py (from /root/reference, authoring container only) and by
`bench.py --impl reference` / `cpu_baseline` (from the pip-installed copy under baseline/_ref,
git-ignored, which travels to the GPU box; `use_root`).
"""
import copy
import os
import sys
import time
import types

import ast

import yaml

REFERENCE_ROOT = os.environ.get("LVT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vidgen"))


def use_root(path):
    """Point the shim at another copy of the unmodified reference (bench.py --impl reference uses the
    pip-installed one under baseline/_ref, which travels to the GPU box).  Call before install()."""
    global REFERENCE_ROOT
    assert not _installed or path == REFERENCE_ROOT, "reference already imported from another root"
    REFERENCE_ROOT = path


# --------------------------------------------------------------------------------------------
def _decode(v):
    """yacs `_decode_cfg_value`: strings that are Python literals ("(16, 1, 1)") become values."""
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


class CfgNode(dict):
    """Attribute-access config tree with the subset of the yacs/fvcore API the reference uses."""
    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, CfgNode._FROZEN):
            raise AttributeError(f"config is frozen; cannot set {name}")
        self[name] = value

    def is_frozen(self):
        return object.__getattribute__(self, CfgNode._FROZEN)

    def _set_frozen(self, flag):
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    @staticmethod
    def load_yaml_with_base(filename, allow_unsafe=False):
        with open(filename) as f:
            cfg = yaml.load(f, Loader=yaml.UnsafeLoader if allow_unsafe else yaml.SafeLoader) or {}

        def merge(a, b):  # a into b
            for k, v in a.items():
                if isinstance(v, dict) and isinstance(b.get(k), dict):
                    merge(v, b[k])
                else:
                    b[k] = v

        if "_BASE_" in cfg:
            base = cfg.pop("_BASE_")
            if not os.path.isabs(base):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base, allow_unsafe=allow_unsafe)
            merge(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f"Non-existent config key: {k}")
            if isinstance(v, dict) and isinstance(self[k], CfgNode):
                self[k].merge_from_other_cfg(v)
            else:
                v = _decode(v)
                if isinstance(self[k], tuple) and isinstance(v, list):
                    v = tuple(v)
                if isinstance(self[k], list) and isinstance(v, tuple):
                    v = list(v)
                dict.__setitem__(self, k, v)

    def merge_from_file(self, filename, allow_unsafe=False):
        self.merge_from_other_cfg(CfgNode(self.load_yaml_with_base(filename, allow_unsafe)))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for full_key, v in zip(lst[0::2], lst[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                node = node[k]
            v = _decode(v)
            old = node[keys[-1]]
            if isinstance(old, tuple) and isinstance(v, list):
                v = tuple(v)
            if isinstance(old, float) and isinstance(v, int):
                v = float(v)
            dict.__setitem__(node, keys[-1], v)

    def dump(self, **kwargs):
        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else v for k, v in n.items()}
        return yaml.safe_dump(plain(self), **kwargs)


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]


class Checkpointer:
    def __init__(self, model, save_dir="", **kw):
        self.model, self.save_dir = model, save_dir

    def resume_or_load(self, path, resume=True):
        return {}

    def load(self, path, *a, **k):
        return {}

    def save(self, name, **kw):
        pass

    def has_checkpoint(self):
        return False


class PeriodicCheckpointer:
    def __init__(self, checkpointer, period, max_iter=None):
        self.checkpointer, self.period, self.max_iter = checkpointer, period, max_iter

    def step(self, iteration, **kw):
        pass

    def save(self, name, **kw):
        pass


class _PathManager:
    @staticmethod
    def open(path, mode="r", **kw):
        return open(path, mode)

    @staticmethod
    def mkdirs(path):
        os.makedirs(path, exist_ok=True)

    @staticmethod
    def exists(path):
        return os.path.exists(path)

    @staticmethod
    def isfile(path):
        return os.path.isfile(path)

    @staticmethod
    def isdir(path):
        return os.path.isdir(path)

    @staticmethod
    def ls(path):
        return os.listdir(path)

    @staticmethod
    def get_local_path(path):
        return path


class HistoryBuffer:
    def __init__(self, max_length=1000000):
        self._data, self._count, self._sum = [], 0, 0.0

    def update(self, value, iteration=None):
        self._data.append((value, iteration if iteration is not None else self._count))
        self._count += 1
        self._sum += value

    def latest(self):
        return self._data[-1][0]

    def median(self, window):
        import numpy as np
        return float(np.median([v for v, _ in self._data[-window:]]))

    def avg(self, window):
        import numpy as np
        return float(np.mean([v for v, _ in self._data[-window:]]))

    def global_avg(self):
        return self._sum / max(1, self._count)

    def values(self):
        return self._data


class Timer:
    def __init__(self):
        self.reset()

    def reset(self):
        self._start, self._paused, self._total_paused = time.perf_counter(), None, 0.0

    def pause(self):
        self._paused = time.perf_counter()

    def is_paused(self):
        return self._paused is not None

    def resume(self):
        self._total_paused += time.perf_counter() - self._paused
        self._paused = None

    def seconds(self):
        end = self._paused if self._paused is not None else time.perf_counter()
        return end - self._start - self._total_paused


_installed = False


def install():
    """Register the stand-ins and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("fvcore")
    mod("fvcore.common")
    mod("fvcore.common.config", CfgNode=CfgNode)
    mod("fvcore.common.registry", Registry=Registry)
    mod("fvcore.common.checkpoint", Checkpointer=Checkpointer, PeriodicCheckpointer=PeriodicCheckpointer)
    mod("fvcore.common.file_io", PathManager=_PathManager)
    mod("fvcore.common.history_buffer", HistoryBuffer=HistoryBuffer)
    mod("fvcore.common.timer", Timer=Timer)
    if "termcolor" not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            mod("termcolor", colored=lambda s, *a, **k: s)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_cfg(config_relpath, overrides=()):
    """get_cfg() + merge_from_file(<reference>/configs/...) + overrides (tools/train_net.py:60-69)."""
    install()
    from vidgen.config import get_cfg
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, config_relpath))
    cfg.merge_from_list(["MODEL.DEVICE", "cpu"] + list(overrides))
    cfg.freeze()
    return cfg
