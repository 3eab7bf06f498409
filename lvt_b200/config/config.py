"""Config tree with the yacs/fvcore surface the reference uses (vidgen/config/config.py:9-106):
attribute access, `_BASE_` YAML inheritance, literal decoding of string values ("(16, 1, 1)"),
merge_from_file / merge_from_list, freeze / clone / dump.  `get_cfg()` returns the default tree
(`defaults.py`), so the reference's configs/vqvae/*.yaml and configs/vt/*.yaml load unchanged."""
import ast
import copy
import os

import yaml


def _decode(v):
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


class CfgNode(dict):
    _IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None):
        super().__init__()
        object.__setattr__(self, CfgNode._IMMUTABLE, False)
        for k, v in (init_dict or {}).items():
            dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v)

    # attribute access -----------------------------------------------------------------
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def is_frozen(self):
        return object.__getattribute__(self, CfgNode._IMMUTABLE)

    def _immutable(self, flag):
        object.__setattr__(self, CfgNode._IMMUTABLE, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    # loading / merging ----------------------------------------------------------------
    @staticmethod
    def load_yaml_with_base(filename, allow_unsafe=True):
        with open(filename) as f:
            cfg = yaml.load(f, Loader=yaml.UnsafeLoader if allow_unsafe else yaml.SafeLoader) or {}
        if "_BASE_" in cfg:
            base = cfg.pop("_BASE_")
            if not os.path.isabs(base):
                base = os.path.join(os.path.dirname(filename), base)
            merged = CfgNode.load_yaml_with_base(base, allow_unsafe)

            def rec(src, dst):
                for k, v in src.items():
                    if isinstance(v, dict) and isinstance(dst.get(k), dict):
                        rec(v, dst[k])
                    else:
                        dst[k] = v
            rec(cfg, merged)
            return merged
        return cfg

    def merge_from_other_cfg(self, other, _path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f"Non-existent config key: {_path}{k}")
            if isinstance(v, dict) and isinstance(self[k], CfgNode):
                self[k].merge_from_other_cfg(v, _path + k + ".")
                continue
            v = _decode(v)
            old = self[k]
            if isinstance(old, tuple) and isinstance(v, list):
                v = tuple(v)
            elif isinstance(old, list) and isinstance(v, tuple):
                v = list(v)
            elif isinstance(old, float) and isinstance(v, int) and not isinstance(v, bool):
                v = float(v)
            dict.__setitem__(self, k, v)

    def merge_from_file(self, cfg_filename, allow_unsafe=True):
        loaded = CfgNode(self.load_yaml_with_base(cfg_filename, allow_unsafe))
        ver = loaded.get("VERSION", None)
        assert ver is None or ver <= self.VERSION, f"Cannot merge a v{ver} config into a v{self.VERSION} config."
        self.merge_from_other_cfg(loaded)

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0, "Override list has odd length"
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                if k not in node:
                    raise KeyError(f"Non-existent key: {full_key}")
                node = node[k]
            if keys[-1] not in node:
                raise KeyError(f"Non-existent key: {full_key}")
            node.merge_from_other_cfg({keys[-1]: v})

    def dump(self, **kwargs):
        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v)
                    for k, v in n.items()}
        return yaml.safe_dump(plain(self), **kwargs)


def get_cfg() -> CfgNode:
    """A copy of the default config (reference config/config.py:76-85)."""
    from .defaults import defaults
    return CfgNode(defaults())
