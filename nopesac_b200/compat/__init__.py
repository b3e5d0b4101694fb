"""Minimal stand-ins for the detectron2 / yacs / fvcore surface the hot path touches, so the reference's
registries and `configs/inference_mp3d.yaml` work unchanged where those packages are not installed
(they are not in this image).  If detectron2 is importable its own Registry / CfgNode are used instead.

  Registry      detectron2.utils.registry.Registry (fvcore.common.registry)
  CfgNode       yacs.config.CfgNode semantics: attribute access, _BASE_ inheritance,
                merge_from_file / merge_from_list / freeze / clone
  ShapeSpec     detectron2.layers.ShapeSpec
"""
from __future__ import annotations

import copy
import os
from ast import literal_eval
from collections import namedtuple
from typing import Any, Dict, Iterable, Optional

import yaml

try:  # pragma: no cover - not installed in this image
    from detectron2.utils.registry import Registry as _D2Registry  # type: ignore
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    _D2Registry = None
    HAVE_DETECTRON2 = False


class Registry:
    """name -> object mapping with the `@REG.register()` decorator form the reference uses
    (camera_head.py:21-34, matching_head.py:15-23)."""

    def __init__(self, name: str):
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any):
        if name in self._obj_map:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj: Any = None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())


ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=(None, None, None, None))

BASE_KEY = "_BASE_"


class CfgNode(dict):
    """yacs-style config node.  Differences from yacs, on purpose: merging a key that does not exist yet
    is allowed (the full detectron2 default tree is not replicated here), type mismatches still raise."""

    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict: Optional[dict] = None):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access ---------------------------------------------------------------------
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value

    def is_frozen(self) -> bool:
        return self.__dict__[CfgNode.IMMUTABLE]

    def _set_immutable(self, flag: bool):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return out

    # merging --------------------------------------------------------------------------------
    @staticmethod
    def load_yaml_with_base(filename: str) -> dict:
        """detectron2's `_BASE_` inheritance: the base file is loaded first, the child overrides it."""
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}

        def merge_a_into_b(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and isinstance(b.get(k), dict):
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if BASE_KEY in cfg:
            base = cfg.pop(BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not base.startswith("/"):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename: str):
        self.merge_from_other_cfg(CfgNode(self.load_yaml_with_base(cfg_filename)))

    def merge_from_other_cfg(self, other: "CfgNode"):
        if self.is_frozen():
            raise AttributeError("CfgNode is immutable")
        _merge(other, self, [])

    def merge_from_list(self, cfg_list: Iterable):
        cfg_list = list(cfg_list)
        if len(cfg_list) % 2 != 0:
            raise AssertionError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        if self.is_frozen():
            raise AttributeError("CfgNode is immutable")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            d = self
            keys = full_key.split(".")
            for sub in keys[:-1]:
                if sub not in d:
                    raise KeyError(f"Non-existent key: {full_key}")
                d = d[sub]
            if keys[-1] not in d:
                raise KeyError(f"Non-existent key: {full_key}")
            value = _decode(v)
            d[keys[-1]] = _coerce(value, d[keys[-1]], full_key)

    def dump(self) -> str:
        def to_dict(n):
            return {k: to_dict(v) if isinstance(v, CfgNode) else v for k, v in n.items()}
        return yaml.safe_dump(to_dict(self))


def _decode(v):
    if isinstance(v, str):
        try:
            return literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _coerce(new, old, key):
    if old is None or new is None or type(new) is type(old):
        return new
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, (tuple, list)) and isinstance(new, (tuple, list)):
        return type(old)(new)
    if isinstance(old, str) and not isinstance(new, str):
        raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) for config key: {key}")
    if isinstance(old, bool) != isinstance(new, bool) or not isinstance(new, type(old)):
        raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) with values ({old} vs. {new}) for config key: {key}")
    return new


def _merge(a: CfgNode, b: CfgNode, stack):
    for k, v in a.items():
        full = ".".join(stack + [k])
        if isinstance(v, str):
            v = _decode(v) if k in b and not isinstance(b[k], str) else v
        if k in b:
            if isinstance(v, CfgNode) and isinstance(b[k], CfgNode):
                _merge(v, b[k], stack + [k])
            else:
                dict.__setitem__(b, k, _coerce(copy.deepcopy(v), b[k], full))
        else:
            dict.__setitem__(b, k, copy.deepcopy(v))
