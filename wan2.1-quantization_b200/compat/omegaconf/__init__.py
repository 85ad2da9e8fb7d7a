"""Minimal stand-in for the `omegaconf` package (2.3.0 in the reference's
environment.yml:217), used ONLY when the real package is not installed.

The reference's qdiff imports `omegaconf.ListConfig` at module import time
(quant_utils/qdiff/base/base_quantizer.py:9, quant_layer.py:6,
mixed_precision_quantizer.py:9, quant_attn.py:6) and its scripts build the
quant_config with `OmegaConf.load(path)` (examples/Wan2.1/quant_generate.py:347).
This shim provides the subset those call sites use: attribute + item + `.get`
access on mappings, `ListConfig` as a list subclass, `OmegaConf.create/load/
to_container`.  It is a config container, not part of the numeric path.
"""
from __future__ import annotations

import collections.abc as _abc

__version__ = "0.0-b200q-shim"


class ListConfig(_abc.MutableSequence):
    """Like the real ListConfig this is a Sequence but NOT a `list` subclass: the
    reference relies on `isinstance(n_bits, list)` being False for a ListConfig
    (quant_utils/qdiff/base/base_quantizer.py:23-24)."""

    def __init__(self, items=()):
        self._items = [_wrap(v) for v in items]

    def __getitem__(self, i):
        return self._items[i]

    def __setitem__(self, i, v):
        self._items[i] = _wrap(v)

    def __delitem__(self, i):
        del self._items[i]

    def __len__(self):
        return len(self._items)

    def insert(self, i, v):
        self._items.insert(i, _wrap(v))

    def __eq__(self, other):
        return list(self._items) == list(other) if isinstance(other, (list, tuple, ListConfig)) else NotImplemented

    def __repr__(self):
        return repr(self._items)


class DictConfig(dict):
    def __init__(self, mapping=None, **kw):
        super().__init__()
        for k, v in dict(mapping or {}, **kw).items():
            dict.__setitem__(self, k, _wrap(v))

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        try:
            return self[key]
        except KeyError:
            # real OmegaConf (struct mode off) returns None for missing keys
            return None

    def __setattr__(self, key, value):
        self[key] = _wrap(value)

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, _wrap(value))

    def get(self, key, default=None):
        v = dict.get(self, key, default)
        return default if v is None else v


def _wrap(v):
    if isinstance(v, (DictConfig, ListConfig)):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(v)
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, (list, ListConfig)):
        return [_unwrap(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            import yaml
            obj = yaml.safe_load(obj)
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def load(path):
        import yaml
        with open(path, "r") as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)

    @staticmethod
    def to_yaml(cfg):
        import yaml
        return yaml.safe_dump(_unwrap(cfg))

    @staticmethod
    def is_list(obj):
        return isinstance(obj, ListConfig)

    @staticmethod
    def is_dict(obj):
        return isinstance(obj, DictConfig)
