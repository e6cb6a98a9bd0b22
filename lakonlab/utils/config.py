"""A small stand-in for `mmcv.Config` / `DictAction` (mmcv is not installable offline; train.py:29,145-147 uses them).

Supports what the reference's config files use: python files evaluated top to bottom, `_base_` lists (relative paths)
whose dicts are merged recursively with the child winning, `_delete_=True` to replace instead of merge, attribute access,
`merge_from_dict` with dotted keys (the `--cfg-options a.b=1` form), `pretty_text` and `dump`.
"""
from __future__ import annotations

import argparse
import ast
import os
import pprint
from typing import Any, Dict


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


def _merge(base: Dict[str, Any], child: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"} if isinstance(v, dict) else v
    return out


def _load_file(path: str) -> Dict[str, Any]:
    path = os.path.abspath(path)
    with open(path) as f:
        src = f.read()
    scope: Dict[str, Any] = {"__file__": path}
    exec(compile(src, path, "exec"), scope)
    own = {k: v for k, v in scope.items() if not k.startswith("__") and not callable(v) and not hasattr(v, "__loader__")}
    bases = own.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged: Dict[str, Any] = {}
    for b in bases:
        merged = _merge(merged, _load_file(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, own)


class Config:
    def __init__(self, cfg_dict: Dict[str, Any], filename: str = None):
        object.__setattr__(self, "_cfg", _wrap(cfg_dict))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename: str) -> "Config":
        return Config(_load_file(filename), filename)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_cfg"), name)

    def __setattr__(self, name, value):
        self._cfg[name] = _wrap(value)

    def __getitem__(self, k):
        return self._cfg[k]

    def __contains__(self, k):
        return k in self._cfg

    def get(self, k, default=None):
        return self._cfg.get(k, default)

    def merge_from_dict(self, options: Dict[str, Any]):
        for dotted, v in options.items():
            d = self._cfg
            keys = dotted.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, ConfigDict())
            d[keys[-1]] = _wrap(v)

    def to_dict(self) -> Dict[str, Any]:
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(un(x) for x in v)
            return v
        return un(self._cfg)

    @property
    def pretty_text(self) -> str:
        return "\n".join(f"{k} = {pprint.pformat(v, width=110)}" for k, v in self.to_dict().items())

    def dump(self, path: str):
        with open(path, "w") as f:
            f.write(self.pretty_text + "\n")


class DictAction(argparse.Action):
    """`--cfg-options key=value [key=value ...]`; values parsed as python literals when possible."""

    @staticmethod
    def _parse(v: str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v

    def __call__(self, parser, namespace, values, option_string=None):
        opts = {}
        for kv in values:
            k, _, v = kv.partition("=")
            opts[k] = self._parse(v)
        setattr(namespace, self.dest, opts)
