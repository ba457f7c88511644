"""Minimal stand-in for the ``profane`` module system the reference builds on (``capreolus/__init__.py:5``).

``profane`` is a third-party package that is not installed in this image; the rerankers here only need
its declarative surface -- ``ConfigOption`` / ``Dependency`` class attributes, ``@Base.register``,
``Base.create(name, config, provide)`` and a ``config`` dict filled from ``config_spec`` defaults -- so that
is what this file provides.  When the real capreolus is importable, INTEGRATION.md shows how the same
classes are registered with it instead.
"""
from __future__ import annotations

import numpy as np


class ConfigOption:
    def __init__(self, key, default_value=None, description="", value_type=None):
        self.key, self.default_value, self.description, self.value_type = key, default_value, description, value_type


class Dependency:
    def __init__(self, key=None, module=None, name=None, default_config_overrides=None, provide_this=False, provide_children=None):
        self.key, self.module, self.name = key, module, name
        self.default_config_overrides = default_config_overrides or {}


module_registry: dict[str, dict[str, type]] = {}


class ModuleBase:
    module_type: str | None = None
    module_name: str | None = None
    requires_random_seed = False
    config_spec: list = []
    dependencies: list = []

    def __init__(self, config=None, provide=None, **kwargs):
        cfg = {opt.key: opt.default_value for opt in type(self).config_spec}
        for k, v in (config or {}).items():
            if k not in cfg and k not in ("name", "seed"):
                raise ValueError(f"unknown config option {k!r} for {type(self).module_type} {type(self).module_name!r}")
            cfg[k] = v
        cfg.setdefault("name", type(self).module_name)
        if type(self).requires_random_seed:
            cfg.setdefault("seed", 42)
            self.rng = np.random.Generator(np.random.PCG64(cfg["seed"]))
        self.config = cfg
        for k, v in (provide or {}).items():
            setattr(self, k, v)
        if hasattr(self, "build"):
            self.build()

    @classmethod
    def register(cls, sub):
        """Class decorator, used as ``@Reranker.register`` (capreolus/reranker/KNRM.py:58)."""
        if not sub.module_name:
            raise ValueError(f"{sub} has no module_name")
        module_registry.setdefault(sub.module_type, {})[sub.module_name] = sub
        return sub

    @classmethod
    def create(cls, name, config=None, provide=None):
        try:
            sub = module_registry[cls.module_type][name]
        except KeyError:
            raise ValueError(f"no {cls.module_type} module named {name!r}; known: {sorted(module_registry.get(cls.module_type, {}))}")
        return sub(config, provide=provide)
