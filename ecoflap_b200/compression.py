"""``load_pruner`` with the reference's calling convention (LAVIS/lavis/compression/__init__.py:29-46):
looks the class up in the registry and calls ``cls(model=, data_loader=, **cfg)``; an unknown name or a bad
keyword surfaces as ``TypeError`` -> message + ``exit(1)``, as in the reference."""
from __future__ import annotations

from . import pruners  # noqa: F401  (registers the classes)
from .pruners.base import BasePruner
from .registry import registry

__all__ = ["BasePruner", "load_pruner", "registry"]


def load_pruner(name, model, data_loader, cfg_path=None, cfg=None):
    if cfg_path is None and cfg is None:
        cfg = None
    elif cfg_path is not None:
        import yaml

        with open(cfg_path, "r") as f:
            cfg = yaml.safe_load(f)
    try:
        pruner = registry.get_pruner_class(name)(model=model, data_loader=data_loader, **cfg)
    except TypeError:
        print(f"Pruner {name} not found. Available pruners:\n" + ", ".join(registry.list_pruners()))
        exit(1)
    return pruner
