"""Pruner registry with the reference's semantics (LAVIS/lavis/common/registry.py:113-137, 270):
``@registry.register_pruner(name)`` stores the class, ``registry.get_pruner_class(name)`` returns it
(``None`` when unknown); registering a name twice raises ``KeyError``."""
from __future__ import annotations


class Registry:
    mapping = {"pruner_name_mapping": {}}

    @classmethod
    def register_pruner(cls, name):
        def wrap(pruner_cls):
            from .pruners.base import BasePruner

            assert issubclass(pruner_cls, BasePruner), "All pruners must inherit BasePruner class"
            table = cls.mapping["pruner_name_mapping"]
            if name in table:
                raise KeyError("Name '{}' already registered for {}.".format(name, table[name]))
            table[name] = pruner_cls
            return pruner_cls

        return wrap

    @classmethod
    def get_pruner_class(cls, name):
        return cls.mapping["pruner_name_mapping"].get(name, None)

    @classmethod
    def list_pruners(cls):
        return sorted(cls.mapping["pruner_name_mapping"].keys())


registry = Registry()
