"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference pruner files for golden generation.

The reference (ylsung/ECoFLaP) is pure Python.  Its pruner files import a handful of
``lavis.*`` modules that pull in omegaconf/timm (absent here), so those names are replaced by
stub modules in ``sys.modules`` before the files are executed with importlib.  Nothing from the
reference is copied into this repository: the files are executed where they lie under
``$ECOFLAP_REFERENCE_ROOT`` (default ``/root/reference``).

This module can only be used where the reference tree exists (the build container).  It is used
by ``tests/gen_golden.py`` to write the fixtures in ``tests/golden/``; no ``-m gpu`` test,
``smoke()`` or ``bench.py`` imports it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ECOFLAP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "LAVIS", "lavis", "compression", "pruners"))


class _Registry:
    """Minimal stand-in for lavis.common.registry.registry (registry.py:113-137,270)."""

    mapping = {"pruner_name_mapping": {}}

    @classmethod
    def register_pruner(cls, name):
        def wrap(pruner_cls):
            cls.mapping["pruner_name_mapping"][name] = pruner_cls
            return pruner_cls

        return wrap

    @classmethod
    def get_pruner_class(cls, name):
        return cls.mapping["pruner_name_mapping"].get(name, None)


def _prepare_sample(samples, cuda_enabled=True):
    # the reference passes ``device != "cpu"`` (always True for a torch.device); stay on CPU.
    return samples


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _exec(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_LAVIS = None


def load_lavis_pruners():
    """Returns a namespace with the LAVIS copy: WrappedGPT, SparseGPT, LayerSparsity, pruner classes."""
    global _LAVIS
    if _LAVIS is not None:
        return _LAVIS
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    import torch

    # SparseGPT.fasterprune calls torch.cuda.synchronize() unconditionally (sparsegpt_pruner.py:215)
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None

    class _Dummy:  # Blip2T5 / T5 / EVA_CLIP are only imported, never used, by the pruner files
        pass

    _stub("lavis")
    _stub("lavis.common")
    _stub("lavis.common.registry", registry=_Registry)
    _stub("lavis.datasets")
    _stub("lavis.datasets.data_utils", prepare_sample=_prepare_sample)
    _stub("lavis.models")
    _stub("lavis.models.blip2_models")
    _stub("lavis.models.blip2_models.blip2_t5", Blip2T5=_Dummy)
    _stub("lavis.models.t5_models")
    _stub("lavis.models.t5_models.t5", T5=_Dummy)
    _stub("lavis.models.clip_models")
    _stub("lavis.models.clip_models.eva_model", EVA_CLIP=_Dummy)
    _stub("lavis.compression")
    _stub("lavis.compression.pruners")
    base = os.path.join(REF_ROOT, "LAVIS", "lavis", "compression", "pruners")
    ns = types.SimpleNamespace()
    ns.utils = _exec("lavis.compression.pruners.utils", os.path.join(base, "utils.py"))
    ns.base = _exec("lavis.compression.pruners.base_pruner", os.path.join(base, "base_pruner.py"))
    sys.modules["lavis.compression"].BasePruner = ns.base.BasePruner
    ns.layer = _exec(
        "lavis.compression.pruners.layer_single_base_pruner",
        os.path.join(base, "layer_single_base_pruner.py"),
    )
    ns.wanda = _exec("lavis.compression.pruners.wanda_pruner", os.path.join(base, "wanda_pruner.py"))
    ns.sparsegpt = _exec(
        "lavis.compression.pruners.sparsegpt_pruner", os.path.join(base, "sparsegpt_pruner.py")
    )
    ns.glob = _exec("lavis.compression.pruners.global_pruner", os.path.join(base, "global_pruner.py"))
    ns.registry = _Registry
    ns.WrappedGPT = ns.wanda.WrappedGPT
    ns.SparseGPT = ns.sparsegpt.SparseGPT
    ns.LayerSparsity = ns.layer.LayerSparsity
    _LAVIS = ns
    return ns


_COOP = None


def load_coop_pruners():
    """CoOp copy (CLIP): imports as-is, it only needs torch + transformers.Conv1D."""
    global _COOP
    if _COOP is not None:
        return _COOP
    import torch

    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    root = os.path.join(REF_ROOT, "CoOp")
    if root not in sys.path:
        sys.path.insert(0, root)
    # ``trainers/__init__`` pulls in dassl; load the sub-package by hand instead.
    pkg = types.ModuleType("ecf_ref_coop")
    pkg.__path__ = [os.path.join(root, "trainers", "pruners")]
    sys.modules["ecf_ref_coop"] = pkg
    ns = types.SimpleNamespace()
    d = pkg.__path__[0]
    ns.utils = _exec("ecf_ref_coop.utils", os.path.join(d, "utils.py"))
    ns.base = _exec("ecf_ref_coop.base_pruner", os.path.join(d, "base_pruner.py"))
    ns.layer = _exec("ecf_ref_coop.layer_single_base_pruner", os.path.join(d, "layer_single_base_pruner.py"))
    ns.wanda = _exec("ecf_ref_coop.wanda_pruner", os.path.join(d, "wanda_pruner.py"))
    ns.sparsegpt = _exec("ecf_ref_coop.sparsegpt_pruner", os.path.join(d, "sparsegpt_pruner.py"))
    _COOP = ns
    return ns


_UPOP = None


def load_upop_pruners():
    global _UPOP
    if _UPOP is not None:
        return _UPOP
    root = os.path.join(REF_ROOT, "UPop")
    pkg = types.ModuleType("ecf_ref_upop")
    pkg.__path__ = [os.path.join(root, "pruners")]
    sys.modules["ecf_ref_upop"] = pkg
    ns = types.SimpleNamespace()
    d = pkg.__path__[0]
    ns.utils = _exec("ecf_ref_upop.utils", os.path.join(d, "utils.py"))
    ns.base = _exec("ecf_ref_upop.base_pruner", os.path.join(d, "base_pruner.py"))
    ns.layer = _exec("ecf_ref_upop.layer_single_base_pruner", os.path.join(d, "layer_single_base_pruner.py"))
    ns.wanda = _exec("ecf_ref_upop.wanda_pruner", os.path.join(d, "wanda_pruner.py"))
    _UPOP = ns
    return ns
