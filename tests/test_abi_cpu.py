"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ecoflap_b200.h
declares, refuses to compute without an sm_100 device (no CPU fallback), and the host-side mirror of the reference
API keeps the reference's names.  No compute calls are made here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ecoflap_b200.h")).read()
    return sorted(set(re.findall(r"ECF_API\s+[\w\s\*]+?\b(ecf_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ecoflap_b200 import _abi

    declared = _declared_symbols()
    assert len(declared) >= 17, declared
    lib = ctypes.CDLL(_abi.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_abi.EXPORTED) == declared, "the ctypes binding and the header disagree"
    header = open(os.path.join(ROOT, "include", "ecoflap_b200.h")).read()
    assert _abi.lib.ecf_version() == int(re.search(r"#define\s+ECF_ABI_VERSION\s+(\d+)", header).group(1))


def test_header_cites_the_reference_for_every_entry_point():
    text = open(os.path.join(ROOT, "include", "ecoflap_b200.h")).read()
    for name in ("wanda_pruner.py", "sparsegpt_pruner.py", "layer_single_base_pruner.py"):
        assert name in text
    # every compute entry point carries a file:line citation in the comment block above it
    blocks = re.split(r"\n\s*\n", text)
    for blk in blocks:
        m = re.search(r"ECF_API\s+int\s+(ecf_(?:sqnorm|wanda|group|zo|count|hessian|obs)\w*)", blk)
        if m:
            assert re.search(r"\.py:\d+", blk), f"{m.group(1)} has no reference citation"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a device")
def test_no_cpu_fallback_without_a_device():
    from ecoflap_b200 import _abi, ops

    assert _abi.lib.ecf_device_sm_count() == _abi.ERR_NO_DEVICE
    x = torch.zeros(4, 8)
    s = torch.zeros(8)
    with pytest.raises(RuntimeError):
        ops.sqnorm_accum(x, s, 0.0, 1.0)  # CPU tensors are rejected before the ABI is reached
    buf = (ctypes.c_float * 8)()
    rc = _abi.lib.ecf_sqnorm_accum(ctypes.addressof(buf), 0, 1, 8, 8, ctypes.addressof(buf), 0.0, 1.0, ctypes.addressof(buf), 32, None)
    assert rc == _abi.ERR_NO_DEVICE and b"no CPU fallback" in _abi.lib.ecf_last_error()


def test_workspace_queries_are_pure_host_functions():
    from ecoflap_b200 import _abi

    lib = _abi.lib
    assert lib.ecf_workspace_bytes(_abi.OP_SQNORM, 2056, 1408) >= 65536
    assert lib.ecf_workspace_bytes(_abi.OP_LAYER_THRESH, 6144, 1408) > 32768 * 4
    assert lib.ecf_workspace_bytes(_abi.OP_OBS, 768, 768) > 0
    d = (_abi.SqnormDesc * 2)()
    for i in range(2):
        d[i].x, d[i].scaler_row, d[i].T, d[i].C, d[i].ld, d[i].dtype = 4096, 8192 + 64 * i, 512, 2048, 2048, 2
    # two hooks on the SAME input (q/k share x) are computed once: no extra partial rows
    assert lib.ecf_sqnorm_batched_workspace_bytes(d, 2) == lib.ecf_sqnorm_batched_workspace_bytes(d, 1) > 65536
    d[1].x = 4096 + (1 << 22)
    assert lib.ecf_sqnorm_batched_workspace_bytes(d, 2) > lib.ecf_sqnorm_batched_workspace_bytes(d, 1)
    assert lib.ecf_sqnorm_batched_workspace_bytes(d, 0) == 0


def test_host_mirror_keeps_the_reference_names():
    """registry strings and class names of LAVIS/lavis/compression/pruners (wanda_pruner.py:87,378,660;
    sparsegpt_pruner.py:225,494,752; global_pruner.py:246,254,303) and the CoOp / UPop classes."""
    import ecoflap_b200.compression  # noqa: F401  registers everything
    from ecoflap_b200 import registry

    for name in ("t5_wanda_pruner", "vit_wanda_pruner", "blipt5_wanda_pruner", "t5_sparsegpt_pruner",
                 "vit_sparsegpt_pruner", "blipt5_sparsegpt_pruner", "blipt5_global_mag_pruner",
                 "blipt5_global_gradmagabs_pruner", "blipt5_global_mezo_pruner"):
        assert registry.registry.get_pruner_class(name) is not None, name
    from ecoflap_b200.pruners import coop, upop

    assert hasattr(coop, "CLIPLayerWandaPruner") and hasattr(coop, "CLIPLayerSparseGPTPruner")
    assert hasattr(upop, "BLIPBertLayerWandaPruner")
    from ecoflap_b200.accumulators import SparseGPT, WrappedGPT

    for cls, methods in ((WrappedGPT, ("add_batch",)), (SparseGPT, ("add_batch", "fasterprune", "free"))):
        for m in methods:
            assert callable(getattr(cls, m))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under ecoflap_b200/ may reference it"""
    pkg = os.path.join(ROOT, "ecoflap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert ("ecoflap_oracle" not in text and "c_oracle" not in text and "oracle/" not in text
                        and "aten_reference" not in text), os.path.join(dirpath, f)
