"""Pruner entry points with the reference's names."""
from .base import BasePruner, LayerWiseBasePruner  # noqa: F401
from . import lavis, coop, upop, llama  # noqa: F401
from .lavis import (  # noqa: F401
    BLIPT5LayerSparseGPTPruner, BLIPT5LayerWandaPruner, T5LayerSparseGPTPruner, T5LayerWandaPruner,
    VITLayerSparseGPTPruner, VITLayerWandaPruner,
)
from .coop import (  # noqa: F401
    CLIPLayerSparseGPTPruner, CLIPLayerWandaPruner, TransformerLayerSparseGPTPruner, TransformerLayerWandaPruner,
)
from .upop import BertLayerWandaPruner, BLIPBertLayerWandaPruner  # noqa: F401
from . import global_pruner  # noqa: F401
from .global_pruner import (  # noqa: F401
    BLIPT5GlobalGradMagAbsPruner, BLIPT5GlobalMagPruner, BLIPT5GlobalMeZoPruner, BLIPT5GlobalPruner,
)
