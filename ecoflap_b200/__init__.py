"""ecoflap_b200 -- B200-native (sm_100a) implementation of ECoFLaP's coarse-to-fine pruning hot path.

Python host code mirrors the reference's pruner interface (``ecoflap_b200.pruners``) and calls
hand-written CUDA kernels through the C ABI in ``include/ecoflap_b200.h`` (``ecoflap_b200._abi``).
"""
__version__ = "0.1.0"
