"""CUDA-graph capture without the allocator flush.

``torch.cuda.graph.__enter__`` runs ``gc.collect()`` and ``torch.cuda.empty_cache()`` before every capture.  The sweep and the
zeroth-order loop capture one graph per block / per prefix variant (87-90 per run): each flush hands every cached block back
with cudaFree and the next allocations pay cudaMalloc again -- measured on the full-size BLIP-2 SparseGPT run: 2.0-5.5 s in
``_cuda_emptyCache`` alone.  ``capture`` uses the raw begin / end calls on a side stream instead."""
from __future__ import annotations

import os

import torch


def capture(graph: "torch.cuda.CUDAGraph", fn, pool=None, device=None):
    """Runs ``fn()`` under stream capture into ``graph`` and returns its result."""
    if os.environ.get("ECF_GRAPH_CAPTURE_FLUSH", "0") == "1":  # A/B switch: torch's own context manager (with the flush)
        with torch.cuda.graph(graph, **({"pool": pool} if pool is not None else {})):
            return fn()
    cur = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        graph.capture_begin(**({"pool": pool} if pool is not None else {}))
        try:
            out = fn()
        except BaseException:
            try:
                graph.capture_end()
            except Exception:
                pass
            raise
        graph.capture_end()
    cur.wait_stream(side)
    return out
