"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's hot path restated with the very torch (ATen) calls it makes,
device-agnostic.  On a CUDA tensor this is "what a user of ylsung/ECoFLaP runs today" (last row): the
honest GPU competitor of the hand-written kernels.  Imported by tests/ (as a checker for entry points whose reference
file is absent, e.g. LLaMA/lib) and by bench.py's ``aten_gpu_baseline`` leg only; the product never imports it.

Each function cites the reference lines it follows (LAVIS/lavis/compression/pruners/...).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class AtenWrappedGPT:
    """wanda_pruner.py:54-84 -- per-input-channel running mean of the squared activation norm."""

    def __init__(self, layer=None, columns=None, device=None):
        if layer is not None:
            columns, device = layer.weight.data.shape[1], layer.weight.device
        self.scaler_row = torch.zeros((columns), device=device)
        self.nsamples = 0

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:                                     # :72-73
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]                                          # :74
        if len(inp.shape) == 3:                                     # :76-77
            inp = inp.reshape((-1, inp.shape[-1]))
        inp = inp.t()                                               # :78
        self.scaler_row *= self.nsamples / (self.nsamples + tmp)    # :80
        self.nsamples += tmp                                        # :81
        inp = inp.type(torch.float32)                               # :83
        self.scaler_row += torch.norm(inp, p=2, dim=1) ** 2 / self.nsamples  # :84


def wanda_metric(W, scaler_row):
    """wanda_pruner.py:260 / :541."""
    return torch.abs(W) * torch.sqrt(scaler_row.reshape((1, -1)))


def prune_rows_(W, scaler_row, sparsity):
    """Per-row select + apply, wanda_pruner.py:260-279 (prune_n == 0): stable sort, first int(C*s) indices, scatter_."""
    W_metric = wanda_metric(W, scaler_row)
    W_mask = (torch.zeros_like(W_metric) == 1)
    sort_res = torch.sort(W_metric, dim=-1, stable=True)
    indices = sort_res[1][:, :int(W_metric.shape[1] * sparsity)]
    W_mask.scatter_(1, indices, True)
    W[W_mask] = 0
    return W_mask


def prune_layer_(W, scaler_row, sparsity):
    """Per-layer threshold select + apply, wanda_pruner.py:541-558."""
    W_metric = wanda_metric(W, scaler_row)
    thres = torch.sort(W_metric.flatten())[0][int(W_metric.numel() * sparsity)]
    W_mask = (W_metric <= thres)
    W[W_mask] = 0
    return W_mask


def prune_nm_(W, scaler_row, n, m):
    """n:m branch, wanda_pruner.py:265-270."""
    W_metric = wanda_metric(W, scaler_row)
    W_mask = (torch.zeros_like(W_metric) == 1)
    for ii in range(W_metric.shape[1]):
        if ii % m == 0:
            tmp = W_metric[:, ii:(ii + m)].float()
            W_mask.scatter_(1, ii + torch.topk(tmp, n, dim=1, largest=False)[1], True)
    W[W_mask] = 0
    return W_mask


def find_layers(module, layers=(nn.Linear,), name=""):
    """wanda_pruner.py:33-52."""
    if type(module) in layers:
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_layers(child, layers=layers, name=name + "." + name1 if name != "" else name1))
    return res


def sweep_blocks_(layers, inps, caches, sparsity_of, select="row", output_index=None, prune_n=0, prune_m=0):
    """The block loop of `_prune` (wanda_pruner.py:217-290 / :499-568) on already captured block-0 inputs:
    hooks on -> forward all batches -> metric / select / apply per Linear -> forward again -> swap.
    ``sparsity_of(i, name)`` gives the ratio of Linear ``name`` in block ``i``.  In place; returns the final ``inps``."""
    outs = [None] * len(inps)
    for i in range(len(layers)):
        layer = layers[i]
        subset = find_layers(layer)
        wrapped = {name: AtenWrappedGPT(subset[name]) for name in subset}

        def add_batch(name):
            def tmp(_, inp, out):
                wrapped[name].add_batch(inp[0].data, out.data)
            return tmp

        handles = [subset[name].register_forward_hook(add_batch(name)) for name in wrapped]

        def run(j):
            out = layer(inps[j], **caches[j])
            return out[output_index] if (output_index is not None and isinstance(out, (tuple, list))) else out

        for j in range(len(inps)):
            with torch.no_grad():
                outs[j] = run(j)
        for h in handles:
            h.remove()
        for name in subset:
            W = subset[name].weight.data
            if prune_n != 0:
                prune_nm_(W, wrapped[name].scaler_row, prune_n, prune_m)
            elif select == "row":
                prune_rows_(W, wrapped[name].scaler_row, sparsity_of(i, name))
            else:
                prune_layer_(W, wrapped[name].scaler_row, sparsity_of(i, name))
        for j in range(len(inps)):
            with torch.no_grad():
                outs[j] = run(j)
        inps, outs = outs, inps
    return inps
