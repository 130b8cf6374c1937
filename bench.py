#!/usr/bin/env python
"""Benchmark of the ECoFLaP pruning hot path on B200 (contract: see the task description / DESIGN.md).

Workload (BASELINE.json configs[3], named in config.workload): BLIP-2 = EVA ViT-g (39 blocks, fp16) +
FlanT5-XL (24 encoder + 24 decoder blocks, bf16), 588 Linears / 3.70 G prunable parameters, 128 synthetic
calibration samples in 16 batches of 8, 50 % sparsity (Wanda: per-layer threshold on the ViT, per-row on T5).

One STEP = one pass of the Wanda stage-2 hot path over every block: for each Linear 16 calibration-norm
accumulations (A1) followed by the fused score / select / apply (A3-A5, A7).  Activations are synthetic tensors
of the shapes / dtypes the hooks see (the model forward is outside the hot path).

  metric `calib_tokens_per_s`  = sum over hooked Linears of tokens processed / step time (whole job, all ranks)
  value      inputs resident in HBM when the timed region starts (fresh, unpruned weights every step)
  e2e        the same pass through the public API (WrappedGPT.add_batch + prune) from pinned HOST buffers:
             every activation batch and weight is copied H2D, every pruned weight D2H, inside the timed region
  roofline   dominant kernel family, algorithmic bytes / CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the oracle (numpy + plain-C/OpenMP port of the reference's path) on the
             host cores over a bounded sample of the same workload.

N > 1 (torchrun): every rank accumulates norms over its OWN 128-sample shard (weak scaling: per-GPU
calibration work fixed), one NCCL all-reduce per block merges the norm vectors, the per-row select is sharded
over output rows and all-gathered.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BATCHES = 16
BATCH = 8
SPARSITY = 0.5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-s", type=float, default=15.0, help="target seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--prune-wall", default=os.environ.get("ECF_BENCH_PRUNE_WALL", "wanda,sparsegpt,ecoflap,ecoflap_first"),
                    help="comma list of full-size BLIP-2 prune() runs to time (wanda, sparsegpt, ecoflap) or 'none'; "
                         "ecoflap = 4 704 BLIP-2 forwards, ~3 min on one B200")
    ap.add_argument("--no-sparsegpt-kernels", action="store_true")
    ap.add_argument("--no-aten", action="store_true")
    return ap.parse_args()


def workload_config(summ):
    """The `config` object, IDENTICAL in both arms (the driver compares them): what one step is, independent of who runs it."""
    return {"workload": "BLIP-2 (EVA ViT-g + FlanT5-XL) Wanda 50% hot path, 128 samples / 16 batches of 8",
            "arithmetic": "fp32 norms and scores over fp16 / bf16 weights and fp32 / fp16 / bf16 hook inputs",
            "linears": summ["linears"], "params": summ["params"],
            "algorithmic_bytes_per_step": summ["norm_bytes"] + summ["select_bytes"],
            "hook_inputs": "q/k/v, wi_0/wi_1 and cross-attention k/v share one input tensor per block as in the model "
                           f"(distinct norm input bytes per step {summ['unique_norm_input_bytes']})",
            "l2": "GPU arm: inputs larger than L2 -- 47 GB touched per step, no buffer re-read within 126 MB (the per-kernel "
                  "roofline timings flush L2 explicitly); reference arm: host memory",
            "timing": "GPU arm: CUDA events per step, max over ranks, weights restored between steps outside the events; "
                      "reference arm: perf_counter around the C/OpenMP calls"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_pass(blocks, n_batches, budget_s, threads):
    """The reference's path on the host: the plain-C (OpenMP) oracle port over a bounded sample of the workload -- blocks are
    taken round-robin over the three towers until ~budget_s of CPU work is done (at least one block of every tower) -- and
    extrapolated to the whole workload WITH THE WORKLOAD'S TOWER MIX (39 ViT-g : 24 T5-encoder : 24 T5-decoder blocks):
    seconds(workload) = sum over towers of blocks(tower) * mean seconds per sampled block of that tower.
    Returns (tokens of the whole workload, estimated seconds for it, description)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import ecoflap_oracle as orc

    os.environ["OMP_NUM_THREADS"] = str(threads)  # torchrun exports OMP_NUM_THREADS=1 to its workers
    used_threads = c_oracle.set_threads(threads)
    rng = np.random.default_rng(0)
    towers = {}
    for b in blocks:
        towers.setdefault(b.name.rsplit(".", 1)[0], []).append(b)
    order = []
    depth = 0
    while True:
        added = False
        for t in towers.values():
            if depth < len(t):
                order.append(t[depth])
                added = True
        if not added:
            break
        depth += 1
    t_total, used = 0.0, []
    per_tower = {k: [0.0, 0] for k in towers}  # seconds, blocks sampled
    cache = {}
    for blk in order:
        if t_total >= budget_s and all(v[1] > 0 for v in per_tower.values()):
            break
        t_blk = 0.0
        for l in blk.linears:
            key = (l.tokens, l.cols, l.x_dtype)
            if key not in cache:
                cache[key] = c_oracle.to_storage(rng.standard_normal((l.tokens, l.cols)).astype(np.float32), l.x_dtype)
            x = cache[key]
            wkey = (l.rows, l.cols, l.w_dtype)
            if wkey not in cache:
                cache[wkey] = c_oracle.to_storage((rng.standard_normal((l.rows, l.cols)) * 0.02).astype(np.float32), l.w_dtype)
            W = cache[wkey].copy()
            t1 = time.perf_counter()
            s = np.zeros(l.cols, dtype=np.float32)
            n = 0
            for _ in range(n_batches):
                c_oracle.sqnorm_accum(x, l.x_dtype, s, n, BATCH)
                n += BATCH
            if l.select == "row":
                c_oracle.wanda_row_prune(W, l.w_dtype, s, orc.row_k(l.cols, SPARSITY))
            else:
                c_oracle.wanda_layer_prune(W, l.w_dtype, s, orc.layer_kth_index(l.rows * l.cols, SPARSITY))
            t_blk += time.perf_counter() - t1
        t_total += t_blk
        tw = per_tower[blk.name.rsplit(".", 1)[0]]
        tw[0] += t_blk
        tw[1] += 1
        used.append(blk.name)
    est = sum(len(towers[k]) * v[0] / v[1] for k, v in per_tower.items())
    tokens = sum(l.tokens * n_batches for b in blocks for l in b.linears)
    mix = ", ".join(f"{v[1]} of {len(towers[k])} {k.split('.')[-2] if '.' in k else k}" for k, v in per_tower.items())
    desc = (f"{len(used)} of {len(blocks)} blocks timed ({mix}; {t_total:.1f} s of CPU work on {used_threads} OpenMP threads), "
            f"extrapolated per tower to all {len(blocks)} blocks; all Linears, {n_batches} batches of {BATCH}")
    return tokens, est, desc


def run_reference(args):
    """--impl reference: the CPU path only.  Rank 0 works; other ranks exit 0 (contract)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from ecoflap_b200 import workload as wl

    blocks = wl.blip2_blocks(BATCH)
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    per_step = float(os.environ.get("ECF_REF_STEP_S", per_step))  # tests shrink the bounded CPU sample
    vals, desc = [], ""
    for i in range(args.warmup + args.steps):
        tok, sec, desc = cpu_reference_pass(blocks, N_BATCHES, per_step, threads)
        if i >= args.warmup:
            vals.append((tok, sec))
    tok = sum(v[0] for v in vals)
    sec = sum(v[1] for v in vals)
    value = tok / sec
    summ = wl.summarize(blocks, N_BATCHES)
    line = {
        "impl": "reference", "metric": "calib_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(summ),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_cpus(index):
    """Pin this process (and therefore the first-touch placement of its pinned host buffers) to the CPU cores NVML reports
    as local to the GPU: the e2e leg streams ~57 GB per step from host memory, and a rank scheduled on the far socket
    pulls all of it across the inter-socket link.  Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores local to GPU {index} (NVML)"
    except Exception as exc:  # pragma: no cover - NVML or the syscall unavailable: run unbound, say so
        return f"unbound ({type(exc).__name__})"
    return "unbound"


def prune_wall(which=("wanda", "ecoflap"), verbose=False):
    """`pruner.prune()` wall seconds -- the first item of BASELINE.json's metric -- on a full-size random-init BLIP-2
    (EVA ViT-g fp16 + FlanT5-XL bf16, 588 Linears / 3.70 G parameters; synthetic.blip2_full) with 128 synthetic samples
    in batches of 8, through the reference's entry point `load_pruner("blipt5_wanda_pruner", ...).prune()`:
      wanda     uniform 50 % (reference: 240.16 s, LAVIS/training_statistics/cc3m-blipt5_wanda_pruner_0.5-1.0-1.0.yaml)
      ecoflap   zeroth-order stage 1 (MEZO-GradOnly_sum, block granularity, max 0.6, 32 first-stage samples: 588 layers x
                4 batches x 2 forwards) + Wanda (reference: 5 985-6 115 s, ..._olmezo-gradient_sum0.6_block_nd32.yaml)
      sparsegpt blipt5_sparsegpt_pruner 50 %, batch size 1 as the reference asserts (reference: 802.63 s)
      ecoflap_first first-order stage 1 (GradMagAbs_sum, 16 backward passes) + Wanda (reference: 450.31 s)
    Unlike the hot-path step this INCLUDES the model forwards (two per block and batch in stage 2, two per
    layer and batch in stage 1), i.e. mostly cuBLAS / SDPA time.  Under torchrun the calibration batches (stage 2) and the
    layers (stage 1) are sharded over the ranks; the wall time is the max over ranks."""
    import contextlib

    import numpy as np
    import torch
    import torch.distributed as dist

    from ecoflap_b200 import ops
    from ecoflap_b200 import synthetic as syn
    from ecoflap_b200.compression import load_pruner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    if world > 1:  # the communicator (and its peer mappings) is set up by the first collective: not part of prune()
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    # Untimed warm-up, like the W warm-up steps of the hot-path bench: one prune() of a toy BLIP-2 per registry name loads our
    # kernels' modules (CUDA loads a kernel on its first launch), cuSOLVER / cuBLAS handles and the lazily imported Python
    # modules -- with 8 ranks on one host that first touch once cost the first pruner of the list 18 s (N = 8, round 3).
    with open(os.devnull, "w") as sink, contextlib.redirect_stdout(sink):
        for reg, bs in (("blipt5_wanda_pruner", 4), ("blipt5_sparsegpt_pruner", 1)):
            if reg == "blipt5_sparsegpt_pruner" and "sparsegpt" not in which:
                continue
            toy = syn.init_weights_(syn.Blip2Model(
                vit_kw=dict(img_size=32, patch=8, dim=64, depth=2, heads=4, mlp_hidden=128),
                t5_kw=dict(vocab=128, d_model=64, heads=4, d_kv=16, d_ff=128, depth=2), n_query=5, autocast=False), seed=5).to(dev).eval()
            n_toy = 16 * world  # (batches are sharded over the ranks: every rank gets some)
            load_pruner(reg, toy, syn.text_batches(n_toy, bs, 10, 6, 128, seed=6, with_image=32), cfg=dict(
                t5_prune_spec="2-0.5-1.0-1.0", vit_prune_spec="2-0.5-1.0-1.0", t5_pruning_method="x", vit_pruning_method="x",
                num_samples=n_toy)).prune()
            del toy
    torch.cuda.synchronize()
    out = {}
    for name in which:
        torch.manual_seed(0)
        np.random.seed(42)
        model = syn.blip2_full(dev)
        cfg = dict(t5_prune_spec="24-0.5-1.0-1.0", vit_prune_spec="39-0.5-1.0-1.0", t5_pruning_method="x",
                   vit_pruning_method="x", num_samples=128)
        loader = syn.blip2_full_loader(128, 8)
        reg = "blipt5_wanda_pruner"
        if name == "ecoflap":
            cfg.update(sparsity_ratio_granularity="block", max_sparsity_per_layer=0.6, score_method="MEZO-GradOnly_sum",
                       num_data_first_stage=32, num_noise=1, noise_eps=1e-3)
        elif name == "ecoflap_first":  # LAVIS/scripts/blip2/ecoflap_first.py: GradMagAbs_sum, 128 first-stage samples
            cfg.update(sparsity_ratio_granularity="block", max_sparsity_per_layer=0.6, score_method="GradMagAbs_sum",
                       num_data_first_stage=128)
        elif name == "sparsegpt":
            reg = "blipt5_sparsegpt_pruner"
            loader = syn.blip2_full_loader(128, 1)
        pruner = load_pruner(reg, model, loader, cfg=cfg)
        torch.cuda.reset_peak_memory_stats(dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with open(os.devnull, "w") as sink, contextlib.redirect_stdout(sink if not verbose else sys.stderr):
            _, sd = pruner.prune()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        zeros = total = 0
        for k, p in model.named_parameters():
            if p.dim() == 2 and (".blocks." in k or ".block." in k) and "relative_attention_bias" not in k:
                zeros += int(ops.count_zero(p.data).item())
                total += p.numel()
        entry = {"wall_s": round(wall, 3), "sparsity": zeros / total, "linears_params": total,
                 "peak_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)}
        if isinstance(sd, dict):
            vals = list(sd.values())
            entry["ratio_min_max"] = [float(min(vals)), float(max(vals))]
        out[name] = entry
        del pruner, model
        torch.cuda.empty_cache()
    return out


def aten_gpu_baseline(lins, summ_tokens, budget_s=6.0):
    """SURVEY 8(d), last row: the reference's own ATen path ON THE SAME B200 -- what a user of ylsung/ECoFLaP runs today
    (torch.norm hooks, |W|*sqrt(norm) materialised in fp32, torch.sort(stable) / sort(flatten), scatter_, masked store;
    oracle/aten_reference.py restates wanda_pruner.py:54-84,260-279,541-558 call for call).  One block per tower is timed
    with CUDA events (more while the budget lasts) and extrapolated with the workload's tower mix, like the CPU arm."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aten_reference as aten

    towers = {}
    for pb in lins:
        towers.setdefault((len(pb), tuple(pb[0].W.shape)), []).append(pb)
    per_tower = {k: [0.0, 0] for k in towers}
    spent, depth = 0.0, 0
    while depth < max(len(t) for t in towers.values()):
        for key, t in towers.items():
            if depth >= len(t) or (spent >= budget_s and per_tower[key][1] > 0):
                continue
            pb = t[depth]
            Ws = [o.W0.clone() for o in pb]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for o, W in zip(pb, Ws):
                acc = aten.AtenWrappedGPT(columns=W.shape[1], device=W.device)
                for x in o.acts:
                    acc.add_batch(x.view(BATCH, -1, x.shape[-1]))
                if o.spec.select == "row":
                    aten.prune_rows_(W, acc.scaler_row, SPARSITY)
                else:
                    aten.prune_layer_(W, acc.scaler_row, SPARSITY)
            e1.record()
            torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) * 1e-3
            if depth > 0 or per_tower[key][1] == 0:
                per_tower[key][0] += sec
                per_tower[key][1] += 1
            spent += sec
            del Ws
        depth += 1
        if spent >= budget_s and all(v[1] > 0 for v in per_tower.values()):
            break
    torch.cuda.empty_cache()
    est = sum(len(towers[k]) * v[0] / v[1] for k, v in per_tower.items())
    return {"value": summ_tokens / est, "unit": "tokens/s", "ms_per_step": 1e3 * est, "kind": "restatement of the reference's ATen calls on cuda:0",
            "sample": f"{sum(v[1] for v in per_tower.values())} of {len(lins)} blocks timed with CUDA events, extrapolated per tower"}


def sparsegpt_roofline(dev):
    """Tensor-pipe side of the path (north_star: Hessian and OBS kernels against the bf16 tensor peak), on the SparseGPT
    shapes of BASELINE.json configs 2 (CLIP ViT-B/16: T = 128 x 197, C = 768 / 3072) and 4 (EVA ViT-g: T = 128 x 257,
    C = 1408 / 6144; FlanT5-XL wo: C = 5120).  Hessian: one `ecf_hessian_accum` launch over the concatenated calibration
    batches (what HessianBatch issues), algorithmic flops 2*T*C^2 over CUDA-event time.  OBS: `ecf_obs_prune`
    with its phases timed separately through ECF_OBS_PHASES (tile threshold | mask + 128-step sweep | tcgen05 trailing
    update); trailing flops R*C^2 algorithmic (the bf16 hi/mid split executes 3x that).  Prologue (A9, cuSOLVER) beside it."""
    import torch

    from ecoflap_b200 import ops
    from ecoflap_b200.accumulators import SparseGPT

    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak_sus = float(pk.get("bf16_tflops_sustained", 1388.0))
    peak_burst = float(pk.get("bf16_tflops_burst", pk.get("bf16_tflops", 1640.0)))
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps=3, prepare=None):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            if prepare is not None:
                prepare()
            scratch.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b)
            best = t if best is None else min(best, t)
        return best

    out = {"peak_tflops_sustained": peak_sus, "peak_tflops_burst": peak_burst, "hessian": [], "obs": []}
    torch.manual_seed(7)
    for name, T, C, dt in (("cfg4 ViT-g fc2, 128 x 257 tokens", 128 * 257, 6144, torch.float16),
                           ("cfg4 ViT-g qkv/fc1 (fp32 LayerNorm output)", 128 * 257, 1408, torch.float32),
                           ("cfg4 T5-XL wo, 128 x 56 tokens", 128 * 56, 5120, torch.bfloat16),
                           ("cfg2 CLIP ViT-B/16 c_proj, 128 x 197 tokens", 128 * 197, 3072, torch.float16),
                           ("cfg2 CLIP ViT-B/16 in_proj/c_fc", 128 * 197, 768, torch.float16)):
        X = torch.randn(T, C, device=dev).to(dt)
        H = torch.zeros(C, C, device=dev)
        ms = timed(lambda: ops.hessian_accum(X, H, 2.0 / 128, 0.0))
        tf = 2.0 * T * C * C / (ms * 1e-3) / 1e12
        # what the tensor pipe executes: only the 128 x 256 tiles that touch the upper triangle (the mirror kernel fills the
        # rest), times three bf16 products per element for fp32 inputs (hi*hi + hi*mid + mid*hi)
        mt, nt = -(-C // 128), -(-C // 256)
        tiles = sum(1 for i in range(mt) for j in range(nt) if (j + 1) * 256 > i * 128)
        ex = tf * tiles / (mt * nt) * (3.0 if dt == torch.float32 else 1.0)
        out["hessian"].append({"shape": name, "T": T, "C": C, "dtype": str(dt).split(".")[-1], "ms": ms, "tflops_algorithmic": tf,
                               "tflops_executed": ex, "tensor_pipe_frac_of_sustained": ex / peak_sus,
                               "tensor_pipe_frac_of_burst": ex / peak_burst,
                               "note": "algorithmic = 2*T*C^2 (SURVEY 8d counts the full product); executed = upper-triangle tiles only"})
        del X, H
    for name, R, C in (("cfg4 ViT-g fc2 1408x6144", 1408, 6144), ("cfg4 ViT-g fc1 6144x1408", 6144, 1408),
                       ("cfg4 T5-XL wo 2048x5120", 2048, 5120), ("cfg2 CLIP c_proj 768x3072", 768, 3072),
                       ("cfg2 CLIP c_fc 3072x768", 3072, 768)):
        Xh = torch.randn(4 * C, C, device=dev)
        Xh[:, 1] *= 8.0
        H0 = (2.0 / (4 * C)) * (Xh.t() @ Xh)
        del Xh
        lin = torch.nn.Linear(C, R, bias=False).to(dev)
        acc = SparseGPT(lin)
        acc.H = H0.clone()
        t0 = time.perf_counter()
        torch.cuda.synchronize()
        Hinv, _ = acc.prepare_hinv(0.01)
        torch.cuda.synchronize()
        prologue_ms = 1e3 * (time.perf_counter() - t0)
        W0 = torch.randn(R, C, device=dev) * 0.02
        W = W0.clone()
        kth = [int(R * (min(i1 + 128, C) - i1) * 0.5) for i1 in range(0, C, 128)]
        ph = {}
        for tag, mask in (("all", 7), ("threshold", 1), ("threshold+sweep", 3)):
            os.environ["ECF_OBS_PHASES"] = str(mask)
            ph[tag] = timed(lambda: ops.obs_prune(W, Hinv, kth), prepare=lambda: W.copy_(W0))
        os.environ.pop("ECF_OBS_PHASES", None)
        trailing = max(ph["all"] - ph["threshold+sweep"], 1e-6)
        tf = float(R) * C * C / (trailing * 1e-3) / 1e12
        out["obs"].append({"shape": name, "R": R, "C": C, "total_ms": ph["all"], "threshold_ms": ph["threshold"],
                           "sweep_ms": ph["threshold+sweep"] - ph["threshold"], "trailing_ms": trailing,
                           "trailing_tflops_algorithmic": tf, "trailing_frac_of_sustained": tf / peak_sus,
                           "trailing_tflops_executed_x3_split": 3 * tf, "prologue_cusolver_ms": prologue_ms})
        del H0, Hinv, W, W0, acc, lin
    del scratch
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from ecoflap_b200 import _abi, ops
    from ecoflap_b200 import dist as edist
    from ecoflap_b200 import workload as wl
    from ecoflap_b200.accumulators import NormBatch, WrappedGPT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ecoflap_b200 has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    TD = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}

    blocks = wl.blip2_blocks(BATCH)
    summ = wl.summarize(blocks, N_BATCHES)
    torch.manual_seed(1234 + rank)

    # ---- resident data: pristine + working weights, activation pools -------------------------------------
    class Lin:
        pass

    lins = []
    act_pool = {}
    for b in blocks:
        per_block = []
        for li, l in enumerate(b.linears):
            o = Lin()
            o.spec = l
            o.W0 = (torch.randn(l.rows, l.cols, device=dev) * 0.02).to(TD[l.w_dtype])
            o.W = o.W0.clone()
            key = (l.src or li, l.tokens, l.cols, l.x_dtype)  # Linears fed by the same tensor in the model (q/k/v, wi_0/wi_1) share it here too
            if key not in act_pool:
                xs = []
                for _ in range(N_BATCHES):
                    x = torch.randn(l.tokens, l.cols, device=dev)
                    x[:, 3] *= 30.0  # outlier channel
                    x[:, 7] = 0.0    # dead channel
                    xs.append(x.to(TD[l.x_dtype]))
                act_pool[key] = xs
            o.acts = act_pool[key]
            o.layer = torch.nn.Module()
            o.layer.weight = torch.nn.Parameter(o.W, requires_grad=False)
            o.k = int(l.cols * SPARSITY)
            o.idx = int(l.rows * l.cols * SPARSITY)
            per_block.append(o)
        lins.append(per_block)
    flat = [o for pb in lins for o in pb]

    def restore():
        for o in flat:
            o.W.copy_(o.W0)

    ROW_SHARD = os.environ.get("ECF_ROW_SHARD", "0") == "1"  # row-sharded select + all-gather instead of replication
    # exchange step of the norms: our peer-memory kernel over NVSwitch (22 us per block at 8 GPUs) unless it is
    # unavailable or switched off (ECF_P2P_EXCHANGE=0), then one NCCL all-reduce per block (36 us at 8 GPUs)
    pex = None
    if world > 1 and os.environ.get("ECF_P2P_EXCHANGE", "1") != "0" and edist.PeerNormExchange.available():
        try:
            pex = edist.PeerNormExchange(max(sum(o.spec.cols for o in pb) for pb in lins), dev)
        except Exception as exc:  # pragma: no cover - rendezvous unsupported on this box: say so, use NCCL
            print(f"[bench] peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            pex = None

    def step_device(batches=range(N_BATCHES)):
        """one pass, everything resident in HBM (``batches``: the calibration batches this rank accumulates -- all 16 of its
        own shard under weak scaling, every world-th of the one fixed set under strong scaling).  Per block: the hook calls of its 16 calibration batches are
        deferred into batched norm launches (<= 256 hook calls / 32 accumulators each), then the fused selects of the
        block's Linears: one launch per select family (per-layer: all Linears; per-row: one per row length and dtype)."""
        main = torch.cuda.current_stream()
        for pb in lins:
            nb = NormBatch()
            accs = [WrappedGPT(o.layer, batch=nb) for o in pb]
            flat = edist.pack_block_norms(accs) if world > 1 else None  # one buffer per block: one in-place all-reduce
            for j in batches:
                for o, acc in zip(pb, accs):
                    acc.add_batch(o.acts[j])
            nb.flush()
            layer_items = [(o.W, acc.scaler_row, o.idx) for o, acc in zip(pb, accs) if o.spec.select == "layer"]
            if world > 1:
                # equal shards: the global mean is the average of the rank means -- one NCCL all-reduce (AVG) per block on
                # the packed norm buffer, nothing read back, so the whole step stays capturable in one CUDA graph
                if pex is not None:
                    pex.sync(flat, accs)
                else:
                    edist.sync_packed_norms(flat, accs)
                if ROW_SHARD:
                    for o, acc in zip(pb, accs):
                        if o.spec.select == "row":
                            edist.row_sharded_select(o.W, lambda sh, a=acc, k=o.k: ops.wanda_row_select_apply(sh, a.scaler_row, k))
                    if layer_items:
                        ops.wanda_layer_thresh_apply_batched(layer_items)
                    continue
            if layer_items:  # the per-layer selects of a block: one cooperative launch
                ops.wanda_layer_thresh_apply_batched(layer_items)
            rows = [(o.W, acc.scaler_row, o.k) for o, acc in zip(pb, accs) if o.spec.select == "row"]
            if rows:  # the per-row selects of a block: one persistent launch per distinct (row length, dtype)
                ops.wanda_row_select_apply_batched(rows)

    launches_per_step = 0
    for pb in lins:
        launches_per_step += -(-(N_BATCHES * len(pb)) // _abi.SQNORM_MAX_BATCH)  # batched norm launches per block
        launches_per_step += len({(o.W.shape[1], o.W.dtype) for o in pb if o.spec.select == "row"})
        layer_sel = [o for o in pb if o.spec.select == "layer"]
        if layer_sel:
            # per-layer select of the block: split path (sample, count, refine, apply) + the cooperative kernel behind it as
            # the fallback when every matrix is 16-bit; the cooperative kernel alone otherwise
            split = os.environ.get("ECF_LT_SPLIT", "1") != "0" and all(o.W.dtype != torch.float32 for o in layer_sel)
            launches_per_step += 5 if split else 1

    if pex is not None:
        launches_per_step += len(lins)  # one peer-memory exchange kernel per block

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident steps ---------------------------------------------------------------------
    for _ in range(2):
        restore()
        step_device()
    barrier()
    launch_mode = "cuda-graph"
    graph = None
    try:
        restore()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_device()
        run_step = graph.replay
    except Exception as exc:  # pragma: no cover - keep the benchmark alive, say what happened
        launch_mode = f"eager ({type(exc).__name__}: graph capture unavailable)"
        graph = None
        run_step = step_device
        torch.cuda.synchronize()
    for _ in range(max(3, args.warmup)):
        restore()
        run_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times = []
    t_bracket0 = time.perf_counter()
    for _ in range(args.steps):
        restore()  # fresh, unpruned weights for every step (not part of the hot path: outside the event pair)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    barrier()
    bracket_ms = 1e3 * (time.perf_counter() - t_bracket0)
    clocks = sampler.stop() if rank == 0 else None
    step_ms = sum(times) / len(times)
    if world > 1:
        t = torch.tensor([step_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    tokens_per_step = summ["calib_tokens_per_step"] * world
    value = tokens_per_step / (step_ms * 1e-3)

    # ---- strong scaling (N > 1): BASELINE.json configs[3] as written -- the ONE 128-sample calibration set sharded over the
    # ranks (rank r accumulates batches r, r + N, ...), exchange, every rank selects.  Reported next to the weak-scaling
    # headline; the replicated select is what bounds it (Amdahl).
    strong = None
    if world > 1:
        mine = range(rank, N_BATCHES, world)
        restore()
        step_device(mine)
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            step_device(mine)
        st = []
        for i in range(3 + args.steps):
            restore()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g2.replay()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                st.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(st) / len(st)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong_ms = float(t.item())
        strong = {"ms_per_step": strong_ms, "value": summ["calib_tokens_per_step"] / (strong_ms * 1e-3), "unit": "tokens/s",
                  "samples": N_BATCHES * BATCH, "batches_per_rank": len(mine),
                  "limiter": "replicated per-row / per-layer select (every rank prunes every Linear) + one exchange per block"}
        del g2

    # ---- roofline of the kernel families ------------------------------------------------------------------------
    # The sampled launches of a family (every sixth block of every tower) are captured back to back in ONE CUDA graph and
    # replayed between two events on the launching stream: device time only, and the ~8 us between an event record and the
    # first kernel of a graph launch is paid once per family, not once per launch (round 1 timed every launch as its own
    # graph, which added that constant to each of them).  Before each replay the weights are restored and the 126 MB L2 is
    # evicted by a 512 MB write followed by a 512 MB read -- the read leaves CLEAN lines, as the norm kernel that precedes a
    # select in the real step does (dirty lines would make the first launch pay for their write-back).  Every launch of
    # the graph works on its own block's weights and activations (> 126 MB apart), so all of them start cold.
    peak, peak_src = peaks()
    fam = {"sqnorm": [0.0, 0, 0], "row_select": [0.0, 0, 0], "layer_thresh": [0.0, 0, 0]}  # ms, bytes, launches
    scratch = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def timed_graph(fns, prepare=None, reps=3):
        if prepare is not None:  # (the last timed step left the weights pruned: selecting them again would be the heavy-tie path)
            prepare()
        for fn in fns:  # warm-up (workspaces, function attributes) outside the capture
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for fn in fns:
                fn()
        best = None
        for _ in range(reps):
            if prepare is not None:
                prepare()
            scratch.zero_()
            scratch.view(torch.int64).sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b)
            best = t if best is None else min(best, t)
        return best

    sq_fns, lt_fns, rs_fns, lt_lins, rs_lins = [], [], [], [], []
    sq_distinct = 0
    sampled = lins[::6]  # every sixth block: 7 ViT-g, 4 T5 encoder, 4 T5 decoder blocks, ~100 Linears
    # blocks of one tower share their (synthetic) activation tensors: interleave the towers so that two launches reading the
    # same tensors are > 1 GB of other traffic apart
    towers = {}
    for pb in sampled:
        towers.setdefault((len(pb), pb[0].W.shape), []).append(pb)
    order = [t[i] for i in range(max(len(t) for t in towers.values())) for t in towers.values() if i < len(t)]
    for pb in order:
        accs = [WrappedGPT(o.layer) for o in pb]
        items = []
        for j in range(N_BATCHES):
            n0 = j * BATCH
            for o, acc in zip(pb, accs):
                items.append((o.acts[j], acc.scaler_row, n0 / (n0 + BATCH), 1.0 / (n0 + BATCH)))
        sq_fns.append(lambda items=items: ops.sqnorm_accum_batched(items))
        fam["sqnorm"][1] += sum(wl.norm_bytes(o.spec, N_BATCHES) for o in pb)
        seen_x = {}
        for o in pb:  # distinct input bytes: hooks that see the very same tensor (q/k/v ...) are read once by the kernel
            for x in o.acts:
                seen_x[x.data_ptr()] = x.numel() * x.element_size()
        sq_distinct += sum(seen_x.values()) + sum(8 * o.spec.cols for o in pb)
        fam["sqnorm"][2] += -(-len(items) // _abi.SQNORM_MAX_BATCH)
        layer = [(o, acc) for o, acc in zip(pb, accs) if o.spec.select == "layer"]
        if layer:
            items_l = [(o.W, acc.scaler_row, o.idx) for o, acc in layer]
            lt_fns.append(lambda items_l=items_l: ops.wanda_layer_thresh_apply_batched(items_l))
            lt_lins += [o for o, _ in layer]
            fam["layer_thresh"][1] += sum(wl.select_bytes(o.spec) for o, _ in layer)
            fam["layer_thresh"][2] += 1
        rows = [(o, acc) for o, acc in zip(pb, accs) if o.spec.select == "row"]
        if rows:
            items_r = [(o.W, acc.scaler_row, o.k) for o, acc in rows]
            rs_fns.append(lambda items_r=items_r: ops.wanda_row_select_apply_batched(items_r))
            rs_lins += [o for o, _ in rows]
            fam["row_select"][1] += sum(wl.select_bytes(o.spec) for o, _ in rows)
            fam["row_select"][2] += len({(o.W.shape[1], o.W.dtype) for o, _ in rows})
    fam["sqnorm"][0] = timed_graph(sq_fns)  # (the norms of the first graph feed the selects below)
    if lt_fns:
        fam["layer_thresh"][0] = timed_graph(lt_fns, prepare=lambda: [o.W.copy_(o.W0) for o in lt_lins])
    if rs_fns:
        fam["row_select"][0] = timed_graph(rs_fns, prepare=lambda: [o.W.copy_(o.W0) for o in rs_lins])
    del scratch
    kernels = {}
    for name, (ms, nbytes, n) in fam.items():
        if n:
            kernels[name] = {"launches": n, "avg_us": 1e3 * ms / n, "achieved_gbs": nbytes / ms / 1e6,
                             "frac": nbytes / ms / 1e6 / peak, "time_share": ms}
    if "sqnorm" in kernels:
        # frac counts every hook's input (T*C*sizeof(x) + 8*C per hook call); shared inputs are read once, so the
        # bytes the kernel actually has to move are fewer: both fractions are reported
        kernels["sqnorm"]["achieved_gbs_distinct_bytes"] = sq_distinct / fam["sqnorm"][0] / 1e6
        kernels["sqnorm"]["frac_distinct_bytes"] = sq_distinct / fam["sqnorm"][0] / 1e6 / peak
    tot_ms = sum(v["time_share"] for v in kernels.values())
    for v in kernels.values():
        v["time_share"] = v["time_share"] / tot_ms
    dominant = max(kernels, key=lambda k: kernels[k]["time_share"])
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (per launch; the capture's launch
    # is named next to it -- traffic ~= algorithmic bytes means no wasted re-reads)
    traffic, traffic_of = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1", "traffic.json")) as fh:
            tr = json.load(fh).get(dominant)
        if tr:
            traffic = tr["dram_bytes_per_launch"]
            traffic_of = {"launch": tr["launch"], "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_launch"], "source": tr["source"]}
    except (OSError, ValueError, KeyError):
        pass
    # headline of the dominant kernel: for the norms the bytes the kernel has to move (shared hook inputs once); the per-hook
    # figure of SURVEY 8(d) is next to it and can exceed 1 because it counts a shared input once per hook
    head = kernels[dominant]
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": head.get("achieved_gbs_distinct_bytes", head["achieved_gbs"]),
                "peak": peak, "unit": "GB/s", "frac": head.get("frac_distinct_bytes", head["frac"]),
                "frac_per_hook_bytes": head["frac"], "traffic": traffic, "traffic_of": traffic_of,
                "peak_source": peak_src + "; the peak is a COPY (read + write) bandwidth, a read-only stream can exceed it",
                "method": "one CUDA graph per kernel family over the sampled launches, CUDA events around the replay, clean L2 flush",
                "kernels": kernels}

    # ---- e2e: host buffers through the public API ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_act, host_w = {}, {}
        for o in flat:
            l = o.spec
            ka = (l.tokens, l.cols, l.x_dtype)
            if ka not in host_act:
                host_act[ka] = o.acts[0].cpu().pin_memory()
            kw = (l.rows, l.cols, l.w_dtype)
            if kw not in host_w:
                host_w[kw] = o.W0.cpu().pin_memory()
        d2h = sum(l.rows * l.cols * wl.BYTES[l.w_dtype] for l in (o.spec for o in flat))
        dev_act = {k: torch.empty_like(v, device=dev) for k, v in host_act.items()}
        out_w = {k: torch.empty_like(v).pin_memory() for k, v in host_w.items()}

        # three streams: H2D staging, compute, D2H.  Activation batches go through a ring of device staging buffers
        # (events: slot filled -> kernel may read, kernel done -> slot may be refilled); a hook input shared by several
        # Linears of a block (q/k/v, wi_0/wi_1, cross-attention k/v) is the same host tensor in the reference and is
        # staged once per batch.
        s_comp = torch.cuda.current_stream()
        s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        RING = 6
        max_act = max(v.numel() * v.element_size() for v in host_act.values())
        ring = [torch.empty(max_act, dtype=torch.uint8, device=dev) for _ in range(RING)]
        ev_full = [torch.cuda.Event() for _ in range(RING)]
        ev_free = [torch.cuda.Event() for _ in range(RING)]
        h2d = 0
        for pb in lins:
            seen = set()
            for o in pb:
                l = o.spec
                h2d += l.rows * l.cols * wl.BYTES[l.w_dtype]
                if (l.src or l.name) not in seen:
                    seen.add(l.src or l.name)
                    h2d += l.tokens * l.cols * wl.BYTES[l.x_dtype] * N_BATCHES

        def step_e2e(batches=range(N_BATCHES)):
            slot_i = 0
            for e in ev_free:
                e.record(s_comp)
            for pb in lins:
                groups = {}
                for o in pb:
                    groups.setdefault(o.spec.src or o.spec.name, []).append(o)
                for members in groups.values():
                    l0 = members[0].spec
                    ka = (l0.tokens, l0.cols, l0.x_dtype)
                    accs = []
                    w_ready = torch.cuda.Event()
                    with torch.cuda.stream(s_h2d):
                        for o in members:
                            l = o.spec
                            o.W.copy_(host_w[(l.rows, l.cols, l.w_dtype)], non_blocking=True)  # H2D weight
                        w_ready.record(s_h2d)
                    for o in members:
                        accs.append(WrappedGPT(o.layer))
                    src = host_act[ka]
                    for j in batches:
                        k = slot_i % RING
                        slot_i += 1
                        stage = ring[k][:src.numel() * src.element_size()].view(src.dtype).view(src.shape)
                        with torch.cuda.stream(s_h2d):
                            s_h2d.wait_event(ev_free[k])
                            stage.copy_(src, non_blocking=True)  # H2D activation batch (once per distinct hook input)
                            ev_full[k].record(s_h2d)
                        s_comp.wait_event(ev_full[k])
                        for acc in accs:
                            acc.add_batch(stage)  # per-hook launch on the compute stream
                        ev_free[k].record(s_comp)
                    if world > 1:  # the exchange step: global running means over the batches of all ranks
                        edist.sync_block_norms(accs)
                    s_comp.wait_event(w_ready)
                    done = torch.cuda.Event()
                    for o, acc in zip(members, accs):
                        if o.spec.select == "row":
                            ops.wanda_row_select_apply(o.W, acc.scaler_row, o.k)
                        else:
                            ops.wanda_layer_thresh_apply(o.W, acc.scaler_row, o.idx)
                    done.record(s_comp)
                    with torch.cuda.stream(s_d2h):
                        s_d2h.wait_event(done)
                        for o in members:
                            l = o.spec
                            out_w[(l.rows, l.cols, l.w_dtype)].copy_(o.W, non_blocking=True)  # D2H pruned weight
            torch.cuda.synchronize()

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": tokens_per_step / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": n_e2e}
        if world > 1:  # strong scaling through the public API: this rank copies in only its share of the batches
            mine = range(rank, N_BATCHES, world)
            step_e2e(mine)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                step_e2e(mine)
            barrier()
            t = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            strong["e2e_value"] = summ["calib_tokens_per_step"] / float(t.item())
            strong["e2e_ms_per_step"] = 1e3 * float(t.item())

    # ---- the reference's ATen path on this GPU, the tensor-core kernels, prune() wall seconds ------------------------
    aten = None
    if rank == 0 and world == 1 and not args.no_aten:
        aten = aten_gpu_baseline(lins, summ["calib_tokens_per_step"])
    # free the resident hot-path data before the full-size model runs
    graph = None
    run_step = None
    for o in flat:
        o.W = o.W0 = o.acts = o.layer = None
    lins, flat, act_pool = [], [], {}
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    sgpt = None
    if rank == 0 and not args.no_sparsegpt_kernels:
        sgpt = sparsegpt_roofline(dev)
        roofline["tensor"] = sgpt
    walls = None
    which = [w for w in args.prune_wall.split(",") if w and w != "none"]
    if which:
        walls = prune_wall(which)
        ref_s = {"wanda": 240.16, "ecoflap": 5985.24, "sparsegpt": 802.63, "ecoflap_first": 450.31}
        for k, v in walls.items():
            v["reference_wall_s"] = ref_s.get(k)
            v["reference_hardware"] = "one GPU, model unstated (LAVIS/training_statistics/*.yaml; BASELINE.md section 1)"
            v["n_gpus"] = world

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)  # the CPU arm gets every host core back
        threads = os.cpu_count() or 1
        tok, sec, desc = cpu_reference_pass(blocks, N_BATCHES, args.cpu_sample_s, threads)
        cpu = {"value": tok / sec, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": desc, "seconds": sec}

    if rank == 0:
        line = {
            "metric": "calib_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(summ),
            "run": {"bracket_ms": bracket_ms, "launch": launch_mode,
                    "parallelism": f"dp{world}: batch-sharded norms + " + ("peer-memory exchange kernel (NVSwitch P2P)" if pex is not None else "NCCL all-reduce") + " per block; select " + ("row-sharded + all-gather" if ROW_SHARD and world > 1 else "replicated per rank")},
            "prune_wall_s_hot_path": step_ms * 1e-3,
            "prune_wall_s": walls, "aten_gpu_baseline": aten, "strong_scaling": strong,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "abi_version": _abi.lib.ecf_version(), "cpu_affinity": affinity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured step holds NCCL work; tearing the communicator down under live CUDA graphs can block for
        # minutes.  Everything is measured and printed: synchronise, meet the other ranks, and leave without running
        # the destructors.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
