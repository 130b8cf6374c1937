"""Shape tables of the BASELINE.json configurations (SURVEY.md section 8, table at the top): which Linears a
block holds, their weight / hook-input dtypes, the per-batch token counts and the selection variant.  Shared by
``bench.py``, ``__graft_entry__.smoke()`` and the tests so that every consumer measures the same workload."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List


@dataclass(frozen=True)
class LinearSpec:
    name: str
    rows: int          # out features
    cols: int          # in features (the hooked activation width)
    w_dtype: str       # "fp16" | "bf16" | "fp32"
    x_dtype: str       # dtype of the hook input
    tokens: int        # tokens per calibration batch fed to this Linear (batch * seq)
    select: str        # "row" | "layer"
    src: str = ""      # name of the hook-input tensor inside the block: Linears with the same src see the SAME tensor


@dataclass(frozen=True)
class BlockSpec:
    name: str
    linears: tuple


def vit_g_block(i, batch=8, tokens=257, prefix="visual_encoder.blocks"):
    T = batch * tokens
    # LayerNorm outputs reach qkv / fc1 as fp32 under autocast, attention/GELU outputs as fp16
    return BlockSpec(f"{prefix}.{i}", (
        LinearSpec("attn.qkv", 4224, 1408, "fp16", "fp32", T, "layer", "norm1"),
        LinearSpec("attn.proj", 1408, 1408, "fp16", "fp16", T, "layer", "attn_ctx"),
        LinearSpec("mlp.fc1", 6144, 1408, "fp16", "fp32", T, "layer", "norm2"),
        LinearSpec("mlp.fc2", 1408, 6144, "fp16", "fp16", T, "layer", "gelu"),
    ))


def t5_xl_block(i, decoder, batch=8, enc_tokens=64, dec_tokens=32, prefix="t5_model"):
    d, ff = 2048, 5120
    Tq = batch * (dec_tokens if decoder else enc_tokens)
    Tkv = batch * enc_tokens
    # q/k/v are fed the same normed hidden states, wi_0/wi_1 the same FF input, cross-attention k/v the encoder output
    lin = [LinearSpec(f"layer.0.SelfAttention.{n}", d, d, "bf16", "bf16", Tq, "row", "sa_ctx" if n == "o" else "sa_in")
           for n in "qkvo"]
    j = 1
    if decoder:
        lin += [LinearSpec("layer.1.EncDecAttention.q", d, d, "bf16", "bf16", Tq, "row", "ca_in"),
                LinearSpec("layer.1.EncDecAttention.k", d, d, "bf16", "bf16", Tkv, "row", "enc_out"),
                LinearSpec("layer.1.EncDecAttention.v", d, d, "bf16", "bf16", Tkv, "row", "enc_out"),
                LinearSpec("layer.1.EncDecAttention.o", d, d, "bf16", "bf16", Tq, "row", "ca_ctx")]
        j = 2
    lin += [LinearSpec(f"layer.{j}.DenseReluDense.wi_0", ff, d, "bf16", "bf16", Tq, "row", "ff_in"),
            LinearSpec(f"layer.{j}.DenseReluDense.wi_1", ff, d, "bf16", "bf16", Tq, "row", "ff_in"),
            LinearSpec(f"layer.{j}.DenseReluDense.wo", d, ff, "bf16", "bf16", Tq, "row", "ff_mid")]
    stack = "decoder" if decoder else "encoder"
    return BlockSpec(f"{prefix}.{stack}.block.{i}", tuple(lin))


def blip2_blocks(batch=8) -> List[BlockSpec]:
    """BLIP-2 (EVA ViT-g + FlanT5-XL), BASELINE.json configs[3]: 39 + 24 + 24 blocks, 588 Linears, 3.70 G params."""
    blocks = [vit_g_block(i, batch) for i in range(39)]
    blocks += [t5_xl_block(i, False, batch) for i in range(24)]
    blocks += [t5_xl_block(i, True, batch) for i in range(24)]
    return blocks


def llama7b_blocks(batch=1, tokens=2048) -> List[BlockSpec]:
    T = batch * tokens
    out = []
    for i in range(32):
        lin = [LinearSpec(f"self_attn.{n}_proj", 4096, 4096, "fp16", "fp16", T, "row", "attn_ctx" if n == "o" else "norm1")
               for n in "qkvo"]
        lin += [LinearSpec("mlp.gate_proj", 11008, 4096, "fp16", "fp16", T, "row", "norm2"),
                LinearSpec("mlp.up_proj", 11008, 4096, "fp16", "fp16", T, "row", "norm2"),
                LinearSpec("mlp.down_proj", 4096, 11008, "fp16", "fp16", T, "row", "act")]
        out.append(BlockSpec(f"model.layers.{i}", tuple(lin)))
    return out


BYTES = {"fp32": 4, "fp16": 2, "bf16": 2}


def norm_bytes(l: LinearSpec, n_batches: int) -> int:
    """A1 algorithmic bytes: T*C*sizeof(x) + 8*C per hook call."""
    return n_batches * (l.tokens * l.cols * BYTES[l.x_dtype] + 8 * l.cols)


def select_bytes(l: LinearSpec) -> int:
    """A4/A5 algorithmic bytes: 2*R*C*sizeof(w) + 4*C per Linear."""
    return 2 * l.rows * l.cols * BYTES[l.w_dtype] + 4 * l.cols


def unique_norm_bytes(blocks, n_batches: int) -> int:
    """Bytes of DISTINCT hook-input tensors per pass (what HBM must deliver when q/k/v etc. share their input)."""
    tot = 0
    for b in blocks:
        seen = {}
        for l in b.linears:
            seen[l.src or l.name] = l.tokens * l.cols * BYTES[l.x_dtype]
        tot += n_batches * sum(seen.values())
    return tot


def summarize(blocks, n_batches):
    lin = [l for b in blocks for l in b.linears]
    return {
        "unique_norm_input_bytes": unique_norm_bytes(blocks, n_batches),
        "linears": len(lin),
        "params": sum(l.rows * l.cols for l in lin),
        "calib_tokens_per_step": sum(l.tokens * n_batches for l in lin),
        "norm_bytes": sum(norm_bytes(l, n_batches) for l in lin),
        "select_bytes": sum(select_bytes(l) for l in lin),
    }
