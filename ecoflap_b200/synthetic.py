"""Random-init stand-ins for the model families the pruners operate on, with the REFERENCE'S MODULE NAMES and
call conventions ("a ModuleList of blocks whose nn.Linear children can be hooked").

The reference's model zoo (LAVIS EVA-ViT / vendored HF T5 / BLIP-2, OpenAI CLIP) cannot be imported here and
needs checkpoints; the pruners are duck-typed, so these small modules are enough to drive them end to end in
tests, ``smoke()`` and ``bench.py`` with synthetic data:

  EvaClipModel   model.visual.blocks[i].{attn.qkv, attn.proj, mlp.fc1, mlp.fc2}      (eva_vit.py:44-185,250-400)
  T5Model        model.t5_model.{encoder,decoder}.block[i].layer[j].{SelfAttention,EncDecAttention}.{q,k,v,o},
                 .DenseReluDense.{wi_0,wi_1,wo}                                      (modeling_t5.py)
  Blip2Model     model.visual_encoder.blocks + model.t5_model                         (blip2_t5.py:21-172)
  ClipModel      model.visual.transformer.resblocks / model.transformer.resblocks with nn.MultiheadAttention
                                                                                     (CoOp/clip/model.py:165-300)
"""
from __future__ import annotations

import contextlib
import math
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def _maybe_autocast(model, dtype):
    dev = next(model.parameters()).device
    if dev.type == "cpu" or dtype is None:
        return contextlib.nullcontext()
    return torch.autocast(device_type="cuda", dtype=dtype)


# =====================================================================================  EVA ViT
class EvaAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, rel_pos_bias=None):
        B, N, C = x.shape
        qkv_bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias, requires_grad=False), self.v_bias))
        qkv = self.qkv(x) + qkv_bias  # called as a module so the forward hooks fire (eva_vit.py:123-128)
        qkv = qkv.reshape(B, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=rel_pos_bias)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class EvaMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class EvaBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = EvaAttention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = EvaMlp(dim, mlp_hidden)

    def forward(self, x, rel_pos_bias=None):
        x = x + self.attn(self.norm1(x), rel_pos_bias=rel_pos_bias)
        return x + self.mlp(self.norm2(x))


class EvaVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch=16, dim=768, depth=12, heads=12, mlp_hidden=3072, num_classes=0):
        super().__init__()
        self.num_features = dim
        self.patch_embed = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        n = (img_size // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, dim) * 0.02)
        self.blocks = nn.ModuleList([EvaBlock(dim, heads, mlp_hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        self.head = nn.Linear(dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward_features(self, x):
        x = self.patch_embed(x).flatten(2).transpose(1, 2)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype), x), dim=1) + self.pos_embed.to(x.dtype)
        rel_pos_bias = None
        for blk in self.blocks:
            x = blk(x, rel_pos_bias)
        return self.norm(x)

    def forward(self, x):
        return self.head(self.forward_features(x)[:, 0])


class EvaClipModel(nn.Module):
    """EVA_CLIP-like wrapper: ``encode_image``, ``predict`` (logits * 100 + targets), ``maybe_autocast``."""

    def __init__(self, num_classes=16, autocast_dtype=None, prefix="visual", **vit_kw):
        super().__init__()
        setattr(self, prefix, EvaVisionTransformer(num_classes=num_classes, **vit_kw))
        self._prefix = prefix
        self._autocast_dtype = autocast_dtype

    def maybe_autocast(self, dtype=None):
        return _maybe_autocast(self, dtype if dtype is not None else self._autocast_dtype)

    def encode_image(self, image):
        vit = getattr(self, self._prefix)
        p = next(vit.parameters())
        x = image.to(p.device)
        if p.dtype != torch.float32:
            x = x.to(p.dtype)
        with self.maybe_autocast():
            return vit(x)

    def predict(self, samples):
        logits = self.encode_image(samples["image"]).float()
        return {"predictions": logits * 100, "targets": samples["label"].to(logits.device)}


# =====================================================================================  T5
class T5LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps

    def forward(self, x):
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        return self.weight * x


class T5Attention(nn.Module):
    def __init__(self, d_model, heads, d_kv, has_relative_attention_bias=False, num_buckets=32):
        super().__init__()
        self.n_heads, self.d_kv = heads, d_kv
        inner = heads * d_kv
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)
        self.has_relative_attention_bias = has_relative_attention_bias
        self.num_buckets = num_buckets
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    def compute_bias(self, q_len, k_len, device):
        ctx = torch.arange(q_len, device=device)[:, None]
        mem = torch.arange(k_len, device=device)[None, :]
        bucket = (mem - ctx).clamp(-(self.num_buckets // 2), self.num_buckets // 2 - 1) + self.num_buckets // 2
        return self.relative_attention_bias(bucket).permute(2, 0, 1).unsqueeze(0)  # [1, H, q, k]

    def forward(self, hidden, mask=None, key_value_states=None, position_bias=None):
        B, L, _ = hidden.shape
        kv = hidden if key_value_states is None else key_value_states
        q = self.q(hidden).view(B, L, self.n_heads, self.d_kv).transpose(1, 2)
        k = self.k(kv).view(B, kv.shape[1], self.n_heads, self.d_kv).transpose(1, 2)
        v = self.v(kv).view(B, kv.shape[1], self.n_heads, self.d_kv).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2))  # T5 does not scale by sqrt(d)
        if position_bias is None:
            if self.has_relative_attention_bias:
                position_bias = self.compute_bias(L, kv.shape[1], hidden.device)
            else:  # blocks >= 1 replayed with position_bias=None get a ZERO bias (modeling_t5.py:565-587)
                position_bias = torch.zeros((1, self.n_heads, L, kv.shape[1]), device=hidden.device, dtype=scores.dtype)
            if mask is not None:
                position_bias = position_bias + mask
        scores = scores + position_bias.to(scores.dtype)
        attn = F.softmax(scores.float(), dim=-1).type_as(scores)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(B, L, -1)
        return self.o(out), position_bias


class T5LayerSelfAttention(nn.Module):
    def __init__(self, d_model, heads, d_kv, has_bias):
        super().__init__()
        self.SelfAttention = T5Attention(d_model, heads, d_kv, has_relative_attention_bias=has_bias)
        self.layer_norm = T5LayerNorm(d_model)

    def forward(self, hidden, attention_mask=None, position_bias=None):
        out, pb = self.SelfAttention(self.layer_norm(hidden), mask=attention_mask, position_bias=position_bias)
        return hidden + out, pb


class T5LayerCrossAttention(nn.Module):
    def __init__(self, d_model, heads, d_kv):
        super().__init__()
        self.EncDecAttention = T5Attention(d_model, heads, d_kv)
        self.layer_norm = T5LayerNorm(d_model)

    def forward(self, hidden, key_value_states, attention_mask=None, position_bias=None):
        out, pb = self.EncDecAttention(self.layer_norm(hidden), mask=attention_mask, key_value_states=key_value_states,
                                       position_bias=position_bias)
        return hidden + out, pb


class T5DenseGatedActDense(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi_0 = nn.Linear(d_model, d_ff, bias=False)
        self.wi_1 = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)

    def forward(self, x):
        return self.wo(F.gelu(self.wi_0(x), approximate="tanh") * self.wi_1(x))


class T5LayerFF(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.DenseReluDense = T5DenseGatedActDense(d_model, d_ff)
        self.layer_norm = T5LayerNorm(d_model)

    def forward(self, hidden):
        return hidden + self.DenseReluDense(self.layer_norm(hidden))


class T5Block(nn.Module):
    def __init__(self, d_model, heads, d_kv, d_ff, is_decoder, has_bias):
        super().__init__()
        self.is_decoder = is_decoder
        self.layer = nn.ModuleList([T5LayerSelfAttention(d_model, heads, d_kv, has_bias)])
        if is_decoder:
            self.layer.append(T5LayerCrossAttention(d_model, heads, d_kv))
        self.layer.append(T5LayerFF(d_model, d_ff))

    def forward(self, hidden_states, attention_mask=None, position_bias=None, encoder_hidden_states=None,
                encoder_attention_mask=None, encoder_decoder_position_bias=None, layer_head_mask=None,
                cross_attn_layer_head_mask=None, **kwargs):
        hidden_states, position_bias = self.layer[0](hidden_states, attention_mask=attention_mask,
                                                     position_bias=position_bias)
        if self.is_decoder and encoder_hidden_states is not None:
            hidden_states, encoder_decoder_position_bias = self.layer[1](
                hidden_states, encoder_hidden_states, attention_mask=encoder_attention_mask,
                position_bias=encoder_decoder_position_bias)
        hidden_states = self.layer[-1](hidden_states)
        return (hidden_states, position_bias, encoder_decoder_position_bias)


class T5Stack(nn.Module):
    def __init__(self, embed, d_model, heads, d_kv, d_ff, depth, is_decoder):
        super().__init__()
        self.embed_tokens = embed
        self.is_decoder = is_decoder
        self.block = nn.ModuleList([T5Block(d_model, heads, d_kv, d_ff, is_decoder, has_bias=(i == 0)) for i in range(depth)])
        self.final_layer_norm = T5LayerNorm(d_model)

    def forward(self, input_ids=None, inputs_embeds=None, encoder_hidden_states=None):
        h = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        B, L, _ = h.shape
        mask = None
        if self.is_decoder:
            causal = torch.full((L, L), float("-inf"), device=h.device).triu(1)
            mask = causal[None, None].to(h.dtype)
        position_bias, enc_dec_bias = None, None
        for blk in self.block:
            # every key the reference's Catcher reads must be passed by keyword (wanda_pruner.py:179-195)
            out = blk(h, attention_mask=mask, position_bias=position_bias, encoder_hidden_states=encoder_hidden_states,
                      encoder_attention_mask=None, encoder_decoder_position_bias=enc_dec_bias, layer_head_mask=None,
                      cross_attn_layer_head_mask=None)
            h, position_bias, enc_dec_bias = out[0], out[1], out[2]
        return self.final_layer_norm(h)


class T5ForConditionalGeneration(nn.Module):
    def __init__(self, vocab=512, d_model=64, heads=4, d_kv=16, d_ff=128, depth=2):
        super().__init__()
        self.config = SimpleNamespace(use_cache=True, d_model=d_model)
        self.shared = nn.Embedding(vocab, d_model)
        self.encoder = T5Stack(self.shared, d_model, heads, d_kv, d_ff, depth, is_decoder=False)
        self.decoder = T5Stack(self.shared, d_model, heads, d_kv, d_ff, depth, is_decoder=True)
        self.lm_head = nn.Linear(d_model, vocab, bias=False)

    def forward(self, input_ids=None, inputs_embeds=None, labels=None):
        enc = self.encoder(input_ids=input_ids, inputs_embeds=inputs_embeds)
        start = torch.zeros_like(labels[:, :1])
        dec_in = torch.cat([start, labels[:, :-1]], dim=1)
        dec = self.decoder(input_ids=dec_in, encoder_hidden_states=enc)
        logits = self.lm_head(dec)
        loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), labels.reshape(-1))
        return SimpleNamespace(loss=loss, logits=logits)


class T5Model(nn.Module):
    """LAVIS ``T5`` wrapper: ``model.t5_model``; ``model(samples) -> {"loss"}``; bf16 autocast."""

    def __init__(self, autocast=True, **t5_kw):
        super().__init__()
        self.t5_model = T5ForConditionalGeneration(**t5_kw)
        self._autocast = autocast

    def maybe_autocast(self, dtype=torch.bfloat16):
        return _maybe_autocast(self, dtype if self._autocast else None)

    def forward(self, samples):
        dev = self.t5_model.shared.weight.device
        with self.maybe_autocast(dtype=torch.bfloat16):
            out = self.t5_model(input_ids=samples["input_ids"].to(dev), labels=samples["labels"].to(dev))
        return {"loss": out.loss}


class Blip2Model(nn.Module):
    """BLIP-2-like: frozen-ViT features -> linear projection to ``n_query`` prefix embeddings -> T5."""

    def __init__(self, vit_kw=None, t5_kw=None, n_query=8, autocast=True):
        super().__init__()
        vit_kw = dict(vit_kw or {})
        t5_kw = dict(t5_kw or {})
        self.visual_encoder = EvaVisionTransformer(**vit_kw)
        self.ln_vision = nn.LayerNorm(self.visual_encoder.num_features)
        self.t5_model = T5ForConditionalGeneration(**t5_kw)
        self.t5_proj = nn.Linear(self.visual_encoder.num_features, self.t5_model.config.d_model)
        self.n_query = n_query
        self._autocast = autocast

    def maybe_autocast(self, dtype=torch.float16):
        return _maybe_autocast(self, dtype if self._autocast else None)

    def forward(self, samples):
        p = next(self.visual_encoder.parameters())
        image = samples["image"].to(device=p.device)
        with self.maybe_autocast():
            feats = self.ln_vision(self.visual_encoder.forward_features(image.to(p.dtype) if not self._autocast else image))
        with self.maybe_autocast(dtype=torch.bfloat16):
            prefix = self.t5_proj(feats[:, : self.n_query].to(self.t5_proj.weight.dtype))
            tok = self.t5_model.shared(samples["input_ids"].to(p.device))
            out = self.t5_model(inputs_embeds=torch.cat([prefix.to(tok.dtype), tok], dim=1),
                                labels=samples["labels"].to(p.device))
        return {"loss": out.loss}


# =====================================================================================  CLIP (CoOp)
class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ClipLayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.type(torch.float32)).type(x.dtype) if self.weight.dtype == torch.float32 else super().forward(x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x):
        self.attn_mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]

    def forward(self, x):
        x = x + self.attention(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class ClipTransformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class ClipVisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = ClipTransformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x):
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x).permute(1, 0, 2)  # NLD -> LND
        x = self.transformer(x).permute(1, 0, 2)
        return self.ln_post(x[:, 0, :]) @ self.proj


class ClipModel(nn.Module):
    def __init__(self, embed_dim=64, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=8,
                 context_length=16, vocab_size=256, transformer_width=32, transformer_heads=4, transformer_layers=2,
                 vision_heads=None):
        super().__init__()
        self.context_length = context_length
        self.visual = ClipVisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                            vision_heads or max(1, vision_width // 64), embed_dim)
        mask = torch.empty(context_length, context_length).fill_(float("-inf")).triu_(1)
        self.transformer = ClipTransformer(transformer_width, transformer_layers, transformer_heads, attn_mask=mask)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.randn(context_length, transformer_width) * 0.01)
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.randn(transformer_width, embed_dim) * transformer_width ** -0.5)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype) + self.positional_embedding.type(self.dtype)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection


def clip_forward_to_cache(class_tokens):
    """The closure CoOp assigns to ``pruner.forward_to_cache`` (CoOp/trainers/zsclip.py:73-93): symmetric CLIP
    cross entropy between image features and the text features of the batch labels; returns (loss, batch_len)."""

    def forward_to_cache(model, batch, device):
        dev = next(model.parameters()).device
        image, label = batch["img"].to(dev), batch["label"].to(dev)
        img = model.encode_image(image)
        txt = model.encode_text(class_tokens.to(dev)[label])
        img = img / img.norm(dim=-1, keepdim=True)
        txt = txt / txt.norm(dim=-1, keepdim=True)
        logits = model.logit_scale.exp() * img @ txt.t()
        target = torch.arange(len(image), device=dev)
        loss = (F.cross_entropy(logits.float(), target) + F.cross_entropy(logits.float().t(), target)) / 2
        return loss, len(image)

    return forward_to_cache


# =====================================================================================  LLaMA-style decoder
class LlamaRMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return (self.weight * (x.float() * torch.rsqrt(v + self.eps))).to(x.dtype)


class LlamaDecoderLayer(nn.Module):
    """``self_attn.{q,k,v,o}_proj`` + ``mlp.{gate,up,down}_proj`` (the names upstream Wanda's LLaMA loop finds).
    ``tuple_output`` selects the transformers < 5 convention (a tuple) or the >= 5 one (a bare tensor)."""

    def __init__(self, dim, heads, ffn, tuple_output=False):
        super().__init__()
        self.heads, self.tuple_output = heads, tuple_output
        self.input_layernorm = LlamaRMSNorm(dim)
        self.self_attn = nn.ModuleDict({k: nn.Linear(dim, dim, bias=False) for k in ("q_proj", "k_proj", "v_proj", "o_proj")})
        self.post_attention_layernorm = LlamaRMSNorm(dim)
        self.mlp = nn.ModuleDict({"gate_proj": nn.Linear(dim, ffn, bias=False), "up_proj": nn.Linear(dim, ffn, bias=False),
                                  "down_proj": nn.Linear(ffn, dim, bias=False)})

    def forward(self, hidden_states, attention_mask=None, position_ids=None, **kwargs):
        B, L, C = hidden_states.shape
        h = self.input_layernorm(hidden_states)
        a = self.self_attn
        q, k, v = (a[n](h).view(B, L, self.heads, -1).transpose(1, 2) for n in ("q_proj", "k_proj", "v_proj"))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, L, C)
        hidden_states = hidden_states + a["o_proj"](o)
        h = self.post_attention_layernorm(hidden_states)
        hidden_states = hidden_states + self.mlp["down_proj"](F.silu(self.mlp["gate_proj"](h)) * self.mlp["up_proj"](h))
        return (hidden_states,) if self.tuple_output else hidden_states


class LlamaModel(nn.Module):
    """``model.model.layers`` decoder stack, ``model(input_ids, labels=...)`` -> object with ``.loss`` / ``.logits``,
    ``model.config.use_cache`` (LLaMA/main.py:60-80 duck type)."""

    def __init__(self, vocab=128, dim=64, heads=4, ffn=160, depth=2, tuple_output=False):
        super().__init__()
        self.config = SimpleNamespace(use_cache=True, hidden_size=dim)
        self.model = nn.Module()
        self.model.embed_tokens = nn.Embedding(vocab, dim)
        self.model.layers = nn.ModuleList([LlamaDecoderLayer(dim, heads, ffn, tuple_output) for _ in range(depth)])
        self.model.norm = LlamaRMSNorm(dim)
        self.lm_head = nn.Linear(dim, vocab, bias=False)

    def forward(self, input_ids, labels=None, **kwargs):
        x = self.model.embed_tokens(input_ids)
        pos = torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        for layer in self.model.layers:
            out = layer(x, attention_mask=None, position_ids=pos)
            x = out[0] if isinstance(out, (tuple, list)) else out
        logits = self.lm_head(self.model.norm(x))
        loss = None
        if labels is not None:
            loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1))
        return SimpleNamespace(loss=loss, logits=logits)


def token_batches(n, batch, seq_len, vocab, seed=0):
    """(input_ids, targets) tuples, the shape upstream Wanda's ``get_loaders`` yields."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n // batch):
        ids = torch.randint(1, vocab, (batch, seq_len), generator=g)
        out.append((ids, ids.clone()))
    return ListLoader(out)


# =====================================================================================  BLIP (UPop) NLVR stand-in
class BlipVitBlock(EvaBlock):
    """UPop's ViT blocks are called ``blk(x, register_hook)`` (UPop/models/vit.py); the flag is ignored here."""

    def forward(self, x, register_hook=False):
        return super().forward(x, None)


class BertSelfOutput(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dense = nn.Linear(dim, dim)
        self.LayerNorm = nn.LayerNorm(dim)

    def forward(self, h, inp):
        return self.LayerNorm(self.dense(h) + inp)


class BertAttention(nn.Module):
    def __init__(self, dim, heads, kv_dim=None):
        super().__init__()
        self.heads = heads
        kv_dim = kv_dim or dim
        self.self = nn.ModuleDict({"query": nn.Linear(dim, dim), "key": nn.Linear(kv_dim, dim), "value": nn.Linear(kv_dim, dim)})
        self.output = BertSelfOutput(dim)

    def forward(self, x, kv=None, mask=None):
        B, L, C = x.shape
        kv = x if kv is None else kv
        q = self.self["query"](x).view(B, L, self.heads, -1).transpose(1, 2)
        k = self.self["key"](kv).view(B, kv.shape[1], self.heads, -1).transpose(1, 2)
        v = self.self["value"](kv).view(B, kv.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(B, L, C)
        return self.output(o, x)


class BertLayer(nn.Module):
    """BERT layer with cross-attention onto the image tokens; returns a tuple like HF's (UPop/models/med.py)."""

    def __init__(self, dim, heads, ffn):
        super().__init__()
        self.attention = BertAttention(dim, heads)
        self.crossattention = BertAttention(dim, heads)
        self.intermediate = nn.ModuleDict({"dense": nn.Linear(dim, ffn)})
        self.output = BertSelfOutput(dim)
        self.output.dense = nn.Linear(ffn, dim)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, output_attentions=False, mode="multimodal", **kwargs):
        h = self.attention(hidden_states, mask=attention_mask)
        if encoder_hidden_states is not None and mode == "multimodal":
            h = self.crossattention(h, kv=encoder_hidden_states, mask=encoder_attention_mask)
        return (self.output(F.gelu(self.intermediate["dense"](h)), h),)


class BlipCaptionModel(nn.Module):
    """``visual_encoder.blocks`` + ``text_decoder.bert.encoder.layer``; ``model(image, caption)`` -> LM loss
    (UPop/models/blip.py duck type used by BLIPBertLayerWandaPruner, task="coco"; ``caption`` is a token tensor here)."""

    def __init__(self, img_size=32, patch=8, dim=64, depth=2, heads=4, mlp_hidden=128, vocab=128, text_depth=2):
        super().__init__()
        self.visual_encoder = EvaVisionTransformer(img_size, patch, dim, 0, heads, mlp_hidden)
        self.visual_encoder.blocks = nn.ModuleList([BlipVitBlock(dim, heads, mlp_hidden) for _ in range(depth)])
        self.text_decoder = nn.Module()
        self.text_decoder.config = SimpleNamespace(use_cache=True)
        self.text_decoder.bert = nn.Module()
        self.text_decoder.bert.embeddings = nn.Embedding(vocab, dim)
        self.text_decoder.bert.encoder = nn.Module()
        self.text_decoder.bert.encoder.layer = nn.ModuleList([BertLayer(dim, heads, mlp_hidden) for _ in range(text_depth)])
        self.text_decoder.cls = nn.Linear(dim, vocab)

    def encode_image(self, image):
        ve = self.visual_encoder
        x = ve.patch_embed(image).flatten(2).transpose(1, 2)
        x = torch.cat((ve.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + ve.pos_embed
        for blk in ve.blocks:
            x = blk(x, False)
        return ve.norm(x)

    def forward(self, image, caption):
        enc = self.encode_image(image)
        caption = caption.to(enc.device)
        h = self.text_decoder.bert.embeddings(caption)
        for layer in self.text_decoder.bert.encoder.layer:
            h = layer(h, attention_mask=None, head_mask=None, encoder_hidden_states=enc, encoder_attention_mask=None,
                      output_attentions=False, mode="multimodal")[0]
        logits = self.text_decoder.cls(h)
        return F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), caption[:, 1:].reshape(-1))


def caption_batches(n, batch, res, seq_len, vocab, seed=0):
    """(image, caption, image_id) tuples (UPop coco caption loaders; ``caption`` is a token tensor in this stand-in)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n // batch):
        out.append((torch.randn(batch, 3, res, res, generator=g), torch.randint(1, vocab, (batch, seq_len), generator=g),
                    torch.arange(batch)))
    return ListLoader(out)


# =====================================================================================  synthetic loaders
class ListLoader:
    """Re-iterable list of batches (the reference re-iterates its calibration loader once per tower and,
    for the zeroth-order score, once per layer)."""

    def __init__(self, batches):
        self.batches = list(batches)

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def image_batches(n, batch, res, classes=16, seed=0, key="image", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n // batch):
        out.append({key: torch.randn(batch, 3, res, res, generator=g).to(dtype),
                    "label": torch.randint(0, classes, (batch,), generator=g)})
    return ListLoader(out)


def text_batches(n, batch, src_len, tgt_len, vocab, seed=0, with_image=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n // batch):
        b = {"input_ids": torch.randint(1, vocab, (batch, src_len), generator=g),
             "labels": torch.randint(1, vocab, (batch, tgt_len), generator=g),
             "text_input": ["synthetic"] * batch}
        if with_image is not None:
            b["image"] = torch.randn(batch, 3, with_image, with_image, generator=g)
        out.append(b)
    return ListLoader(out)


def init_weights_(model, std=0.02, seed=0):
    """N(0, 0.02^2) for every matrix (SURVEY 8d synthetic inputs), deterministic.  LayerNorm gains/biases are
    randomised too: with the default (1, 0) affine every LayerNorm output sums to zero over the channels, which
    makes the SparseGPT Hessian of the following Linear exactly singular -- trained models never are."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * std)
        for m in model.modules():
            if isinstance(m, nn.LayerNorm) and m.elementwise_affine:
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.3 * torch.randn(m.bias.shape, generator=g))
    return model


# =====================================================================================  full-size BLIP-2 (bench)
BLIP2_VIT_G = dict(img_size=224, patch=14, dim=1408, depth=39, heads=16, mlp_hidden=6144)        # eva_vit.py:444-470
FLAN_T5_XL = dict(vocab=32128, d_model=2048, heads=32, d_kv=64, d_ff=5120, depth=24)             # google/flan-t5-xl config


def blip2_full(device, seed=0):
    """Random-init BLIP-2 at the real size (EVA ViT-g fp16 + FlanT5-XL bf16: 588 prunable Linears, 3.70 G parameters),
    built and initialised on the device (N(0, 0.02^2) matrices, randomised LayerNorm affines as in init_weights_)."""
    with torch.device(device):
        model = Blip2Model(vit_kw=BLIP2_VIT_G, t5_kw=FLAN_T5_XL, n_query=32, autocast=True)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02, generator=g)
        for m in model.modules():
            if isinstance(m, nn.LayerNorm) and m.elementwise_affine:
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=g, device=device))
                m.bias.copy_(0.3 * torch.randn(m.bias.shape, generator=g, device=device))
    model.visual_encoder.half()
    model.ln_vision.half()
    model.t5_model.bfloat16()
    model.t5_proj.bfloat16()
    return model.eval()


def blip2_full_loader(n=128, batch=8, text_len=24, tgt_len=32, seed=0):
    """128 synthetic (image, prompt, target) samples in batches of 8: 32 query + 24 prompt tokens
    into the T5 encoder, 32 target tokens into the decoder."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n // batch):
        out.append({"image": torch.randn(batch, 3, 224, 224, generator=g).half(),
                    "input_ids": torch.randint(1, 32128, (batch, text_len), generator=g),
                    "labels": torch.randint(1, 32128, (batch, tgt_len), generator=g),
                    "text_input": ["synthetic"] * batch})
    return ListLoader(out)
