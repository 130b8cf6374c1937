"""Shared definitions of the end-to-end pruner cases: tiny random-init models + seeded synthetic calibration
loaders.  Used by tests/gen_golden_e2e.py (reference, CPU) and tests/test_gpu_pruners.py (product, GPU) so both
sides see bit-identical initial weights and batches."""
import torch

from ecoflap_b200 import synthetic as syn

VIT_KW = dict(img_size=32, patch=8, dim=64, depth=3, heads=4, mlp_hidden=128)
T5_KW = dict(vocab=128, d_model=64, heads=4, d_kv=16, d_ff=128, depth=2)


def vit_model():
    torch.manual_seed(0)
    return syn.init_weights_(syn.EvaClipModel(num_classes=16, **VIT_KW), seed=1).eval()


def vit_loader(batch=4, n=16):
    return syn.image_batches(n, batch, 32, classes=16, seed=2)


def t5_model():
    torch.manual_seed(0)
    return syn.init_weights_(syn.T5Model(autocast=False, **T5_KW), seed=3).eval()


def t5_loader(batch=4, n=16):
    return syn.text_batches(n, batch, 12, 6, 128, seed=4, with_image=8)


def blip2_model():
    torch.manual_seed(0)
    return syn.init_weights_(syn.Blip2Model(vit_kw=VIT_KW, t5_kw=T5_KW, n_query=5, autocast=False), seed=5).eval()


def blip2_loader(batch=4, n=16):
    return syn.text_batches(n, batch, 10, 6, 128, seed=6, with_image=32)


def clip_model():
    torch.manual_seed(0)
    return syn.init_weights_(syn.ClipModel(vision_heads=4), seed=7).eval()


def clip_loader(batch=4, n=16):
    return syn.image_batches(n, batch, 32, classes=64, seed=8, key="img")


def clip_class_tokens():
    g = torch.Generator().manual_seed(9)
    t = torch.randint(1, 200, (64, 16), generator=g)
    t[:, 10] = 255  # argmax position = "EOT" token
    return t


def prunable_state(model):
    return {k: v.detach().float().cpu().numpy() for k, v in model.state_dict().items() if v.dim() == 2}


def caption_model():
    torch.manual_seed(0)
    return syn.init_weights_(syn.BlipCaptionModel(), seed=11).eval()


def caption_loader(batch=4, n=16):
    return syn.caption_batches(n, batch, 32, 6, 128, seed=12)


LLAMA_KW = dict(vocab=128, dim=64, heads=4, ffn=160, depth=3)


def llama_model(tuple_output=False):
    torch.manual_seed(0)
    return syn.init_weights_(syn.LlamaModel(tuple_output=tuple_output, **LLAMA_KW), seed=13).eval()


def llama_loader(batch=1, n=16, seq_len=24):
    return syn.token_batches(n, batch, seq_len, 128, seed=14)
