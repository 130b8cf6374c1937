"""Loss closures used by stage 1 (LAVIS/lavis/compression/pruners/utils.py:21-66).  Each returns
``(loss, batch_len)``.  ``prepare_sample`` mirrors lavis.datasets.data_utils.prepare_sample: tensors of the
batch dict are moved to the GPU when ``cuda_enabled`` is truthy (the reference passes ``device != "cpu"``,
which is always True for a torch.device)."""
from __future__ import annotations

import torch


def prepare_sample(samples, cuda_enabled=True):
    if not cuda_enabled or not torch.cuda.is_available():
        return samples

    def move(x):
        if torch.is_tensor(x):
            return x.cuda(non_blocking=True)
        if isinstance(x, dict):
            return {k: move(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(move(v) for v in x)
        return x

    return move(samples)


def loss_vision_language(model, samples, cuda_enabled):
    samples = prepare_sample(samples, cuda_enabled=cuda_enabled)
    loss = model(samples)["loss"]
    return loss, len(samples["text_input"])


def loss_language(model, samples, cuda_enabled):
    samples = prepare_sample(samples, cuda_enabled=cuda_enabled)
    loss = model(samples)["loss"]
    return loss, len(samples["text_input"])


def loss_vision(model, samples, cuda_enabled):
    """Cross entropy over model.predict() logits / 100 (utils.py:47-66)."""
    samples = prepare_sample(samples, cuda_enabled=cuda_enabled)
    outputs = model.predict(samples)
    logits = outputs["predictions"] / 100
    targets = outputs["targets"]
    probs = torch.nn.functional.softmax(logits, -1)
    picked = probs[torch.arange(len(targets)).to(targets.device), targets]
    return -picked.log().mean(), len(targets)
