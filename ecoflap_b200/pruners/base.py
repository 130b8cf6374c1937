"""Base classes of the pruner API (LAVIS/lavis/compression/pruners/base_pruner.py:17-92 and
layer_single_base_pruner.py:19-117): constructor kwargs, attribute names and the
``prune(importance_scores=None, keep_indices_or_masks=None) -> (model, sparsity_dict)`` contract are kept."""
from __future__ import annotations

from time import time


def print_time(func):
    """Wall-clock decorator the reference puts on prune/_prune/return_sparsity (pruners/utils.py:6-18)."""

    def wrapper(*args, **kwargs):
        start = time()
        ret = func(*args, **kwargs)
        print(f"{func.__name__} spent {time() - start:.3f} s")
        return ret

    wrapper.__name__ = func.__name__
    return wrapper


class BasePruner:
    def __init__(self, model, data_loader, is_strct_pruning, keep_indices_or_masks_cache, importance_scores_cache,
                 is_global, num_samples):
        self.model = model
        self.data_loader = data_loader
        self.is_strct_pruning = is_strct_pruning
        self.is_global = is_global
        self.num_samples = num_samples
        self.keep_indices_or_masks_cache = keep_indices_or_masks_cache
        self.importance_scores_cache = importance_scores_cache

    def compute_importance_scores(self, model, data_loader, loss_func):
        raise NotImplementedError

    def get_params(self, model):
        names, params = [], []
        for name, param in model.named_parameters():
            names.append(name)
            params.append(param)
        return names, params

    def convert_spec_to_list(self, spec):
        """'<nlayers>-<keep>-<attn>-<ffn>'; only the second field is used by the layer-wise pruners."""
        num_layers, res_keep, attn_keep, ffn_keep = spec.split("-")
        return int(num_layers), float(res_keep), float(attn_keep), float(ffn_keep)

    def create_pruned_arch(self, *args, **kwargs):
        return NotImplementedError

    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        raise NotImplementedError


class LayerWiseBasePruner(BasePruner):
    def __init__(self, model, data_loader, prune_spec=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 model_prefix="t5_model", sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8,
                 score_method="GradMagSquare_avg", num_data_first_stage=128, num_noise=1, sparsity_dict=None,
                 noise_eps=1e-3, prune_per_model=False, **kwargs):
        super().__init__(model=model, data_loader=data_loader, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples)
        self.sparsity_ratio_granularity = sparsity_ratio_granularity
        self.max_sparsity_per_layer = max_sparsity_per_layer
        self.score_method = score_method
        self.num_data_first_stage = num_data_first_stage
        self.num_noise = num_noise
        self.sparsity_dict = sparsity_dict
        self.noise_eps = noise_eps
        self.prune_per_model = prune_per_model
        self.prune_spec = prune_spec
        self.model_prefix = model_prefix
        self.prune_n = 0
        self.prune_m = 0
        self.model_stem = getattr(self.model, model_prefix, None)

    def model_setup_and_record_attributes(self, model):
        """Record dtypes/requires_grad, switch every parameter to requires_grad=True (:79-95)."""
        dtype_record, requires_grad_record = {}, {}
        for n, p in model.named_parameters():
            dtype_record[n] = p.data.dtype
        for n, p in model.named_parameters():
            requires_grad_record[n] = p.requires_grad
            p.requires_grad = True
        device = next(iter(model.parameters())).device
        return dtype_record, requires_grad_record, device

    def model_reset(self, model, dtype_record, requires_grad_record, device):
        for n, p in model.named_parameters():
            p.requires_grad = requires_grad_record[n]
        for n, p in model.named_parameters():
            p.data = p.data.type(dtype_record[n])
        model.to(device)

    @staticmethod
    def _load_sparsity_yaml(path):
        import yaml

        with open(path, "r") as f:
            return yaml.load(f, Loader=yaml.FullLoader)
