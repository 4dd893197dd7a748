"""`info_nce_loss`: the loss the reference's headline configs train with (model/loss.py:52-57, selected by
`"loss": "info_nce_loss"` in config_files/config.mag.json / config.wordnet.json and applied to the
`prediction.reshape(n_batches, -1)` of trainer/trainer.py:52-56), on the library's fused kernel (tx_info_nce_fwd/bwd,
SURVEY.md section 8 row f1).  Same name, arguments and value as the reference function; no CPU fallback.

The reference's other losses (bce / margin-rank / nll variants, model/loss.py:1-50) are plain torch one-liners outside the
hot path and are not restated here.
"""
from __future__ import annotations

import torch

from . import functional as txf


def info_nce_loss(output: torch.Tensor, target: torch.Tensor = None) -> torch.Tensor:
    """
    output: a (batch_size, 1+negative_size) tensor of matching scores
    target: a (batch_size, ) tensor of dtype long - all zeros in the reference (the positive comes first,
            data_loader/dataset.py:308-313); None means exactly that and skips the index upload
    returns sum_q cross_entropy(output[q], target[q])  (reduction="sum", loss.py:57)
    """
    if output.dim() != 2:
        raise ValueError(f"info_nce_loss expects a (batch_size, 1 + negative_size) tensor, got {tuple(output.shape)}")
    t32 = None
    if target is not None:
        if target.shape != (output.shape[0],):
            raise ValueError(f"info_nce_loss: target shape {tuple(target.shape)} for {output.shape[0]} queries")
        t32 = target.to(device=output.device, dtype=torch.int32).contiguous()
    return txf.InfoNCE.apply(output, t32)
