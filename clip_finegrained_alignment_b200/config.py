"""Hyper-parameter dataclass with the reference's field names and defaults (finetune/config.py:4-28).

SPARCLoss only reads `similarity_threshold`, `global_loss_weight`, `local_loss_weight` and
`inverse_temperature`; any object exposing those four attributes works.
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass
class CLIPFineTuneConfig:
    lr: float = 1e-5
    batch_size: int = 32
    max_grad_norm: float = 1.0
    warmup_steps: int = 1000
    max_epochs: int = 400
    save_every: int = 1
    weight_decay: float = 0.2
    use_amp: bool = True
    clip_model: str = "ViT-B/32"
    max_length: int = 77
    experiment_name: str = "clip_default"
    gradient_accumulation_steps: int = 4
    loss_type: str = "count"
    similarity_threshold: float = 0.5
    global_loss_weight: float = 1.0
    local_loss_weight: float = 1.0
    inverse_temperature: float = 1.0
    optimizer_type: str = "adamw"
    betas: Tuple[float, float] = (0.9, 0.98)
    eps: float = 5e-6
    amsgrad: bool = False
    count_alpha: float = 1.0
