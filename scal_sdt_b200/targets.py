"""Programmatic builders for the target-selection config of the benchmark workloads.

Users of the reference pass ``configs/optim_targets/<name>.yaml`` (load it with ``scal_sdt_b200.config.load_yaml`` and
hand ``cfg['unet']['targets']`` to ``config_module``).  The benchmark cannot read the reference tree on the GPU box, so
the same nested ``{index, targets, recurse_conf}`` structure is produced here from the workload description in
BASELINE.json ("attention q/k/v/out", "attention + FF", all 12 targets of the stock ``lora`` target).
"""
from __future__ import annotations

ATTN_BLOCKS = ["down_blocks.0", "down_blocks.1", "down_blocks.2", "mid_block", "up_blocks.1", "up_blocks.2", "up_blocks.3"]


def lora_unet_targets(rank: int = 16, alpha=1, dropout: float = 0.0, attention: bool = True, feed_forward: bool = True,
                      projections: bool = True, lr: float = 5e-4, weight_decay: float = 2e-2) -> list:
    inner = []
    if attention:
        inner.append({"index": ["attn1", "attn2"], "targets": [{"index": ["to_q", "to_k", "to_v", "to_out.0"]}]})
    if feed_forward:
        inner.append({"index": ["ff.net.0.proj", "ff.net.2"]})
    per_transformer = []
    if inner:
        per_transformer.append({"index": ["transformer_blocks"], "targets": [{"targets": inner}]})
    if projections:
        per_transformer.append({"index": ["proj_in", "proj_out"]})
    return [{
        "index": list(ATTN_BLOCKS),
        "recurse_conf": {"lora": {"rank": rank, "alpha": alpha, "dropout": dropout},
                         "optimizer": {"lr": lr, "weight_decay": weight_decay}},
        "targets": [{"index": ["attentions"], "targets": [{"targets": per_transformer}]}],
    }]


def full_unet_targets(lr: float = 5e-6, weight_decay: float = 1e-2) -> list:
    """Native full fine-tune: every child of the UNet is selected with its own parameters (no ``lora`` key)."""
    return [{"recurse_conf": {"optimizer": {"lr": lr, "weight_decay": weight_decay}}}]
