"""Hot-path configuration fields of the reference ``config.py`` (same names, same defaults).

``VQ_model`` accepts the reference's own argparse namespace (``config.read_arguments``); this helper builds the same
namespace without the reference on the path.  Only the fields the hot path reads are listed (config.py:40-132); the
defaults are the reference's, except ``encoder`` / ``in_channels`` which default to the synthetic Swin-3D setup the
benchmark is quoted on (the reference asks for ``--encoder Swin_3D --in_channels 1`` there, config.py:40,50).
"""
from __future__ import annotations

import argparse

DEFAULTS = dict(
    seed=0, batch_size=1,
    encoder="Swin_3D", classifier="CNN_3D", codebook="LFQ",
    in_channels_dynamic=6, in_channels=1,
    en_embed_dim=[16, 16], en_depths=[2, 1], en_patch_size=(1, 1, 1), en_window_size=[(2, 4, 4), (8, 1, 1)],
    en_mlp_ratio=4., en_drop_rate=0., en_drop_path_rate=0., en_patch_norm=False, en_use_checkpoint=False,
    en_n_heads=[2, 2], en_attn_drop_rate=0.0, en_qkv_bias=True, en_qk_scale=None,
    codebook_size=2, codebook_dim=16, cls_dim=16, cls_drop_rate=0., en_de_pretrained=None,
    delta_t=8, x_min=0, x_max=200, y_min=0, y_max=200,
    n_epochs=100, optimizer="Adam", lr=1e-3, weight_decay=0.003, beta1=0.9, beta2=0.999,
    lambda_commitment=3.0, lambda_anomaly=100.0, lambda_entropy=0.1, diversity_gamma=0.1,
)


def default_config(**overrides) -> argparse.Namespace:
    cfg = dict(DEFAULTS)
    unknown = set(overrides) - set(cfg)
    if unknown:
        raise KeyError(f"unknown config fields: {sorted(unknown)}")
    cfg.update(overrides)
    return argparse.Namespace(**cfg)
