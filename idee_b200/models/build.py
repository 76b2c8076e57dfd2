"""Model assembly: drop-in for the reference ``models/build.py`` (``import_class``, ``VQ_model``).

``VQ_model(config)`` takes the reference's ``config.py`` namespace, builds encoder -> quantiser -> classifier with the
reference's module/parameter names and initialisation (build.py:23-127), and ``forward(x_d[N,V,C,T,H,W])`` returns the
reference 5-tuple ``(z, y, anomaly[N,V,T,H,W], z_q[N,V,C,T,H,W], loss_z_q[1])`` (build.py:130-159).  Internally every
activation stays in the channel-last token layout [N,V,T,H,W,16]; the reference's two 983 MB permute copies around the
quantiser (build.py:150,153) are views here.
"""
from __future__ import annotations

import importlib

import torch
import torch.nn as nn


def import_class(en_de, name):
    """'encoder'|'codebook'|'classifier', class name -> class (build.py:17-20), resolved inside idee_b200.models."""
    module = importlib.import_module(f"{__package__}.{en_de}.{name}")
    return getattr(module, name)


class VQ_model(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.encoder == "Swin_3D":
            self.encoder = import_class('encoder', config.encoder)(
                in_vars=config.in_channels_dynamic, in_chans=config.in_channels, embed_dim=config.en_embed_dim,
                window_size=config.en_window_size, depths=config.en_depths, num_heads=config.en_n_heads,
                mlp_ratio=config.en_mlp_ratio, drop_rate=config.en_drop_rate, attn_drop_rate=config.en_attn_drop_rate,
                drop_path_rate=config.en_drop_path_rate, qkv_bias=config.en_qkv_bias, qk_scale=config.en_qk_scale,
                patch_size=config.en_patch_size, patch_norm=config.en_patch_norm, use_checkpoint=config.en_use_checkpoint)
        elif config.encoder == "CNN_3D":
            self.encoder = import_class('encoder', config.encoder)(
                in_vars=config.in_channels_dynamic, in_channels=config.in_channels, out_channels=config.en_embed_dim,
                drop_path_rate=config.en_drop_path_rate, drop_rate=config.en_drop_rate)
        else:
            raise NotImplementedError(f"idee_b200: encoder {config.encoder} is not built (built: Swin_3D, CNN_3D; Mamba needs the "
                                      f"un-vendored mamba_ssm package)")
        self.cls = import_class('classifier', 'CNN_3D')(in_var=config.in_channels_dynamic, embed_dim=config.codebook_dim,
                                                        dim=config.cls_dim, drop_rate=config.cls_drop_rate)
        self.vq = import_class('codebook', 'LFQ')(dim=config.codebook_dim, codebook_size=config.codebook_size,
                                                  entropy_loss_weight=config.lambda_entropy,
                                                  diversity_gamma=config.diversity_gamma,
                                                  commitment_loss_weight=config.lambda_commitment)
        self.pretrained = config.en_de_pretrained
        self._init_weights()

    def _init_weights(self, init_type='normal', gain=.02):
        """Every Conv*/Linear* weight ~ N(0.02, gain), biases 0 (build.py:96-118); then optional checkpoint load (:120-127)."""
        for m in self.modules():
            name = m.__class__.__name__
            if 'BatchNorm2d' in name or 'BatchNorm3d' in name or 'LayerNorm' in name:
                if getattr(m, 'weight', None) is not None:
                    nn.init.constant_(m.weight.data, 1.0)
                if getattr(m, 'bias', None) is not None:
                    nn.init.constant_(m.bias.data, 0.0)
            elif hasattr(m, 'weight') and ('Conv' in name or 'Linear' in name):
                if init_type == 'normal':
                    nn.init.normal_(m.weight.data, 0.02, gain)
                elif init_type == 'xavier':
                    nn.init.xavier_normal_(m.weight.data, gain=gain)
                else:
                    raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
                if getattr(m, 'bias', None) is not None:
                    nn.init.constant_(m.bias.data, 0.0)
        if self.pretrained:
            print('initialize weights from pretrained model {} ...'.format(self.pretrained))
            checkpoint = torch.load(self.pretrained, map_location='cpu')
            state_dict = {k.replace('module.', ''): v for k, v in checkpoint['model_state_dict'].items()}
            self.load_state_dict(state_dict, strict=True)

    def forward(self, x_d):
        from .. import _lib
        one_bit = self.vq.codebook_dim == 1        # the IDEE configuration: rank-1 z_q, project_in folded into the encoder's last conv
        if _lib.PRECISION == "bf16" and getattr(self.encoder, "forward_tokens", None) is not None and self.vq.dim == 16 and one_bit:
            # the quantiser's project_in (Linear 16 -> 1) is the only consumer of the encoder output: fold it into the encoder's
            # last 3x3x3 conv (one 16 -> 1 conv instead of 16 -> 16 followed by a dot product; forward, data and weight gradient)
            s = self.encoder.forward_tokens(x_d, fold_last=(self.vq.project_in.weight, self.vq.project_in.bias))   # [N,V,T,H,W]
            N, V, T, H, W = s.shape
            C = self.vq.dim
            z_q, anomaly, loss_z_q = self.vq.forward_projected(s.reshape(N, V * T * H * W), want_bf16=True)
        else:
            tok = self.encoder.forward_tokens(x_d)                         # [N,V,T,H,W,C] channel-last
            N, V, T, H, W, C = tok.shape
            z_q, anomaly, loss_z_q = self.vq(tok.view(N, V * T * H * W, C))  # token order (v,t,h,w) as build.py:150
        z_q = z_q.view(N, V, T, H, W, C).permute(0, 1, 5, 2, 3, 4)      # logical [N,V,C,T,H,W]
        anomaly = anomaly.view(N, V, T, H, W)
        if not one_bit:                                                     # codebook_size 2^K, K > 1: z_q has rank K, generic paths
            z, y = self.cls(z_q)
            return z, y, anomaly, z_q, loss_z_q.unsqueeze(0)
        # the joint head consumes the rank-1 form of z_q (x * w_out + b_out): identical result, 1/6 of the conv1 work
        rank1 = (self.vq.last_scalar.view(N, V, T, H, W), self.vq.project_out.weight, self.vq.project_out.bias)
        zq16 = getattr(self.vq, "last_zq_bf16", None)
        if zq16 is not None:               # bf16 copy written by the quantiser kernel: input of the per-variable heads' first conv
            z_q._idee_bf16 = zq16.view(N, V, T, H, W, C)
            self.vq.last_zq_bf16 = None
        z, y = self.cls(z_q, rank1=rank1)
        z_q._idee_rank1 = rank1            # lets Anomaly_L1_loss_synthetic evaluate the same loss on the scalar plane
        return z, y, anomaly, z_q, loss_z_q.unsqueeze(0)
