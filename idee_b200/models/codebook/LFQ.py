"""B200-native lookup-free quantiser: drop-in for the reference ``models/codebook/LFQ.py``.

Same constructor, parameters (``project_in``/``project_out``), buffers (``mask`` persistent, ``zero``/``codebook`` not),
``forward(x[b,n,d]) -> Return(quantized, indices, entropy_aux_loss)`` and ``indices_to_codes`` (LFQ.py:70-190).
The forward/backward arithmetic (projection, sign, straight-through estimator, index, entropy + commitment losses,
output projection) is ONE fused streaming CUDA kernel each way (idee_b200/csrc/lfq.cu); always fp32, like the reference
which pins the quantiser to fp32 with ``autocast(enabled=False)`` (LFQ.py:183,199).
"""
from __future__ import annotations

from collections import namedtuple
from math import log2

import torch
from torch import nn

from ... import ops

Return = namedtuple('Return', ['quantized', 'indices', 'entropy_aux_loss'])
LossBreakdown = namedtuple('LossBreakdown', ['per_sample_entropy', 'batch_entropy', 'commitment'])


class LFQ(nn.Module):
    def __init__(self, *, dim=16, dim_out=16, codebook_size=2, entropy_loss_weight=0.1, commitment_loss_weight=1.5,
                 diversity_gamma=1., straight_through_activation=nn.Identity(), num_codebooks=1,
                 keep_num_codebooks_dim=None, codebook_scale=1., frac_per_sample_entropy=1.):
        super().__init__()
        assert dim is not None or codebook_size is not None, 'either dim or codebook_size must be specified for LFQ'
        assert codebook_size is None or log2(codebook_size).is_integer(), 'your codebook size must be a power of 2'
        codebook_size = codebook_size if codebook_size is not None else 2 ** dim
        codebook_dim = int(log2(codebook_size))
        codebook_dims = codebook_dim * num_codebooks
        dim = dim if dim is not None else codebook_dims
        if not (codebook_size in (2, 4, 8, 16) and num_codebooks == 1 and dim == 16 and codebook_scale == 1.
                and frac_per_sample_entropy == 1. and isinstance(straight_through_activation, nn.Identity)):
            raise NotImplementedError("idee_b200: the LFQ kernels are built for dim=16, codebook_size 2 (the IDEE configuration, "
                                      "build.py:87-91), 4, 8 or 16, one codebook, scale 1, identity activation")
        has_projections = dim != codebook_dims
        self.project_in = nn.Linear(dim, codebook_dims) if has_projections else nn.Identity()
        self.project_out = nn.Linear(codebook_dims, dim) if has_projections else nn.Identity()
        self.has_projections = has_projections
        self.dim, self.dim_out, self.codebook_dim, self.num_codebooks = dim, dim_out, codebook_dim, num_codebooks
        self.codebook_size = codebook_size
        self.keep_num_codebooks_dim = keep_num_codebooks_dim if keep_num_codebooks_dim is not None else num_codebooks > 1
        self.activation = straight_through_activation
        self.frac_per_sample_entropy = frac_per_sample_entropy
        self.diversity_gamma = diversity_gamma
        self.entropy_loss_weight = entropy_loss_weight
        self.codebook_scale = codebook_scale
        self.commitment_loss_weight = commitment_loss_weight
        self.register_buffer('mask', 2 ** torch.arange(codebook_dim - 1, -1, -1))
        self.register_buffer('zero', torch.tensor(0.), persistent=False)
        all_codes = torch.arange(codebook_size)
        bits = ((all_codes[..., None].int() & self.mask) != 0).float()
        self.register_buffer('codebook', self.bits_to_codes(bits), persistent=False)

    def bits_to_codes(self, bits):
        return bits * self.codebook_scale * 2 - self.codebook_scale

    @property
    def dtype(self):
        return self.codebook.dtype

    def indices_to_codes(self, indices, project_out=True):
        """indices -> {-1,+1} codes (-> project_out).  Tiny host-visible helper (LFQ.py:152-181); torch ops."""
        is_img_or_video = indices.ndim >= (3 + int(self.keep_num_codebooks_dim))
        if not self.keep_num_codebooks_dim:
            indices = indices.unsqueeze(-1)
        bits = ((indices[..., None].int() & self.mask) != 0).to(self.dtype)
        codes = self.bits_to_codes(bits).flatten(-2)
        if project_out:
            if self.has_projections and codes.is_cuda:
                # same arithmetic as the quantiser kernels (b + sum_i W[:, i] c_i as a chain of fused multiply-adds), so the code of an
                # index is bit-identical to the z_q the kernels emit for it (Anomaly_L1 compares the two, losses.py:147-168)
                w, acc = self.project_out.weight, self.project_out.bias.expand(*codes.shape[:-1], -1)
                for i in range(w.shape[1]):
                    acc = torch.addcmul(acc, codes[..., i:i + 1], w[:, i])
                codes = acc
            else:
                codes = self.project_out(codes)
        if is_img_or_video:
            codes = codes.movedim(-1, 1)
        return codes

    def forward_projected(self, s, inv_temperature=100., want_bf16=False):
        """s [...] = project_in(z) already evaluated by the producer (VQ_model folds project_in into the encoder's last conv)
        -> Return(quantized [..., dim], indices int64 [...], aux loss).  Same arithmetic as ``forward`` from s onwards.
        want_bf16: the kernel also writes a bf16 copy of the quantized tensor (``self.last_zq_bf16``)."""
        out = ops.LFQScalarFn.apply(s, self.project_out.weight, self.project_out.bias, self.training,
                                    float(inv_temperature), float(self.commitment_loss_weight),
                                    float(self.entropy_loss_weight), float(self.diversity_gamma), self.codebook_size, bool(want_bf16))
        zq, idx, aux, xq = out[:4]
        self.last_zq_bf16 = out[4] if want_bf16 else None
        self.last_scalar = xq
        if not self.training:
            aux = self.zero
        if self.keep_num_codebooks_dim:
            idx = idx.unsqueeze(-1)
        return Return(zq, idx, aux)

    def forward(self, x, inv_temperature=100., return_loss_breakdown=False, mask=None):
        """x [b, n, d] (or image/video [b, d, ...]) -> Return(quantized, indices int64, aux loss)."""
        if mask is not None:
            raise NotImplementedError("idee_b200: LFQ token masks are not built (unused by IDEE)")
        is_img_or_video = x.ndim >= 4
        if is_img_or_video:
            x = x.movedim(1, -1)
        assert x.shape[-1] == self.dim, f'expected dimension of {self.dim} but received {x.shape[-1]}'
        if self.codebook_dim > 1:
            # K-bit sign codes (codebook_size 4, 8, 16): distances to the 2^K codes, argmin, STE and the loss sums in one kernel each way
            if return_loss_breakdown:
                raise NotImplementedError("idee_b200: return_loss_breakdown is not built (unused by IDEE, build.py:151)")
            zq, idx, aux = ops.LFQGeneralFn.apply(x, self.project_in.weight, self.project_in.bias, self.project_out.weight,
                                                  self.project_out.bias, self.training, float(inv_temperature),
                                                  float(self.commitment_loss_weight), float(self.entropy_loss_weight),
                                                  float(self.diversity_gamma), self.codebook_dim)
            self.last_scalar = None
            if not self.training:
                aux = self.zero
            if is_img_or_video:
                zq = zq.movedim(-1, 1)
            if self.keep_num_codebooks_dim:
                idx = idx.unsqueeze(-1)
            return Return(zq, idx, aux)
        zq, idx, aux, xq = ops.LFQFn.apply(x, self.project_in.weight, self.project_in.bias, self.project_out.weight,
                                           self.project_out.bias, self.training, float(inv_temperature),
                                           float(self.commitment_loss_weight), float(self.entropy_loss_weight),
                                           float(self.diversity_gamma), self.codebook_size)
        # the quantised scalar x of the last call (z_q = x * w_out + b_out): lets a consumer use the rank-1 form of z_q
        self.last_scalar = xq
        if not self.training:
            aux = self.zero
        if is_img_or_video:
            zq = zq.movedim(-1, 1)
        if self.keep_num_codebooks_dim:
            idx = idx.unsqueeze(-1)
        if return_loss_breakdown:
            raise NotImplementedError("idee_b200: return_loss_breakdown is not built (unused by IDEE, build.py:151)")
        return Return(zq, idx, aux)
