"""CPU oracle for the IDEE hot path (Swin_3D encoder -> LFQ quantiser -> CNN_3D classifier + losses).

TEST INFRASTRUCTURE ONLY.  Nothing under ``idee_b200/`` may import this file; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do,
and there only as the checker / the timed CPU baseline, never as the product path.

This is an independent *restatement* of the reference algorithm in plain fp32 PyTorch CPU ops, written
as pure functions over a ``state_dict`` that uses the reference's own parameter names.  Every function
cites the reference ``file:line`` it follows (paths relative to ``/root/reference``).

Parity pinning: the reference repo has no tests and no golden vectors.  The oracle
is therefore pinned against outputs of the reference itself, imported read-only in the build container by
``tests/golden/make_golden.py`` (with an in-memory ``timm`` shim); the resulting fixtures are committed
under ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file against them (outputs,
losses and every parameter gradient).  Because the restatement is written in differentiable torch ops,
autograd over it is the gradient oracle for the hand-written CUDA backward kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration (the hot-path fields of config.py:50-132)
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    encoder: str = "Swin_3D"              # config.py:40  'Swin_3D' | 'CNN_3D'
    in_vars: int = 6                      # config.py:49  in_channels_dynamic
    in_chans: int = 1                     # config.py:50  (1 synthetic, 2 real)
    embed_dim: Sequence[int] = (16, 16)   # config.py:51
    depths: Sequence[int] = (2, 1)        # config.py:52
    patch_size: Tuple[int, int, int] = (1, 1, 1)            # config.py:53
    window_size: Sequence[Tuple[int, int, int]] = ((2, 4, 4), (8, 1, 1))  # config.py:55
    mlp_ratio: float = 4.0                # config.py:56
    num_heads: Sequence[int] = (2, 2)     # config.py:66
    qk_scale: float | None = None         # config.py:70
    codebook_size: int = 2                # config.py:81
    codebook_dim: int = 16                # config.py:82
    cls_dim: int = 16                     # config.py:85
    lambda_commitment: float = 3.0        # config.py:129
    lambda_anomaly: float = 100.0         # config.py:130
    lambda_entropy: float = 0.1           # config.py:131
    diversity_gamma: float = 0.1          # config.py:132
    delta_t: int = 8                      # config.py:101
    # checker-only switch: evaluate the straight-through value as q + (s - s.detach()) (== q exactly) instead of the reference's
    # s + (q - s).detach() (== q up to one ulp, LFQ.py:226).  The one-ulp residue makes z_q differ from the code of index 0 at
    # rounding level for tokens quantised to code 0, and Anomaly_L1's |z_q - vq0| then has a +-1 gradient of arbitrary sign there
    # (losses.py:147-168): a rounding artefact of the reference, not part of its algorithm.  See DESIGN.md section 4.
    exact_ste: bool = False


# ----------------------------------------------------------------------------------------------
# Swin_3D helpers
# ----------------------------------------------------------------------------------------------
def get_window_size(x_size, window_size, shift_size=None):
    """Clamp window (and zero the shift) on axes where the tensor is not larger than the window.
    Swin_3D.py:77-90."""
    ws = list(window_size)
    ss = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            ws[i] = x_size[i]
            if ss is not None:
                ss[i] = 0
    if ss is None:
        return tuple(ws)
    return tuple(ws), tuple(ss)


def window_partition(x: Tensor, ws) -> Tensor:
    """[B,D,H,W,C] -> [B*nW, Wd*Wh*Ww, C]; window order (b,dw,hw,ww), token order (d,h,w).
    Swin_3D.py:45-57."""
    B, D, H, W, C = x.shape
    x = x.reshape(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws[0] * ws[1] * ws[2], C)


def window_reverse(win: Tensor, ws, B, D, H, W) -> Tensor:
    """Inverse of window_partition.  Swin_3D.py:60-74."""
    x = win.reshape(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, D, H, W, -1)


def relative_position_index(ws) -> Tensor:
    """[N,N] int64 index into the (2Wd-1)(2Wh-1)(2Ww-1) bias table.  Swin_3D.py:121-135
    (meshgrid default indexing == 'ij')."""
    cd, ch, cw = torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2])
    coords = torch.stack(torch.meshgrid(cd, ch, cw, indexing="ij")).flatten(1)  # 3,N
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 2] += ws[2] - 1
    rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
    rel[:, :, 1] *= (2 * ws[2] - 1)
    return rel.sum(-1)


def region_ids(S: int, ws: int, ss: int) -> Tensor:
    """Per-axis region id used by compute_mask (Swin_3D.py:343-347): the three slices
    [0,S-ws), [S-ws,S-ss), [S-ss,S); with ss==0 the last slice(-0,None) covers the whole axis."""
    r = torch.zeros(S, dtype=torch.long)
    if ss == 0:
        r[:] = 2
        return r
    r[S - ws:S - ss] = 1
    r[S - ss:] = 2
    return r


def compute_mask(Dp, Hp, Wp, ws, ss) -> Tensor:
    """[nW,N,N] additive mask, 0 where both tokens share a region else -100.  Swin_3D.py:340-352."""
    rd, rh, rw = region_ids(Dp, ws[0], ss[0]), region_ids(Hp, ws[1], ss[1]), region_ids(Wp, ws[2], ss[2])
    img = (rd[:, None, None] * 9 + rh[None, :, None] * 3 + rw[None, None, :]).float()
    mw = window_partition(img[None, ..., None], ws).squeeze(-1)  # nW,N
    diff = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def layer_norm_noaffine(x: Tensor) -> Tensor:
    """LayerNorm over the last dim, eps 1e-5, no affine.  Swin_3D.py:214,220,469."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5)


def window_attention(sd: Dict[str, Tensor], p: str, xw: Tensor, mask, num_heads: int, cfg_ws, qk_scale) -> Tensor:
    """WindowAttention3D.forward, Swin_3D.py:145-178.  xw [B_,N,C]."""
    B_, N, C = xw.shape
    hd = C // num_heads
    scale = qk_scale or hd ** -0.5
    qkv = xw @ sd[p + "qkv.weight"].t()
    if (p + "qkv.bias") in sd:
        qkv = qkv + sd[p + "qkv.bias"]
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index(cfg_ws)[:N, :N].reshape(-1).to(xw.device)  # :158-160 (slice of configured window)
    bias = sd[p + "relative_position_bias_table"][idx].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return out @ sd[p + "proj.weight"].t() + sd[p + "proj.bias"]


def swin_block(sd, p, x: Tensor, cfg_ws, cfg_ss, num_heads, qk_scale, layer_ss) -> Tensor:
    """SwinTransformerBlock3D.forward, Swin_3D.py:224-287.  x [B,D,H,W,C] channel-last.
    ``cfg_ss`` is the block's own shift, ``layer_ss`` the layer shift used to build the mask (:433-438)."""
    B, D, H, W, C = x.shape
    ws, ss = get_window_size((D, H, W), cfg_ws, cfg_ss)
    shortcut = x
    xn = layer_norm_noaffine(x)
    pd = (ws[0] - D % ws[0]) % ws[0]
    pb = (ws[1] - H % ws[1]) % ws[1]
    pr = (ws[2] - W % ws[2]) % ws[2]
    xn = F.pad(xn, (0, 0, 0, pr, 0, pb, 0, pd))                              # zeros AFTER LN1 (:233-238)
    _, Dp, Hp, Wp, _ = xn.shape
    if any(s > 0 for s in ss):
        xs = torch.roll(xn, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        _, mss = get_window_size((D, H, W), cfg_ws, layer_ss)
        mask = compute_mask(Dp, Hp, Wp, ws, mss).to(x.device)
    else:
        xs, mask = xn, None
    aw = window_attention(sd, p + "attn.", window_partition(xs, ws), mask, num_heads, cfg_ws, qk_scale)
    xs = window_reverse(aw, ws, B, Dp, Hp, Wp)
    if any(s > 0 for s in ss):
        xs = torch.roll(xs, shifts=(ss[0], ss[1], ss[2]), dims=(1, 2, 3))
    xs = xs[:, :D, :H, :W, :]
    y = shortcut + xs
    yn = layer_norm_noaffine(y)                                              # :264-265,285
    h = F.gelu(yn @ sd[p + "mlp.fc1.weight"].t() + sd[p + "mlp.fc1.bias"])   # exact erf GELU (:27,32)
    return y + (h @ sd[p + "mlp.fc2.weight"].t() + sd[p + "mlp.fc2.bias"])


def patch_embed(sd, p, x: Tensor, patch) -> Tensor:
    """PatchEmbed3D.forward (norm always on: BasicLayer passes norm_layer=nn.LayerNorm, Swin_3D.py:418),
    Swin_3D.py:473-491.  x [B,Cin,D,H,W] -> [B,C,D',H',W']."""
    _, _, D, H, W = x.shape
    if W % patch[2]:
        x = F.pad(x, (0, patch[2] - W % patch[2]))
    if H % patch[1]:
        x = F.pad(x, (0, 0, 0, patch[1] - H % patch[1]))
    if D % patch[0]:
        x = F.pad(x, (0, 0, 0, 0, 0, patch[0] - D % patch[0]))
    x = F.conv3d(x, sd[p + "proj.weight"], sd[p + "proj.bias"], stride=patch)
    B, C, D2, H2, W2 = x.shape
    x = layer_norm_noaffine(x.flatten(2).transpose(1, 2))
    return x.transpose(1, 2).reshape(B, C, D2, H2, W2)


def basic_layer(sd, p, x: Tensor, in_dim, dim, depth, heads, cfg_ws, patch, qk_scale) -> Tensor:
    """BasicLayer.forward, Swin_3D.py:422-446.  x [B,C,D,H,W]."""
    if in_dim != dim or tuple(patch) != (1, 1, 1):                           # :417-420
        x = patch_embed(sd, p + "downsample.", x, patch)
    x = x.permute(0, 2, 3, 4, 1)
    layer_ss = tuple(i // 2 for i in cfg_ws)                                 # :393
    for b in range(depth):
        ss = (0, 0, 0) if b % 2 == 0 else layer_ss                           # :405
        x = swin_block(sd, f"{p}blocks.{b}.", x, cfg_ws, ss, heads, qk_scale, layer_ss)
    return x.permute(0, 4, 1, 2, 3)


def swin3d_forward(sd: Dict[str, Tensor], x: Tensor, cfg: OracleConfig, prefix: str = "encoder.") -> Tensor:
    """Swin_3D.forward, Swin_3D.py:616-636.  x [N,V,C,D,H,W] -> [N,V,E,D,H,W]."""
    outs = []
    for v in range(cfg.in_vars):
        xv = x[:, v]
        for l in range(len(cfg.embed_dim)):
            xv = basic_layer(sd, f"{prefix}layers_var.{v}.{l}.", xv,
                             cfg.embed_dim[l - 1] if l > 0 else cfg.in_chans, cfg.embed_dim[l], cfg.depths[l],
                             cfg.num_heads[l], tuple(cfg.window_size[l]),
                             cfg.patch_size if l == 0 else (1, 1, 1), cfg.qk_scale)
        pv = f"{prefix}proj_var.{v}."
        xv = F.conv3d(F.pad(xv, (1,) * 6, mode="replicate"), sd[pv + "0.weight"], sd[pv + "0.bias"])  # :586-592
        xv = F.relu(xv)
        xv = F.conv3d(F.pad(xv, (1,) * 6, mode="replicate"), sd[pv + "2.weight"], sd[pv + "2.bias"])
        outs.append(xv.unsqueeze(1))
    return torch.cat(outs, dim=1)


def cnn3d_forward(sd: Dict[str, Tensor], x: Tensor, cfg: OracleConfig, prefix: str = "encoder.") -> Tensor:
    """encoder CNN_3D.forward, models/encoder/CNN_3D.py:214-237 with conv_block.forward (:119-147) and its PatchEmbed3D
    (:44-71: conv k=1 without bias + LayerNorm without affine).  x [N,V,C,D,H,W] -> [N,V,E,D,H,W]."""
    outs = []
    chans = [cfg.in_chans] + list(cfg.embed_dim[:-1])
    for v in range(cfg.in_vars):
        xv = x[:, v]
        for l, dim in enumerate(cfg.embed_dim):
            p = f"{prefix}layers_var.{v}.{l}."
            if chans[l] != dim:                                                           # :113-117
                xv = F.conv3d(xv, sd[p + "downsample.proj.weight"])
                B, C, D, H, W = xv.shape
                xv = layer_norm_noaffine(xv.flatten(2).transpose(1, 2)).transpose(1, 2).reshape(B, C, D, H, W)
            B, C, D, H, W = xv.shape
            for conv, norm in (("conv1", "norm1"), ("conv2", "norm2")):                   # :133-145
                y = F.conv3d(F.pad(xv, (1,) * 6, mode="replicate"), sd[p + conv + ".weight"])
                y = y.reshape(B, C, D * H * W).permute(0, 2, 1)
                y = F.layer_norm(y, (C,), sd[p + norm + ".weight"], sd[p + norm + ".bias"], eps=1e-5)
                xv = xv + F.relu(y.permute(0, 2, 1).reshape(B, C, D, H, W))
        pv = f"{prefix}proj_var.{v}."
        xv = F.conv3d(F.pad(xv, (1,) * 6, mode="replicate"), sd[pv + "0.weight"], sd[pv + "0.bias"])
        xv = F.relu(xv)
        xv = F.conv3d(F.pad(xv, (1,) * 6, mode="replicate"), sd[pv + "2.weight"], sd[pv + "2.bias"])
        outs.append(xv.unsqueeze(1))
    return torch.cat(outs, dim=1)


def encoder_forward(sd, x, cfg: OracleConfig) -> Tensor:
    return cnn3d_forward(sd, x, cfg) if cfg.encoder == "CNN_3D" else swin3d_forward(sd, x, cfg)


# ----------------------------------------------------------------------------------------------
# LFQ (codebook_size 2^k; IDEE uses k=1)
# ----------------------------------------------------------------------------------------------
def _entropy(p: Tensor) -> Tensor:
    """LFQ.py:52-56."""
    return (-p * p.clamp(min=1e-5).log()).sum(-1)


def lfq_forward(sd, z: Tensor, cfg: OracleConfig, training: bool, prefix: str = "vq.", inv_temperature: float = 100.0):
    """LFQ.forward, LFQ.py:183-307, for num_codebooks=1, codebook_scale=1, identity activation.
    z [b,n,d] -> (quantized [b,n,d], indices [b,n] int64, aux_loss scalar, breakdown)."""
    kbits = int(math.log2(cfg.codebook_size))
    z = z.float()
    has_proj = cfg.codebook_dim != kbits                                     # LFQ.py:98
    s = z @ sd[prefix + "project_in.weight"].t() + sd[prefix + "project_in.bias"] if has_proj else z
    q = torch.where(s > 0, torch.ones_like(s), -torch.ones_like(s))         # :221-222 (s==0 -> -1)
    if training:
        x = q + (s - s.detach()) if cfg.exact_ste else s + (q - s).detach()  # :226-230
    else:
        x = q
    bitmask = 2 ** torch.arange(kbits - 1, -1, -1, device=z.device)          # :134
    indices = ((x > 0).int() * bitmask.int()).sum(-1).long()                 # :234
    zero = torch.zeros((), device=z.device)
    if training:
        codes = torch.arange(cfg.codebook_size, device=z.device)
        codebook = ((codes[:, None].int() & bitmask.int()) != 0).float() * 2 - 1   # :139-146
        dist = -2 * torch.einsum("bnd,jd->bnj", s, codebook)                 # :239
        prob = (-dist * inv_temperature).softmax(-1).reshape(-1, cfg.codebook_size)  # :240,246
        per_sample_entropy = _entropy(prob).mean()                           # :258
        codebook_entropy = _entropy(prob.mean(0))                            # :260-261
        ent_aux = cfg.lambda_entropy * per_sample_entropy - cfg.diversity_gamma * codebook_entropy  # :262
        commit = ((s - q.detach()) ** 2).mean()                              # :270-274
    else:
        per_sample_entropy = codebook_entropy = ent_aux = commit = zero      # :264-266,275-276
    zq = x @ sd[prefix + "project_out.weight"].t() + sd[prefix + "project_out.bias"] if has_proj else x  # :284
    aux = commit * cfg.lambda_commitment + ent_aux                           # :300
    return zq, indices, aux, (per_sample_entropy, codebook_entropy, commit)


def lfq_indices_to_codes(sd, indices: Tensor, cfg: OracleConfig, prefix: str = "vq.") -> Tensor:
    """LFQ.indices_to_codes(project_out=True) for <3-D indices, LFQ.py:152-181."""
    kbits = int(math.log2(cfg.codebook_size))
    bitmask = 2 ** torch.arange(kbits - 1, -1, -1, device=indices.device)
    bits = ((indices[..., None].int() & bitmask.int()) != 0).float()
    codes = bits * 2 - 1
    return codes @ sd[prefix + "project_out.weight"].t() + sd[prefix + "project_out.bias"]


# ----------------------------------------------------------------------------------------------
# classifier (models/classifier/CNN_3D.py)
# ----------------------------------------------------------------------------------------------
def _head(sd, p, x: Tensor) -> Tensor:
    """conv(2,3,3)/s(2,1,1)/p(0,1,1) x3 with ReLU between; Dropout(p=0) omitted.
    classifier/CNN_3D.py:36-38,51-57 and :83-85,131-137."""
    for i, name in enumerate(("conv1", "conv2", "conv3")):
        x = F.conv3d(x, sd[p + name + ".weight"], sd[p + name + ".bias"], stride=(2, 1, 1), padding=(0, 1, 1))
        if i < 2:
            x = F.relu(x)
    return x.squeeze(2)


def classifier_forward(sd, zq: Tensor, cfg: OracleConfig, prefix: str = "cls."):
    """CNN_3D.forward, classifier/CNN_3D.py:112-139.  zq [N,V,C,T,H,W] -> (z [N,1,H,W], [y_v])."""
    N, V, C, T, H, W = zq.shape
    ys = [_head(sd, f"{prefix}layers.{i}.", zq[:, i]) for i in range(V)]
    z = _head(sd, prefix, zq.reshape(N, V * C, T, H, W))
    return z, ys


# ----------------------------------------------------------------------------------------------
# full model + losses + train-step loss
# ----------------------------------------------------------------------------------------------
def vq_model_forward(sd, x: Tensor, cfg: OracleConfig, training: bool):
    """VQ_model.forward, build.py:130-159."""
    z = encoder_forward(sd, x, cfg)
    N, V, C, T, H, W = z.shape
    zt = z.permute(0, 2, 1, 3, 4, 5).reshape(N, C, V * T * H * W).permute(0, 2, 1)      # :150
    zq, idx, aux, _ = lfq_forward(sd, zt, cfg, training)
    zq = zq.permute(0, 2, 1).reshape(N, C, V, T, H, W).permute(0, 2, 1, 3, 4, 5)        # :153
    anomaly = idx.view(N, V, T, H, W)                                                   # :154
    zc, ys = classifier_forward(sd, zq, cfg)
    return zc, ys, anomaly, zq, aux.unsqueeze(0), z


def bce_loss_synthetic(pred: Tensor, target: Tensor) -> Tensor:
    """BCE_loss_synthetic.forward, losses.py:105-124."""
    w = torch.histc(target, bins=2)
    w[torch.isinf(w)] = 1
    w = (w / w.sum()) ** -0.5
    w = torch.log(w + 1.1)
    w = w[target.long()]
    return (F.binary_cross_entropy_with_logits(pred, target, reduction="none") * w).mean()


def anomaly_l1_loss_synthetic(zq: Tensor, mask_extreme: Tensor, vq0: Tensor, snap: bool = False) -> Tensor:
    """Anomaly_L1_loss_synthetic.forward, losses.py:147-168, without materialising the three
    full-size broadcasts: target := pred where mask==1 (zero loss & grad), weight = 1-mask."""
    N, V, C, T, H, W = zq.shape
    m = mask_extreme.view(N, 1, 1, 1, H, W)
    tgt = vq0.reshape(1, 1, C, 1, 1, 1)
    d = zq - tgt
    if snap:
        # checker-only (OracleConfig.exact_ste): a token quantised to code 0 has z_q == vq0 mathematically, but z_q comes out of a
        # batched matmul and vq0 out of a one-row matmul (LFQ.py:152-181,284) whose roundings differ in the last bit; |d| at rounding
        # level then carries a +-1 gradient of arbitrary sign.  Snap such residues to the exact zero they stand for.
        d = torch.where(d.abs() <= 4e-7 * torch.maximum(zq.abs(), tgt.abs()), torch.zeros((), dtype=zq.dtype, device=zq.device), d)
    diff = torch.where(m == 1, torch.zeros((), dtype=zq.dtype, device=zq.device), d.abs())
    wsum = (1 - mask_extreme).sum() * (V * C * T)
    return (diff * (1 - m)).sum() / wsum


def train_step_loss(sd, x, mask_extreme, mask_extreme_loss, cfg: OracleConfig):
    """Loss assembly of one optimisation step, train_synthetic.py:175-201."""
    zc, ys, anomaly, zq, aux, z_enc = vq_model_forward(sd, x, cfg, training=True)
    tgt = mask_extreme.unsqueeze(1).float()
    loss = bce_loss_synthetic(zc, tgt)
    vq0 = lfq_indices_to_codes(sd, torch.tensor([0], device=x.device), cfg).detach()                     # :188-194
    loss_anom = anomaly_l1_loss_synthetic(zq, mask_extreme_loss.float(), vq0, snap=cfg.exact_ste)
    loss_var = sum(bce_loss_synthetic(y, tgt) for y in ys)
    total = loss + loss_anom * cfg.lambda_anomaly + loss_var + aux
    return total, dict(pred=zc, pred_y=ys, anomaly=anomaly, z_q=zq, loss_z_q=aux, z_enc=z_enc,
                       loss_bce=loss, loss_anomaly=loss_anom, loss_var=loss_var)


# ----------------------------------------------------------------------------------------------
# deterministic weights / inputs (shared by tests, golden generation and bench)
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameter names/shapes of models.build.VQ_model for the Swin_3D config, in state_dict order
    (SURVEY.md section 9 'Shapes at defaults'; checked against the reference by make_golden.py)."""
    shp: Dict[str, Tuple[int, ...]] = {}
    hid = lambda d: int(d * cfg.mlp_ratio)
    if cfg.encoder == "CNN_3D":                                    # models/encoder/CNN_3D.py:97-117, 176-183 (module order)
        chans = [cfg.in_chans] + list(cfg.embed_dim[:-1])
        for v in range(cfg.in_vars):
            for l, dim in enumerate(cfg.embed_dim):
                p = f"encoder.layers_var.{v}.{l}."
                shp[p + "conv1.weight"] = (dim, dim, 3, 3, 3)
                shp[p + "norm1.weight"] = (dim,)
                shp[p + "norm1.bias"] = (dim,)
                shp[p + "conv2.weight"] = (dim, dim, 3, 3, 3)
                shp[p + "norm2.weight"] = (dim,)
                shp[p + "norm2.bias"] = (dim,)
                if chans[l] != dim:
                    shp[p + "downsample.proj.weight"] = (dim, chans[l], 1, 1, 1)
    for v in range(cfg.in_vars if cfg.encoder != "CNN_3D" else 0):
        for l, dim in enumerate(cfg.embed_dim):
            p = f"encoder.layers_var.{v}.{l}."
            ws = cfg.window_size[l]
            for b in range(cfg.depths[l]):
                q = f"{p}blocks.{b}."
                shp[q + "attn.relative_position_bias_table"] = ((2 * ws[0] - 1) * (2 * ws[1] - 1) * (2 * ws[2] - 1), cfg.num_heads[l])
                shp[q + "attn.qkv.weight"] = (3 * dim, dim)
                shp[q + "attn.qkv.bias"] = (3 * dim,)
                shp[q + "attn.proj.weight"] = (dim, dim)
                shp[q + "attn.proj.bias"] = (dim,)
                shp[q + "mlp.fc1.weight"] = (hid(dim), dim)
                shp[q + "mlp.fc1.bias"] = (hid(dim),)
                shp[q + "mlp.fc2.weight"] = (dim, hid(dim))
                shp[q + "mlp.fc2.bias"] = (dim,)
            in_dim = cfg.embed_dim[l - 1] if l > 0 else cfg.in_chans
            patch = cfg.patch_size if l == 0 else (1, 1, 1)
            if in_dim != dim or tuple(patch) != (1, 1, 1):
                shp[p + "downsample.proj.weight"] = (dim, in_dim) + tuple(patch)
                shp[p + "downsample.proj.bias"] = (dim,)
    E = cfg.embed_dim[-1]
    for v in range(cfg.in_vars):
        for i in (0, 2):
            shp[f"encoder.proj_var.{v}.{i}.weight"] = (E, E, 3, 3, 3)
            shp[f"encoder.proj_var.{v}.{i}.bias"] = (E,)
    V, cd, d = cfg.in_vars, cfg.codebook_dim, cfg.cls_dim
    for name, ci, co in (("conv1", V * cd, V * d), ("conv2", V * d, V * d), ("conv3", V * d, 1)):
        shp[f"cls.{name}.weight"] = (co, ci, 2, 3, 3)
        shp[f"cls.{name}.bias"] = (co,)
    for v in range(V):
        for name, ci, co in (("conv1", cd, d), ("conv2", d, d), ("conv3", d, 1)):
            shp[f"cls.layers.{v}.{name}.weight"] = (co, ci, 2, 3, 3)
            shp[f"cls.layers.{v}.{name}.bias"] = (co,)
    kbits = int(math.log2(cfg.codebook_size))
    shp["vq.project_in.weight"] = (kbits, cd)
    shp["vq.project_in.bias"] = (kbits,)
    shp["vq.project_out.weight"] = (cd, kbits)
    shp["vq.project_out.bias"] = (cd,)
    return shp


def make_state_dict(cfg: OracleConfig, seed: int = 0, kind: str = "random") -> Dict[str, Tensor]:
    """Deterministic parameters.
    kind='reference' mimics VQ_model._init_weights (build.py:96-118): weights N(0.02,0.02), biases 0,
      rpb tables trunc_normal(.02) -- NOT bit-identical to the reference RNG stream, same distribution;
    kind='random' draws zero-mean weights with non-zero biases so that every term of every kernel is
      exercised and the quantiser mask is ~50/50 (default init gives mask == 1)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if kind == "reference":
            if name.endswith("relative_position_bias_table"):
                t = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
            elif name.endswith("bias"):
                t = torch.zeros(shape)
            else:
                t = 0.02 + 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if name.endswith("relative_position_bias_table"):
                t = torch.randn(shape, generator=g) * 0.5
            elif name.endswith("bias"):
                t = torch.randn(shape, generator=g) * 0.1
            else:
                t = torch.randn(shape, generator=g) / math.sqrt(max(fan_in, 1))
        sd[name] = t.float()
    return sd


def make_inputs(cfg: OracleConfig, N: int, T: int, H: int, W: int, seed: int = 0):
    """Synthetic-CERRA-shaped inputs: z-scored cube clipped to [-10,10]
    (Synthetic_dataset.py:206-215), binary extreme masks (:342-349)."""
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(N, cfg.in_vars, cfg.in_chans, T, H, W, generator=g).clamp_(-10, 10)
    mask_extreme = (torch.rand(N, H, W, generator=g) < 0.05).float()
    mask_extreme_loss = (torch.rand(N, H, W, generator=g) < 0.10).float()
    return x, mask_extreme, mask_extreme_loss
