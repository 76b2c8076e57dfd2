"""B200-native Video-Swin-3D encoder: drop-in for the reference ``models/encoder/Swin_3D.py``.

Same constructor arguments, module tree, parameter names/shapes (``state_dict`` compatible, and bit-identical
initialisation under the same torch seed because parameters are created in the reference's order) and the same
``forward(x[N,V,C,D,H,W]) -> [N,V,E,D,H,W]`` contract (Swin_3D.py:499-515, 616-636).  The compute is different: the
V per-variable encoders run as ONE batched sequence of fused CUDA kernels on channel-last tokens

    embed+LN  ->  SwinBlock x sum(depths)  ->  conv3^3(replicate)+ReLU  ->  conv3^3(replicate)

(libidee_b200.so via idee_b200.ops); window partition / cyclic shift / mask / bias gather are index math inside the
block kernel, so none of the reference's roll / permute / compute_mask tensors exist.  The returned tensor is a
permuted view of the channel-last buffer (logical shape and values as the reference, no copy).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib, ops


def get_window_size(x_size, window_size, shift_size=None):
    """Window/shift clamped to the tensor extent (Swin_3D.py:77-90)."""
    ws = [min(w, s) if s <= w else w for w, s in zip(window_size, x_size)]
    if shift_size is None:
        return tuple(ws)
    ss = [0 if s <= w else sh for w, s, sh in zip(window_size, x_size, shift_size)]
    return tuple(ws), tuple(ss)


def _win_tokens(blk, D, H, W) -> int:
    ws = get_window_size((D, H, W), blk.window_size)
    return ws[0] * ws[1] * ws[2]


def _relative_position_index(ws) -> torch.Tensor:
    """Pairwise relative-position index of the tokens of one window (Swin_3D.py:121-135)."""
    grid = torch.stack(torch.meshgrid(*[torch.arange(w) for w in ws], indexing="ij")).flatten(1)   # 3, N
    rel = grid[:, :, None] - grid[:, None, :]                                                      # 3, N, N
    mult = ((2 * ws[1] - 1) * (2 * ws[2] - 1), 2 * ws[2] - 1, 1)
    return sum((rel[a] + ws[a] - 1) * mult[a] for a in range(3))


class Mlp(nn.Module):
    """Parameter holder for fc1/fc2 (Swin_3D.py:24-42); the arithmetic is fused into the block kernel."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if drop != 0.:
            raise NotImplementedError("idee_b200: dropout inside the Swin block is not built (reference default 0)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("idee_b200: only exact-erf GELU is built")
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        raise NotImplementedError("Mlp is fused into SwinTransformerBlock3D in idee_b200")


class WindowAttention3D(nn.Module):
    """Parameter holder for W-MSA (Swin_3D.py:93-143); the arithmetic is fused into the block kernel."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("idee_b200: attention dropout is not built (reference default 0)")
        if not qkv_bias:
            raise NotImplementedError("idee_b200: qkv_bias=False is not built (reference default True)")
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        n_rel = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_rel, num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)

    def packed_parameters(self):
        return [self.relative_position_bias_table, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias]

    def forward(self, x, mask=None):
        raise NotImplementedError("WindowAttention3D is fused into SwinTransformerBlock3D in idee_b200")


class SwinTransformerBlock3D(nn.Module):
    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        if drop_path > 0.:
            raise NotImplementedError("idee_b200: stochastic depth is not built (reference default 0)")
        assert all(0 <= s < w for s, w in zip(shift_size, window_size)), "shift_size must in 0-window_size"
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size = tuple(window_size), tuple(shift_size)
        self.mlp_ratio, self.use_checkpoint = mlp_ratio, use_checkpoint   # the kernels always recompute in backward
        self.norm1 = norm_layer(dim, elementwise_affine=False)
        self.attn = WindowAttention3D(dim, self.window_size, num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim, elementwise_affine=False)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self._pack1 = None

    def packed_parameters(self):
        return self.attn.packed_parameters() + [self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias]

    def kernel_args(self, D, H, W):
        """(window, shift, rel_index[int32 G,G], rpb_rows, scale, heads, hidden) for a [.., D,H,W, C] token tensor."""
        ws, ss = get_window_size((D, H, W), self.window_size, self.shift_size)
        G = ws[0] * ws[1] * ws[2]
        idx = self.attn.relative_position_index[:G, :G].to(torch.int32).contiguous()     # Swin_3D.py:158-160
        return ws, ss, idx, self.attn.relative_position_bias_table.shape[0], self.attn.scale, self.num_heads, \
            int(self.dim * self.mlp_ratio)

    def forward(self, x, mask_matrix=None):
        """x [B,D,H,W,C] (one variable); ``mask_matrix`` is ignored: the shift mask is computed inside the kernel."""
        if self._pack1 is None:
            self._pack1 = ops.ParamPack([self.packed_parameters()])
        B, D, H, W, C = x.shape
        ws, ss, idx, rows, scale, heads, hidden = self.kernel_args(D, H, W)
        y = ops.swin_block(x.unsqueeze(1), self._pack1, idx, ws, ss, rows, scale, heads, hidden)
        return y.squeeze(1).to(x.dtype)      # the tcgen05 kernels keep tokens in bf16; the module contract returns x's dtype


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size=(2, 4, 4), in_chans=16, embed_dim=64, norm_layer=None):
        super().__init__()
        self.patch_size, self.in_chans, self.embed_dim = tuple(patch_size), in_chans, embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        self.norm = norm_layer(embed_dim, elementwise_affine=False) if norm_layer is not None else None
        if self.patch_size != (1, 1, 1):
            raise NotImplementedError("idee_b200: only patch_size (1,1,1) is built (reference default, config.py:53)")
        if self.norm is None:
            raise NotImplementedError("idee_b200: PatchEmbed3D without norm is not built (BasicLayer always passes LayerNorm)")
        self._packs = None

    def forward(self, x):
        """x [B,Cin,D,H,W] -> [B,E,D,H,W] (permuted view of channel-last tokens)."""
        if self._packs is None:
            self._packs = (ops.ParamPack([[self.proj.weight]]), ops.ParamPack([[self.proj.bias]]))
        tok = ops.embed_ln(x.unsqueeze(1), *self._packs)           # [B,1,D,H,W,E]
        return tok.squeeze(1).permute(0, 4, 1, 2, 3)


class BasicLayer(nn.Module):
    def __init__(self, in_dim, patch_size, dim, depth, num_heads, window_size=(4, 4, 4), mlp_ratio=4., qkv_bias=False,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None,
                 patch_norm=None, use_checkpoint=False):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth, self.use_checkpoint, self.dim, self.in_dim = depth, use_checkpoint, dim, in_dim
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim=dim, num_heads=num_heads, window_size=self.window_size,
                                   shift_size=(0, 0, 0) if i % 2 == 0 else self.shift_size, mlp_ratio=mlp_ratio,
                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                   drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                   norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        if in_dim != dim or tuple(patch_size) != (1, 1, 1):
            self.downsample = downsample(patch_size=patch_size, in_chans=in_dim, embed_dim=dim, norm_layer=nn.LayerNorm)
        else:
            self.downsample = None

    def forward(self, x):
        """x [B,C,D,H,W] -> [B,C',D,H,W] for ONE variable (Swin_3D.py:422-446)."""
        if self.downsample is not None:
            x = self.downsample(x)
        x = x.permute(0, 2, 3, 4, 1)
        for blk in self.blocks:
            x = blk(x)
        return x.permute(0, 4, 1, 2, 3)


class Swin_3D(nn.Module):
    def __init__(self, in_vars: int = 6, in_chans: int = 1, embed_dim: list = None, window_size: list = None,
                 depths: list = None, num_heads: list = None, mlp_ratio: int = 4, drop_rate: float = 0.,
                 attn_drop_rate: float = 0., drop_path_rate: float = 0., qkv_bias: bool = True, qk_scale: float = None,
                 patch_size: tuple = (1, 1, 1), patch_norm: bool = False, use_checkpoint: bool = False):
        super().__init__()
        self.in_vars, self.in_chans = in_vars, in_chans
        self.embed_dim = list(embed_dim) if embed_dim is not None else [16, 16]
        self.num_layers = len(self.embed_dim)
        self.window_size = [tuple(w) for w in window_size] if window_size is not None else [(2, 4, 4), (8, 1, 1)]
        self.depths = list(depths) if depths is not None else [2, 1]
        self.num_heads = list(num_heads) if num_heads is not None else [2, 2]
        self.mlp_ratio, self.drop_rate, self.attn_drop_rate, self.drop_path_rate = mlp_ratio, drop_rate, attn_drop_rate, drop_path_rate
        self.qkv_bias, self.qk_scale, self.patch_size = qkv_bias, qk_scale, tuple(patch_size)
        self.patch_norm, self.use_checkpoint = patch_norm, use_checkpoint
        if any(d != 16 for d in self.embed_dim) or any(h != 2 for h in self.num_heads) or int(16 * mlp_ratio) != 64:
            raise NotImplementedError("idee_b200: kernels are built for embed_dim 16, 2 heads, mlp_ratio 4 (config.py:51-66)")
        dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, sum(self.depths))]
        self.layers_var, self.proj_var = nn.ModuleList(), nn.ModuleList()
        E = self.embed_dim[-1]
        for _ in range(self.in_vars):
            layers = nn.ModuleList()
            for i in range(self.num_layers):
                layers.append(BasicLayer(
                    in_dim=self.embed_dim[i - 1] if i > 0 else self.in_chans,
                    patch_size=self.patch_size if i == 0 else (1, 1, 1), dim=self.embed_dim[i], depth=self.depths[i],
                    num_heads=self.num_heads[i], window_size=self.window_size[i], mlp_ratio=self.mlp_ratio,
                    qkv_bias=self.qkv_bias, qk_scale=self.qk_scale, drop=self.drop_rate, attn_drop=self.attn_drop_rate,
                    drop_path=dpr[sum(self.depths[:i]):sum(self.depths[:i + 1])], norm_layer=nn.LayerNorm,
                    downsample=PatchEmbed3D, patch_norm=nn.LayerNorm if self.patch_norm and i == 0 else None,
                    use_checkpoint=self.use_checkpoint))
            self.layers_var.append(layers)
            self.proj_var.append(nn.Sequential(
                nn.Conv3d(E, E, kernel_size=3, stride=1, padding=1, padding_mode='replicate', bias=True), nn.ReLU(),
                nn.Conv3d(E, E, kernel_size=3, stride=1, padding=1, padding_mode='replicate', bias=True)))
        self.init_weights()
        self._packs = None

    def init_weights(self):
        """trunc_normal(.02) on every Linear/Conv3d weight, zero biases (Swin_3D.py:596-613)."""
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv3d)):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    # ---- packed (variable-batched) view of the parameters ----
    def _build_packs(self):
        V = self.in_vars
        first = self.layers_var[0][0]
        if first.downsample is None:
            raise NotImplementedError("idee_b200: the first stage must embed in_chans -> 16 channels (in_chans != 16)")
        packs = {"embed_w": ops.ParamPack([[self.layers_var[v][0].downsample.proj.weight] for v in range(V)]),
                 "embed_b": ops.ParamPack([[self.layers_var[v][0].downsample.proj.bias] for v in range(V)]),
                 "blocks": []}
        for l in range(self.num_layers):
            if l > 0 and self.layers_var[0][l].downsample is not None:
                raise NotImplementedError("idee_b200: a channel-changing downsample after stage 1 is not built")
            for b in range(self.depths[l]):
                packs["blocks"].append((l, b, ops.ParamPack([self.layers_var[v][l].blocks[b].packed_parameters() for v in range(V)])))
        for i in (0, 2):
            packs[f"proj{i}_w"] = ops.ParamPack([[self.proj_var[v][i].weight] for v in range(V)])
            packs[f"proj{i}_b"] = ops.ParamPack([[self.proj_var[v][i].bias] for v in range(V)])
        self._packs = packs

    def packs(self):
        """Every ParamPack of the variable-batched path (for model-wide flat parameter / gradient buffers)."""
        if self._packs is None:
            self._build_packs()
        pk = self._packs
        return [pk["embed_w"], pk["embed_b"]] + [p for _, _, p in pk["blocks"]] + \
            [pk["proj0_w"], pk["proj0_b"], pk["proj2_w"], pk["proj2_b"]]

    def forward_tokens(self, x: torch.Tensor, fold_last=None) -> torch.Tensor:
        """x [N,V,C,D,H,W] -> channel-last encoder output [N,V,D,H,W,E] (contiguous).

        fold_last = (w_in [1,E], b_in [1]) (bf16 mode only): the caller applies Linear(E -> 1) to the encoder output and nothing
        else reads it (VQ_model: LFQ.project_in), so the last proj conv and that Linear are evaluated as ONE 16 -> 1 conv with
        weights sum_o w_in[o] W2[o, c, tap] and bias w_in . b2 + b_in; returns s [N,V,D,H,W].  The folded weights are built with
        differentiable torch ops, so gradients reach proj_var[2].weight / bias and w_in / b_in through autograd."""
        if x.dim() == 5 and self.in_chans == 1:
            x = x.unsqueeze(2)
        if self._packs is None:
            self._build_packs()
        N, V, Cin, D, H, W = x.shape
        assert V == self.in_vars and Cin == self.in_chans, "input must be [N, in_vars, in_chans, D, H, W]"
        pk = self._packs
        E = self.embed_dim[-1]
        # bf16 mode, in_chans == 1: the patch embedding (1x1x1 conv + LayerNorm) is evaluated inside the first block's kernels
        fuse_embed = _lib.PRECISION == "bf16" and Cin == 1 and E == 16 and len(pk["blocks"]) > 1
        tok = None if fuse_embed else ops.embed_ln(x, pk["embed_w"], pk["embed_b"])
        # bf16 mode: the tensor-core proj convs round their operands to bf16 when they load them, so the tensors only they
        # consume (the last block's output and the hidden ReLU activation, plus the hidden gradient in backward) are kept in
        # HBM as bf16 -- bit-identical results, half the traffic, and no conversion pass in the conv kernels.
        bf16_io = _lib.PRECISION == "bf16" and E == 16
        tok16 = None
        if _lib.swin_umma() and bf16_io and all(_win_tokens(self.layers_var[0][l].blocks[b], D, H, W) >= 8 for l, b, _ in pk["blocks"]):
            # tcgen05 / TMEM kernels: every block in one autograd node (fp32 residual stream, bf16 saved activations / gradients)
            specs = []
            for l, b, pack in pk["blocks"]:
                ws, ss, idx, rows, scale, heads, hidden = self.layers_var[0][l].blocks[b].kernel_args(D, H, W)
                specs.append((pack, idx, tuple(ws), tuple(ss), rows, float(scale), heads, hidden))
            if fuse_embed:
                xin = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
                tok = ops.swin_stack(xin, specs, embed=(pk["embed_w"], pk["embed_b"]))
            else:
                tok = ops.swin_stack(tok, specs)
            return self._proj(tok, None, pk, V, E, bf16_io, fold_last)
        for i, (l, b, pack) in enumerate(pk["blocks"]):
            blk = self.layers_var[0][l].blocks[b]
            ws, ss, idx, rows, scale, heads, hidden = blk.kernel_args(D, H, W)
            if fuse_embed and i == 0:
                xin = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
                tok = ops.swin_block_embed(xin, pk["embed_w"], pk["embed_b"], pack, idx, ws, ss, rows, scale, heads, hidden)
            elif bf16_io and i == len(pk["blocks"]) - 1:
                # the proj conv reads only the bf16 copy: the fp32 tokens of the last block are never written
                tok, tok16 = ops.swin_block(tok, pack, idx, ws, ss, rows, scale, heads, hidden, want_bf16="only")
            else:
                tok = ops.swin_block(tok, pack, idx, ws, ss, rows, scale, heads, hidden)
        return self._proj(tok, tok16, pk, V, E, bf16_io, fold_last)

    def _proj(self, tok, tok16, pk, V, E, bf16_io, fold_last):
        """proj_var: conv3^3(replicate) + ReLU + conv3^3(replicate) on the channel-last tokens (Swin_3D.py:586-592, 631)."""
        w0, b0 = ops.packed(pk["proj0_w"], (V, E, E, 3, 3, 3)), ops.packed(pk["proj0_b"], (V, E))
        w2, b2 = ops.packed(pk["proj2_w"], (V, E, E, 3, 3, 3)), ops.packed(pk["proj2_b"], (V, E))
        # conv -> ReLU -> conv: the second conv is the only consumer of the ReLU output, so its data-gradient epilogue applies
        # the ReLU backward mask and the first conv skips the separate pass
        if tok.dtype == torch.bfloat16 and not bf16_io:
            tok = tok.float()
        tok = ops.conv3d_cl(tok, w0, b0, proj=True, relu=True, consumer_masks=True, x16=tok16, out_bf16=bf16_io)
        if fold_last is not None and bf16_io:
            w_in, b_in = fold_last
            wf = torch.einsum('vocthw,o->vcthw', w2, w_in.reshape(-1)).unsqueeze(1)          # [V,1,E,3,3,3]
            bf = (b2 @ w_in.reshape(-1) + b_in.reshape(())).unsqueeze(1)                        # [V,1]
            s = ops.conv3d_cl(tok, wf, bf, proj=True, relu=False, input_is_relu=True)           # [N,V,D,H,W,1]
            return s.squeeze(-1)
        tok = ops.conv3d_cl(tok, w2, b2, proj=True, relu=False, input_is_relu=True)
        return tok

    def forward(self, x):
        """x [N,V,C,D,H,W] -> [N,V,E,D,H,W] (a permuted view of the channel-last result)."""
        return self.forward_tokens(x).permute(0, 1, 5, 2, 3, 4)
