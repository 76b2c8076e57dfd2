"""torch.autograd.Function wrappers over the C ABI (one per fused op) + parameter packing.

Everything here is plumbing: pointer/stride extraction, workspace allocation through torch's caching allocator and
autograd routing.  All arithmetic happens in libidee_b200.so; there is no PyTorch fallback for any op.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib as L


# ----------------------------------------------------------------------------------------------------------------
# parameter packing: the kernels take per-variable weight blocks [V][P]; the modules keep the reference's individual
# nn.Parameters (same state_dict).  A ParamPack owns one flat buffer and re-points each parameter's .data at a view
# of it, so no per-step gather/scatter is needed and in-place optimiser updates / load_state_dict stay visible.
# ----------------------------------------------------------------------------------------------------------------
class ParamPack:
    def __init__(self, groups: Sequence[Sequence[torch.nn.Parameter]]):
        self.groups = [list(g) for g in groups]
        self.sizes = [p.numel() for p in self.groups[0]]
        self.shapes = [tuple(p.shape) for p in self.groups[0]]
        for g in self.groups:
            assert [tuple(p.shape) for p in g] == self.shapes, "all variables must share parameter shapes"
        self.P = sum(self.sizes)
        self.V = len(self.groups)
        self.flat = None
        self.grad_flat = None     # optional [V,P] destination for the op's weight gradient (a slice of a model-wide buffer)

    def grad_out(self, like: torch.Tensor) -> torch.Tensor:
        """Where the backward kernel writes this pack's gradient: the bound destination if any, else a fresh buffer."""
        if self.grad_flat is not None and self.grad_flat.device == like.device:
            return self.grad_flat
        return torch.empty(self.V, self.P, device=like.device, dtype=torch.float32)

    def params(self) -> List[torch.nn.Parameter]:
        return [p for g in self.groups for p in g]

    def _aliased(self) -> bool:
        base = self.flat.data_ptr()
        for v, g in enumerate(self.groups):
            off = 0
            for p, n in zip(g, self.sizes):
                if p.data_ptr() != base + 4 * (v * self.P + off) or not p.is_contiguous():
                    return False
                off += n
        return True

    @torch.no_grad()
    def tensor(self) -> torch.Tensor:
        p0 = self.groups[0][0]
        if self.flat is not None and self.flat.device == p0.device and self._aliased():
            return self.flat
        return self.bind(torch.empty(self.V, self.P, device=p0.device, dtype=torch.float32))

    @torch.no_grad()
    def bind(self, flat: torch.Tensor) -> torch.Tensor:
        """Move the parameters into ``flat`` ([V,P] fp32, possibly a slice of a model-wide buffer) and alias them to it."""
        p0 = self.groups[0][0]
        if p0.dtype != torch.float32:
            raise RuntimeError("idee_b200 keeps fp32 master parameters; got %s" % p0.dtype)
        assert tuple(flat.shape) == (self.V, self.P) and flat.is_contiguous() and flat.dtype == torch.float32
        for v, g in enumerate(self.groups):
            off = 0
            for p, n in zip(g, self.sizes):
                view = flat[v, off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                off += n
        self.flat = flat
        return flat

    def split_grad(self, gflat: torch.Tensor):
        out = []
        for v in range(self.V):
            off = 0
            for n, shp in zip(self.sizes, self.shapes):
                out.append(gflat[v, off:off + n].view(shp))
                off += n
        return out


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------
# PatchEmbed + LN
# ----------------------------------------------------------------------------------------------------------------
class EmbedLN(torch.autograd.Function):
    """x [N,V,Cin,T,H,W] (any strides) -> tokens [N,V,T,H,W,16].  Swin_3D.py:473-491."""

    @staticmethod
    def forward(ctx, x, wpack: ParamPack, bpack: ParamPack, *params):
        L.require_cuda(x)
        lib = L.load()
        x = x if x.dtype == torch.float32 else x.float()
        N, V, Cin, T, H, W = x.shape
        w, b = wpack.tensor(), bpack.tensor()
        L.require_cuda(w, b)
        y = torch.empty(N, V, T, H, W, 16, device=x.device, dtype=torch.float32)
        xs = (C.c_int64 * 6)(*x.stride())
        L.run("embed_ln_fwd", lib.idee_embed_ln_fwd, x.data_ptr(), C.cast(xs, C.c_void_p), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                      N, V, Cin, T, H, W, 16, L.stream())
        ctx.save_for_backward(x)
        ctx.packs = (wpack, bpack)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.load()
        (x,) = ctx.saved_tensors
        wpack, bpack = ctx.packs
        N, V, Cin, T, H, W = x.shape
        w, b = wpack.tensor(), bpack.tensor()
        gy = _f32c(gy)
        gw, gb = wpack.grad_out(w), bpack.grad_out(b)
        nws = lib.idee_embed_ln_bwd_workspace_bytes(V)
        ws = L.workspace(nws, x.device)
        xs = (C.c_int64 * 6)(*x.stride())
        L.run("embed_ln_bwd", lib.idee_embed_ln_bwd, x.data_ptr(), C.cast(xs, C.c_void_p), w.data_ptr(), b.data_ptr(), gy.data_ptr(),
                                      gw.data_ptr(), gb.data_ptr(), N, V, Cin, T, H, W, 16, ws.data_ptr(), nws, L.stream())
        return (None, None, None, *wpack.split_grad(gw), *bpack.split_grad(gb))


def embed_ln(x, wpack: ParamPack, bpack: ParamPack):
    return EmbedLN.apply(x, wpack, bpack, *wpack.params(), *bpack.params())


# ----------------------------------------------------------------------------------------------------------------
# Swin block
# ----------------------------------------------------------------------------------------------------------------
def _swin_desc(x, window, shift, rpb_rows, scale, pack: ParamPack, heads, hidden):
    N, V, T, H, W, Cc = x.shape
    d = L.SwinDesc()
    d.N, d.V, d.T, d.H, d.W, d.C, d.heads, d.hidden = N, V, T, H, W, Cc, heads, hidden
    d.wd, d.wh, d.ww = window
    d.st, d.sh, d.sw = shift
    d.rpb_rows, d.scale, d.param_stride = rpb_rows, scale, pack.P
    d.precision = 1 if L.PRECISION == "bf16" else 0
    d.act_dtype, d.x_dtype, d.out_dtype = 0, 0, 0
    return d


def _umma_desc(shape_like, x_dtype, out_dtype, window, shift, rpb_rows, scale, pack, heads, hidden):
    """descriptor of the tcgen05 kernels: bf16 saved activation / gradients, block input / output fp32 or bf16"""
    d = _swin_desc(shape_like, window, shift, rpb_rows, scale, pack, heads, hidden)
    d.precision, d.act_dtype = 1, 1
    d.x_dtype, d.out_dtype = int(x_dtype == torch.bfloat16), int(out_dtype == torch.bfloat16)
    return d


def _bf16c(t):
    return t.contiguous() if t.dtype == torch.bfloat16 else t.to(torch.bfloat16).contiguous()


class SwinBlock(torch.autograd.Function):
    """tokens [N,V,T,H,W,16] -> same; one launch for all V variables.  Swin_3D.py:224-287."""

    @staticmethod
    def forward(ctx, x, pack: ParamPack, rel_index, window, shift, rpb_rows, scale, heads, hidden, want_bf16, *params):
        """want_bf16 (bf16 mode only): also return a non-differentiable bf16 copy of the output, written by the same kernel."""
        L.require_cuda(x)
        lib = L.load()
        flat = pack.tensor()
        L.require_cuda(flat, rel_index)
        if lib.idee_swin_block_packed_floats(rpb_rows) != pack.P:
            raise RuntimeError("swin_block: packed parameter size mismatch")
        if L.swin_umma() and window[0] * window[1] * window[2] >= 8:
            # tcgen05 / TMEM kernels: the residual stream stays in the caller's dtype (fp32, or bf16 with L.STREAM_BF16 / for the
            # last block whose only consumer rounds to bf16 anyway); the saved mid residual and all token gradients are bf16
            in_dtype = x.dtype
            x = x.contiguous() if x.dtype in (torch.float32, torch.bfloat16) else x.float().contiguous()
            out_dtype = torch.bfloat16 if (want_bf16 == "only" or x.dtype == torch.bfloat16) else torch.float32
            d = _umma_desc(x, x.dtype, out_dtype, window, shift, rpb_rows, scale, pack, heads, hidden)
            out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
            need_bwd = any(ctx.needs_input_grad)
            ymid = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if need_bwd else None
            L.run("swin_block_fwd", lib.idee_swin_block_fwd, C.byref(d), x.data_ptr(), out.data_ptr(), L.ptr(ymid), None,
                  flat.data_ptr(), rel_index.data_ptr(), L.stream(), tag=f"w{window} s{shift}")
            if need_bwd:
                ctx.save_for_backward(x, ymid, rel_index)
                ctx.pack, ctx.args = pack, (window, shift, rpb_rows, scale, heads, hidden)
                ctx.in_dtype, ctx.umma = in_dtype, True
            return out
        x = _f32c(x)
        d = _swin_desc(x, window, shift, rpb_rows, scale, pack, heads, hidden)
        # want_bf16 == "only": the caller reads nothing but the bf16 copy, so the fp32 result is not written; the returned fp32
        # tensor only routes the gradient and is poisoned with NaN so that an accidental consumer cannot read stale memory
        out = torch.full_like(x, float("nan")) if want_bf16 == "only" else torch.empty_like(x)
        need_bwd = any(ctx.needs_input_grad)
        ymid = torch.empty_like(x) if need_bwd else None
        out16 = torch.empty_like(x, dtype=torch.bfloat16) if want_bf16 else None
        L.run("swin_block_fwd", lib.idee_swin_block_fwd, C.byref(d), x.data_ptr(), None if want_bf16 == "only" else out.data_ptr(),
              L.ptr(ymid), L.ptr(out16),
              flat.data_ptr(), rel_index.data_ptr(), L.stream(), tag=f"w{window} s{shift}")
        if need_bwd:
            ctx.save_for_backward(x, ymid, rel_index)
            ctx.pack, ctx.args = pack, (window, shift, rpb_rows, scale, heads, hidden)
        if want_bf16:
            ctx.mark_non_differentiable(out16)
            ctx.set_materialize_grads(False)     # no 0.5 GB zero tensor for the non-differentiable bf16 copy in every backward
            return out, out16
        return out

    @staticmethod
    def backward(ctx, gout, *_unused):
        lib = L.load()
        x, ymid, rel_index = ctx.saved_tensors
        if gout is None:
            gout = torch.zeros_like(x)
        pack = ctx.pack
        window, shift, rpb_rows, scale, heads, hidden = ctx.args
        flat = pack.tensor()
        d = _swin_desc(x, window, shift, rpb_rows, scale, pack, heads, hidden)
        if getattr(ctx, "umma", False):        # tcgen05 kernels: bf16 token gradients in and out
            d = _umma_desc(x, x.dtype, x.dtype, window, shift, rpb_rows, scale, pack, heads, hidden)
            gout = _bf16c(gout)
            gx = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
            gflat = pack.grad_out(flat)
            nws = lib.idee_swin_block_bwd_workspace_bytes(C.byref(d))
            ws = L.workspace(nws, x.device)
            L.run("swin_block_bwd", lib.idee_swin_block_bwd, C.byref(d), x.data_ptr(), ymid.data_ptr(), gout.data_ptr(), gx.data_ptr(),
                  flat.data_ptr(), rel_index.data_ptr(), gflat.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=f"w{window} s{shift}")
            if ctx.in_dtype != torch.bfloat16:
                gx = gx.to(ctx.in_dtype)
            return (gx, None, None, None, None, None, None, None, None, None, *pack.split_grad(gflat))
        gout = _f32c(gout)
        gx = torch.empty_like(x)
        gflat = pack.grad_out(flat)
        nws = lib.idee_swin_block_bwd_workspace_bytes(C.byref(d))
        ws = L.workspace(nws, x.device)
        L.run("swin_block_bwd", lib.idee_swin_block_bwd, C.byref(d), x.data_ptr(), ymid.data_ptr(), gout.data_ptr(), gx.data_ptr(), flat.data_ptr(),
                                        rel_index.data_ptr(), gflat.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=f"w{window} s{shift}")
        return (gx, None, None, None, None, None, None, None, None, None, *pack.split_grad(gflat))


class SwinBlockEmbed(torch.autograd.Function):
    """First Swin block with the patch embedding fused in (bf16 mode, in_chans == 1): raw input x [N,V,1,T,H,W] (contiguous) ->
    block output tokens [N,V,T,H,W,16].  The kernels evaluate LayerNorm(w * x + b) on the fly wherever the block reads its input
    (forward and the recompute of the attention backward), so the embedded tokens are never written to or read from HBM;
    the token gradient of the block goes straight into the embedding's backward kernel.  Swin_3D.py:473-491 + 224-287."""

    @staticmethod
    def forward(ctx, x, wpack: ParamPack, bpack: ParamPack, pack: ParamPack, rel_index, window, shift, rpb_rows, scale, heads, hidden,
                *params):
        L.require_cuda(x)
        lib = L.load()
        assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[2] == 1
        N, V, _, T, H, W = x.shape
        flat, w, b = pack.tensor(), wpack.tensor(), bpack.tensor()
        L.require_cuda(flat, w, b, rel_index)
        if lib.idee_swin_block_packed_floats(rpb_rows) != pack.P:
            raise RuntimeError("swin_block: packed parameter size mismatch")
        out = torch.empty(N, V, T, H, W, 16, device=x.device, dtype=torch.float32)
        umma = L.swin_umma()
        d = _umma_desc(out, torch.float32, torch.float32, window, shift, rpb_rows, scale, pack, heads, hidden) if umma else \
            _swin_desc(out, window, shift, rpb_rows, scale, pack, heads, hidden)
        d.embed_x, d.embed_w, d.embed_b = x.data_ptr(), w.data_ptr(), b.data_ptr()
        need_bwd = any(ctx.needs_input_grad)
        ymid = torch.empty(out.shape, device=x.device, dtype=torch.bfloat16 if umma else torch.float32) if need_bwd else None
        L.run("swin_block_fwd", lib.idee_swin_block_fwd, C.byref(d), None, out.data_ptr(), L.ptr(ymid), None, flat.data_ptr(),
              rel_index.data_ptr(), L.stream(), tag=f"w{window} s{shift} +embed")
        if need_bwd:
            ctx.save_for_backward(x, ymid, rel_index)
            ctx.packs, ctx.args = (pack, wpack, bpack), (window, shift, rpb_rows, scale, heads, hidden)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        x, ymid, rel_index = ctx.saved_tensors
        pack, wpack, bpack = ctx.packs
        window, shift, rpb_rows, scale, heads, hidden = ctx.args
        N, V, _, T, H, W = x.shape
        flat, w, b = pack.tensor(), wpack.tensor(), bpack.tensor()
        if ymid.dtype == torch.bfloat16:
            d = _umma_desc(ymid, torch.float32, torch.float32, window, shift, rpb_rows, scale, pack, heads, hidden)
            gout = _bf16c(gout)
        else:
            d = _swin_desc(ymid, window, shift, rpb_rows, scale, pack, heads, hidden)
            gout = _f32c(gout)
        d.embed_x, d.embed_w, d.embed_b = x.data_ptr(), w.data_ptr(), b.data_ptr()
        gtok = torch.empty_like(ymid)                      # scratch between the MLP and attention halves of the backward
        gflat = pack.grad_out(flat)
        gw, gb = wpack.grad_out(w), bpack.grad_out(b)
        d.embed_gw, d.embed_gb = gw.data_ptr(), gb.data_ptr()  # the attention half also runs the embedding's backward
        nws = lib.idee_swin_block_bwd_workspace_bytes(C.byref(d))
        ws = L.workspace(nws, x.device)
        L.run("swin_block_bwd", lib.idee_swin_block_bwd, C.byref(d), None, ymid.data_ptr(), gout.data_ptr(), gtok.data_ptr(), flat.data_ptr(),
              rel_index.data_ptr(), gflat.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=f"w{window} s{shift} +embed")
        return (None, None, None, None, None, None, None, None, None, None, None,
                *pack.split_grad(gflat), *wpack.split_grad(gw), *bpack.split_grad(gb))


def swin_block_embed(x, wpack: ParamPack, bpack: ParamPack, pack: ParamPack, rel_index, window, shift, rpb_rows, scale, heads, hidden):
    return SwinBlockEmbed.apply(x, wpack, bpack, pack, rel_index, tuple(window), tuple(shift), rpb_rows, float(scale), heads, hidden,
                                *pack.params(), *wpack.params(), *bpack.params())


def swin_block(x, pack: ParamPack, rel_index, window, shift, rpb_rows, scale, heads, hidden, want_bf16: bool = False):
    """-> out, or (out, out_bf16) when want_bf16 (bf16 mode; the copy feeds a bf16-storage conv, see Conv3dCL).
    With the tcgen05 kernels (L.swin_umma()) the tokens themselves are bf16: out is bf16 and doubles as the copy."""
    if L.swin_umma() and window[0] * window[1] * window[2] >= 8:
        out = SwinBlock.apply(x, pack, rel_index, tuple(window), tuple(shift), rpb_rows, float(scale), heads, hidden,
                              "only" if want_bf16 == "only" else False, *pack.params())
        if want_bf16 == "only":
            return out, out
        return (out, out.detach().to(torch.bfloat16)) if want_bf16 else out
    return SwinBlock.apply(x, pack, rel_index, tuple(window), tuple(shift), rpb_rows, float(scale), heads, hidden,
                           want_bf16 if want_bf16 == "only" else bool(want_bf16), *pack.params())


class SwinStack(torch.autograd.Function):
    """All Swin blocks of the encoder as ONE autograd node on the tcgen05 kernels (bf16 mode): raw input x [N,V,1,T,H,W] (patch
    embedding fused into the first block) or fp32 tokens [N,V,T,H,W,16] -> bf16 tokens of the last block (the proj conv rounds
    its input to bf16 anyway).  Between blocks the residual stream is fp32 in HBM; the saved mid residuals and every token
    gradient are bf16 and are handed from block to block without passing through autograd (no dtype casts, one node instead of
    three).  Swin_3D.py:224-287, 422-446, 473-491."""

    @staticmethod
    def forward(ctx, x, blocks, embed, *params):
        """blocks: list of (pack, rel_index, window, shift, rpb_rows, scale, heads, hidden); embed: (wpack, bpack) or None."""
        L.require_cuda(x)
        lib = L.load()
        need_bwd = any(ctx.needs_input_grad)
        if embed is not None:
            assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[2] == 1
            N, V, _, T, H, W = x.shape
            ew, eb = embed[0].tensor(), embed[1].tensor()
            L.require_cuda(ew, eb)
        else:
            x = _f32c(x)
            N, V, T, H, W, _ = x.shape
        shape = (N, V, T, H, W, 16)
        cur, ins, ymids = (None if embed is not None else x), [], []
        for i, (pack, rel_index, window, shift, rpb_rows, scale, heads, hidden) in enumerate(blocks):
            flat = pack.tensor()
            L.require_cuda(flat, rel_index)
            last = i == len(blocks) - 1
            out_dtype = torch.bfloat16 if last else torch.float32
            out = torch.empty(shape, device=x.device, dtype=out_dtype)
            d = _umma_desc(out, torch.float32, out_dtype, window, shift, rpb_rows, scale, pack, heads, hidden)
            tag = f"w{window} s{shift}"
            if i == 0 and embed is not None:
                d.embed_x, d.embed_w, d.embed_b = x.data_ptr(), ew.data_ptr(), eb.data_ptr()
                tag += " +embed"
            ymid = torch.empty(shape, device=x.device, dtype=torch.bfloat16) if need_bwd else None
            L.run("swin_block_fwd", lib.idee_swin_block_fwd, C.byref(d), L.ptr(cur), out.data_ptr(), L.ptr(ymid), None, flat.data_ptr(),
                  rel_index.data_ptr(), L.stream(), tag=tag)
            ins.append(cur)
            ymids.append(ymid)
            cur = out
        if need_bwd:
            ctx.blocks, ctx.embed, ctx.x = blocks, embed, x
            ctx.ins, ctx.ymids = ins, ymids              # plain attributes: intermediate tensors that never leave this node
        return cur

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        blocks, embed, x = ctx.blocks, ctx.embed, ctx.x
        g = _bf16c(gout)
        grads = []
        for i in range(len(blocks) - 1, -1, -1):
            pack, rel_index, window, shift, rpb_rows, scale, heads, hidden = blocks[i]
            flat = pack.tensor()
            ymid, xin = ctx.ymids[i], ctx.ins[i]
            d = _umma_desc(ymid, torch.float32, torch.float32, window, shift, rpb_rows, scale, pack, heads, hidden)
            gflat = pack.grad_out(flat)
            tag = f"w{window} s{shift}"
            gx = torch.empty(ymid.shape, device=ymid.device, dtype=torch.bfloat16)
            if i == 0 and embed is not None:
                ew, eb = embed[0].tensor(), embed[1].tensor()
                gw, gb = embed[0].grad_out(ew), embed[1].grad_out(eb)
                d.embed_x, d.embed_w, d.embed_b = x.data_ptr(), ew.data_ptr(), eb.data_ptr()
                d.embed_gw, d.embed_gb = gw.data_ptr(), gb.data_ptr()
                tag += " +embed"
            nws = lib.idee_swin_block_bwd_workspace_bytes(C.byref(d))
            ws = L.workspace(nws, ymid.device)
            L.run("swin_block_bwd", lib.idee_swin_block_bwd, C.byref(d), L.ptr(xin), ymid.data_ptr(), g.data_ptr(), gx.data_ptr(),
                  flat.data_ptr(), rel_index.data_ptr(), gflat.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=tag)
            grads.append(pack.split_grad(gflat))
            g = gx
        ctx.ins, ctx.ymids = None, None
        out = []
        for gs in reversed(grads):
            out.extend(gs)
        if embed is not None:
            out.extend(embed[0].split_grad(gw))
            out.extend(embed[1].split_grad(gb))
            return (None, None, None, *out)
        return (g.float(), None, None, *out)


def swin_stack(x, blocks, embed=None):
    params = [p for b in blocks for p in b[0].params()]
    if embed is not None:
        params += embed[0].params() + embed[1].params()
    return SwinStack.apply(x, blocks, embed, *params)


# ----------------------------------------------------------------------------------------------------------------
# channel-last Conv3d (proj_var and classifier geometries)
# ----------------------------------------------------------------------------------------------------------------
def _conv_desc(x_dims, x_strides, y_strides, Vw, Cin, Cout, proj, relu, in_cpg, out_cpg, x_sg, y_sg):
    """x_dims = (N, V, Ti, Hi, Wi); strides = (sn, sv, st, sh, sw) in elements."""
    N, V, Ti, Hi, Wi = x_dims
    d = L.ConvDesc()
    d.N, d.V, d.Vw, d.Cin, d.Cout = N, V, Vw, Cin, Cout
    d.Ti, d.Hi, d.Wi = Ti, Hi, Wi
    d.To = Ti if proj else (Ti - 2) // 2 + 1
    d.Ho, d.Wo = Hi, Wi
    d.proj, d.relu = int(proj), int(relu)
    d.x_sn, d.x_sv, d.x_st, d.x_sh, d.x_sw = x_strides
    d.y_sn, d.y_sv, d.y_st, d.y_sh, d.y_sw = y_strides
    d.x_sg, d.y_sg, d.in_cpg, d.out_cpg = x_sg, y_sg, in_cpg, out_cpg
    d.precision = 1 if L.PRECISION == "bf16" else 0
    d.umma16 = int(L.UMMA16 and L.PRECISION == "bf16")
    d.umma96 = int(L.UMMA96 and L.PRECISION == "bf16")
    return d


class Conv3dCL(torch.autograd.Function):
    """x: channel-last storage viewed as [N,V,Ti,Hi,Wi,Cg] (Cg contiguous); `groups` > 1 reads dim 1 as channel groups
    (joint classifier head over the V planes of z_q).  Output [N,Vimg,To,Ho,Wo,Cout] contiguous.

    bf16 activation storage (bf16 mode, 16 -> 16 proj conv only): the tensor-core kernels round their operands to bf16 as they
    load them, so an activation that only such kernels consume can live in HBM as bf16 with bit-identical results:
      * ``x`` itself may be a bf16 tensor (then its gradient is produced in bf16 too);
      * ``x16`` is a bf16 copy of an fp32 ``x`` made by the producer; it replaces ``x`` as the kernels' input (the gradient
        w.r.t. ``x`` stays fp32);
      * ``out_bf16`` stores the output as bf16 (and its incoming gradient is then bf16)."""

    @staticmethod
    def forward(ctx, x, w, b, proj, relu, groups, input_is_relu=False, consumer_masks=False, x16=None, out_bf16=False, cin_real=0):
        """input_is_relu: x is the fused-ReLU output of a conv whose ONLY consumer is this op -> this op's data gradient is
        multiplied by (x > 0) in the kernel epilogue.  consumer_masks: the (only) consumer of this op's ReLU output does
        exactly that, so the incoming gradient is already masked and no separate ReLU-backward pass is needed."""
        L.require_cuda(x, w, b, x16)
        lib = L.load()
        gx_dtype = torch.bfloat16 if (x.dtype == torch.bfloat16 and x16 is None) else torch.float32
        xin = x16 if x16 is not None else x
        if x16 is not None:
            assert x16.dtype == torch.bfloat16 and x16.shape == x.shape and x16.stride() == x.stride()
        elif xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        if xin.stride(5) != 1 and xin.shape[5] != 1:
            xin = xin.contiguous()
        N, V, Ti, Hi, Wi, Cg = xin.shape
        w = _f32c(w)
        b = _f32c(b)
        Vw = w.shape[0]                      # w: [Vw][Cout][Cin][kt][3][3]
        Cout, Cin = w.shape[1], w.shape[2]
        if groups > 1:
            assert groups == V and Cin == V * Cg and Cg == 16 and Vw == 1
            Vimg, x_sv, x_sg, in_cpg = 1, 0, xin.stride(1), 1
        else:
            assert Cin == Cg
            Vimg, x_sv, x_sg, in_cpg = V, xin.stride(1), 0, max(Cg // 16, 1)
        To = Ti if proj else (Ti - 2) // 2 + 1
        y = torch.empty(N, Vimg, To, Hi, Wi, Cout, device=xin.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
        d = _conv_desc((N, Vimg, Ti, Hi, Wi), (xin.stride(0), x_sv, xin.stride(2), xin.stride(3), xin.stride(4)),
                       (y.stride(0), y.stride(1), y.stride(2), y.stride(3), y.stride(4)), Vw, Cin, Cout, proj, relu,
                       in_cpg, max(Cout // 16, 1), x_sg, 0)
        d.x_dtype, d.y_dtype, d.gx_dtype = int(xin.dtype == torch.bfloat16), int(out_bf16), int(gx_dtype == torch.bfloat16)
        d.cin_real = int(cin_real)       # leading input channels that carry data; the gradient of the rest is left unwritten
        nws = lib.idee_conv3d_fwd_workspace_bytes(C.byref(d))
        ws = L.workspace(nws, xin.device)
        L.run("conv3d_fwd_bf16" if d.precision else "conv3d_fwd", lib.idee_conv3d_fwd, C.byref(d), xin.data_ptr(), w.data_ptr(),
              b.data_ptr(), y.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=_conv_tag(d))
        ctx.save_for_backward(xin, w, y if relu else None)
        ctx.desc, ctx.relu, ctx.groups, ctx.gx_dtype = d, relu, groups, gx_dtype
        ctx.input_is_relu, ctx.consumer_masks = bool(input_is_relu), bool(consumer_masks)
        ctx.grad_dst = (getattr(w, "_idee_grad_dst", None), getattr(b, "_idee_grad_dst", None))
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.load()
        x, w, y = ctx.saved_tensors
        d = ctx.desc
        gy = gy.contiguous() if (d.y_dtype and gy.dtype == torch.bfloat16) else (gy.to(torch.bfloat16).contiguous() if d.y_dtype else _f32c(gy))
        if ctx.relu and not ctx.consumer_masks:
            gy = torch.ops.aten.threshold_backward(gy, y, 0.0)
        gw_dst, gb_dst = ctx.grad_dst
        gw = gw_dst if gw_dst is not None and gw_dst.shape == w.shape else torch.empty_like(w)
        gb = gb_dst if gb_dst is not None and gb_dst.shape == (w.shape[0], w.shape[1]) else \
            torch.empty(w.shape[0], w.shape[1], device=w.device, dtype=torch.float32)
        if ctx.needs_input_grad[0]:
            # both gradients through one entry point (one fused kernel for the 16 -> 1 proj conv, two launches otherwise)
            gx = torch.empty_strided(x.shape, x.stride(), device=x.device, dtype=ctx.gx_dtype) if _dense(x) else None
            if gx is None:
                raise RuntimeError("conv3d_bwd: input must be a dense channel-last tensor")
            relu_src = x.data_ptr() if ctx.input_is_relu else None       # dL/d(pre-activation) = dL/dx * (x > 0), fused
            nws = lib.idee_conv3d_bwd_workspace_bytes(C.byref(d))
            ws = L.workspace(nws, x.device)
            L.run("conv3d_bwd_bf16" if d.precision else "conv3d_bwd", lib.idee_conv3d_bwd, C.byref(d), x.data_ptr(), gy.data_ptr(),
                  w.data_ptr(), relu_src, gx.data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=_conv_tag(d))
        else:
            gx = None
            nws = lib.idee_conv3d_wgrad_workspace_bytes(C.byref(d))
            ws = L.workspace(nws, x.device)
            L.run("conv3d_wgrad_bf16" if d.precision else "conv3d_wgrad", lib.idee_conv3d_wgrad, C.byref(d), x.data_ptr(), gy.data_ptr(),
                  gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), nws, L.stream(), tag=_conv_tag(d))
        return gx, gw, gb, None, None, None, None, None, None, None, None


def _conv_tag(d) -> str:
    return f"{'proj' if d.proj else 'cls'} {d.Cin}->{d.Cout} T{d.Ti}"


def _dense(t: torch.Tensor) -> bool:
    """True if t's strides address every element of its storage span exactly once (a permuted contiguous tensor)."""
    dims = sorted(((s, n) for s, n in zip(t.stride(), t.shape) if n > 1), key=lambda p: p[0])
    expect = 1
    for s, n in dims:
        if s != expect:
            return False
        expect *= n
    return True


def conv3d_cl(x, w, b, proj: bool, relu: bool, groups: int = 1, input_is_relu: bool = False, consumer_masks: bool = False,
              x16=None, out_bf16: bool = False, cin_real: int = 0):
    return Conv3dCL.apply(x, w, b, bool(proj), bool(relu), int(groups), bool(input_is_relu), bool(consumer_masks), x16,
                          bool(out_bf16), int(cin_real))


class PackedWB(torch.autograd.Function):
    """Identity view of a ParamPack's flat [V,P] buffer as a differentiable function of the individual parameters."""

    @staticmethod
    def forward(ctx, pack: ParamPack, shape, *params):
        ctx.pack = pack
        out = pack.tensor().view(shape)
        if pack.grad_flat is not None:
            out._idee_grad_dst = pack.grad_flat.view(shape)      # lets the consumer's wgrad kernel write in place
        return out

    @staticmethod
    def backward(ctx, g):
        pack = ctx.pack
        return (None, None, *pack.split_grad(g.reshape(pack.V, pack.P)))


def packed(pack: ParamPack, shape):
    return PackedWB.apply(pack, tuple(shape), *pack.params())


# ----------------------------------------------------------------------------------------------------------------
# LFQ
# ----------------------------------------------------------------------------------------------------------------
class LFQFn(torch.autograd.Function):
    """z [..., 16] -> (z_q [..., 16], indices int64 [...], aux scalar, xq [...]).  LFQ.py:183-307.
    xq is the quantised scalar x (+-1, with the straight-through gradient) such that z_q = x * w_out + b_out."""

    @staticmethod
    def forward(ctx, z, w_in, b_in, w_out, b_out, training, inv_temp, lam_commit, lam_ent, gamma, codebook_size):
        L.require_cuda(z)
        lib = L.load()
        z = _f32c(z)
        w_in, b_in, w_out, b_out = _f32c(w_in), _f32c(b_in), _f32c(w_out), _f32c(b_out)
        dim = z.shape[-1]
        ntok = z.numel() // dim
        zq = torch.empty_like(z)
        idx = torch.empty(z.shape[:-1], device=z.device, dtype=torch.int64)
        xq = torch.empty(z.shape[:-1], device=z.device, dtype=torch.float32)
        stats = torch.zeros(8, device=z.device, dtype=torch.float32)
        nws = lib.idee_lfq_workspace_bytes(ntok)
        ws = L.workspace(nws, z.device)
        L.run("lfq_fwd" if training else "lfq_fwd_eval", lib.idee_lfq_fwd, z.data_ptr(), w_in.data_ptr(), b_in.data_ptr(),
              w_out.data_ptr(), b_out.data_ptr(), zq.data_ptr(), idx.data_ptr(), xq.data_ptr(), stats.data_ptr(), ntok, dim,
              codebook_size, int(training), inv_temp, lam_commit, lam_ent, gamma, ws.data_ptr(), nws, None, L.stream())
        ctx.save_for_backward(z, w_in, b_in, w_out, stats)
        ctx.hyper = (inv_temp, lam_commit, lam_ent, gamma)
        ctx.training = training
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return zq, idx, stats[0], xq

    @staticmethod
    def backward(ctx, gzq, _gidx, gaux, gxq):
        lib = L.load()
        z, w_in, b_in, w_out, stats = ctx.saved_tensors
        inv_temp, lam_commit, lam_ent, gamma = ctx.hyper
        ntok = z.numel() // z.shape[-1]
        gzq = torch.zeros_like(z) if gzq is None else _f32c(gzq)
        if not ctx.training:
            gaux, gxq = None, None   # eval mode: aux is a constant 0 and x = q has no gradient path to s
        gaux_t = None if gaux is None else _f32c(gaux).reshape(1)
        gxq_t = None if gxq is None else _f32c(gxq)
        gz = torch.empty_like(z)
        grads = torch.empty(49, device=z.device, dtype=torch.float32)
        nws = lib.idee_lfq_workspace_bytes(ntok)
        ws = L.workspace(nws, z.device)
        L.run("lfq_bwd", lib.idee_lfq_bwd, z.data_ptr(), gzq.data_ptr(), L.ptr(gxq_t), L.ptr(gaux_t), stats.data_ptr(), w_in.data_ptr(),
              b_in.data_ptr(), w_out.data_ptr(), gz.data_ptr(), grads.data_ptr(), ntok, inv_temp, lam_commit, lam_ent, gamma,
              ws.data_ptr(), nws, L.stream())
        if not ctx.training:
            # x = q (LFQ.py:229-230): only project_out receives gradient
            gz = torch.zeros_like(z)
            grads[:17] = 0
        return (gz, grads[0:16].view(1, 16), grads[16:17], grads[17:33].view(16, 1), grads[33:49],
                None, None, None, None, None, None)


class LFQScalarFn(torch.autograd.Function):
    """Pre-projected form of LFQFn: s [...] = project_in(z) already computed by the producer (the encoder's last conv folded
    with project_in into one 16 -> 1 conv) -> (z_q [..., 16], indices int64 [...], aux scalar, xq [...]).  The backward pass
    returns the gradient w.r.t. s only; project_in's gradients flow through the producer's folded weights."""

    @staticmethod
    def forward(ctx, s, w_out, b_out, training, inv_temp, lam_commit, lam_ent, gamma, codebook_size, want_bf16=False):
        """want_bf16: also return a non-differentiable bf16 copy of z_q (for consumers that round it to bf16 anyway)."""
        L.require_cuda(s)
        lib = L.load()
        s = _f32c(s)
        w_out, b_out = _f32c(w_out), _f32c(b_out)
        ntok = s.numel()
        zq = torch.empty(*s.shape, 16, device=s.device, dtype=torch.float32)
        idx = torch.empty(s.shape, device=s.device, dtype=torch.int64)
        xq = torch.empty(s.shape, device=s.device, dtype=torch.float32)
        stats = torch.zeros(8, device=s.device, dtype=torch.float32)
        zq16 = torch.empty(*s.shape, 16, device=s.device, dtype=torch.bfloat16) if want_bf16 else None
        nws = lib.idee_lfq_workspace_bytes(ntok)
        ws = L.workspace(nws, s.device)
        L.run("lfq_fwd" if training else "lfq_fwd_eval", lib.idee_lfq_fwd, s.data_ptr(), None, None, w_out.data_ptr(), b_out.data_ptr(),
              zq.data_ptr(), idx.data_ptr(), xq.data_ptr(), stats.data_ptr(), ntok, 1, codebook_size, int(training), inv_temp,
              lam_commit, lam_ent, gamma, ws.data_ptr(), nws, L.ptr(zq16), L.stream())
        ctx.save_for_backward(s, w_out, stats)
        ctx.hyper = (inv_temp, lam_commit, lam_ent, gamma)
        ctx.training = training
        ctx.set_materialize_grads(False)
        if want_bf16:
            ctx.mark_non_differentiable(idx, zq16)
            return zq, idx, stats[0], xq, zq16
        ctx.mark_non_differentiable(idx)
        return zq, idx, stats[0], xq

    @staticmethod
    def backward(ctx, gzq, _gidx, gaux, gxq, *_unused):
        lib = L.load()
        s, w_out, stats = ctx.saved_tensors
        inv_temp, lam_commit, lam_ent, gamma = ctx.hyper
        ntok = s.numel()
        gzq = torch.zeros(*s.shape, 16, device=s.device, dtype=torch.float32) if gzq is None else _f32c(gzq)
        if not ctx.training:
            gaux, gxq = None, None
        gaux_t = None if gaux is None else _f32c(gaux).reshape(1)
        gxq_t = None if gxq is None else _f32c(gxq)
        gs = torch.empty_like(s)
        grads = torch.empty(49, device=s.device, dtype=torch.float32)
        nws = lib.idee_lfq_workspace_bytes(ntok)
        ws = L.workspace(nws, s.device)
        L.run("lfq_bwd", lib.idee_lfq_bwd, s.data_ptr(), gzq.data_ptr(), L.ptr(gxq_t), L.ptr(gaux_t), stats.data_ptr(), None, None,
              w_out.data_ptr(), gs.data_ptr(), grads.data_ptr(), ntok, inv_temp, lam_commit, lam_ent, gamma, ws.data_ptr(), nws, L.stream())
        if not ctx.training:
            gs = torch.zeros_like(s)
        return gs, grads[17:33].view(16, 1), grads[33:49], None, None, None, None, None, None, None


class LFQGeneralFn(torch.autograd.Function):
    """LFQ with a 2^K-entry codebook, K = 2..4: z [..., 16] -> (z_q [..., 16], indices int64 [...], aux scalar).  LFQ.py:183-307."""

    @staticmethod
    def forward(ctx, z, w_in, b_in, w_out, b_out, training, inv_temp, lam_commit, lam_ent, gamma, bits):
        L.require_cuda(z, w_in, w_out)
        lib = L.load()
        z = _f32c(z)
        w_in, b_in, w_out, b_out = _f32c(w_in), _f32c(b_in), _f32c(w_out), _f32c(b_out)
        ntok = z.numel() // z.shape[-1]
        zq = torch.empty_like(z)
        idx = torch.empty(z.shape[:-1], device=z.device, dtype=torch.int64)
        stats = torch.zeros(4 + 2 ** bits, device=z.device, dtype=torch.float32)
        nws = lib.idee_lfqk_workspace_bytes(bits)
        ws = L.workspace(nws, z.device)
        L.run("lfqk_fwd" if training else "lfqk_fwd_eval", lib.idee_lfqk_fwd, z.data_ptr(), w_in.data_ptr(), b_in.data_ptr(), w_out.data_ptr(),
              b_out.data_ptr(), zq.data_ptr(), idx.data_ptr(), stats.data_ptr(), ntok, z.shape[-1], bits, int(training), inv_temp,
              lam_commit, lam_ent, gamma, ws.data_ptr(), nws, L.stream())
        ctx.save_for_backward(z, w_in, b_in, w_out, stats)
        ctx.hyper = (inv_temp, lam_commit, lam_ent, gamma, bits, training)
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return zq, idx, stats[0]

    @staticmethod
    def backward(ctx, gzq, _gidx, gaux):
        lib = L.load()
        z, w_in, b_in, w_out, stats = ctx.saved_tensors
        inv_temp, lam_commit, lam_ent, gamma, bits, training = ctx.hyper
        ntok = z.numel() // z.shape[-1]
        gzq = torch.zeros_like(z) if gzq is None else _f32c(gzq)
        gaux_t = None if (gaux is None or not training) else _f32c(gaux).reshape(1)
        gz = torch.empty_like(z)
        K = bits
        grads = torch.empty(K * 16 + K + 16 * K + 16, device=z.device, dtype=torch.float32)
        nws = lib.idee_lfqk_workspace_bytes(bits)
        ws = L.workspace(nws, z.device)
        L.run("lfqk_bwd", lib.idee_lfqk_bwd, z.data_ptr(), gzq.data_ptr(), L.ptr(gaux_t), stats.data_ptr(), w_in.data_ptr(), b_in.data_ptr(),
              w_out.data_ptr(), gz.data_ptr(), grads.data_ptr(), ntok, bits, int(training), inv_temp, lam_commit, lam_ent, gamma,
              ws.data_ptr(), nws, L.stream())
        o = K * 16
        return (gz, grads[:o].view(K, 16), grads[o:o + K], grads[o + K:o + K + 16 * K].view(16, K), grads[o + K + 16 * K:],
                None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------------
class BCEWeighted(torch.autograd.Function):
    """K logit maps sharing one binary target -> losses [K].  losses.py:105-124.
    pred: [K?, N, 1, H, W]-like storage addressed as k*sk + n*sn + i."""

    @staticmethod
    def forward(ctx, pred, target, K, sk, sn, N, HW, kdim=0):
        L.require_cuda(pred, target)
        lib = L.load()
        target = _f32c(target)
        loss = torch.empty(K, device=pred.device, dtype=torch.float32)
        wts = torch.empty(2, device=pred.device, dtype=torch.float32)
        dpred = torch.empty_strided(pred.shape, pred.stride(), device=pred.device, dtype=torch.float32)
        nws = lib.idee_bce_loss_workspace_bytes(K)
        ws = L.workspace(nws, pred.device)
        L.run("bce_loss_fwd", lib.idee_bce_loss_fwd, pred.data_ptr(), sk, sn, K, N, HW, target.data_ptr(), wts.data_ptr(), loss.data_ptr(),
              dpred.data_ptr(), ws.data_ptr(), nws, L.stream())
        ctx.save_for_backward(dpred)
        ctx.K, ctx.kdim = K, kdim
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (dpred,) = ctx.saved_tensors
        if ctx.K == 1:
            return dpred * gloss.reshape(()), None, None, None, None, None, None, None
        shape = [1] * dpred.dim()
        shape[ctx.kdim] = ctx.K
        return dpred * gloss.view(shape), None, None, None, None, None, None, None


def bce_loss_map(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """single map, pred/target [N,1,H,W] (pred may be a strided view with contiguous H,W)."""
    if pred.dtype != torch.float32:
        pred = pred.float()
    N = pred.shape[0]
    HW = pred[0].numel()
    inner_ok = pred[0].is_contiguous() if N > 0 else True
    if not inner_ok:
        pred = pred.contiguous()
    return BCEWeighted.apply(pred, target, 1, 0, pred.stride(0), N, HW)[0]


def bce_loss_maps(stacked: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """K logit maps sharing one target in ONE launch set: stacked [N,K,H,W]-like storage (contiguous H,W), target [N,1,H,W]
    -> losses [K] (the same values as K bce_loss_map calls)."""
    assert stacked.dtype == torch.float32 and stacked.dim() == 4 and stacked[0, 0].is_contiguous()
    N, K, H, W = stacked.shape
    return BCEWeighted.apply(stacked, target, K, stacked.stride(1), stacked.stride(0), N, H * W, 1)


class AnomalyL1(torch.autograd.Function):
    """z_q tokens [N,V,T,H,W,16], mask [N,H,W], vq0 [16] -> scalar.  losses.py:147-168."""

    @staticmethod
    def forward(ctx, zq, mask, vq0):
        L.require_cuda(zq, mask, vq0)
        lib = L.load()
        zq, mask, vq0 = _f32c(zq), _f32c(mask), _f32c(vq0)
        N, V, T, H, W, Cc = zq.shape
        out = torch.empty(2, device=zq.device, dtype=torch.float32)
        ntok = N * V * T * H * W
        nws = lib.idee_anomaly_l1_workspace_bytes(ntok)
        ws = L.workspace(nws, zq.device)
        L.run("anomaly_l1_fwd", lib.idee_anomaly_l1_fwd, zq.data_ptr(), mask.data_ptr(), vq0.data_ptr(), N, V, T, H * W, Cc, out.data_ptr(),
                                        ws.data_ptr(), nws, L.stream())
        ctx.save_for_backward(zq, mask, vq0, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        zq, mask, vq0, out = ctx.saved_tensors
        N, V, T, H, W, Cc = zq.shape
        g = _f32c(g).reshape(1)
        gzq = torch.empty_like(zq)
        L.run("anomaly_l1_bwd", lib.idee_anomaly_l1_bwd, zq.data_ptr(), mask.data_ptr(), vq0.data_ptr(), N, V, T, H * W, Cc, out.data_ptr(),
                                        g.data_ptr(), gzq.data_ptr(), L.stream())
        return gzq, None, None


class AnomalyRank1(torch.autograd.Function):
    """The same loss evaluated on the rank-1 form of z_q: xq [N,V,T,H,W] (+-1), w_out [16,1] / [16], b_out [16], mask [N,H,W],
    vq0 [16] -> scalar.  z_q[c] = xq * w_out[c] + b_out[c] (LFQ.py:284), so the per-token L1 takes one of two values and the pass
    reads 1/16 of the bytes; the gradient reaches the encoder through xq (straight-through, idee_lfq_bwd's gxq input)."""

    @staticmethod
    def forward(ctx, xq, w_out, b_out, mask, vq0):
        L.require_cuda(xq, mask, vq0)
        lib = L.load()
        xq, mask, vq0 = _f32c(xq), _f32c(mask), _f32c(vq0)
        w, b = _f32c(w_out.reshape(-1)), _f32c(b_out.reshape(-1))
        N, V, T, H, W = xq.shape
        out = torch.empty(4, device=xq.device, dtype=torch.float32)
        ntok = N * V * T * H * W
        nws = lib.idee_anomaly_rank1_workspace_bytes(ntok)
        ws = L.workspace(nws, xq.device)
        L.run("anomaly_rank1_fwd", lib.idee_anomaly_rank1_fwd, xq.data_ptr(), mask.data_ptr(), w.data_ptr(), b.data_ptr(), vq0.data_ptr(),
              N, V, T, H * W, 16, out.data_ptr(), ws.data_ptr(), nws, L.stream())
        ctx.save_for_backward(xq, mask, w, b, vq0, out)
        ctx.shapes = (w_out.shape, b_out.shape)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        xq, mask, w, b, vq0, out = ctx.saved_tensors
        N, V, T, H, W = xq.shape
        g = _f32c(g).reshape(1)
        gxq = torch.empty_like(xq)
        gw, gb = torch.empty_like(w), torch.empty_like(b)
        L.run("anomaly_rank1_bwd", lib.idee_anomaly_rank1_bwd, xq.data_ptr(), mask.data_ptr(), w.data_ptr(), b.data_ptr(), vq0.data_ptr(),
              N, V, T, H * W, 16, out.data_ptr(), g.data_ptr(), gxq.data_ptr(), gw.data_ptr(), gb.data_ptr(), L.stream())
        return gxq, gw.view(ctx.shapes[0]), gb.view(ctx.shapes[1]), None, None


class Rank1Planes(torch.autograd.Function):
    """xq [N,V,T,H,W] -> planes [N,T,H,W,16] (channel v < V = xq_v, channel V = 1, rest 0): the image the joint classifier's folded
    first conv reads (CNN_3D._joint_conv1_rank1).  One pass each way instead of a cat of three strided copies."""

    @staticmethod
    def forward(ctx, xq):
        L.require_cuda(xq)
        lib = L.load()
        xq = _f32c(xq)
        N, V, T, H, W = xq.shape
        planes = torch.empty(N, T, H, W, 16, device=xq.device, dtype=torch.float32)
        L.run("rank1_planes_fwd", lib.idee_rank1_planes_fwd, xq.data_ptr(), planes.data_ptr(), N, V, T * H * W, L.stream())
        ctx.shape = xq.shape
        return planes

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        N, V, T, H, W = ctx.shape
        g = _f32c(g)
        gxq = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        L.run("rank1_planes_bwd", lib.idee_rank1_planes_bwd, g.data_ptr(), gxq.data_ptr(), N, V, T * H * W, L.stream())
        return gxq


class LnActRes(torch.autograd.Function):
    """out = shortcut + ReLU(LayerNorm(y) * gamma + beta) on channel-last tokens [N,V,T,H,W,16]; gamma / beta packs [V][16].
    models/encoder/CNN_3D.py:129-147.  want_bf16: also return a non-differentiable bf16 copy of out (input of the next conv)."""

    @staticmethod
    def forward(ctx, y, shortcut, gpack: ParamPack, bpack: ParamPack, want_bf16, *params):
        L.require_cuda(y, shortcut)
        lib = L.load()
        y, shortcut = _f32c(y), _f32c(shortcut)
        gamma, beta = gpack.tensor(), bpack.tensor()
        L.require_cuda(gamma, beta)
        N, V, T, H, W, Cc = y.shape
        out = torch.empty_like(y)
        out16 = torch.empty_like(y, dtype=torch.bfloat16) if want_bf16 else None
        L.run("ln_act_res_fwd", lib.idee_ln_act_res_fwd, y.data_ptr(), shortcut.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
              out.data_ptr(), L.ptr(out16), N, V, T * H * W, Cc, L.stream())
        ctx.save_for_backward(y)
        ctx.packs = (gpack, bpack)
        if want_bf16:
            ctx.mark_non_differentiable(out16)
            ctx.set_materialize_grads(False)
            return out, out16
        return out

    @staticmethod
    def backward(ctx, gout, *_unused):
        lib = L.load()
        (y,) = ctx.saved_tensors
        gpack, bpack = ctx.packs
        gamma, beta = gpack.tensor(), bpack.tensor()
        N, V, T, H, W, Cc = y.shape
        gout = _f32c(gout)
        gy = torch.empty_like(y)
        dg, db = gpack.grad_out(gamma), bpack.grad_out(beta)
        nws = lib.idee_ln_act_res_bwd_workspace_bytes(V)
        ws = L.workspace(nws, y.device)
        L.run("ln_act_res_bwd", lib.idee_ln_act_res_bwd, y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), gout.data_ptr(), gy.data_ptr(),
              dg.data_ptr(), db.data_ptr(), N, V, T * H * W, Cc, ws.data_ptr(), nws, L.stream())
        return (gy, gout, None, None, None, *gpack.split_grad(dg), *bpack.split_grad(db))


def ln_act_res(y, shortcut, gpack: ParamPack, bpack: ParamPack, want_bf16: bool = False):
    return LnActRes.apply(y, shortcut, gpack, bpack, bool(want_bf16), *gpack.params(), *bpack.params())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    lib = L.load()
    L.run("adam_step", lib.idee_adam_step, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                               weight_decay, step, L.stream())


def adam_step_state(p, g, m, v, state, beta1, beta2, eps, weight_decay):
    """Adam with (step, lr) in the 2-float device tensor ``state``: CUDA-graph capturable."""
    lib = L.load()
    L.run("adam_step_state", lib.idee_adam_step_state, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), state.data_ptr(),
          beta1, beta2, eps, weight_decay, L.stream())
