"""tcgen05 / TMEM Swin block kernels (swin_umma.cuh: bf16 tokens, TMA bulk loads, A operands from TMEM, weight gradients in TMEM)
against the mma.sync kernels on fp32 tokens (swin_tc.cuh) and against the fp32 exact kernels, block by block: every window
shape that is built, shifted and unshifted, sizes that need padding, several variables."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # window, shift, (N, V, T, H, W)
    ((2, 4, 4), (0, 0, 0), (2, 3, 8, 16, 24)),
    ((2, 4, 4), (1, 2, 2), (2, 3, 8, 16, 24)),
    ((2, 4, 4), (1, 2, 2), (1, 2, 7, 10, 14)),      # padding on every axis
    ((8, 1, 1), (0, 0, 0), (2, 3, 8, 12, 20)),
    ((8, 1, 1), (0, 0, 0), (1, 2, 12, 9, 7)),       # T = 12 -> padded to 16
    ((2, 2, 2), (1, 1, 1), (1, 2, 4, 6, 10)),
    ((2, 4, 4), (0, 0, 0), (1, 1, 8, 200, 200)),    # benchmark grid: wide rows, many tiles per CTA
]


def _blocks(V, window, shift, seed):
    from idee_b200 import ops
    from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D
    torch.manual_seed(seed)
    blocks = [SwinTransformerBlock3D(16, 2, window, shift, 4., True).cuda() for _ in range(V)]
    with torch.no_grad():
        for b in blocks:
            for p in b.parameters():
                p.normal_(0.0, 0.25)
    return blocks, ops.ParamPack([b.packed_parameters() for b in blocks])


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("window,shift,dims", CASES)
def test_umma_forward_matches_mma_sync_and_fp32(window, shift, dims):
    from idee_b200 import _lib, ops
    N, V, T, H, W = dims
    blocks, pack = _blocks(V, window, shift, 3)
    ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)     # same rounded input for every path
    old = (_lib.PRECISION, _lib.SWIN_UMMA)
    try:
        _lib.set_precision("fp32")
        with torch.no_grad():
            y32 = ops.swin_block(x.float(), pack, idx, ws, ss, rows, scale, heads, hidden)
        _lib.set_precision("bf16")
        _lib.set_swin_umma(False)
        with torch.no_grad():
            ytc = ops.swin_block(x.float(), pack, idx, ws, ss, rows, scale, heads, hidden)
        _lib.set_swin_umma(True)
        with torch.no_grad():
            yu = ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden)                 # bf16 stream
            yu32 = ops.swin_block(x.float(), pack, idx, ws, ss, rows, scale, heads, hidden)       # fp32 stream (the encoder's default)
    finally:
        _lib.set_precision(old[0]); _lib.set_swin_umma(old[1])
    assert yu.dtype == torch.bfloat16 and yu.shape == x.shape
    assert torch.isfinite(yu.float()).all()
    assert yu32.dtype == torch.float32
    e_tc, e_u, e_u32 = _rel(ytc, y32), _rel(yu.float(), y32), _rel(yu32, y32)
    print(f"window {window} shift {shift} dims {dims}: mma.sync vs fp32 {e_tc:.3e}, tcgen05 vs fp32 {e_u:.3e} (bf16 out) {e_u32:.3e} (fp32 out)")
    assert e_u < 2e-2 and e_u < 3 * e_tc + 8e-3, (e_tc, e_u)     # bf16 output rounding (2^-8) on top of the bf16-operand error
    assert e_u32 < 2 * e_tc + 1e-3, (e_tc, e_u32)                # same operand rounding as the mma.sync kernels


def _rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("window,shift,dims", CASES[:6] + [((2, 4, 4), (1, 2, 2), (1, 1, 8, 64, 72))])
def test_umma_backward_matches_fp32(window, shift, dims):
    """g_x and every packed parameter gradient (qkv, proj, fc1, fc2 weights and biases, relative-position-bias table) of the tcgen05
    backward kernels (weight gradients accumulated in TMEM) against the exact fp32 kernels and the mma.sync bf16 kernels."""
    from idee_b200 import _lib, ops
    N, V, T, H, W = dims
    blocks, pack = _blocks(V, window, shift, 5)
    ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)
    old = (_lib.PRECISION, _lib.SWIN_UMMA)
    res = {}
    try:
        for name, prec, umma in (("fp32", "fp32", False), ("tc", "bf16", False), ("umma", "bf16", True)):
            _lib.set_precision(prec)
            _lib.set_swin_umma(umma)
            for p in pack.params():
                p.grad = None
            xin = (x if umma else x.float()).clone().requires_grad_(True)
            y = ops.swin_block(xin, pack, idx, ws, ss, rows, scale, heads, hidden)
            y.backward(gy if umma else gy.float())
            res[name] = (xin.grad.float().clone(), torch.cat([p.grad.reshape(-1) for p in pack.params()]).clone(),
                         [p.grad.clone() for p in pack.params()])
    finally:
        _lib.set_precision(old[0]); _lib.set_swin_umma(old[1])
    e_gx_tc, e_gx_u = _rel_l2(res["tc"][0], res["fp32"][0]), _rel_l2(res["umma"][0], res["fp32"][0])
    e_gp_tc, e_gp_u = _rel_l2(res["tc"][1], res["fp32"][1]), _rel_l2(res["umma"][1], res["fp32"][1])
    names = ["rpb", "qkv.w", "qkv.b", "proj.w", "proj.b", "fc1.w", "fc1.b", "fc2.w", "fc2.b"]
    per = {n: _rel_l2(a, b) for n, a, b in zip(names * V, res["umma"][2], res["fp32"][2])}
    print(f"window {window} shift {shift} dims {dims}: gx mma.sync {e_gx_tc:.3e} tcgen05 {e_gx_u:.3e}; params mma.sync {e_gp_tc:.3e} tcgen05 {e_gp_u:.3e}; {per}")
    assert torch.isfinite(res["umma"][0]).all() and torch.isfinite(res["umma"][1]).all()
    assert e_gx_u < 2e-2 and e_gp_u < 2e-2, (e_gx_u, e_gp_u)
    assert max(per.values()) < 5e-2, per


@pytest.mark.parametrize("dims", [(2, 3, 8, 16, 24), (1, 2, 7, 10, 14)])
def test_umma_fused_embedding_block(dims):
    """First block with the patch embedding fused in (raw scalar input): forward, and the embedding's weight / bias gradients
    that come out of the same TMEM weight-gradient accumulator, against embed_ln + block on the exact fp32 kernels."""
    from idee_b200 import _lib, ops
    N, V, T, H, W = dims
    window, shift = (2, 4, 4), (0, 0, 0)
    blocks, pack = _blocks(V, window, shift, 7)
    ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
    g = torch.Generator(device="cuda").manual_seed(4)
    ew = [torch.nn.Parameter(torch.randn(16, 1, 1, 1, 1, device="cuda", generator=g)) for _ in range(V)]
    eb = [torch.nn.Parameter(torch.randn(16, device="cuda", generator=g) * 0.3) for _ in range(V)]
    wpack, bpack = ops.ParamPack([[w] for w in ew]), ops.ParamPack([[b] for b in eb])
    x = torch.randn(N, V, 1, T, H, W, device="cuda", generator=g)
    gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)
    old = (_lib.PRECISION, _lib.SWIN_UMMA)
    res = {}
    try:
        for name, prec, umma in (("fp32", "fp32", False), ("umma", "bf16", True)):
            _lib.set_precision(prec)
            _lib.set_swin_umma(umma)
            for p in pack.params() + wpack.params() + bpack.params():
                p.grad = None
            if umma:
                y = ops.swin_block_embed(x, wpack, bpack, pack, idx, ws, ss, rows, scale, heads, hidden)
                y.backward(gy)
            else:
                tok = ops.embed_ln(x, wpack, bpack)
                y = ops.swin_block(tok, pack, idx, ws, ss, rows, scale, heads, hidden)
                y.backward(gy.float())
            res[name] = (y.detach().float().clone(), torch.cat([p.grad.reshape(-1) for p in pack.params()]).clone(),
                         torch.cat([p.grad.reshape(-1) for p in wpack.params()]).clone(), torch.cat([p.grad.reshape(-1) for p in bpack.params()]).clone())
    finally:
        _lib.set_precision(old[0]); _lib.set_swin_umma(old[1])
    errs = [_rel(res["umma"][0], res["fp32"][0])] + [_rel_l2(res["umma"][i], res["fp32"][i]) for i in (1, 2, 3)]
    print(f"fused embedding dims {dims}: out {errs[0]:.3e} block params {errs[1]:.3e} embed w {errs[2]:.3e} embed b {errs[3]:.3e}")
    assert errs[0] < 2e-2 and errs[1] < 2e-2 and errs[2] < 3e-2 and errs[3] < 3e-2, errs
