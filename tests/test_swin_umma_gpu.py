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
            yu = ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden)
    finally:
        _lib.set_precision(old[0]); _lib.set_swin_umma(old[1])
    assert yu.dtype == torch.bfloat16 and yu.shape == x.shape
    assert torch.isfinite(yu.float()).all()
    e_tc, e_u = _rel(ytc, y32), _rel(yu.float(), y32)
    print(f"window {window} shift {shift} dims {dims}: mma.sync vs fp32 {e_tc:.3e}, tcgen05 vs fp32 {e_u:.3e}")
    assert e_u < 2e-2 and e_u < 3 * e_tc + 8e-3, (e_tc, e_u)     # bf16 output rounding (2^-8) on top of the bf16-operand error
