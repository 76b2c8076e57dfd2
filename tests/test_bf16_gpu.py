"""GPU parity tests of the bf16 tensor-core path (precision='bf16': bf16 MMA operands, fp32 accumulate / softmax / LayerNorm /
quantiser / losses, fp32 activations in HBM).  north_star tolerance: 2e-2 relative.

bf16 rounding can flip (a) the quantiser bit of a token whose pre-quantiser scalar s is ~0 and (b) a ReLU of a ~0
pre-activation; both are discontinuities, not kernel errors.  The tests therefore check the continuous quantities directly
(encoder output, logits given the same z_q), demand that every mask flip is a near-tie within the bf16 error band of s, and
compare end-to-end logits / losses / gradients where no bit flipped."""
import pytest
import torch

from oracle import idee_oracle as O
from tests.golden_util import CASES, load_case, rel_err, lfq_scalar, mask_agreement

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(autouse=True)
def bf16_mode():
    from idee_b200 import _lib
    old = _lib.PRECISION
    _lib.set_precision("bf16")
    yield
    _lib.set_precision(old)


def _model(cfg, sd, train=True):
    from tests.test_parity_gpu import build_model
    return build_model(cfg, sd, train)


def rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", CASES)
def test_bf16_encoder_and_mask(name):
    cfg, sd, ins, train, ev, grads = load_case(name)
    model = _model(cfg, sd, train=False)
    with torch.no_grad():
        z_enc = model.encoder(ins["x"].cuda())
        pred, pred_y, anomaly, z_q, _ = model(ins["x"].cuda())
    assert rel_err(z_enc, train["z_enc"]) < TOL
    s_ref = lfq_scalar(sd, train["z_enc"])
    tie = TOL * float(s_ref.abs().max())                       # bf16 error band of the pre-quantiser scalar
    frac, ties_ok = mask_agreement(anomaly, ev["anomaly"], s_ref, tie)
    assert ties_ok, "a driver-mask mismatch is not a quantiser near-tie"
    assert frac >= 0.99
    if frac == 1.0:
        assert rel_err(pred, ev["pred"]) < TOL
        assert rel_err(torch.stack(list(pred_y)), ev["pred_y"]) < TOL


def test_bf16_reference_init_train_step():
    """Reference initialisation (mask far from ties): the whole step must agree, >= 99.9 % (here 100 %) of mask bits."""
    from idee_b200.models.losses import train_step_loss
    cfg, sd, ins, train, ev, grads = load_case("small_default")
    model = _model(cfg, sd)
    total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
    total.backward()
    assert float((out["anomaly"].cpu() == train["anomaly"].long()).float().mean()) >= 0.999
    assert rel_err(out["pred"], train["pred"]) < TOL
    assert rel_err(torch.stack(list(out["pred_y"])), train["pred_y"]) < TOL
    assert rel_err(total, train["total"]) < TOL
    assert rel_err(out["loss_z_q"], train["loss_z_q"]) < TOL
    named = dict(model.named_parameters())
    got = torch.cat([named[k].grad.reshape(-1).cpu() for k in grads])
    want = torch.cat([g.reshape(-1) for g in grads.values()])
    assert rel_l2(got, want) < 5e-2                            # ReLU flips of ~0 pre-activations are the residual


@pytest.mark.parametrize("name", ["six_vars", "small_random"])
def test_bf16_classifier_given_reference_zq(name):
    """Classifier heads alone on the reference's own z_q: no quantiser discontinuity in between."""
    cfg, sd, ins, train, ev, grads = load_case(name)
    model = _model(cfg, sd, train=False)
    with torch.no_grad():
        z, y = model.cls(train["z_q"].cuda())
    assert rel_err(z, train["pred"]) < TOL
    assert rel_err(torch.stack(list(y)), train["pred_y"]) < TOL


def test_bf16_medium_mask_flips_are_near_ties():
    """Larger size (V=6, 2x8x32x40 = 122 880 tokens), reference-style weights with the quantiser bias CENTRED so the mask is
    ~50/50: the worst case for a sign quantiser, because the pre-quantiser scalar is a small difference of large terms.
    Every flipped bit must lie inside the bf16 error band of s; the flip rate is then simply the density of s near zero
    (measured ~0.7 % here, 0 % with the un-centred reference initialisation).  The >= 99.9 % criterion of north_star is
    carried by the fp32 path (tests/test_parity_gpu.py)."""
    cfg = O.OracleConfig()
    sd = O.make_state_dict(cfg, seed=3, kind="reference")
    x, _, _ = O.make_inputs(cfg, 2, 8, 32, 40, seed=3)
    with torch.no_grad():
        z = O.swin3d_forward(sd, x, cfg)
        s = torch.einsum("nvcthw,c->nvthw", z, sd["vq.project_in.weight"][0])
        sd["vq.project_in.bias"] = -s.median().reshape(1)
        want = O.vq_model_forward(sd, x, cfg, training=False)
    model = _model(cfg, sd, train=False)
    with torch.no_grad():
        z_enc = model.encoder(x.cuda())
        pred, pred_y, anomaly, z_q, _ = model(x.cuda())
    assert rel_err(z_enc, want[5]) < TOL
    s_ref = lfq_scalar(sd, want[5])
    s_got = lfq_scalar(sd, z_enc)
    band = float((s_got - s_ref).abs().max())                 # measured error band of the pre-quantiser scalar
    frac, ties_ok = mask_agreement(anomaly, want[2], s_ref, 1.0001 * band + 1e-12)
    assert 0.2 < float(want[2].float().mean()) < 0.8
    assert band < TOL * float(z_enc.abs().max()) * float(sd["vq.project_in.weight"].abs().sum())
    assert ties_ok and frac >= 0.98, frac


@pytest.mark.parametrize("dims", [(2, 4, 20, 28), (1, 8, 37, 19), (1, 5, 18, 9)])      # last: odd T (unpaired data-gradient tiles, unused last slice)
def test_tcgen05_classifier_conv_matches_mma_sync_and_oracle(dims):
    """The warp-specialised tcgen05 / TMEM kernel for the 96->96 classifier conv (conv96_umma.cu: forward + data gradient incl.
    the fused ReLU mask) against the mma.sync kernel and the fp32 oracle (torch conv3d on CPU)."""
    import torch.nn.functional as F
    from idee_b200 import _lib, ops
    N, T, H, W = dims
    To = (T - 2) // 2 + 1
    torch.manual_seed(0)
    x = torch.randn(N, 1, T, H, W, 96).relu_()
    w = torch.randn(1, 96, 96, 2, 3, 3) * 0.05
    b = torch.randn(1, 96)
    g = torch.randn(N, 1, To, H, W, 96)
    xr = x.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = F.conv3d(xr[:, 0].permute(0, 4, 1, 2, 3), wr[0], br[0], stride=(2, 1, 1), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1).unsqueeze(1)
    (want * g).sum().backward()
    want_gx = xr.grad * (x > 0)                                    # input_is_relu: the data gradient carries the ReLU mask
    old96 = _lib.UMMA96
    outs = {}
    try:
        for on in (False, True):
            _lib.set_umma96(on)
            xc = x.cuda().requires_grad_(True)
            wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
            y = ops.conv3d_cl(xc, wc, bc, False, False, input_is_relu=True)
            (y * g.cuda()).sum().backward()
            outs[on] = (y.detach().cpu(), xc.grad.cpu(), wc.grad.cpu(), bc.grad.cpu())
    finally:
        _lib.set_umma96(old96)
    for on in (False, True):
        assert rel_err(outs[on][0], want.detach()) < TOL
        assert rel_err(outs[on][1], want_gx) < TOL
        assert rel_err(outs[on][2], wr.grad) < TOL                 # weight gradient (tcgen05: TMEM accumulation over pixels)
        assert rel_err(outs[on][3], br.grad) < TOL
    for i in range(4):                                             # same bf16 operands, at most a different accumulation order
        assert rel_err(outs[True][i], outs[False][i]) < 2e-3


@pytest.mark.parametrize("dims", [(2, 4, 20, 28), (1, 8, 37, 19), (1, 3, 9, 8)])
def test_tcgen05_joint_conv1(dims):
    """The joint head's first conv on the 16-channel plane image (16 -> 96): forward and data gradient (conv96_umma.cu with resident
    weights) and weight / bias gradient (conv96_wgrad_umma.cu, all 18 taps in one CTA) against the mma.sync kernels and torch
    autograd (fp32 conv3d on CPU)."""
    import torch.nn.functional as F
    from idee_b200 import _lib, ops
    N, T, H, W = dims
    To = (T - 2) // 2 + 1
    torch.manual_seed(1)
    x = torch.randn(N, 1, T, H, W, 16)
    x[..., 7:] = 0                                                  # plane image: 6 scalar planes | ones | zeros
    x[..., 6] = 1
    w = torch.randn(1, 96, 16, 2, 3, 3) * 0.1
    b = torch.randn(1, 96)
    g = torch.randn(N, 1, To, H, W, 96)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = F.conv3d(xr[:, 0].permute(0, 4, 1, 2, 3), wr[0], br[0], stride=(2, 1, 1), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1).unsqueeze(1)
    (want * g).sum().backward()
    old96 = _lib.UMMA96
    outs = {}
    try:
        for on in (False, True):
            _lib.set_umma96(on)
            xc = x.cuda().requires_grad_(True)
            wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
            y = ops.conv3d_cl(xc, wc, bc, False, False, cin_real=7)     # no ReLU: a sign flip of a near-zero output would move whole gradient terms
            (y * g.cuda()).sum().backward()
            outs[on] = (y.detach().cpu(), xc.grad.cpu()[..., :7], wc.grad.cpu(), bc.grad.cpu())
    finally:
        _lib.set_umma96(old96)
    for on in (False, True):
        assert rel_err(outs[on][0], want.detach()) < TOL
        assert rel_err(outs[on][1], xr.grad[..., :7]) < TOL             # channels >= cin_real carry no gradient by contract
        assert rel_err(outs[on][2][:, :, :7], wr.grad[:, :, :7]) < TOL
        assert rel_err(outs[on][3], br.grad) < TOL
    for i in range(4):
        assert rel_err(outs[True][i], outs[False][i]) < 2e-3


@pytest.mark.parametrize("shape", [(2, 3, 4, 21, 37), (1, 6, 8, 40, 48)])
def test_bf16_activation_storage_is_bit_identical(shape):
    """proj conv -> ReLU -> proj conv with the intermediate tensors stored as bf16 (Conv3dCL x16 / out_bf16 / bf16 input,
    idee_conv_desc.x_dtype / y_dtype / gx_dtype) against the same kernels on fp32 storage: the tensor-core kernels round their
    operands to bf16 when they load them, so every output and gradient must agree bit for bit.  (Same mma.sync kernels on
    both sides: the tcgen05 kernel accumulates in a different order and is compared separately below.)"""
    from idee_b200 import _lib, ops
    old_umma16 = _lib.UMMA16
    _lib.set_umma16(False)
    N, V, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g)
    w0 = torch.randn(V, 16, 16, 3, 3, 3, device="cuda", generator=g) * 0.08
    w1 = torch.randn(V, 16, 16, 3, 3, 3, device="cuda", generator=g) * 0.08
    b0 = torch.randn(V, 16, device="cuda", generator=g) * 0.1
    b1 = torch.randn(V, 16, device="cuda", generator=g) * 0.1
    gz = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g)

    def run(bf16_io):
        leaves = [t.clone().requires_grad_(True) for t in (x, w0, b0, w1, b1)]
        xx, a0, c0, a1, c1 = leaves
        x16 = xx.detach().to(torch.bfloat16) if bf16_io else None
        h = ops.conv3d_cl(xx, a0, c0, proj=True, relu=True, consumer_masks=True, x16=x16, out_bf16=bf16_io)
        z = ops.conv3d_cl(h, a1, c1, proj=True, relu=False, input_is_relu=True)
        z.backward(gz)
        return h.detach(), z.detach(), [t.grad for t in leaves]

    try:
        h32, z32, g32 = run(False)
        h16, z16, g16 = run(True)
    finally:
        _lib.set_umma16(old_umma16)
    assert h16.dtype == torch.bfloat16 and h32.dtype == torch.float32
    assert torch.equal(h16, h32.to(torch.bfloat16))
    assert torch.equal(z16, z32)
    for a, b in zip(g16, g32):
        assert a.dtype == torch.float32 and torch.equal(a, b)


def test_bf16_swin_block_side_output_matches():
    """The bf16 copy written by the last Swin block's forward kernel is the rounded fp32 output."""
    from idee_b200 import ops
    from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D
    torch.manual_seed(3)
    blocks = [SwinTransformerBlock3D(16, 2, (2, 4, 4), (1, 2, 2), 4., True).cuda() for _ in range(3)]
    pack = ops.ParamPack([b.packed_parameters() for b in blocks])
    ws, ss, idx, rows, scale, heads, hidden = blocks[0].kernel_args(4, 18, 22)
    x = torch.randn(2, 3, 4, 18, 22, 16, device="cuda")
    y = ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden)
    y2, y16 = ops.swin_block(x, pack, idx, ws, ss, rows, scale, heads, hidden, want_bf16=True)
    assert torch.equal(y, y2) and y16.dtype == torch.bfloat16 and torch.equal(y16, y.to(torch.bfloat16))


@pytest.mark.parametrize("shape", [(1, 2, 4, 21, 37), (2, 6, 8, 40, 48)])
def test_tcgen05_proj_conv_matches_mma_sync(shape):
    """conv16_umma.cu (tcgen05.mma + TMEM accumulators, zero-copy halo descriptors) against the mma.sync kernel on the same bf16
    operands: forward (bf16 output) and padded-domain data gradient differ only by fp32 accumulation order."""
    from idee_b200 import _lib, ops
    N, V, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g)
    w = torch.randn(V, 16, 16, 3, 3, 3, device="cuda", generator=g) * 0.08
    b = torch.randn(V, 16, device="cuda", generator=g) * 0.1
    gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)

    def run(umma):
        old = _lib.UMMA16
        _lib.set_umma16(umma)
        try:
            xx, ww, bb = (t.clone().requires_grad_(True) for t in (x, w, b))
            y = ops.conv3d_cl(xx, ww, bb, proj=True, relu=True, consumer_masks=True, x16=xx.detach().to(torch.bfloat16), out_bf16=True)
            y.backward(gy)
            return y.detach().float(), xx.grad, ww.grad, bb.grad
        finally:
            _lib.set_umma16(old)

    y0, gx0, gw0, gb0 = run(False)
    y1, gx1, gw1, gb1 = run(True)
    assert float((y0 - y1).abs().max()) <= 2 ** -7 * float(y0.abs().max())        # at most one bf16 ulp of the largest value
    assert float((y0 != y1).float().mean()) < 1e-3                                # ... and only where the fp32 sums straddle a tie
    assert rel_l2(gx1, gx0) < 1e-6
    assert torch.equal(gw0, gw1) and torch.equal(gb0, gb1)                        # weight gradient kernel is shared


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shape", [(1, 2, 1, 9, 11), (1, 2, 4, 21, 37), (2, 6, 8, 40, 48)])
def test_tcgen05_proj_dgrad_direct_matches_mma_sync(shape, masked):
    """bf16 input gradient of the 16 -> 16 proj conv: the tcgen05 kernel on the [T][H+2][W+2] domain (replicate adjoint along t as
    extra MMAs on the first / last slice, final pixels written straight to gx with the fused ReLU mask, only the h / w ring through
    the fp32 buffer + fold_ring_kernel) against the mma.sync kernel on the fully padded domain + fold_pad_kernel, and against
    autograd through torch's replicate-padded conv3d on the same bf16-rounded operands."""
    import torch.nn.functional as F
    from idee_b200 import _lib, ops
    N, V, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g)
    if masked:
        x = x.relu()
    x16 = x.to(torch.bfloat16)
    w = torch.randn(V, 16, 16, 3, 3, 3, device="cuda", generator=g) * 0.08
    b = torch.randn(V, 16, device="cuda", generator=g) * 0.1
    gy = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).to(torch.bfloat16)

    def run(umma):
        old = _lib.UMMA16
        _lib.set_umma16(umma)
        try:
            xx = x16.clone().requires_grad_(True)
            y = ops.conv3d_cl(xx, w, b, proj=True, relu=True, consumer_masks=True, input_is_relu=masked, out_bf16=True)
            y.backward(gy)
            assert xx.grad.dtype == torch.bfloat16
            return xx.grad.float()
        finally:
            _lib.set_umma16(old)

    gx0, gx1 = run(False), run(True)
    # oracle: fp32 autograd on the bf16-rounded operands
    xr = x16.float().requires_grad_(True)
    wr = w.to(torch.bfloat16).float()
    want = torch.zeros_like(xr)
    for v in range(V):
        xi = xr[:, v].permute(0, 4, 1, 2, 3)
        yi = F.conv3d(F.pad(xi, (1, 1, 1, 1, 1, 1), mode="replicate"), wr[v])
        want += torch.autograd.grad(yi, xr, gy[:, v].float().permute(0, 4, 1, 2, 3), retain_graph=False)[0]
    if masked:
        want = want * (x16 > 0)
    scale = float(want.abs().max())
    for got in (gx0, gx1):
        assert float((got - want).abs().max()) <= 2 ** -7 * scale                 # one bf16 rounding of the result
    assert float((gx0 - gx1).abs().max()) <= 2 ** -7 * scale
    assert float((gx0 != gx1).float().mean()) < 2e-3                              # same operands: only ties of the final rounding differ


@pytest.mark.parametrize("shape", [(1, 2, 1, 9, 11), (2, 6, 8, 40, 48), (1, 3, 4, 21, 37)])
def test_fused_conv_backward_matches_separate_entry_points(shape):
    """idee_conv3d_bwd on the folded 16 -> 1 proj conv (bf16 input, ReLU mask = the input): the ONE fused kernel must reproduce
    idee_conv3d_wgrad + idee_conv3d_dgrad called separately (same bf16 operands; the weight gradient bit for bit per split, the
    data gradient up to the final bf16 rounding) and torch autograd on the bf16-rounded operands."""
    import ctypes as C
    import torch.nn.functional as F
    from idee_b200 import _lib as L, ops
    N, V, T, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn(N, V, T, H, W, 16, device="cuda", generator=g).relu().to(torch.bfloat16)
    w = (torch.randn(V, 1, 16, 3, 3, 3, device="cuda", generator=g) * 0.1).contiguous()
    gy = torch.randn(N, V, T, H, W, 1, device="cuda", generator=g).contiguous()
    lib = L.load()
    old = L.PRECISION
    L.set_precision("bf16")
    try:
        d = ops._conv_desc((N, V, T, H, W), (x.stride(0), x.stride(1), x.stride(2), x.stride(3), x.stride(4)),
                           (gy.stride(0), gy.stride(1), gy.stride(2), gy.stride(3), gy.stride(4)), V, 16, 1, True, False, 1, 1, 0, 0)
        d.x_dtype, d.y_dtype, d.gx_dtype, d.cin_real = 1, 0, 1, 0

        def buffers():
            return (torch.empty_like(x), torch.empty_like(w), torch.empty(V, 1, device="cuda"))

        gx0, gw0, gb0 = buffers()
        nws = max(lib.idee_conv3d_wgrad_workspace_bytes(C.byref(d)), lib.idee_conv3d_dgrad_workspace_bytes(C.byref(d)))
        ws = L.workspace(nws, x.device)
        L.check(lib.idee_conv3d_wgrad(C.byref(d), x.data_ptr(), gy.data_ptr(), gw0.data_ptr(), gb0.data_ptr(), ws.data_ptr(), nws, L.stream()), "wgrad")
        L.check(lib.idee_conv3d_dgrad(C.byref(d), gy.data_ptr(), w.data_ptr(), x.data_ptr(), gx0.data_ptr(), ws.data_ptr(), nws, L.stream()), "dgrad")
        gx1, gw1, gb1 = buffers()
        nws = lib.idee_conv3d_bwd_workspace_bytes(C.byref(d))
        ws = L.workspace(nws, x.device)
        L.check(lib.idee_conv3d_bwd(C.byref(d), x.data_ptr(), gy.data_ptr(), w.data_ptr(), x.data_ptr(), gx1.data_ptr(), gw1.data_ptr(),
                                    gb1.data_ptr(), ws.data_ptr(), nws, L.stream()), "bwd")
        torch.cuda.synchronize()
    finally:
        L.set_precision(old)
    # oracle: fp32 autograd on the bf16-rounded operands
    xr = x.float().requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    tot = 0
    for v in range(V):
        yi = F.conv3d(F.pad(xr[:, v].permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1), mode="replicate"), wr[v])
        tot = tot + (yi * gy[:, v].permute(0, 4, 1, 2, 3)).sum()
    tot.backward()
    want_gx = xr.grad * (x > 0)
    scale = float(want_gx.abs().max())
    for got in (gx0, gx1):
        assert float((got.float() - want_gx).abs().max()) <= 2 ** -6 * scale       # bf16 operands (gy rounded) + bf16 result
    assert float((gx0.float() - gx1.float()).abs().max()) <= 2 ** -7 * scale
    assert rel_err(gw1, wr.grad) < TOL and rel_err(gw0, wr.grad) < TOL
    assert rel_err(gw1, gw0) < 1e-5                                                 # same products, different split boundaries
    assert rel_err(gb1, gy.sum(dim=(0, 2, 3, 4, 5)).reshape(V, 1)) < 1e-5 and rel_err(gb0, gb1) < 1e-5


def test_folded_last_conv_matches_conv_then_project_in():
    """VQ_model's bf16 path evaluates proj_var[2] followed by LFQ.project_in as ONE 16 -> 1 conv (Swin_3D.forward_tokens(fold_last=...)):
    the scalar it feeds to the quantiser must equal project_in(encoder output), and so must the gradients that reach the last
    conv's weights, project_in and the encoder input."""
    from idee_b200.config import default_config
    from idee_b200.models.build import VQ_model
    cfg = default_config()
    torch.manual_seed(2)
    model = VQ_model(cfg).cuda().train()
    with torch.no_grad():                                  # spread the quantiser input around 0 so both signs occur
        model.vq.project_in.bias.fill_(-0.35)
    x = torch.randn(2, cfg.in_channels_dynamic, cfg.in_channels, 8, 24, 40, device="cuda")
    w_in, b_in = model.vq.project_in.weight, model.vq.project_in.bias
    gs = torch.randn(2, cfg.in_channels_dynamic, 8, 24, 40, device="cuda")

    def grads():
        ps = [model.encoder.proj_var[0][2].weight, w_in, b_in, model.encoder.layers_var[0][0].blocks[0].attn.qkv.weight]
        out = [p.grad.clone() for p in ps]
        model.zero_grad(set_to_none=True)
        return out

    s_fold = model.encoder.forward_tokens(x, fold_last=(w_in, b_in))
    (s_fold * gs).sum().backward()
    g_fold = grads()
    z = model.encoder.forward_tokens(x)
    s_ref = z @ w_in.reshape(-1) + b_in
    (s_ref * gs).sum().backward()
    g_ref = grads()
    assert rel_l2(s_fold, s_ref) < 5e-3
    for a, b in zip(g_fold, g_ref):
        assert rel_l2(a, b) < TOL
