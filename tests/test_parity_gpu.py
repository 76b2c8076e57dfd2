"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors of the unmodified reference and
against the CPU oracle on the same seeded inputs.  Tolerances follow BASELINE.json north_star: logits within 1e-4
relative (fp32 path); driver masks identical on >= 99.9 % with every mismatch a quantiser near-tie."""
import pytest
import torch

from oracle import idee_oracle as O
from tests.golden_util import CASES, load_case, rel_err, lfq_scalar, mask_agreement

pytestmark = pytest.mark.gpu

TOL = 1e-4        # north_star fp32 tolerance (max-norm relative)
TOL_GRAD = 5e-4   # parameter gradients: our own criterion (sums over up to 1e5 tokens in different order)
TIE = 1e-4        # |s| below which a mask flip counts as a quantiser tie


def build_model(cfg, sd, train=True):
    from idee_b200.config import default_config
    from idee_b200.models.build import VQ_model
    torch.manual_seed(0)
    model = VQ_model(default_config(encoder=cfg.encoder, in_channels_dynamic=cfg.in_vars, in_channels=cfg.in_chans,
                                    codebook_size=cfg.codebook_size))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    assert all(k.endswith("relative_position_index") or k == "vq.mask" for k in missing), missing
    model = model.cuda()
    return model.train() if train else model.eval()


@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference_golden(name):
    from idee_b200.models.losses import train_step_loss
    cfg, sd, ins, train, ev, grads = load_case(name)
    model = build_model(cfg, sd)
    x, m_ext, m_loss = ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda()
    total, out = train_step_loss(model, x, m_ext, m_loss)
    total.backward()
    torch.cuda.synchronize()
    assert tuple(out["pred"].shape) == tuple(train["pred"].shape)
    assert tuple(out["z_q"].shape) == tuple(train["z_q"].shape)
    assert out["anomaly"].dtype == torch.int64 and tuple(out["anomaly"].shape) == tuple(train["anomaly"].shape)
    s = lfq_scalar(sd, train["z_enc"])
    frac, ties_ok = mask_agreement(out["anomaly"], train["anomaly"], s, TIE)
    assert frac >= 0.999 and ties_ok, (frac, ties_ok)
    flips = (out["anomaly"].cpu() != train["anomaly"].long())
    if not flips.any():   # values downstream of the mask are only comparable when no bit flipped
        assert rel_err(out["pred"], train["pred"]) < TOL
        assert rel_err(torch.stack(list(out["pred_y"])), train["pred_y"]) < TOL
        assert rel_err(out["z_q"], train["z_q"]) < TOL
        assert rel_err(out["loss_anomaly"], train["loss_anomaly"]) < TOL
        assert rel_err(total, train["total"]) < TOL
    assert rel_err(out["loss_z_q"], train["loss_z_q"]) < TOL
    if cfg.codebook_size > 2:
        # K-bit codebooks: for tokens quantised to code 0 the reference's Anomaly_L1 gradient is the sign of a pure rounding
        # residue (see OracleConfig.exact_ste); the kernels return the exact 0 there.  Gradient oracle = autograd over the CPU
        # restatement with the straight-through value evaluated exactly; the reference's own forward values were checked above.
        import dataclasses
        ocfg = dataclasses.replace(cfg, exact_ste=True)
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        want_total, _ = O.train_step_loss(sdg, ins["x"], ins["mask_extreme"], ins["mask_extreme_loss"], ocfg)
        want_total.backward()
        assert rel_err(want_total, train["total"]) < 1e-6
        grads = {k: v.grad for k, v in sdg.items()}
    if not flips.any():
        named = dict(model.named_parameters())
        worst = max((rel_err(named[k].grad, g), k) for k, g in grads.items())
        assert worst[0] < TOL_GRAD, worst


@pytest.mark.parametrize("name", CASES)
def test_eval_matches_reference_golden(name):
    cfg, sd, ins, train, ev, grads = load_case(name)
    model = build_model(cfg, sd, train=False)
    with torch.no_grad():
        pred, pred_y, anomaly, z_q, loss_z_q = model(ins["x"].cuda())
    assert float(loss_z_q) == 0.0
    frac, ties_ok = mask_agreement(anomaly, ev["anomaly"], lfq_scalar(sd, train["z_enc"]), TIE)
    assert frac >= 0.999 and ties_ok
    if frac == 1.0:
        assert rel_err(pred, ev["pred"]) < TOL
        assert rel_err(torch.stack(list(pred_y)), ev["pred_y"]) < TOL


def test_encoder_matches_oracle_medium():
    """Encoder alone against the CPU oracle at a size with many windows (V=3, 8x24x28: shifted windows wrap)."""
    cfg = O.OracleConfig(in_vars=3, in_chans=1)
    sd = O.make_state_dict(cfg, seed=5, kind="random")
    x, _, _ = O.make_inputs(cfg, 2, 8, 24, 28, seed=5)
    with torch.no_grad():
        want = O.swin3d_forward(sd, x, cfg)
    from idee_b200.models.encoder.Swin_3D import Swin_3D
    enc = Swin_3D(in_vars=3, in_chans=1)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=False)
    enc = enc.cuda()
    with torch.no_grad():
        got = enc(x.cuda())
    assert tuple(got.shape) == tuple(want.shape)
    assert rel_err(got, want) < TOL


def test_full_step_matches_oracle_medium():
    """Whole train step (loss + every gradient) against autograd over the CPU oracle, V=6 at 8x16x20."""
    from idee_b200.models.losses import train_step_loss
    cfg = O.OracleConfig(in_vars=6, in_chans=1)
    sd = O.make_state_dict(cfg, seed=11, kind="random")
    x, m_ext, m_loss = O.make_inputs(cfg, 2, 8, 16, 20, seed=11)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want_total, want = O.train_step_loss(sdg, x, m_ext, m_loss, cfg)
    want_total.backward()
    model = build_model(cfg, sd)
    total, out = train_step_loss(model, x.cuda(), m_ext.cuda(), m_loss.cuda())
    total.backward()
    s = lfq_scalar(sd, want["z_enc"].detach())
    frac, ties_ok = mask_agreement(out["anomaly"], want["anomaly"], s, TIE)
    assert frac >= 0.999 and ties_ok
    if frac == 1.0:
        assert rel_err(out["pred"], want["pred"]) < TOL
        assert rel_err(total, want_total) < TOL
        named = dict(model.named_parameters())
        worst = max((rel_err(named[k].grad, sdg[k].grad), k) for k in sd)
        assert worst[0] < TOL_GRAD, worst


def test_anomaly_rank1_matches_full_loss():
    """Anomaly_L1 on the rank-1 factors (xq, w_out, b_out) of z_q against the 16-channel kernel and the oracle: same loss, and the
    gradients agree after the chain rule z_q = xq * w_out + b_out."""
    from idee_b200 import ops
    from oracle import idee_oracle as O
    g = torch.Generator(device="cuda").manual_seed(5)
    N, V, T, H, W = 2, 3, 4, 9, 13
    xq = torch.where(torch.rand(N, V, T, H, W, device="cuda", generator=g) > 0.4, 1.0, -1.0).requires_grad_(True)
    w_out = (torch.randn(16, 1, device="cuda", generator=g) * 0.5).requires_grad_(True)
    b_out = (torch.randn(16, device="cuda", generator=g) * 0.5).requires_grad_(True)
    mask = (torch.rand(N, H, W, device="cuda", generator=g) > 0.7).float()
    vq0 = (-w_out.detach().reshape(-1) + b_out.detach()).clone()            # code of index 0
    loss1 = ops.AnomalyRank1.apply(xq, w_out, b_out, mask, vq0)
    loss1.backward()
    g1 = (xq.grad.clone(), w_out.grad.clone(), b_out.grad.clone())
    for t in (xq, w_out, b_out):
        t.grad = None
    zq = xq.unsqueeze(-1) * w_out.reshape(-1) + b_out                       # [N,V,T,H,W,16]
    loss2 = ops.AnomalyL1.apply(zq, mask, vq0)
    loss2.backward()
    want = O.anomaly_l1_loss_synthetic(zq.detach().permute(0, 1, 5, 2, 3, 4).cpu(), mask.cpu(), vq0.cpu())
    assert abs(float(loss1) - float(want)) <= 1e-6 * abs(float(want)) and abs(float(loss1) - float(loss2)) <= 1e-6 * abs(float(loss2))
    for a, b in zip(g1, (xq.grad, w_out.grad, b_out.grad)):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("V", [1, 6, 11, 15])
def test_rank1_planes_layout_and_gradient(V):
    """idee_rank1_planes_fwd / _bwd against the torch statement of the same layout (bit-exact: pure data movement), ragged THW."""
    from idee_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(V)
    N, T, H, W = 3, 2, 7, 37
    xq = torch.where(torch.rand(N, V, T, H, W, device="cuda", generator=g) > 0.5, 1.0, -1.0).requires_grad_(True)
    planes = ops.Rank1Planes.apply(xq)
    want = torch.cat([xq.detach().permute(0, 2, 3, 4, 1), xq.new_ones(N, T, H, W, 1), xq.new_zeros(N, T, H, W, 15 - V)], dim=-1)
    assert planes.shape == (N, T, H, W, 16) and torch.equal(planes, want)
    gp = torch.randn(N, T, H, W, 16, device="cuda", generator=g)
    planes.backward(gp)
    assert torch.equal(xq.grad, gp[..., :V].permute(0, 4, 1, 2, 3))
