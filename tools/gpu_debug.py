"""Per-op GPU diagnostics against the CPU oracle (development aid; the formal tests are tests/test_*_gpu.py).

    python tools/gpu_debug.py            # runs every stage, each in its own process (a CUDA fault poisons a context)
    python tools/gpu_debug.py <stage>    # one stage in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

from oracle import idee_oracle as O


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


BF16 = os.environ.get("IDEE_B200_PRECISION", "fp32") == "bf16"


def report(name, err, tol=1e-4):
    if BF16:
        tol = 2e-2
    print(f"  {'OK  ' if err < tol else 'FAIL'} {name:58s} rel_err={err:.3e}", flush=True)


def stage_embed():
    from idee_b200 import ops
    for Cin in (1, 2):
        V, N, T, H, W = 3, 2, 8, 6, 10
        torch.manual_seed(Cin)
        ws = [torch.randn(16, Cin, 1, 1, 1, requires_grad=True) for _ in range(V)]
        bs = [torch.randn(16, requires_grad=True) for _ in range(V)]
        x = torch.randn(N, V, Cin, T, H, W)
        g = torch.randn(N, V, T, H, W, 16)
        want = torch.stack([O.patch_embed({"proj.weight": ws[v], "proj.bias": bs[v]}, "", x[:, v], (1, 1, 1)) for v in range(V)], 1)
        want_tok = want.permute(0, 1, 3, 4, 5, 2)
        (want_tok * g).sum().backward()
        wc = [torch.nn.Parameter(w.detach().cuda()) for w in ws]
        bc = [torch.nn.Parameter(b.detach().cuda()) for b in bs]
        pw, pb = ops.ParamPack([[w] for w in wc]), ops.ParamPack([[b] for b in bc])
        got = ops.embed_ln(x.cuda(), pw, pb)
        (got * g.cuda()).sum().backward()
        report(f"embed_ln fwd Cin={Cin}", rel(got, want_tok))
        report(f"embed_ln dW Cin={Cin}", max(rel(wc[v].grad, ws[v].grad) for v in range(V)), 5e-4)
        report(f"embed_ln db Cin={Cin}", max(rel(bc[v].grad, bs[v].grad) for v in range(V)), 5e-4)


def _block_case(ws_cfg, shift_on, dims, V=2, N=2, seed=0, tol=1e-4):
    from idee_b200 import ops
    from idee_b200.models.encoder.Swin_3D import SwinTransformerBlock3D
    T, H, W = dims
    torch.manual_seed(seed)
    layer_ss = tuple(i // 2 for i in ws_cfg)
    ss = layer_ss if shift_on else (0, 0, 0)
    blocks = [SwinTransformerBlock3D(16, 2, ws_cfg, ss, 4., True) for _ in range(V)]
    for b in blocks:
        for p in b.parameters():
            torch.nn.init.normal_(p, 0, 0.3)
    x = torch.randn(N, V, T, H, W, 16)
    g = torch.randn(N, V, T, H, W, 16)
    # oracle
    wants, sds = [], []
    xo = x.clone().requires_grad_(True)
    for v, b in enumerate(blocks):
        sd = {k: p.detach().clone().requires_grad_(True) for k, p in b.named_parameters()}
        sds.append(sd)
        wants.append(O.swin_block(sd, "", xo[:, v], ws_cfg, ss, 2, None, layer_ss))
    want = torch.stack(wants, 1)
    (want * g).sum().backward()
    # ours
    for b in blocks:
        b.cuda()
    pack = ops.ParamPack([b.packed_parameters() for b in blocks])
    wsz, ssz, idx, rows, scale, heads, hidden = blocks[0].kernel_args(T, H, W)
    xc = x.cuda().requires_grad_(True)
    got = ops.swin_block(xc, pack, idx, wsz, ssz, rows, scale, heads, hidden)
    tag = f"swin ws={ws_cfg} shift={int(shift_on)} dims={dims}"
    report(tag + " fwd", rel(got, want), tol)
    (got * g.cuda()).sum().backward()
    report(tag + " gx", rel(xc.grad, xo.grad), 5e-4)
    worst = (0.0, "")
    for v, b in enumerate(blocks):
        for k, p in b.named_parameters():
            worst = max(worst, (rel(p.grad, sds[v][k].grad), k))
    report(tag + f" dparams (worst {worst[1]})", worst[0], 5e-4)


def stage_swin_a():
    _block_case((2, 4, 4), False, (8, 8, 12))
    _block_case((2, 4, 4), True, (8, 8, 12))


def stage_swin_b():
    _block_case((8, 1, 1), False, (8, 6, 10))
    _block_case((2, 4, 4), True, (8, 10, 14))      # padded H, W
    _block_case((8, 1, 1), False, (12, 4, 6))      # padded T (12 -> 16)
    _block_case((2, 4, 4), True, (2, 8, 9))        # clamped window on T (2 <= 2 -> no T shift), padded W
    _block_case((2, 4, 4), True, (8, 24, 28), V=3)


def _conv_case(proj, Cin, Cout, V, Vw, dims, relu, groups=1, N=2, seed=0):
    from idee_b200 import ops
    T, H, W = dims
    torch.manual_seed(seed)
    kt = 3 if proj else 2
    w = (torch.randn(Vw, Cout, Cin, kt, 3, 3) * 0.2).requires_grad_(True)
    b = torch.randn(Vw, Cout).requires_grad_(True)
    Cg = Cin // groups
    x = torch.randn(N, V, T, H, W, Cg).requires_grad_(True)
    outs = []
    if groups > 1:
        xi = x.permute(0, 1, 5, 2, 3, 4).reshape(N, V * Cg, T, H, W)
        imgs = [(xi, 0)]
    else:
        imgs = [(x[:, v].permute(0, 4, 1, 2, 3), v if Vw > 1 else 0) for v in range(V)]
    for xi, wv in imgs:
        if proj:
            y = F.conv3d(F.pad(xi, (1,) * 6, mode="replicate"), w[wv], b[wv])
        else:
            y = F.conv3d(xi, w[wv], b[wv], stride=(2, 1, 1), padding=(0, 1, 1))
        outs.append(F.relu(y) if relu else y)
    want = torch.stack(outs, 1).permute(0, 1, 3, 4, 5, 2)
    g = torch.randn_like(want)
    (want * g).sum().backward()
    xc = x.detach().cuda().requires_grad_(True)
    wc = w.detach().cuda().requires_grad_(True)
    bc = b.detach().cuda().requires_grad_(True)
    got = ops.conv3d_cl(xc, wc, bc, proj, relu, groups)
    tag = f"conv proj={int(proj)} {Cin}->{Cout} V={V} Vw={Vw} g={groups} dims={dims} relu={int(relu)}"
    report(tag + " fwd", rel(got, want))
    if BF16 and relu:
        # bf16 rounding flips the ReLU mask of near-zero pre-activations, which is not a kernel error: check the
        # backward kernels on the linear conv instead
        return _conv_case(proj, Cin, Cout, V, Vw, dims, False, groups, N, seed)
    (got * g.cuda()).sum().backward()
    report(tag + " dgrad", rel(xc.grad, x.grad), 5e-4)
    report(tag + " wgrad", rel(wc.grad, w.grad), 5e-4)
    report(tag + " bgrad", rel(bc.grad, b.grad), 5e-4)


def stage_conv_a():
    _conv_case(True, 16, 16, 3, 3, (8, 6, 10), True)
    _conv_case(True, 16, 16, 2, 2, (1, 5, 1), False)
    _conv_case(False, 16, 16, 3, 3, (8, 6, 10), True)
    _conv_case(False, 16, 1, 3, 3, (2, 6, 10), False)


def stage_conv_b():
    _conv_case(False, 96, 96, 6, 1, (8, 6, 10), True, groups=6)
    _conv_case(False, 96, 96, 1, 1, (4, 9, 7), True)
    _conv_case(False, 96, 1, 1, 1, (3, 9, 7), False)
    _conv_case(True, 16, 16, 2, 2, (8, 40, 52), True, N=1)


def stage_lfq():
    from idee_b200.models.codebook.LFQ import LFQ
    cfg = O.OracleConfig()
    for training in (True, False):
        torch.manual_seed(3)
        m = LFQ(dim=16, codebook_size=2, entropy_loss_weight=0.1, diversity_gamma=0.1, commitment_loss_weight=3.0)
        for p in m.parameters():
            torch.nn.init.normal_(p, 0, 0.5)
        z = (torch.randn(2, 5000, 16) * 0.05).requires_grad_(True)
        sd = {"vq." + k: p.detach().clone().requires_grad_(True) for k, p in m.named_parameters()}
        zq, idx, aux, _ = O.lfq_forward(sd, z, cfg, training)
        g = torch.randn_like(zq)
        ((zq * g).sum() + 7.0 * aux).backward()
        m = m.cuda().train(training)
        zc = z.detach().cuda().requires_grad_(True)
        r = m(zc)
        ((r.quantized * g.cuda()).sum() + 7.0 * r.entropy_aux_loss).backward()
        tag = f"lfq train={int(training)}"
        report(tag + " zq", rel(r.quantized, zq))
        report(tag + " idx mismatch frac", float((r.indices.cpu() != idx).float().mean()), 1e-3)
        report(tag + " aux", rel(r.entropy_aux_loss, aux) if training else float(r.entropy_aux_loss.abs()))
        if training:
            report(tag + " gz", rel(zc.grad, z.grad), 5e-4)
        for k, p in m.named_parameters():
            if sd["vq." + k].grad is not None:
                report(tag + " d" + k, rel(p.grad, sd["vq." + k].grad), 5e-4)


def stage_losses():
    from idee_b200.models.losses import BCE_loss_synthetic, Anomaly_L1_loss_synthetic
    torch.manual_seed(0)
    pred = torch.randn(3, 1, 20, 24, requires_grad=True)
    tgt = (torch.rand(3, 1, 20, 24) < 0.1).float()
    want = O.bce_loss_synthetic(pred, tgt)
    want.backward()
    pc = pred.detach().cuda().requires_grad_(True)
    got = BCE_loss_synthetic()(pc, tgt.cuda())
    got.backward()
    report("bce loss", rel(got, want))
    report("bce dpred", rel(pc.grad, pred.grad), 5e-4)
    zq = torch.randn(2, 3, 16, 8, 6, 10, requires_grad=True)
    mask = (torch.rand(2, 6, 10) < 0.3).float()
    vq0 = torch.randn(1, 16)
    want = O.anomaly_l1_loss_synthetic(zq, mask, vq0)
    want.backward()
    zc = zq.detach().cuda().requires_grad_(True)
    got = Anomaly_L1_loss_synthetic(3, 8, 16)(zc, mask.cuda(), vq0.cuda())
    got.backward()
    report("anomaly l1 loss", rel(got, want))
    report("anomaly l1 dzq", rel(zc.grad, zq.grad), 5e-4)


def stage_model():
    from tests.golden_util import CASES, load_case, lfq_scalar, mask_agreement
    from tests.test_parity_gpu import build_model
    from idee_b200.models.losses import train_step_loss
    for name in CASES:
        cfg, sd, ins, train, ev, grads = load_case(name)
        model = build_model(cfg, sd)
        total, out = train_step_loss(model, ins["x"].cuda(), ins["mask_extreme"].cuda(), ins["mask_extreme_loss"].cuda())
        total.backward()
        s_ref = lfq_scalar(sd, train["z_enc"])
        tie = 2e-2 * float(s_ref.abs().max()) if BF16 else 1e-4
        frac, ties = mask_agreement(out["anomaly"], train["anomaly"], s_ref, tie)
        print(f" [{name}] mask agreement {frac:.5f} ties_ok={ties} (tie threshold {tie:.2e})")
        with torch.no_grad():
            report(f"{name} z_enc", rel(model.encoder(ins["x"].cuda()), train["z_enc"]))
        report(f"{name} pred", rel(out["pred"], train["pred"]))
        report(f"{name} pred_y", rel(torch.stack(list(out["pred_y"])), train["pred_y"]))
        report(f"{name} z_q", rel(out["z_q"], train["z_q"]))
        report(f"{name} loss_z_q", rel(out["loss_z_q"], train["loss_z_q"]))
        report(f"{name} loss_anomaly", rel(out["loss_anomaly"], train["loss_anomaly"]))
        report(f"{name} total", rel(total, train["total"]))
        named = dict(model.named_parameters())
        errs = sorted(((rel(named[k].grad, g), k) for k, g in grads.items()), reverse=True)
        for e, k in errs[:4]:
            report(f"{name} grad {k}", e, 5e-4)


STAGES = {k[6:]: v for k, v in list(globals().items()) if k.startswith("stage_")}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        t0 = time.time()
        print(f"== stage {sys.argv[1]}", flush=True)
        STAGES[sys.argv[1]]()
        torch.cuda.synchronize()
        print(f"== stage {sys.argv[1]} done in {time.time() - t0:.1f}s", flush=True)
    else:
        for s in STAGES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=600)
            if r.returncode != 0:
                print(f"== stage {s} exited with {r.returncode}", flush=True)
