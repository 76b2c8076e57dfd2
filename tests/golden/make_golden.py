"""Generate golden vectors for the IDEE hot path by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (``models.build.VQ_model`` + ``models.losses``) is imported read-only from
``/root/reference``.  ``timm`` is not installed here; the reference needs only
``timm.models.layers.{DropPath, trunc_normal_}`` (Swin_3D.py:16, classifier/CNN_3D.py:13), which a tiny
in-memory shim supplies (DropPath is Identity at rate 0, the only rate the fixtures use).

For every case we store: the config, the full ``state_dict``, the inputs, the reference outputs of one
training step (``train_synthetic.py:175-201``: logits, driver mask, z_q, aux loss, total loss), the
gradient of the total loss w.r.t. every parameter, and the eval-mode outputs.  Fixtures are small
(a few hundred KB) ``.npz`` files committed under ``tests/golden/``.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def install_timm_shim():
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    layers.DropPath = DropPath
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models = models
    models.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.models"] = models
    sys.modules["timm.models.layers"] = layers


def import_reference():
    install_timm_shim()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    build = importlib.import_module("models.build")
    losses = importlib.import_module("models.losses")
    config = importlib.import_module("config")
    return build, losses, config


def reference_config(config_mod, **over):
    argv = sys.argv
    sys.argv = ["x"]
    try:
        cfg = config_mod.read_arguments(train=True, print=False, save=False)
    finally:
        sys.argv = argv
    cfg.encoder = "Swin_3D"
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


CASES = {
    # name: (config overrides, N, T, H, W, init kind)
    "small_default": (dict(in_channels_dynamic=2, in_channels=1), 1, 8, 8, 12, "reference"),
    "small_random": (dict(in_channels_dynamic=2, in_channels=1), 2, 8, 8, 12, "random"),
    "pad_shift": (dict(in_channels_dynamic=1, in_channels=1), 1, 8, 10, 14, "random"),      # H,W not multiples of 4
    "two_chan": (dict(in_channels_dynamic=2, in_channels=2), 1, 8, 8, 8, "random"),         # real-data C=2
    "six_vars": (dict(in_channels_dynamic=6, in_channels=1), 1, 8, 8, 8, "random"),         # full V, joint head 96ch
    "long_t": (dict(in_channels_dynamic=1, in_channels=1), 1, 12, 8, 8, "random"),          # T=12: stage-2 pad 12->16
    # BASELINE.json configs[3]: 3D-CNN backbone (models/encoder/CNN_3D.py) instead of the Swin encoder
    "cnn_encoder": (dict(encoder="CNN_3D", in_channels_dynamic=2, in_channels=1), 1, 8, 8, 12, "random"),
    "cnn_encoder_2ch": (dict(encoder="CNN_3D", in_channels_dynamic=2, in_channels=2), 1, 8, 10, 14, "reference"),
    # general LFQ: 2^K-entry codebooks of K-bit sign codes (LFQ.py:92-101,134-146)
    "lfq_4_codes": (dict(in_channels_dynamic=2, in_channels=1, codebook_size=4), 1, 8, 8, 12, "random"),
    "lfq_16_codes": (dict(in_channels_dynamic=1, in_channels=1, codebook_size=16), 2, 8, 8, 8, "random"),
}


def run_case(name, build, losses, config_mod):
    sys.path.insert(0, ROOT)
    from oracle import idee_oracle as O

    over, N, T, H, W, kind = CASES[name]
    cfg = reference_config(config_mod, **over)
    torch.manual_seed(0)
    model = build.VQ_model(cfg)
    ocfg = O.OracleConfig(encoder=cfg.encoder, in_vars=cfg.in_channels_dynamic, in_chans=cfg.in_channels, codebook_size=cfg.codebook_size)
    # parameter inventory of the oracle must equal the reference's
    ref_shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    assert ref_shapes == O.param_shapes(ocfg), "oracle param_shapes() disagrees with the reference"
    if kind == "random":
        sd_new = O.make_state_dict(ocfg, seed=hash(name) % 1000 if False else len(name), kind="random")
        model.load_state_dict(sd_new, strict=False)
    x, m_ext, m_loss = O.make_inputs(ocfg, N, T, H, W, seed=len(name))
    out = {}
    out["cfg_in_vars"] = np.int64(cfg.in_channels_dynamic)
    out["cfg_in_chans"] = np.int64(cfg.in_channels)
    out["cfg_encoder"] = np.array(cfg.encoder)
    out["cfg_codebook_size"] = np.int64(cfg.codebook_size)
    for k, v in model.state_dict().items():
        if k.endswith("relative_position_index") or k == "vq.mask":
            continue
        out["sd/" + k] = v.detach().numpy().copy()
    out["in/x"], out["in/mask_extreme"], out["in/mask_extreme_loss"] = x.numpy(), m_ext.numpy(), m_loss.numpy()

    # ---- one training step, train_synthetic.py:175-203 ----
    model.train()
    crit = losses.BCE_loss_synthetic()
    crit_an = losses.Anomaly_L1_loss_synthetic(n_dynamic=cfg.in_channels_dynamic, delta_t=T, dim=cfg.codebook_dim)
    pred, pred_y, anomaly, z_q, loss_z_q = model(x)
    tgt = m_ext.unsqueeze(1).float()
    loss = crit(pred, tgt)
    vq0 = model.vq.indices_to_codes(torch.Tensor([0]).long()).clone().detach()
    loss_an = crit_an(z_q, m_loss.float(), vq0)
    loss_var = 0
    for k in range(cfg.in_channels_dynamic):
        loss_var = loss_var + crit(pred_y[k], tgt)
    total = loss + loss_an * cfg.lambda_anomaly + loss_var + loss_z_q
    total.backward()
    out["train/pred"] = pred.detach().numpy()
    out["train/pred_y"] = torch.stack([p.detach() for p in pred_y]).numpy()
    out["train/anomaly"] = anomaly.numpy().astype(np.uint8)
    out["train/z_q"] = z_q.detach().numpy()
    out["train/loss_z_q"] = loss_z_q.detach().numpy()
    out["train/loss_bce"] = loss.detach().numpy()
    out["train/loss_anomaly"] = loss_an.detach().numpy()
    out["train/loss_var"] = loss_var.detach().numpy()
    out["train/total"] = total.detach().numpy()
    out["train/vq0"] = vq0.numpy()
    for k, p in model.named_parameters():
        out["grad/" + k] = p.grad.detach().numpy().copy()
    # encoder output (pre-quantiser) for near-tie attribution
    with torch.no_grad():
        out["train/z_enc"] = model.encoder(x).numpy()

    # ---- eval forward, test_synthetic.py:101-124 ----
    model.eval()
    with torch.no_grad():
        pred, pred_y, anomaly, z_q, loss_z_q = model(x)
    out["eval/pred"] = pred.numpy()
    out["eval/pred_y"] = torch.stack(pred_y).numpy()
    out["eval/anomaly"] = anomaly.numpy().astype(np.uint8)
    out["eval/loss_z_q"] = loss_z_q.numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    frac = float(out["train/anomaly"].mean())
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KB)  total={float(total):.6f}  mask==1 frac={frac:.3f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    build, losses, config_mod = import_reference()
    for name in args.cases:
        run_case(name, build, losses, config_mod)


if __name__ == "__main__":
    main()
