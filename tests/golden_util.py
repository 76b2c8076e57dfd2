"""Helpers shared by the golden-vector tests (CPU oracle tests and GPU parity tests)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from oracle import idee_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_case(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    enc = str(d["cfg_encoder"]) if "cfg_encoder" in d.files else "Swin_3D"
    cbs = int(d["cfg_codebook_size"]) if "cfg_codebook_size" in d.files else 2
    cfg = O.OracleConfig(encoder=enc, in_vars=int(d["cfg_in_vars"]), in_chans=int(d["cfg_in_chans"]), codebook_size=cbs)
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd/")}
    grads = {k[5:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("grad/")}
    ins = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("in/")}
    train = {k[6:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("train/")}
    ev = {k[5:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("eval/")}
    return cfg, sd, ins, train, ev, grads


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-norm relative error ||a-b||_inf / max(||b||_inf, tiny)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def lfq_scalar(sd, z_enc: torch.Tensor) -> torch.Tensor:
    """pre-quantiser scalar s per token from an encoder output [N,V,C,T,H,W] (LFQ.py:211); for a K-bit codebook the
    projection closest to a sign change, min_i |s_i| (signed), since any bit can flip the index."""
    w, b = sd["vq.project_in.weight"].double().cpu(), sd["vq.project_in.bias"].double().cpu()
    s = torch.einsum("nvcthw,kc->nvthwk", z_enc.double().cpu(), w) + b
    if s.shape[-1] == 1:
        return s[..., 0]
    i = s.abs().argmin(-1, keepdim=True)
    return s.gather(-1, i)[..., 0]


def mask_agreement(anom_a, anom_b, s: torch.Tensor, tie_tol: float):
    """fraction of equal mask elements, and whether every mismatch is a near-tie |s| < tie_tol."""
    a = anom_a.cpu().long().reshape(-1)
    b = anom_b.cpu().long().reshape(-1)
    neq = a != b
    frac = 1.0 - float(neq.float().mean())
    ties_ok = bool((s.reshape(-1)[neq].abs() < tie_tol).all()) if neq.any() else True
    return frac, ties_ok
