#!/usr/bin/env python
"""One-process-per-GPU twin of the reference's ``test_synthetic.py`` (:80-140): eval-mode forward over a (synthetic) test split,
extreme-event and driver metrics counted on the device, one host read at the end; loads a checkpoint in the reference's format.

    python test_synthetic_ddp.py --steps 20 [--checkpoint ckpt.pth] [--encoder CNN_3D]
"""
from __future__ import annotations

import argparse
import os
import time

import torch
import torch.distributed as dist

from train_synthetic_ddp import synthetic_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=200)
    ap.add_argument("--encoder", default="Swin_3D", choices=["Swin_3D", "CNN_3D"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--checkpoint", default=None)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from idee_b200 import _lib
    from idee_b200.config import default_config
    from idee_b200.metrics import AnomalyEvaluator, ExtremeEvaluator
    from idee_b200.models.build import VQ_model
    _lib.set_precision(args.precision)
    cfg = default_config(encoder=args.encoder)
    torch.manual_seed(cfg.seed)
    model = VQ_model(cfg)
    if args.checkpoint:
        ck = torch.load(args.checkpoint, map_location="cpu")
        model.load_state_dict({k.replace("module.", ""): v for k, v in ck["model_state_dict"].items()}, strict=True)   # build.py:120-127
    model = model.to(dev).eval()
    names = [f"var_{v}" for v in range(cfg.in_channels_dynamic)]
    ev, ev_an = ExtremeEvaluator("test", dev), AnomalyEvaluator("test", names, dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(args.steps):
            x, me, ml, drivers = synthetic_batch(cfg, args.batch, args.hw, gen, dev)
            pred, _, anomaly, _, _ = model(x)                                                         # test_synthetic.py:101-124
            ev.update(pred, me)
            ev_an.update(anomaly, drivers)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    m_an, _ = ev_an.message()
    m_ex, _ = ev.message(float("nan"), float("nan"))
    if rank == 0:
        print(m_an); print(m_ex)
        print(f"{args.steps * args.batch * world / dt:.1f} test samples/s on {world} GPU(s)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
