#!/usr/bin/env python
"""One-process-per-GPU twin of the reference's ``train_synthetic.py`` (which wraps the model in nn.DataParallel, :134-135).

    python train_synthetic_ddp.py --epochs 2 --steps-per-epoch 20                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train_synthetic_ddp.py ...

Loop structure, loss assembly, optimiser hyper-parameters and the printed epoch summaries follow train_synthetic.py:156-260; what
differs is the machinery: batch-sharded ranks + one flat NCCL gradient all-reduce (idee_b200.trainer.Trainer), the whole step as one
CUDA graph, evaluators on the device (idee_b200.metrics) with ONE host read per epoch instead of `.item()` / `.cpu()` every step.
The synthetic-CERRA data pipeline (dataset/Synthetic_dataset.py) is out of scope (no data offline): batches are drawn on the device
with the shapes / value ranges of SURVEY.md section 8d; every rank draws its own shard from a rank-dependent seed.
"""
from __future__ import annotations

import argparse
import os
import time

import torch
import torch.distributed as dist


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps-per-epoch", type=int, default=20)
    ap.add_argument("--val-steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per step")
    ap.add_argument("--hw", type=int, default=200)
    ap.add_argument("--encoder", default="Swin_3D", choices=["Swin_3D", "CNN_3D"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--save", default=None, help="checkpoint path (reference format: epoch / model_state_dict / optimizer_state_dict / loss)")
    return ap.parse_args()


def synthetic_batch(cfg, B, hw, gen, dev):
    """x ~ N(0,1) clipped to [-10,10]; extreme masks at 5 % / 10 % density; a random 'true driver' cube."""
    x = torch.randn(B, cfg.in_channels_dynamic, cfg.in_channels, cfg.delta_t, hw, hw, generator=gen, device=dev).clamp_(-10, 10)
    me = (torch.rand(B, hw, hw, generator=gen, device=dev) < 0.05).float()
    ml = (torch.rand(B, hw, hw, generator=gen, device=dev) < 0.10).float()
    drivers = (torch.rand(B, cfg.in_channels_dynamic, cfg.delta_t, hw, hw, generator=gen, device=dev) < 0.5).float()
    return x, me, ml, drivers


def main():
    args = parse_args()
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
    from idee_b200.trainer import Trainer
    _lib.set_precision(args.precision)
    cfg = default_config(encoder=args.encoder, lr=args.lr)
    torch.manual_seed(cfg.seed)
    model = VQ_model(cfg).to(dev).train()
    trainer = Trainer(model, lr=cfg.lr, betas=(cfg.beta1, cfg.beta2), weight_decay=cfg.weight_decay, lambda_anomaly=cfg.lambda_anomaly)
    names = [f"var_{v}" for v in range(cfg.in_channels_dynamic)]
    ev_train, ev_val = ExtremeEvaluator("train", dev), ExtremeEvaluator("val", dev)
    ev_train_an, ev_val_an = AnomalyEvaluator("train", names, dev), AnomalyEvaluator("val", names, dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    log = print if rank == 0 else (lambda *a, **k: None)
    log(f"idee_b200 DDP training: {world} rank(s), batch {args.batch}/GPU, {args.hw}x{args.hw}, encoder {args.encoder}, {args.precision}, "
        f"{sum(p.numel() for p in model.parameters())} parameters")
    best_train, best_val = float("inf"), float("inf")
    use_graph = not args.no_graph
    for epoch in range(args.epochs):
        # cosine schedule with one warm-up epoch, stepped per iteration like the reference's lr_scheduler.step_update (train_synthetic.py:131,176)
        model.train()
        loss_sum = torch.zeros(1, device=dev)
        t0 = time.perf_counter()
        for it in range(args.steps_per_epoch):
            k = epoch * args.steps_per_epoch + it
            total_it = max(args.epochs * args.steps_per_epoch, 1)
            warm = args.steps_per_epoch
            lr = cfg.lr * (k + 1) / warm if k < warm else 0.5 * cfg.lr * (1 + torch.cos(torch.tensor((k - warm) / max(total_it - warm, 1) * 3.141592653589793)).item())
            trainer.set_lr(lr)
            x, me, ml, drivers = synthetic_batch(cfg, args.batch, args.hw, gen, dev)
            if use_graph and trainer._graph is None:
                try:
                    trainer.capture(x, me, ml)
                except Exception as e:                      # capture is an optimisation
                    log(f"CUDA graph capture failed ({type(e).__name__}); eager launches")
                    use_graph = False
            loss, out = (trainer.step_graph if use_graph else trainer.step)(x, me, ml)
            loss_sum += loss.reshape(1)
            ev_train.update(out["pred"], me)
            ev_train_an.update(out["anomaly"], drivers)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            dist.all_reduce(loss_sum)
        mean_train = float(loss_sum) / (args.steps_per_epoch * world)
        # validation (train_synthetic.py:237-260): eval mode, no gradients
        model.eval()
        vloss = torch.zeros(1, device=dev)
        from idee_b200.models.losses import train_step_loss
        with torch.no_grad():
            for _ in range(args.val_steps):
                x, me, ml, drivers = synthetic_batch(cfg, args.batch, args.hw, gen, dev)
                total, out = train_step_loss(model, x, me, ml, cfg.lambda_anomaly)
                vloss += total.reshape(1)
                ev_val.update(out["pred"], me)
                ev_val_an.update(out["anomaly"], drivers)
        if world > 1:
            dist.all_reduce(vloss)
        mean_val = float(vloss) / (max(args.val_steps, 1) * world)
        m_an, _ = ev_train_an.message()
        m_ex, r_ex = ev_train.message(mean_train, min(best_train, mean_train))
        v_an, _ = ev_val_an.message()
        v_ex, _ = ev_val.message(mean_val, min(best_val, mean_val))
        best_train, best_val = min(best_train, mean_train), min(best_val, mean_val)
        log(f"**** EPOCH {epoch + 1:03d} ****  lr {lr:.2e}  {args.steps_per_epoch * args.batch * world / dt:.1f} train samples/s")
        log(m_an); log(m_ex); log(v_an); log(v_ex)
        for e in (ev_train, ev_val, ev_train_an, ev_val_an):
            e.reset()
    if args.save and rank == 0:
        # the reference's checkpoint dict (utils_train.py:576-582); 'module.' prefixes are stripped on load there (build.py:123)
        torch.save({"epoch": args.epochs, "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                    "optimizer_state_dict": {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in trainer.state_dict().items()},
                    "loss": best_val}, args.save)
        log(f"saved {args.save}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
