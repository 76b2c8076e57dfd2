"""One eager training step of the benchmark workload between cudaProfilerStart / cudaProfilerStop, for ncu:

    ncu --profile-from-start off --set full --clock-control none --import-source on [-k regex:...] \
        -o gpurun_out/step python tools/profile_step.py [--batch 8] [--hw 200]

The step is the one bench.py times (forward, losses, backward, Adam) launched kernel by kernel (no CUDA graph), after two
untimed warm-up steps.  Numbers printed by a run under ncu are never bench values.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=200)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    from idee_b200 import _lib
    from idee_b200.config import default_config
    from idee_b200.models.build import VQ_model
    from idee_b200.trainer import Trainer
    from oracle import idee_oracle as O            # synthetic input factory only

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.set_precision(args.precision)
    cfg = default_config()
    torch.manual_seed(0)
    model = VQ_model(cfg).to(dev).train()
    trainer = Trainer(model, lr=cfg.lr, betas=(cfg.beta1, cfg.beta2), weight_decay=cfg.weight_decay, lambda_anomaly=cfg.lambda_anomaly)
    x, me, ml = (t.to(dev) for t in O.make_inputs(O.OracleConfig(), args.batch, 8, args.hw, args.hw, seed=0))
    for _ in range(2):
        trainer.step(x, me, ml)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loss, _ = trainer.step(x, me, ml)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("loss", float(loss))


if __name__ == "__main__":
    main()
