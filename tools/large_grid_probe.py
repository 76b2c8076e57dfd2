"""Config-5 style stress (BASELINE.json configs[4]): ERA5-Land-like grid 412x424, C=2 channels, B=1: one train step."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib
from idee_b200.config import default_config
from idee_b200.models.build import VQ_model
from idee_b200.trainer import Trainer
_lib.set_precision("bf16")
out = []
for (H, W, T) in ((412, 424, 8), (804, 776, 8)):
    torch.manual_seed(0)
    cfg = default_config(in_channels=2)
    model = VQ_model(cfg).cuda().train()
    tr = Trainer(model, distributed=False)
    x = torch.randn(1, 6, 2, T, H, W, device="cuda").clamp_(-10, 10)
    me = (torch.rand(1, H, W, device="cuda") < 0.05).float(); ml = (torch.rand(1, H, W, device="cuda") < 0.1).float()
    for _ in range(2):
        loss, _ = tr.step(x, me, ml)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        loss, _ = tr.step(x, me, ml)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    out.append({"grid": [H, W], "T": T, "C": 2, "B": 1, "ms_per_step": dt * 1e3, "loss": float(loss),
                "finite": bool(torch.isfinite(loss)), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30})
    del model, tr, x
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
print(json.dumps(out))
