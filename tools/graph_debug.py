import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idee_b200 import _lib
from idee_b200.config import default_config
from idee_b200.models.build import VQ_model
from idee_b200.trainer import Trainer
from oracle import idee_oracle as O

_lib.set_precision("bf16")
for V, B, HW in ((3, 2, 24), (6, 2, 24), (6, 8, 200)):
    cfg = O.OracleConfig(in_vars=V, in_chans=1)
    torch.manual_seed(0)
    model = VQ_model(default_config(in_channels_dynamic=V)).cuda().train()
    tr = Trainer(model, distributed=False)
    x, me, ml = (t.cuda() for t in O.make_inputs(cfg, B, 8, HW, HW, seed=0))
    for mode in ("global", "thread_local", "relaxed"):
        try:
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    tr.step(x, me, ml)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                tr.step(x, me, ml)
            g.replay(); torch.cuda.synchronize()
            print(f"V={V} B={B} HW={HW} mode={mode}: capture OK, mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB")
            break
        except Exception as e:
            print(f"V={V} B={B} HW={HW} mode={mode}: FAILED {type(e).__name__}: {str(e)[:200]}")
            tb = traceback.format_exc().splitlines()
            print("   ", [l for l in tb if "idee_b200" in l or "ops.py" in l][-4:])
            torch.cuda.synchronize()
    del tr, model
    torch.cuda.empty_cache()
