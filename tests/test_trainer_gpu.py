"""Step-level parity on the GPU: three optimisation steps of idee_b200.Trainer (flat buffers + fused Adam kernel) against the
CPU oracle driven by torch.optim.Adam with the reference's hyper-parameters (train_synthetic.py:127-129, 175-205)."""
import pytest
import torch

from oracle import idee_oracle as O
from tests.golden_util import rel_err

pytestmark = pytest.mark.gpu


def test_three_adam_steps_match_oracle():
    from idee_b200.config import default_config
    from idee_b200.models.build import VQ_model
    from idee_b200.trainer import Trainer
    cfg = O.OracleConfig(in_vars=2, in_chans=1)
    sd = O.make_state_dict(cfg, seed=2, kind="reference")
    x, m_ext, m_loss = O.make_inputs(cfg, 2, 8, 12, 16, seed=2)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(ref.values()), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.003)
    ref_losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        total, _ = O.train_step_loss(ref, x, m_ext, m_loss, cfg)
        total.backward()
        opt.step()
        ref_losses.append(float(total))
    model = VQ_model(default_config(in_channels_dynamic=2))
    model.load_state_dict(sd, strict=False)
    model = model.cuda().train()
    tr = Trainer(model, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.003, distributed=False)
    losses = []
    for _ in range(3):
        loss, _ = tr.step(x.cuda(), m_ext.cuda(), m_loss.cuda())
        losses.append(float(loss))
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) / abs(b) < 1e-4, (losses, ref_losses)
    named = dict(model.named_parameters())
    worst = max((rel_err(named[k], v), k) for k, v in ref.items())
    assert worst[0] < 1e-4, worst
    # the module parameters are still views of the flat buffer and the state_dict keeps the reference keys
    tr._check_flat()
    assert set(sd) <= set(model.state_dict())
