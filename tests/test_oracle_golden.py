"""Pin the CPU oracle against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Runs on CPU (-m "not gpu")."""
import pytest
import torch

from oracle import idee_oracle as O
from tests.golden_util import CASES, load_case, rel_err, lfq_scalar, mask_agreement

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order only


def test_cases_present():
    assert {"small_default", "small_random", "pad_shift", "two_chan", "six_vars", "long_t"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_train_step_matches_reference(name):
    cfg, sd, ins, train, ev, grads = load_case(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    total, out = O.train_step_loss(sd, ins["x"], ins["mask_extreme"], ins["mask_extreme_loss"], cfg)
    total.backward()
    assert rel_err(out["pred"], train["pred"]) < TOL
    assert rel_err(torch.stack(out["pred_y"]), train["pred_y"]) < TOL
    assert rel_err(out["z_q"], train["z_q"]) < TOL
    assert rel_err(out["z_enc"], train["z_enc"]) < TOL
    assert rel_err(out["loss_z_q"], train["loss_z_q"]) < TOL
    assert rel_err(out["loss_anomaly"], train["loss_anomaly"]) < TOL
    assert rel_err(total, train["total"]) < TOL
    s = lfq_scalar(sd, train["z_enc"])
    frac, ties_ok = mask_agreement(out["anomaly"], train["anomaly"], s, 1e-4)
    assert frac >= 0.999 and ties_ok
    worst = max((rel_err(sd[k].grad, g), k) for k, g in grads.items())
    assert worst[0] < 2e-4, worst


@pytest.mark.parametrize("name", CASES)
def test_oracle_eval_matches_reference(name):
    cfg, sd, ins, train, ev, grads = load_case(name)
    with torch.no_grad():
        zc, ys, anomaly, zq, aux, z_enc = O.vq_model_forward(sd, ins["x"], cfg, training=False)
    assert rel_err(zc, ev["pred"]) < TOL
    assert rel_err(torch.stack(ys), ev["pred_y"]) < TOL
    assert float(aux) == 0.0 and float(ev["loss_z_q"]) == 0.0
    frac, ties_ok = mask_agreement(anomaly, ev["anomaly"], lfq_scalar(sd, z_enc), 1e-4)
    assert frac >= 0.999 and ties_ok


def test_indices_to_codes_is_bout_minus_wout():
    cfg, sd, ins, train, ev, grads = load_case("small_random")
    vq0 = O.lfq_indices_to_codes(sd, torch.tensor([0]), cfg)
    assert torch.allclose(vq0, train["vq0"])
    assert torch.allclose(vq0[0], sd["vq.project_out.bias"] - sd["vq.project_out.weight"][:, 0])


def test_anomaly_loss_closed_form():
    """loss == 2*sum|w_out| * #{q=+1 and m=0} / (C*V*T*sum(1-m)) for binary masks."""
    cfg, sd, ins, train, ev, grads = load_case("small_random")
    anomaly = train["anomaly"].float()
    m = ins["mask_extreme_loss"]
    N, V, T, H, W = anomaly.shape
    cnt = (anomaly * (1 - m).view(N, 1, 1, H, W)).sum()
    closed = 2 * sd["vq.project_out.weight"].abs().sum() * cnt / (16 * V * T * (1 - m).sum())
    assert rel_err(closed, train["loss_anomaly"]) < 1e-5
