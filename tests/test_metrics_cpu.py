"""idee_b200.metrics (device-side evaluators) against a numpy restatement of the reference's counters and formulas
(utils/utils_train.py:269-554), on CPU tensors; plus the 2-rank all-reduce of the counters over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idee_b200.metrics import AnomalyCollector, AnomalyEvaluator, ExtremeEvaluator


def _ref_extreme(pred_c, gt):          # utils_train.py:343-351 + 301-309
    correct = np.sum((pred_c[:, 0] == 1) & (gt[:, 0] == 1)); seen = np.sum(gt[:, 0] == 1)
    iou_de = np.sum((pred_c[:, 0] == 1) | (gt[:, 0] == 1)); predicted = np.sum(pred_c[:, 0] == 1)
    precision = correct / float(predicted); accuracy = correct / (float(seen) + 1e-6)
    return dict(precision=precision, accuracy=accuracy, F1=2 * precision * accuracy / (accuracy + precision), IoU=correct / float(iou_de),
                weight=seen / (gt.size / 1))


def test_extreme_evaluator_matches_reference_formulas():
    g = torch.Generator().manual_seed(0)
    ev = ExtremeEvaluator("train", "cpu")
    preds, gts = [], []
    for _ in range(3):
        logit = torch.randn(2, 1, 9, 11, generator=g)
        gt = (torch.rand(2, 9, 11, generator=g) < 0.3).float()
        ev.update(logit, gt)
        p = torch.sigmoid(logit); pc = p.clone(); pc[p > 0.5] = 1; pc[p <= 0.5] = 0          # train_synthetic.py:209-212
        preds.append(pc.numpy()); gts.append(gt.unsqueeze(1).numpy())
    want = _ref_extreme(np.concatenate(preds), np.concatenate(gts))
    msg, got = ev.message(1.25, 1.0)
    for k, v in want.items():
        assert abs(got[k] - v) < 1e-9, (k, got[k], v)
    assert "train mean F1       : %.4f" % want["F1"] in msg and "train mean loss     : 1.2500" in msg
    ev.reset()
    assert ev.results()["seen_all"] == 0


def test_anomaly_evaluator_matches_reference_counts():
    g = torch.Generator().manual_seed(1)
    names = ["var_%d" % i for i in range(3)]
    ev = AnomalyEvaluator("val", names, "cpu")
    P = (torch.rand(4, 3, 8, 5, 6, generator=g) < 0.4).float()
    G = (torch.rand(4, 3, 8, 5, 6, generator=g) < 0.5).float()
    ev.update(P[:2], G[:2]); ev.update(P[2:], G[2:])
    p, gnp = P.numpy(), G.numpy()
    msg, r = ev.message()
    assert abs(r["accuracy"] - np.sum(p == gnp) / gnp.size) < 1e-12                         # utils_train.py:501-504, 392
    for v in range(3):                                                                       # :506-520, 396-404
        cp = np.sum((p[:, v] == 1) & (gnp[:, v] == 1)); sp = np.sum(gnp[:, v] == 1); pp = np.sum(p[:, v] == 1)
        cn = np.sum((p[:, v] == 0) & (gnp[:, v] == 0)); ip = np.sum((p[:, v] == 1) | (gnp[:, v] == 1))
        d = r["vars"][v]
        assert (d["TP"], d["TN"]) == (cp, cn)
        assert d["FP"] == np.sum((p[:, v] == 1) & (gnp[:, v] == 0)) and d["FN"] == np.sum((p[:, v] == 0) & (gnp[:, v] == 1))
        prec, acc = cp / float(pp), cp / (float(sp) + 1e-6)
        assert abs(d["precision_pos"] - prec) < 1e-12 and abs(d["accuracy_pos"] - acc) < 1e-12
        assert abs(d["F1_pos"] - 2 * prec * acc / (acc + prec)) < 1e-12 and abs(d["IoU_pos"] - cp / float(ip)) < 1e-12
        assert abs(d["weight_pos"] - sp / (gnp.size / 3)) < 1e-12
    cpa = np.sum((p == 1) & (gnp == 1))                                                       # :522-526, 406-409
    assert abs(r["all"]["IoU"] - cpa / float(np.sum((p == 1) | (gnp == 1)))) < 1e-12
    assert "val mean F1 positive" in msg and "all var" in msg


def test_anomaly_collector_majority_vote():
    V, T, H, W, dt = 2, 12, 3, 4, 4
    g = torch.Generator().manual_seed(2)
    col = AnomalyCollector((V, T, H, W), dt, "cpu")
    ref_sum, ref_cnt = np.zeros((V, T, H, W)), np.zeros((V, T, H, W))
    for idx in (3, 4, 5, 8, 11):                                                             # utils_train.py:547-554
        a = (torch.rand(1, V, dt, H, W, generator=g) < 0.5).float()
        col.update(a, torch.tensor([idx]))
        ref_sum[:, idx - dt + 1:idx + 1] += np.flip(a[0].numpy(), axis=1)
        ref_cnt[:, idx - dt + 1:idx + 1] += 1
    with np.errstate(invalid="ignore"):
        mean = ref_sum / ref_cnt
    want = np.where(np.isnan(mean), 0.0, (mean >= 0.5).astype(np.float32))                    # :541-545 (unvoted positions: 0 here)
    assert np.array_equal(col.majority_vote().numpy(), want)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        logit = torch.randn(4, 1, 6, 6, generator=g); gt = (torch.rand(4, 6, 6, generator=g) < 0.3).float()
        ev = ExtremeEvaluator("train", "cpu")
        ev.update(logit[rank * 2:(rank + 1) * 2], gt[rank * 2:(rank + 1) * 2])                # each rank sees its shard
        full = ExtremeEvaluator("train", "cpu"); full.update(logit, gt)
        out[rank] = ev.k.reduce() == full.k.c.tolist()          # sum of the shards' counters == counters of the whole batch
    finally:
        dist.destroy_process_group()


def test_counters_all_reduce_over_two_ranks():
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        assert out[0] and out[1]
