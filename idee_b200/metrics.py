"""On-device ports of the reference's per-step evaluators (utils/utils_train.py:269-554).

The reference pulls ``pred.cpu()`` / ``anomaly.cpu().numpy()`` after every step and counts confusion entries with numpy
(train_synthetic.py:207-215): at GPU step times of a few milliseconds those host round trips dominate.  Here the counters are
int64 tensors ON THE DEVICE, updated by a handful of fused torch reductions per step (no host synchronisation), summed across
ranks with ONE all-reduce per evaluator at the end of an epoch, and read by the host once.  Formulas and the log format are the
reference's (file:line cited per method), so the printed epoch summaries are interchangeable.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def _nanmean(xs):
    xs = [x for x in xs if not math.isnan(x)]
    return sum(xs) / len(xs) if xs else float("nan")


def _div(a, b):
    return a / b if b else float("nan")


class _Counters:
    """int64 counters on the device; ``reduce()`` sums them over the process group (no-op on one rank)."""

    def __init__(self, n, device):
        self.c = torch.zeros(n, dtype=torch.int64, device=device)

    def reset(self):
        self.c.zero_()

    def reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.c, op=dist.ReduceOp.SUM, group=group)
        return self.c.tolist()               # the ONE host read of the epoch


class ExtremeEvaluator:
    """evaluator_synthetic (utils_train.py:269-351): one class, the extreme-event map of the last time step."""
    classes = [u' Δt0']

    def __init__(self, mode: str, device):
        self.mode = mode
        self.k = _Counters(5, device)        # correct | seen | iou_de | predicted | seen_all

    def reset(self):
        self.k.reset()

    @torch.no_grad()
    def update(self, pred_logits: torch.Tensor, gt: torch.Tensor):
        """pred_logits [N,1,H,W] (pre-sigmoid), gt [N,H,W] or [N,1,H,W] in {0,1}.  sigmoid(x) > 0.5 <=> x > 0
        (train_synthetic.py:209-212); counts as utils_train.py:343-351."""
        p = pred_logits.reshape(pred_logits.shape[0], -1) > 0
        g = gt.reshape(gt.shape[0], -1) == 1
        self.k.c += torch.stack([(p & g).sum(), g.sum(), (p | g).sum(), p.sum(), torch.tensor(g.numel(), device=g.device)]).to(torch.int64)

    def results(self, group=None):
        correct, seen, iou_de, predicted, seen_all = self.k.reduce(group)
        precision = _div(correct, float(predicted))
        accuracy = correct / (float(seen) + 1e-6)
        f1 = _div(2 * precision * accuracy, accuracy + precision)
        return {"weight": _div(seen, seen_all / 1.0), "precision": precision, "accuracy": accuracy, "F1": f1, "IoU": _div(correct, float(iou_de)),
                "seen_all": seen_all}

    def message(self, mean_loss, best_loss, group=None):
        """The reference's epoch summary (utils_train.py:297-320)."""
        r = self.results(group)
        m = '-----------------   %s   -----------------\n' % self.mode
        m += 'class %s weight: %.4f, precision: %.4f, accuracy: %.4f, F1: %.4f IoU: %.4f \n' % (
            self.classes[0] + ' ' * (14 - len(self.classes[0])), r["weight"], r["precision"], r["accuracy"], r["F1"], r["IoU"])
        m += '\n%s mean accuracy : %.4f' % (self.mode, _nanmean([r["accuracy"]]))
        m += '\n%s mean IoU      : %.4f' % (self.mode, _nanmean([r["IoU"]]))
        m += '\n%s mean F1       : %.4f' % (self.mode, _nanmean([r["F1"]]))
        m += '\n%s mean loss     : %.4f' % (self.mode, mean_loss)
        m += '\n%s best mean loss: %.4f\n' % (self.mode, best_loss)
        return m, r


class AnomalyEvaluator:
    """evaluator_anomaly_synthetic (utils_train.py:354-527): per-variable confusion counts of the binary driver masks."""
    PER = 10     # correct_pos seen_pos iou_de_pos predicted_pos correct_neg seen_neg iou_de_neg predicted_neg FP FN

    def __init__(self, mode: str, variables, device):
        self.mode, self.classes = mode, list(variables)
        self.V = len(self.classes)
        self.k = _Counters(self.PER * self.V + 6, device)   # ... | correct_all seen_all correct_p_all seen_p_all iou_de_all predicted_all

    def reset(self):
        self.k.reset()

    @torch.no_grad()
    def update(self, pred: torch.Tensor, gt: torch.Tensor):
        """pred, gt [N, V, ...] in {0,1} (utils_train.py:499-527)."""
        V = self.V
        p1 = (pred == 1).transpose(0, 1).reshape(V, -1)
        g1 = (gt == 1).transpose(0, 1).reshape(V, -1)
        p0 = (pred == 0).transpose(0, 1).reshape(V, -1)
        g0 = (gt == 0).transpose(0, 1).reshape(V, -1)
        per = torch.stack([(p1 & g1).sum(1), g1.sum(1), (p1 | g1).sum(1), p1.sum(1), (p0 & g0).sum(1), g0.sum(1), (p0 | g0).sum(1), p0.sum(1),
                           (p1 & g0).sum(1), (p0 & g1).sum(1)], dim=1).reshape(-1)
        tot = torch.stack([(pred == gt).sum(), torch.tensor(gt.numel(), device=gt.device), (p1 & g1).sum(), g1.sum(), (p1 | g1).sum(), p1.sum()])
        self.k.c += torch.cat([per, tot]).to(torch.int64)

    def results(self, group=None):
        c = self.k.reduce(group)
        V, P = self.V, self.PER
        correct_all, seen_all, correct_p_all, seen_p_all, iou_de_all, predicted_all = c[P * V:]
        out = {"accuracy": _div(correct_all, float(seen_all)), "vars": []}
        for v in range(V):
            cp, sp, ip, pp, cn, sn, inn, pn, fp, fn = c[P * v:P * (v + 1)]
            prec_p, acc_p = _div(cp, float(pp)), cp / (float(sp) + 1e-6)
            prec_n, acc_n = _div(cn, float(pn)), cn / (float(sn) + 1e-6)
            out["vars"].append({"name": self.classes[v], "weight_pos": _div(sp, seen_all / V), "precision_pos": prec_p, "accuracy_pos": acc_p,
                                "F1_pos": _div(2 * prec_p * acc_p, acc_p + prec_p), "IoU_pos": _div(cp, float(ip)),
                                "weight_neg": _div(sn, seen_all / V), "precision_neg": prec_n, "accuracy_neg": acc_n,
                                "F1_neg": _div(2 * prec_n * acc_n, acc_n + prec_n), "IoU_neg": _div(cn, float(inn)),
                                "TP": cp, "FP": fp, "TN": cn, "FN": fn})
        prec, acc = _div(correct_p_all, float(predicted_all)), correct_p_all / (float(seen_p_all) + 1e-6)
        out["all"] = {"weight": _div(seen_p_all, seen_all), "precision": prec, "accuracy": acc, "F1": _div(2 * prec * acc, acc + prec),
                      "IoU": _div(correct_p_all, float(iou_de_all))}
        return out

    def message(self, group=None):
        """The reference's epoch summary (utils_train.py:406-460)."""
        r = self.results(group)
        m = '-----------------   %s   -----------------\n' % self.mode
        for d in r["vars"]:
            n = d["name"]
            m += 'class %s pos   weight: %.4f, precision: %.4f, accuracy: %.4f, F1: %.4f IoU: %.4f \n' % (
                n + ' ' * (7 - len(n)), d["weight_pos"], d["precision_pos"], d["accuracy_pos"], d["F1_pos"], d["IoU_pos"])
            m += ' ' * (13 + 7 - len(n)) + 'neg   weight: %.4f, precision: %.4f, accuracy: %.4f, F1: %.4f IoU: %.4f \n' % (
                d["weight_neg"], d["precision_neg"], d["accuracy_neg"], d["F1_neg"], d["IoU_neg"])
        m += '\n'
        for d in r["vars"]:
            n = d["name"]
            m += 'class %s weight: %.4f, TP: %i, FP: %i, TN: %i FN: %i, F1: %.4f, IoU: %.4f \n' % (
                n + ' ' * (13 - len(n)), d["weight_pos"], d["TP"], d["FP"], d["TN"], d["FN"], d["F1_pos"], d["IoU_pos"])
        a = r["all"]
        m += '\nall var             weight: %.4f, precision: %.4f, accuracy: %.4f, F1: %.4f IoU: %.4f \n' % (
            a["weight"], a["precision"], a["accuracy"], a["F1"], a["IoU"])
        m += '\n%s accuracy               : %.4f' % (self.mode, r["accuracy"])
        m += '\n%s mean accuracy positive : %.4f' % (self.mode, _nanmean([d["accuracy_pos"] for d in r["vars"]]))
        m += '\n%s mean IoU positive      : %.4f' % (self.mode, _nanmean([d["IoU_pos"] for d in r["vars"]]))
        m += '\n%s mean F1 positive       : %.4f' % (self.mode, _nanmean([d["F1_pos"] for d in r["vars"]]))
        return m, r


class AnomalyCollector:
    """anomaly_collector (utils_train.py:530-554): every sample's driver mask covers delta_t time steps ending at its time stamp;
    overlapping windows vote, ``majority_vote`` thresholds the mean at 0.5.  Sums and counts live on the device; with several
    ranks each rank collects its own samples and ``majority_vote`` all-reduces both before thresholding."""

    def __init__(self, shape, delta_t: int, device):
        """shape = (V, T_total, H, W) of the data set's anomaly cube."""
        self.delta_t = delta_t
        self.sum = torch.zeros(shape, dtype=torch.float32, device=device)
        self.count = torch.zeros(shape, dtype=torch.float32, device=device)

    def reset(self):
        self.sum.zero_()
        self.count.zero_()

    @torch.no_grad()
    def update(self, anomaly: torch.Tensor, time_index: torch.Tensor):
        """anomaly [N,V,delta_t,H,W] (driver masks), time_index [N] = position of each sample's time stamp in the cube; the window
        idx-delta_t+1 .. idx receives the mask flipped along time (utils_train.py:547-554)."""
        a = torch.flip(anomaly.to(self.sum.dtype), dims=(2,))
        for n, idx in enumerate(time_index.tolist()):
            lo = idx - self.delta_t + 1
            self.sum[:, lo:idx + 1] += a[n]
            self.count[:, lo:idx + 1] += 1

    @torch.no_grad()
    def majority_vote(self, group=None) -> torch.Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.sum, group=group)
            dist.all_reduce(self.count, group=group)
        mean = self.sum / self.count                      # 0/0 -> NaN where no window voted (:541-545: stays NaN there; 0 here)
        return (mean >= 0.5).to(torch.float32)
