"""Window-level evaluation without per-sample host round trips.

Same constructor arguments, ``evaluate(loader) -> Dict[str, float]`` contract and metric names as the reference's
``DetectorEvaluator`` (models/window_evaluator.py:226-483; groups ``basic``, ``binary``, ``bucket``, ``atd``,
``success``, ``accuracy_i``).  The reference walks every sample of every batch in Python and calls ``.item()`` about six
times per sample (window_evaluator.py:376-418), which at ``val_steps = 10 000`` costs more than the detector itself; here
every count is a tensor expression on the device, accumulated across batches, and the host reads the totals ONCE at the end.

Distances: the reference builds a Dijkstra pipe-distance oracle from the ``.inp`` geometry (window_evaluator.py:112-224).
That is host-side preprocessing outside the hot path; this evaluator takes its result as a dense matrix
``pipe_dist[P, P]`` (metres; ``inf`` where unreachable) -- e.g. ``oracle.pipe_dist`` of the reference object -- and derives
the rank table the ``accuracy_i`` metric needs.  Without it the three distance groups are skipped.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from .models.utils import build_residual_sequence_from_segment

BUCKETS = ("early", "late", "pre", "noleak")


class DetectorEvaluator:
    def __init__(self, predictor: torch.nn.Module, detector: torch.nn.Module, device: torch.device, *, l_pred: int,
                 l_det: int, topk: int = 5, metric_groups: Sequence[str] = ("basic", "binary", "bucket"),
                 pipe_dist: Optional[Any] = None, success_radii_m: Sequence[float] = (50.0, 100.0, 300.0),
                 accuracy_is: Sequence[int] = (1, 5, 10, 20),
                 residual_builder: Callable = build_residual_sequence_from_segment) -> None:
        self.predictor, self.detector, self.device = predictor, detector, torch.device(device)
        self.l_pred, self.l_det, self.topk = int(l_pred), int(l_det), int(topk)
        self.metric_groups = set(metric_groups) | {"basic"}
        if ({"atd", "success", "accuracy_i"} & self.metric_groups) and pipe_dist is None:
            raise ValueError("Distance metrics requested but no pipe_dist matrix given.")
        self.success_radii_m = tuple(float(r) for r in success_radii_m)
        self.accuracy_is = tuple(int(i) for i in accuracy_is)
        self.residual_builder = residual_builder
        self.pipe_dist = self.rank_of = None
        if pipe_dist is not None:
            d = torch.as_tensor(np.asarray(pipe_dist), dtype=torch.float32)
            self.pipe_dist = d.to(self.device)
            # rank_of[p, y] = position of pipe y among the pipes ordered by distance from p (0 = nearest, p itself);
            # "y in oracle.pipe_rank[p][:i]" (window_evaluator.py:414-418) <=> rank_of[p, y] < i
            order = torch.argsort(d, dim=1, stable=True)
            rank = torch.empty_like(order)
            rank.scatter_(1, order, torch.arange(d.shape[1]).unsqueeze(0).expand_as(order))
            self.rank_of = rank.to(self.device)

    @torch.no_grad()
    def evaluate(self, loader: Iterable[Dict[str, Any]]) -> Dict[str, float]:
        self.predictor.eval()
        self.detector.eval()
        dev = self.device
        names = ["total", "correct1", "correctk", "nl_total", "nl_correct", "leak_correct1", "leak_correctk", "tp", "fp", "fn",
                 "tn", "pre_fa", "noleak_fa"]
        acc = {n: torch.zeros((), dtype=torch.int64, device=dev) for n in names}
        b_cnt = torch.zeros(len(BUCKETS), 4, dtype=torch.int64, device=dev)   # total, correct1, correctk, pred-as-noleak
        ranks, dists = [], []
        success = torch.zeros(len(self.success_radii_m), dtype=torch.int64, device=dev)
        acc_i = torch.zeros(len(self.accuracy_is), dtype=torch.int64, device=dev)
        use_dist = self.pipe_dist is not None and bool({"atd", "success", "accuracy_i"} & self.metric_groups)
        radii = torch.tensor(self.success_radii_m, device=dev)
        iis = torch.tensor(self.accuracy_is, device=dev)

        for batch in loader:
            noisy_seg = batch["noisy_seg"].to(dev)
            time_seg = batch["time_seg"].to(dev)
            label = torch.as_tensor(batch["label"], device=dev, dtype=torch.long)
            num_classes = batch.get("num_classes", None)
            if isinstance(num_classes, (list, tuple)):
                num_classes = int(num_classes[0])
            elif torch.is_tensor(num_classes):
                num_classes = int(num_classes.reshape(-1)[0])          # host tensor from the loader: no device sync
            else:
                num_classes = None
            residual = self.residual_builder(self.predictor, noisy_seg, time_seg, l_pred=self.l_pred, l_det=self.l_det,
                                             device=dev)
            logits = self.detector(residual, time_seg[:, self.l_pred:, :].contiguous())
            if num_classes is None:
                num_classes = int(logits.size(-1))
            nlc = num_classes - 1

            pred1 = logits.argmax(dim=-1)
            k = min(self.topk, logits.size(-1))
            hit1 = pred1 == label
            hitk = (logits.topk(k=k, dim=-1).indices == label.unsqueeze(1)).any(dim=1)
            is_nl, pred_nl = label == nlc, pred1 == nlc
            is_leak, pred_leak = ~is_nl, ~pred_nl
            acc["total"] += label.numel()
            acc["correct1"] += hit1.sum()
            acc["correctk"] += hitk.sum()
            acc["nl_total"] += is_nl.sum()
            acc["nl_correct"] += (hit1 & is_nl).sum()
            acc["leak_correct1"] += (hit1 & is_leak).sum()
            acc["leak_correctk"] += (hitk & is_leak).sum()
            acc["tp"] += (pred_leak & is_leak).sum()
            acc["fp"] += (pred_leak & is_nl).sum()
            acc["fn"] += (pred_nl & is_leak).sum()
            acc["tn"] += (pred_nl & is_nl).sum()

            # average rank of the true pipe among the pipe logits (1 = best), leak samples only
            pipe_logits = logits[:, :nlc]
            true_logit = pipe_logits.gather(1, label.clamp_max(nlc - 1).unsqueeze(1))
            # rank under a descending stable argsort = 1 + #(strictly greater) + #(equal, with a smaller index)
            idx = torch.arange(nlc, device=dev).unsqueeze(0)
            r1 = 1 + (pipe_logits > true_logit).sum(1) + ((pipe_logits == true_logit) & (idx < label.unsqueeze(1))).sum(1)
            ranks.append(r1[is_leak])

            if "bucket" in self.metric_groups:
                bl = batch.get("bucket", None)
                if bl is not None:
                    code = torch.tensor([BUCKETS.index(str(b)) if str(b) in BUCKETS else -1 for b in bl], device=dev)
                    onehot = (code.unsqueeze(1) == torch.arange(len(BUCKETS), device=dev).unsqueeze(0)).to(torch.int64)
                    cols = torch.stack([torch.ones_like(hit1), hit1, hitk, pred_nl], dim=1).to(torch.int64)
                    b_cnt += (onehot.unsqueeze(2) * cols.unsqueeze(1)).sum(0)   # (integer matmul does not exist on CUDA)
                    false_alarm = is_nl & pred_leak
                    acc["pre_fa"] += (false_alarm & (code == BUCKETS.index("pre"))).sum()
                    acc["noleak_fa"] += (false_alarm & (code == BUCKETS.index("noleak"))).sum()

            if use_dist:
                y, p = label[is_leak], pred1[is_leak]
                detected = p != nlc
                d = torch.full(y.shape, float("inf"), device=dev)
                d[detected] = self.pipe_dist[y[detected], p[detected]]
                dists.append(d)
                success += (d.unsqueeze(1) <= radii.unsqueeze(0)).sum(0)
                rk = self.rank_of[p[detected], y[detected]]
                acc_i += ((rk.unsqueeze(1) < iis.unsqueeze(0)) & (iis.unsqueeze(0) > 0)).sum(0)

        # ---- the only host reads ----
        c = {n: int(v.item()) for n, v in acc.items()}
        b_cnt = b_cnt.cpu().numpy()
        ar = torch.cat(ranks).cpu().numpy() if ranks else np.zeros(0)
        leak_total = c["total"] - c["nl_total"]

        def safe_div(a: float, b: float) -> float:
            return float(a / b) if b > 0 else 0.0

        out: Dict[str, float] = {}
        if "basic" in self.metric_groups:
            out.update({
                "acc_top1": safe_div(c["correct1"], c["total"]),
                f"acc_top{self.topk}": safe_div(c["correctk"], c["total"]),
                "noleak_acc": safe_div(c["nl_correct"], c["nl_total"]),
                "leak_acc_top1": safe_div(c["leak_correct1"], leak_total),
                f"leak_acc_top{self.topk}": safe_div(c["leak_correctk"], leak_total),
                "ar_mean": float(np.mean(ar)) if ar.size else float("inf"),
                "ar_median": float(np.median(ar)) if ar.size else float("inf"),
                "ar_n": float(ar.size),
                "n_total": float(c["total"]), "n_leak": float(leak_total), "n_noleak": float(c["nl_total"]),
            })
        if "binary" in self.metric_groups:
            prec, rec = safe_div(c["tp"], c["tp"] + c["fp"]), safe_div(c["tp"], c["tp"] + c["fn"])
            out.update({
                "det_precision": prec, "det_recall": rec,
                "det_f1": safe_div(2 * prec * rec, prec + rec) if (prec + rec) > 0 else 0.0,
                "det_tp": float(c["tp"]), "det_fp": float(c["fp"]), "det_fn": float(c["fn"]), "det_tn": float(c["tn"]),
                "leak_pred_as_noleak_rate": safe_div(c["fn"], leak_total),
            })
            pre_total, noleak_total = int(b_cnt[BUCKETS.index("pre"), 0]), int(b_cnt[BUCKETS.index("noleak"), 0])
            if pre_total > 0:
                out["pre_false_alarm_rate"] = safe_div(c["pre_fa"], pre_total)
            if noleak_total > 0:
                out["noleak_false_alarm_rate"] = safe_div(c["noleak_fa"], noleak_total)
        if "bucket" in self.metric_groups:
            for i, b in enumerate(BUCKETS):
                n = int(b_cnt[i, 0])
                out[f"{b}_n"] = float(n)
                out[f"{b}_acc_top1"] = safe_div(int(b_cnt[i, 1]), n)
                out[f"{b}_acc_top{self.topk}"] = safe_div(int(b_cnt[i, 2]), n)
                out[f"{b}_pred_as_noleak_rate"] = safe_div(int(b_cnt[i, 3]), n)
        if use_dist:
            d_all = torch.cat(dists).cpu().numpy().astype(np.float64) if dists else np.zeros(0)
            finite = d_all[np.isfinite(d_all)]
            if "atd" in self.metric_groups:
                out["atd_mean_m"] = float(np.mean(finite)) if finite.size else float("inf")
                out["atd_median_m"] = float(np.median(finite)) if finite.size else float("inf")
                out["atd_n"] = float(d_all.size)
                out["atd_missed_rate"] = safe_div(float(d_all.size - finite.size), d_all.size)
            if "success" in self.metric_groups:
                hits = success.cpu().numpy()
                for r, h in zip(self.success_radii_m, hits):
                    out[f"success_at_{int(r)}"] = safe_div(int(h), leak_total - c["fn"])
                    out[f"success_at_{int(r)}_e2e"] = safe_div(int(h), leak_total)
            if "accuracy_i" in self.metric_groups:
                hits = acc_i.cpu().numpy()
                for ii, h in zip(self.accuracy_is, hits):
                    out[f"accuracy_{int(ii)}"] = safe_div(int(h), leak_total)
        return out


__all__ = ["DetectorEvaluator", "BUCKETS"]
