"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Per-sample restatement of the metric definitions of the reference's ``DetectorEvaluator.evaluate``
(models/window_evaluator.py:268-483): plain Python loops over numpy logits, one sample at a time, exactly the
bookkeeping the reference does with ``.item()`` calls.  The reference module itself cannot be imported here
(``matplotlib`` missing), so this follows its text; pinned only by that text ("parity unpinned").
Inputs are already-computed logits, so it checks the scoring arithmetic, not the models.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

BUCKETS = ["early", "late", "pre", "noleak"]


def evaluate_logits(batches: Sequence[dict], topk: int, groups: Sequence[str], pipe_dist: Optional[np.ndarray] = None,
                    pipe_rank: Optional[np.ndarray] = None, success_radii_m: Sequence[float] = (50.0, 100.0, 300.0),
                    accuracy_is: Sequence[int] = (1, 5, 10, 20)) -> Dict[str, float]:
    """batches: dicts with ``logits`` (B, C) float, ``label`` (B,) int, ``bucket`` list[str]."""
    groups = set(groups) | {"basic"}
    total = correct1 = correctk = nl_total = nl_correct = leak_total = leak_correct1 = leak_correctk = 0
    tp = fp = fn = tn = leak_pred_as_noleak = pre_fa = noleak_fa = pre_total = noleak_only_total = 0
    b_total = {b: 0 for b in BUCKETS}
    b_c1 = {b: 0 for b in BUCKETS}
    b_ck = {b: 0 for b in BUCKETS}
    b_nl = {b: 0 for b in BUCKETS}
    atd: List[float] = []
    succ = {r: 0 for r in success_radii_m}
    acci = {i: 0 for i in accuracy_is}
    ar: List[int] = []
    for batch in batches:
        logits, label, buckets = np.asarray(batch["logits"]), np.asarray(batch["label"]), batch.get("bucket")
        n_cls = logits.shape[1]
        nlc = n_cls - 1
        for i in range(label.shape[0]):
            y = int(label[i])
            row = logits[i]
            order = sorted(range(n_cls), key=lambda c: (-row[c], c))          # descending, ties by index (torch.topk/argmax)
            p1 = order[0]
            k = min(topk, n_cls)
            ink = y in order[:k]
            total += 1
            correct1 += p1 == y
            correctk += ink
            if y == nlc:
                nl_total += 1
                nl_correct += p1 == y
            else:
                leak_total += 1
                leak_correct1 += p1 == y
                leak_correctk += ink
                pipe_order = sorted(range(nlc), key=lambda c: (-row[c], c))  # stable descending argsort of the pipe logits
                ar.append(pipe_order.index(y) + 1)
            pl, tl = p1 != nlc, y != nlc
            tp += pl and tl
            fp += pl and not tl
            fn += (not pl) and tl
            tn += (not pl) and (not tl)
            leak_pred_as_noleak += (not pl) and tl
            if "bucket" in groups and buckets is not None:
                b = str(buckets[i])
                if b in b_total:
                    b_total[b] += 1
                    b_c1[b] += p1 == y
                    b_ck[b] += ink
                    b_nl[b] += p1 == nlc
                if y == nlc and p1 != nlc:
                    pre_fa += b == "pre"
                    noleak_fa += b == "noleak"
                pre_total += b == "pre"
                noleak_only_total += b == "noleak"
            if pipe_dist is not None and y != nlc:
                d = float("inf") if p1 == nlc else float(pipe_dist[y, p1])
                atd.append(d)
                for r in success_radii_m:
                    succ[r] += d <= r
                if p1 != nlc:
                    for ii in accuracy_is:
                        acci[ii] += ii > 0 and y in list(pipe_rank[p1][:ii])

    def sd(a, b):
        return float(a / b) if b > 0 else 0.0

    out: Dict[str, float] = {
        "acc_top1": sd(correct1, total), f"acc_top{topk}": sd(correctk, total), "noleak_acc": sd(nl_correct, nl_total),
        "leak_acc_top1": sd(leak_correct1, leak_total), f"leak_acc_top{topk}": sd(leak_correctk, leak_total),
        "ar_mean": float(np.mean(ar)) if ar else float("inf"), "ar_median": float(np.median(ar)) if ar else float("inf"),
        "ar_n": float(len(ar)), "n_total": float(total), "n_leak": float(leak_total), "n_noleak": float(nl_total)}
    if "binary" in groups:
        prec, rec = sd(tp, tp + fp), sd(tp, tp + fn)
        out.update({"det_precision": prec, "det_recall": rec, "det_f1": sd(2 * prec * rec, prec + rec) if prec + rec > 0 else 0.0,
                    "det_tp": float(tp), "det_fp": float(fp), "det_fn": float(fn), "det_tn": float(tn),
                    "leak_pred_as_noleak_rate": sd(leak_pred_as_noleak, leak_total)})
        if pre_total > 0:
            out["pre_false_alarm_rate"] = sd(pre_fa, pre_total)
        if noleak_only_total > 0:
            out["noleak_false_alarm_rate"] = sd(noleak_fa, noleak_only_total)
    if "bucket" in groups:
        for b in BUCKETS:
            out[f"{b}_n"] = float(b_total[b])
            out[f"{b}_acc_top1"] = sd(b_c1[b], b_total[b])
            out[f"{b}_acc_top{topk}"] = sd(b_ck[b], b_total[b])
            out[f"{b}_pred_as_noleak_rate"] = sd(b_nl[b], b_total[b])
    if pipe_dist is not None:
        if "atd" in groups:
            finite = [x for x in atd if math.isfinite(x)]
            out["atd_mean_m"] = float(np.mean(finite)) if finite else float("inf")
            out["atd_median_m"] = float(np.median(finite)) if finite else float("inf")
            out["atd_n"] = float(len(atd))
            out["atd_missed_rate"] = sd(sum(1 for x in atd if not math.isfinite(x)), len(atd))
        if "success" in groups:
            for r in success_radii_m:
                out[f"success_at_{int(r)}"] = sd(succ[r], leak_total - leak_pred_as_noleak)
                out[f"success_at_{int(r)}_e2e"] = sd(succ[r], leak_total)
        if "accuracy_i" in groups:
            for ii in accuracy_is:
                out[f"accuracy_{int(ii)}"] = sd(acci[ii], leak_total)
    return out
