"""The vectorised window evaluator (leak_det_gnn_b200/evaluator.py) against the per-sample restatement of the
reference's metric bookkeeping (oracle/evaluator_oracle.py, after models/window_evaluator.py:268-483): every metric of
every group, equal to the last bit of the float division, on random logits with ties, all four buckets and a random
pipe-distance matrix with unreachable pairs."""
import numpy as np
import pytest
import torch

from leak_det_gnn_b200.evaluator import BUCKETS, DetectorEvaluator
from oracle.evaluator_oracle import evaluate_logits


class _Replay(torch.nn.Module):
    """Stands in for predictor + residual builder + detector: returns the prepared logits of each batch in turn."""

    def __init__(self, logits):
        super().__init__()
        self.logits, self.i = logits, 0

    def forward(self, residual, tfeat):
        out = self.logits[self.i]
        self.i += 1
        return out


@pytest.mark.parametrize("n_cls,topk", [(3, 5), (41, 5), (765, 5), (12, 1)])
def test_vectorised_metrics_equal_per_sample_bookkeeping(n_cls, topk):
    rng = np.random.default_rng(n_cls)
    gen = torch.Generator().manual_seed(n_cls)
    batches, feed = [], []
    for b in (7, 32, 1, 19):
        logits = torch.randn(b, n_cls, generator=gen)
        logits[:, ::3] = logits[:, ::3].round()                      # exact ties
        label = torch.randint(0, n_cls, (b,), generator=gen)
        label[rng.random(b) < 0.3] = n_cls - 1                        # a fair share of no-leak windows
        bucket = [str(rng.choice(BUCKETS + ("unknown",))) for _ in range(b)]
        batches.append({"logits": logits.numpy(), "label": label.numpy(), "bucket": bucket})
        feed.append({"noisy_seg": torch.zeros(b, 4, 2), "time_seg": torch.zeros(b, 4, 9), "label": label, "bucket": bucket,
                     "num_classes": torch.full((b,), n_cls)})
    P = n_cls - 1
    dist = rng.random((P, P)).astype(np.float32) * 400.0
    dist = np.minimum(dist, dist.T)
    np.fill_diagonal(dist, 0.0)
    dist[rng.random((P, P)) < 0.05] = np.inf
    rank = np.argsort(dist, axis=1, kind="stable")
    groups = ("basic", "binary", "bucket", "atd", "success", "accuracy_i")
    want = evaluate_logits(batches, topk, groups, dist, rank)
    det = _Replay([torch.from_numpy(b["logits"]) for b in batches])
    ev = DetectorEvaluator(torch.nn.Identity(), det, torch.device("cpu"), l_pred=2, l_det=2, topk=topk, metric_groups=groups,
                           pipe_dist=dist, residual_builder=lambda pred, noisy, tseg, l_pred, l_det, device=None: noisy)
    got = ev.evaluate(feed)
    assert set(got) == set(want)
    for k in want:
        assert got[k] == want[k] or (np.isinf(got[k]) and np.isinf(want[k])), (k, got[k], want[k])
    with pytest.raises(ValueError):
        DetectorEvaluator(torch.nn.Identity(), det, torch.device("cpu"), l_pred=2, l_det=2, metric_groups=("atd",))
