"""N > 1 host logic on CPU: world_size-2 gloo processes.  The data-parallel step (per-rank shard, backward into
the flat bucket, one all-reduce, clip, AdamW) must reproduce the single-process step on the concatenated batch
(SURVEY.md section 8e).  The model here is the CPU oracle of the detector -- the product has no CPU path, and
the DP helper is model-agnostic."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    from oracle.detector_oracle import OracleLeakDetector
    z = np.load(REPO / "tests/golden/graph_LTA.npz")
    torch.manual_seed(42)
    m = OracleLeakDetector(661, torch.from_numpy(z["edge_index"]), torch.from_numpy(z["pipe_ends"][:5]),
                           z["sensor_node_idx"].tolist(), 16, 16, 2, 0.0, True)
    return m.train()


def _batch(n):
    gen = torch.Generator().manual_seed(198)
    return torch.randn(n, 6, 29, generator=gen), torch.randn(n, 6, 9, generator=gen), torch.randint(0, 6, (n,), generator=gen)


def _step(model, bucket, r, t, y, opt):
    bucket.zero()
    torch.nn.functional.cross_entropy(model(r, t), y).backward()
    bucket.allreduce()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leak_det_gnn_b200.parallel import FlatGradBucket, broadcast_parameters, shard_indices
    torch.set_num_threads(1)
    model = _model()
    if rank == 1:  # diverge on purpose: broadcast must repair it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    broadcast_parameters(list(model.parameters()))
    bucket = FlatGradBucket(model.parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    r, t, y = _batch(8)
    idx = list(shard_indices(8, rank, world))
    for _ in range(2):
        _step(model, bucket, r[idx], t[idx], y[idx], opt)
    assert bucket.attached()
    torch.save({k: v.clone() for k, v in model.state_dict().items()}, Path(out_dir) / f"rank{rank}.pt")
    dist.destroy_process_group()


def test_dp2_matches_single_process(tmp_path):
    from leak_det_gnn_b200.parallel import FlatGradBucket
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "rank0.pt")
    b = torch.load(tmp_path / "rank1.pt")
    for k in a:
        assert torch.equal(a[k], b[k]), k  # ranks stay in lock step

    torch.set_num_threads(1)
    model = _model()
    bucket = FlatGradBucket(model.parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    r, t, y = _batch(8)
    for _ in range(2):
        _step(model, bucket, r, t, y, opt)  # mean-reduced loss over the 8 windows == average of two 4-window means
    for k, v in model.state_dict().items():
        assert torch.allclose(a[k], v, rtol=1e-5, atol=1e-7), (k, (a[k] - v).abs().max())


def test_bucket_and_sharding_unit():
    from leak_det_gnn_b200.parallel import FlatGradBucket, shard_indices
    lin = torch.nn.Linear(3, 2)
    bk = FlatGradBucket(lin.parameters())
    assert bk.numel == 8 and bk.nbytes() == 32 and bk.attached()
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(bk.flat[:6].view(2, 3), lin.weight.grad) and bk.flat[6:].tolist() == [1.0, 1.0]
    bk.allreduce()  # no process group: no-op
    lin.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError, match="bucket.zero"):
        bk.allreduce()
    assert list(shard_indices(7, 1, 3)) == [1, 4] and list(shard_indices(2, 1, 2)) == [1]
    assert sorted(i for r in range(4) for i in shard_indices(10, r, 4)) == list(range(10))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
