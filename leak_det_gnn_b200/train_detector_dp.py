"""Data-parallel detector training: the reference's ``python -m models.train_detector`` with one process per GPU.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m leak_det_gnn_b200.train_detector_dp \
        --leak_root DATA --inp_path NET.inp --predictor_ckpt predictor_best.ckpt --out_dir OUT [reference flags ...] \
        --reference_root /path/to/Leak-det-gnn

Same flags, same step (train_detector.py:296-321: residual under no_grad -> detector -> cross-entropy -> backward ->
clip_grad_norm_ -> AdamW), same checkpoint layout (train_detector.py:346-361: keys ``epoch, detector_state, sensor_ids,
pipe_ids_in_order, num_classes, predictor_ckpt, args``; ``detector_best.ckpt`` by validation ``acc_top1``,
``detector_last.ckpt``, ``detector_meta.json``), written by rank 0 only.  What changes:

* the model is this package's drop-in ``LeakDetector`` (sm_100a kernels) and the residual builder its cone evaluation;
* every rank takes the indices ``rank, rank + W, ...`` of the reference's deterministic ``seed + idx`` dataset
  (models/datasets.py:489-541), so W ranks with ``--batch_size B`` reproduce one process with batch ``W * B``; gradients
  live in one flat 242 KB bucket and are averaged by ONE NCCL all-reduce per step, BEFORE the clip (parallel.py);
* validation uses the vectorised evaluator (evaluator.py: no per-sample ``.item()``);
* the step is one fixed launch sequence (the loader drops the last partial batch), so it is captured once and replayed as
  a CUDA graph (graphed.py; ``--cuda_graph 0`` issues it call by call).

The datasets stay the reference's own code (host-side CSV loaders, out of scope here): ``--reference_root`` is put on
``sys.path`` and ``models.datasets.AbruptLeakDetectorDataset`` is imported from it unchanged.  ``--synthetic N`` replaces
the data (and, if no ``--predictor_ckpt`` is given, the predictor) by seeded synthetic segments of the same batch layout,
which is what the tests and the multi-GPU smoke runs use.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.utils.data import DataLoader, Dataset, Subset

from .evaluator import BUCKETS, DetectorEvaluator
from .graph import parse_epanet_inp
from .models import LeakDetector, NormalPredictorGRU, NormalPredictorTCN, build_residual_sequence_from_segment
from .graphed import GraphedTrainStep
from .parallel import FlatGradBucket, broadcast_parameters, shard_indices


def now() -> str:
    return time.strftime("%Y-%m-%d %H:%M:%S")


def build_argparser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    # ---- the reference's flags, names and defaults unchanged (models/train_detector.py:131-155)
    ap.add_argument("--leak_root", type=str, default=None, help="Path to leak dataset root")
    ap.add_argument("--inp_path", type=str, required=True, help="Path to EPANET .inp file")
    ap.add_argument("--predictor_ckpt", type=str, default=None, help="Path to trained predictor checkpoint")
    ap.add_argument("--out_dir", type=str, required=True, help="Output directory for detector checkpoints/logs")
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--steps_per_epoch", type=int, default=120000)
    ap.add_argument("--val_steps", type=int, default=10000)
    ap.add_argument("--test_steps", type=int, default=10000)
    ap.add_argument("--batch_size", type=int, default=128, help="per process (global batch = world size x this)")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--weight_decay", type=float, default=1e-4)
    ap.add_argument("--grad_clip", type=float, default=1.0)
    ap.add_argument("--l_pred", type=int, default=36)
    ap.add_argument("--l_det", type=int, default=36)
    ap.add_argument("--topk", type=int, default=5)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--device", type=str, default="auto")
    ap.add_argument("--num_workers", type=int, default=0)
    ap.add_argument("--log_every", type=int, default=50)
    # ---- additions
    ap.add_argument("--reference_root", type=str, default=None, help="checkout of the reference (for models.datasets)")
    ap.add_argument("--synthetic", type=int, default=0, help="use N synthetic windows per epoch instead of --leak_root")
    ap.add_argument("--synthetic_sensors", type=str, default=None, help="comma-separated sensor node ids (synthetic mode)")
    ap.add_argument("--synthetic_pipes", type=int, default=0, help="number of class pipes, 0 = all [PIPES] (synthetic mode)")
    ap.add_argument("--cuda_graph", type=int, default=1, help="1: replay the training step as a CUDA graph (graphed.py); "
                                                              "0: one call per kernel")
    return ap


class SyntheticSegments(Dataset):
    """Seeded stand-in with the batch layout of ``AbruptLeakDetectorDataset.__getitem__`` (models/datasets.py:489-541):
    ``noisy_seg (l_pred + l_det, S)``, ``time_seg (l_pred + l_det, 9)``, ``label``, ``bucket``, ``num_classes``; sample
    ``idx`` depends only on ``seed + idx``, like the reference's datasets."""

    def __init__(self, n: int, n_sensors: int, n_pipes: int, l_pred: int, l_det: int, seed: int) -> None:
        self.n, self.s, self.p, self.length, self.seed = n, n_sensors, n_pipes, l_pred + l_det, seed

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, idx: int) -> Dict:
        rng = np.random.default_rng(self.seed + idx)
        bucket = BUCKETS[int(rng.integers(0, 4))]
        label = self.p if bucket in ("pre", "noleak") else int(rng.integers(0, self.p))
        noisy = rng.standard_normal((self.length, self.s)).astype(np.float32)
        if label < self.p:                                            # a leak leaves a pipe-dependent drift
            noisy[self.length // 2:, label % self.s] -= 1.5
        minutes = 5 * (int(rng.integers(0, 288 * 7)) + np.arange(self.length))
        hour = ((minutes // 60) % 24 + (minutes % 60) / 60.0).astype(np.float32)
        ang = 2.0 * np.pi * hour / 24.0
        tfeat = np.concatenate([np.sin(ang)[:, None], np.cos(ang)[:, None], np.eye(7, dtype=np.float32)[(minutes // 1440) % 7]],
                               axis=1).astype(np.float32)
        return {"noisy_seg": torch.from_numpy(noisy), "time_seg": torch.from_numpy(tfeat), "label": label, "bucket": bucket,
                "num_classes": self.p + 1}


def load_predictor(ckpt_path: str | Path, device: torch.device) -> Tuple[nn.Module, Dict]:
    """models/train_detector.py:114-128 with this package's (checkpoint-compatible) predictor classes."""
    ckpt = torch.load(ckpt_path, map_location=device, weights_only=False)
    n_s = len(ckpt["sensor_ids"])
    model = (NormalPredictorGRU if ckpt.get("arch", "tcn") == "gru" else NormalPredictorTCN)(num_sensors=n_s, time_dim=9)
    model.load_state_dict(ckpt["model_state"])
    model.to(device).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model, ckpt


def _dist_env() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def train_step(detector, predictor, bucket: FlatGradBucket, opt, loss_fn, batch: Dict, device, l_pred: int, l_det: int,
               grad_clip: float) -> torch.Tensor:
    """One step of train_detector.py:296-317 with the gradient average before the clip.  Returns the (device) loss."""
    noisy_seg = batch["noisy_seg"].to(device, non_blocking=True)
    time_seg = batch["time_seg"].to(device, non_blocking=True)
    label = torch.as_tensor(batch["label"], device=device, dtype=torch.long)
    with torch.no_grad():
        residual = build_residual_sequence_from_segment(predictor, noisy_seg, time_seg, l_pred=l_pred, l_det=l_det,
                                                        device=device)
    logits = detector(residual, time_seg[:, l_pred:, :].contiguous())
    loss = loss_fn(logits, label)
    bucket.zero()                      # instead of opt.zero_grad(set_to_none=True): the grads are views into the bucket
    loss.backward()
    bucket.allreduce()
    if grad_clip and grad_clip > 0:
        torch.nn.utils.clip_grad_norm_(detector.parameters(), grad_clip)
    opt.step()
    return loss.detach()


def main(argv: Optional[List[str]] = None) -> Dict[str, float]:
    args = build_argparser().parse_args(argv)
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("train_detector_dp: needs a CUDA device (the drop-in detector has no CPU path)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    out_dir = Path(args.out_dir)
    if rank == 0:
        out_dir.mkdir(parents=True, exist_ok=True)
        print(f"{now()} [detector] device={device} seed={args.seed} world={world}")

    # ---- data + predictor
    if args.synthetic:
        sec = parse_epanet_inp(args.inp_path)
        pipe_ids_in_order = [ln.split()[0] for ln in sec["PIPES"]]
        if args.synthetic_pipes:
            pipe_ids_in_order = pipe_ids_in_order[: args.synthetic_pipes]
        if args.predictor_ckpt:
            predictor, predictor_ckpt = load_predictor(args.predictor_ckpt, device)
            sensor_ids = list(predictor_ckpt["sensor_ids"])
        else:
            if not args.synthetic_sensors:
                raise SystemExit("--synthetic without --predictor_ckpt needs --synthetic_sensors")
            sensor_ids = args.synthetic_sensors.split(",")
            predictor = NormalPredictorTCN(len(sensor_ids), 9).to(device).eval()
            for p in predictor.parameters():
                p.requires_grad_(False)
        mk = lambda n, seed: SyntheticSegments(n, len(sensor_ids), len(pipe_ids_in_order), args.l_pred, args.l_det, seed)
        train_ds, val_ds = mk(args.synthetic, args.seed), mk(max(args.batch_size, min(args.val_steps, args.synthetic)), args.seed + 1)
    else:
        if not (args.leak_root and args.predictor_ckpt and args.reference_root):
            raise SystemExit("need --leak_root, --predictor_ckpt and --reference_root (or --synthetic N)")
        sys.path.insert(0, str(Path(args.reference_root)))
        from models.datasets import AbruptLeakDetectorDataset, SensorStandardizer  # the reference's own loaders, unchanged

        predictor, predictor_ckpt = load_predictor(args.predictor_ckpt, device)
        sensor_ids = list(predictor_ckpt["sensor_ids"])
        stdzr = SensorStandardizer(mean=np.asarray(predictor_ckpt["standardizer_mean"], dtype=np.float32),
                                   std=np.asarray(predictor_ckpt["standardizer_std"], dtype=np.float32))
        mk = lambda steps, seed: AbruptLeakDetectorDataset(
            leak_root=args.leak_root, l_pred_steps=args.l_pred, l_det_steps=args.l_det, steps_per_epoch=steps, seed=seed,
            sensor_ids=sensor_ids, standardizer=stdzr, cache_size=4096)
        train_ds, val_ds = mk(args.steps_per_epoch, args.seed), mk(args.val_steps, args.seed + 1)
        if sensor_ids != train_ds.get_sensor_node_ids():
            raise ValueError("Sensor IDs do not match between the normal and abrupt datasets.")
        pipe_ids_in_order = train_ds.get_pipe_ids_in_order()

    shard = Subset(train_ds, list(shard_indices(len(train_ds), rank, world)))
    train_loader = DataLoader(shard, batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=True, drop_last=True)
    val_loader = DataLoader(val_ds, batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=True)

    # ---- model, optimiser (train_detector.py:246-258)
    detector = LeakDetector(inp_path=args.inp_path, sensor_node_ids=sensor_ids, pipe_ids_in_order=pipe_ids_in_order,
                            sensor_hidden=64, node_hidden=64, gnn_layers=2, dropout=0.1, use_time=True).to(device)
    broadcast_parameters(list(detector.parameters()))
    bucket = FlatGradBucket(detector.parameters())
    opt = torch.optim.AdamW(detector.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    loss_fn = nn.CrossEntropyLoss()
    evaluator = DetectorEvaluator(predictor, detector, device, l_pred=args.l_pred, l_det=args.l_det, topk=args.topk,
                                  metric_groups=("basic", "binary", "bucket"))

    # the loader drops the last partial batch, so every step has one shape: capture it once (graphed.py)
    graphed = (GraphedTrainStep(detector, opt, bucket, args.batch_size, args.l_det, predictor=predictor, l_pred=args.l_pred,
                                grad_clip=args.grad_clip, loss_fn=loss_fn) if args.cuda_graph else None)
    best_acc, metrics = -1.0, {}
    if rank == 0:
        meta = {"inp_path": str(args.inp_path), "predictor_ckpt": str(args.predictor_ckpt), "sensor_ids": sensor_ids,
                "pipe_ids_in_order": pipe_ids_in_order, "num_classes": len(pipe_ids_in_order) + 1, "world_size": world,
                "args": vars(args)}
        (out_dir / "detector_meta.json").write_text(json.dumps(meta, indent=2, ensure_ascii=False), encoding="utf-8")
        print(f"{now()} [detector] start training: epochs={args.epochs}, steps/epoch/rank={len(train_loader)}, "
              f"batch={args.batch_size} x {world}")
    for epoch in range(1, args.epochs + 1):
        detector.train()
        running = torch.zeros((), device=device)
        seen = 0
        for it, batch in enumerate(train_loader, start=1):
            if graphed is not None:
                loss = graphed(batch["noisy_seg"], batch["time_seg"], torch.as_tensor(batch["label"], dtype=torch.long))
            else:
                loss = train_step(detector, predictor, bucket, opt, loss_fn, batch, device, args.l_pred, args.l_det,
                                  args.grad_clip)
            running += loss * batch["noisy_seg"].size(0)     # stays on the device: the reference's per-step loss.item() sync is gone
            seen += batch["noisy_seg"].size(0)
            if rank == 0 and (it % args.log_every) == 0:
                print(f"{now()} [detector][epoch {epoch:02d}] step {it:05d}/{len(train_loader):05d} "
                      f"loss={running.item() / max(seen, 1):.6f}")
        if rank == 0:
            metrics = evaluator.evaluate(val_loader)
            acc = float(metrics.get("acc_top1", 0.0))
            print(f"{now()} [detector][epoch {epoch:02d}] train_loss={running.item() / max(seen, 1):.6f} val_acc_top1={acc:.4f}")
            ckpt = {"epoch": epoch, "detector_state": detector.state_dict(), "sensor_ids": sensor_ids,
                    "pipe_ids_in_order": pipe_ids_in_order, "num_classes": len(pipe_ids_in_order) + 1,
                    "predictor_ckpt": str(args.predictor_ckpt), "args": vars(args)}
            torch.save(ckpt, out_dir / "detector_last.ckpt")
            if acc > best_acc:
                best_acc = acc
                torch.save(ckpt, out_dir / "detector_best.ckpt")
        if world > 1:
            dist.barrier()
    if world > 1 and argv is None:
        dist.destroy_process_group()
    return metrics


if __name__ == "__main__":
    main()
