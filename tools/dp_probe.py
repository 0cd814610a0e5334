#!/usr/bin/env python
"""Where does a data-parallel step spend its time?  torchrun --nproc-per-node N tools/dp_probe.py"""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200.models import LeakDetector
from leak_det_gnn_b200.parallel import FlatGradBucket

local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = np.load(REPO / "tests/golden/graph_LTA.npz")
torch.manual_seed(42)
m = LeakDetector(REPO / "tests/golden/L-TOWN-A.topo.inp", [str(s) for s in g["sensor_node_ids"]],
                 [str(p) for p in g["pipe_ids"]]).to(dev).train()
b = 4096
h_s = torch.randn(b, 29, 64, device=dev, requires_grad=True)
label = torch.randint(0, 765, (b,), device=dev)
bucket = FlatGradBucket(list(m.parameters()))

def step(ar):
    bucket.zero(); h_s.grad = None
    torch.nn.functional.cross_entropy(m.gnn_stack(h_s), label).backward()
    if ar: bucket.allreduce()

def run(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    if local == 0:
        print(f"{name}: device {e0.elapsed_time(e1)/n:.3f} ms/iter, host enqueue {(t1-t0)*1e3/n:.3f} ms/iter", flush=True)

run("step, no allreduce", lambda: step(False))
if world > 1:
    run("allreduce only", lambda: bucket.allreduce())
    run("step + allreduce", lambda: step(True))

# per-iteration device times of the DP step (events around every iteration)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
if world > 1: dist.barrier()
torch.cuda.synchronize()
ev[0].record()
for i in range(40):
    step(world > 1)
    ev[i + 1].record()
torch.cuda.synchronize()
if local == 0:
    print("per-iter ms:", " ".join(f"{ev[i].elapsed_time(ev[i+1]):.2f}" for i in range(40)), flush=True)
