import sys, time
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200.models.detector import SharedSensorGRUEncoder
enc = SharedSensorGRUEncoder(hidden_size=64).cuda().train()
r = torch.randn(4096, 288, 29, device="cuda"); t = torch.randn(4096, 288, 9, device="cuda")
for chunk in (8192, 16384, 32768, 65536):
    enc.max_seqs_per_call = chunk
    try:
        for _ in range(2):
            enc.zero_grad(); h = enc(r, t); h.sum().backward()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            enc.zero_grad(); h = enc(r, t); h.sum().backward()
        torch.cuda.synchronize()
        print(chunk, f"{(time.perf_counter()-t0)/3*1e3:.1f} ms fwd+bwd", f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    except RuntimeError as e:
        print(chunk, "failed:", str(e)[:80])
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
