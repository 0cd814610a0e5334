"""bench.py contract on the CPU side: the reference arm (the oracle port on the host cores) prints ONE JSON line with
the keys the driver reads.  The GPU arm needs a B200 and is exercised by the driver itself."""
import json
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "2", "--l-det", "36"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "window_graphs_per_sec_fwd_bwd" and d["unit"] == "window-graphs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    # same scopes as the product arm: value = GNN stack from resident sensor embeddings, e2e = the full detector call
    # (sensor GRU included), so value / value and e2e / e2e ratios are like for like
    assert 0 < d["e2e"]["value"] <= d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
