"""Mint the golden fixtures under tests/golden/ -- run in the BUILD container only.

    python tests/golden/make_goldens.py

Needs /root/reference (the unmodified upstream checkout).  Nothing under tests/ reads
/root/reference at test time except the optional cross-checks that skip when absent.

What is produced, and by WHOSE code:

1. ``<net>.topo.inp``  -- topology-only EPANET files (ids + link endpoints of the node and
   link sections; no hydraulics, coordinates, demands or patterns) derived from the two
   network files the reference ships (data/raw/L-TOWN-A/L-TOWN_AreaA.inp,
   data/raw/L-TOWN/L-TOWN.inp; KIOS BattLeDIM 2020 "L-TOWN").  They exist so the GPU box,
   which has no /root/reference, can construct ``LeakDetector(inp_path=...)``.
2. ``graph_<net>.npz`` -- outputs of the REFERENCE's ``build_wdn_graph_from_inp``
   (models/utils.py:84-166) on the ORIGINAL files, called the way the detector calls it
   (detector.py:137-144): node_names, edge_index, pipe_ids (= [PIPES] order), pipe_ends,
   plus the 29 pressure sensors of configs/sim_LTA.yaml:21.  Bit-exact targets.
3. ``detector_<case>.pt`` -- seeded state_dict, inputs, logits, loss and all 18 gradients
   from the REFERENCE's own ``models/detector.py`` ``LeakDetector`` run on CPU in eval()
   mode, with ``torch_geometric.nn`` supplied by oracle/pyg_restatement.py (PyG itself is
   not installed; see that file's "parity unpinned" note).  Also records that
   oracle/detector_oracle.py reproduces those logits bit-for-bit.
"""
from __future__ import annotations

import hashlib
import sys
import types
from pathlib import Path

import numpy as np
import torch
import yaml

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

NETS = {
    "LTA": REF / "data/raw/L-TOWN-A/L-TOWN_AreaA.inp",
    "LT": REF / "data/raw/L-TOWN/L-TOWN.inp",
}
TOPO_NAME = {"LTA": "L-TOWN-A.topo.inp", "LT": "L-TOWN.topo.inp"}


def _import_reference():
    from oracle import pyg_restatement as pyg

    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.GCNConv = pyg.GCNConv
    tgnn.global_mean_pool = pyg.global_mean_pool
    tg.nn = tgnn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tgnn
    sys.path.insert(0, str(REF))
    import models.detector as ref_detector  # noqa: E402  (reference code, unmodified)
    import models.utils as ref_utils  # noqa: E402

    return ref_detector, ref_utils


def write_topo_inp(ref_utils, src: Path, dst: Path, title: str) -> None:
    sec = ref_utils.parse_epanet_inp(src)
    out = ["[TITLE]", f"{title} -- topology only (ids and link endpoints), derived from the",
           "KIOS BattLeDIM 2020 L-TOWN network shipped with Mateng0228/Leak-det-gnn; not a hydraulic model.", ""]
    for name in ("JUNCTIONS", "RESERVOIRS", "TANKS"):
        out.append(f"[{name}]")
        out.append(";ID")
        out += [f" {ln.split()[0]}" for ln in sec.get(name, [])]
        out.append("")
    for name in ("PIPES", "PUMPS", "VALVES"):
        out.append(f"[{name}]")
        out.append(";ID Node1 Node2")
        for ln in sec.get(name, []):
            t = ln.split()
            out.append(f" {t[0]} {t[1]} {t[2]}   ;")
        out.append("")
    out.append("[END]")
    dst.write_text("\n".join(out) + "\n")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_graph_goldens(ref_utils, sensors):
    info = {}
    for net, path in NETS.items():
        write_topo_inp(ref_utils, path, HERE / TOPO_NAME[net], path.stem)
        sec = ref_utils.parse_epanet_inp(path)
        pipe_ids = [ln.split()[0] for ln in sec["PIPES"]]
        g = ref_utils.build_wdn_graph_from_inp(path, sensors, pipe_ids, add_self_loops=False, make_undirected=True)
        # and through the topology-only file: must give the identical graph
        g2 = ref_utils.build_wdn_graph_from_inp(HERE / TOPO_NAME[net], sensors, pipe_ids, add_self_loops=False,
                                                make_undirected=True)
        assert g.node_names == g2.node_names and torch.equal(g.edge_index, g2.edge_index)
        assert np.array_equal(g.pipe_ends, g2.pipe_ends)
        ei = g.edge_index.numpy()
        np.savez_compressed(
            HERE / f"graph_{net}.npz",
            node_names=np.array(g.node_names), edge_index=ei, pipe_ids=np.array(pipe_ids), pipe_ends=g.pipe_ends,
            sensor_node_ids=np.array(sensors),
            sensor_node_idx=np.array([g.node_to_idx[s] for s in sensors], dtype=np.int64),
            sha_edge_index=sha(ei), sha_pipe_ends=sha(g.pipe_ends),
            sha_node_names=hashlib.sha256("\n".join(g.node_names).encode()).hexdigest(),
        )
        info[net] = (g, pipe_ids)
        print(f"[graph] {net}: N={len(g.node_names)} E={ei.shape[1]} P={len(pipe_ids)} sha(edge_index)={sha(ei)[:16]}")
    return info


def time_features(n_steps: int, start_minute: int = 0) -> np.ndarray:
    """(n_steps, 9) hour sin/cos + 7-way day-of-week one-hot on a 5-minute grid starting
    2018-01-02 00:00 (a Tuesday; configs/sim_LTA.yaml:10) -- the arithmetic of
    reference models/datasets.py:49-59 without pandas."""
    minutes = start_minute + 5 * np.arange(n_steps)
    hour = ((minutes // 60) % 24).astype(np.float32) + ((minutes % 60).astype(np.float32) / 60.0)
    angle = (2.0 * np.pi) * (hour / 24.0)
    dow = (1 + minutes // 1440) % 7
    return np.concatenate([np.sin(angle).astype(np.float32)[:, None], np.cos(angle).astype(np.float32)[:, None],
                           np.eye(7, dtype=np.float32)[dow]], axis=1).astype(np.float32)


def make_detector_golden(ref_detector, case: str, net: str, pipe_ids, sensors, batch: int, l_det: int,
                         node_hidden: int, sensor_hidden: int, gnn_layers: int, seed: int) -> None:
    from oracle.detector_oracle import OracleLeakDetector

    torch.manual_seed(seed)  # reference default seed is 42 (train_detector.py:151)
    model = ref_detector.LeakDetector(NETS[net], sensors, pipe_ids, sensor_hidden=sensor_hidden,
                                      node_hidden=node_hidden, gnn_layers=gnn_layers, dropout=0.1, use_time=True)
    # conv biases are zero-initialised by PyG; perturb them so their gradients / effect are exercised
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for conv in model.convs:
            conv.bias.copy_(0.05 * torch.randn(conv.bias.shape, generator=gen))
    model.eval()
    gen = torch.Generator().manual_seed(198)
    residual = torch.randn(batch, l_det, len(sensors), generator=gen)
    tf = time_features(l_det + batch)
    tfeat = torch.from_numpy(np.stack([tf[i:i + l_det] for i in range(batch)]))
    label = torch.randint(0, len(pipe_ids) + 1, (batch,), generator=gen)

    # capture the sensor embeddings that enter the message-passing path, and the gradient that leaves it
    cap = {}

    def _hook(_mod, _inp, out):
        out.retain_grad()
        cap["h_s"] = out

    handle = model.sensor_encoder.register_forward_hook(_hook)
    logits = model(residual, tfeat)
    loss = torch.nn.functional.cross_entropy(logits, label)
    loss.backward()
    handle.remove()
    h_s, grad_h_s = cap["h_s"].detach().clone(), cap["h_s"].grad.detach().clone()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    assert len(state) == 4 + 2 + 2 * gnn_layers + 8, len(state)

    # the restated detector must reproduce the reference module bit-for-bit on CPU
    orc = OracleLeakDetector(len(model.node_names), model.edge_index_single, model.pipe_ends,
                             model.sensor_node_idx, sensor_hidden, node_hidden, gnn_layers, 0.1, True)
    orc.load_state_dict(state, strict=True)
    orc.eval()
    o_logits = orc(residual, tfeat)
    bit_equal = bool(torch.equal(o_logits, logits))
    assert bit_equal, (o_logits - logits).abs().max()

    # fp64 run of the same algorithm: the truth both fp32 implementations are measured against
    orc64 = OracleLeakDetector(len(model.node_names), model.edge_index_single, model.pipe_ends,
                               model.sensor_node_idx, sensor_hidden, node_hidden, gnn_layers, 0.1, True).double()
    orc64.load_state_dict({k: v.double() for k, v in state.items()}, strict=True)
    orc64.eval()
    logits64 = orc64(residual.double(), tfeat.double())
    loss64 = torch.nn.functional.cross_entropy(logits64, label)
    loss64.backward()
    grads64 = {k: p.grad.detach().clone() for k, p in orc64.named_parameters()}
    # fp64 truth of the stack alone, driven by the fp32 embeddings (what the GPU stack is handed)
    orc64.zero_grad()
    hs64 = h_s.double().requires_grad_(True)
    stack_logits64 = orc64.gnn_stack(hs64)
    torch.nn.functional.cross_entropy(stack_logits64, label).backward()
    stack_grads64 = {k: p.grad.detach().clone() for k, p in orc64.named_parameters() if p.grad is not None}
    stack_grad_h_s64 = hs64.grad.detach().clone()

    torch.save({
        "logits64": logits64.detach(), "loss64": loss64.detach(), "grads64": grads64,
        "h_s": h_s, "grad_h_s": grad_h_s, "stack_logits64": stack_logits64.detach(),
        "stack_grads64": stack_grads64, "stack_grad_h_s64": stack_grad_h_s64,
        "case": case, "net": net, "pipe_ids": list(pipe_ids), "sensor_node_ids": list(sensors),
        "hparams": dict(sensor_hidden=sensor_hidden, node_hidden=node_hidden, gnn_layers=gnn_layers, dropout=0.1,
                        use_time=True),
        "state_dict": state, "residual": residual, "tfeat": tfeat, "label": label,
        "logits": logits.detach(), "loss": loss.detach(), "grads": grads,
        "oracle_equals_reference_module_bitwise": bit_equal,
        "torch_version": str(torch.__version__),
    }, HERE / f"detector_{case}.pt")
    print(f"[detector] {case}: logits {tuple(logits.shape)} loss={loss.item():.6f} oracle==reference(bitwise)={bit_equal}")


def make_residual_golden(ref_utils, n_sensors: int = 29) -> None:
    """``residual_TCN.pt``: the REFERENCE's ``build_residual_sequence_from_segment`` (models/utils.py:169-216) run with
    the REFERENCE's ``NormalPredictorTCN`` (models/predictor.py:55-81, imported unmodified; pure torch) on a seeded
    segment batch -- the pinned oracle of SURVEY 8f rank 1.  Separate RNG streams: existing goldens are unaffected."""
    import models.predictor as ref_predictor  # noqa: E402  (reference code, unmodified)

    gen = torch.Generator().manual_seed(77)
    torch.manual_seed(77)
    model = ref_predictor.NormalPredictorTCN(n_sensors, 9).eval()
    with torch.no_grad():  # LayerNorm affine / biases away from their trivial init
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    l_pred = l_det = 36
    noisy = torch.randn(6, l_pred + l_det, n_sensors, generator=gen)
    tf = torch.from_numpy(np.stack([time_features(l_pred + l_det, 5 * 17 * i) for i in range(6)]))
    with torch.no_grad():
        residual = ref_utils.build_residual_sequence_from_segment(model, noisy, tf, l_pred, l_det)
        y_first = model(noisy[:, :l_pred], tf[:, :l_pred])
        res64 = ref_utils.build_residual_sequence_from_segment(model.double(), noisy.double(), tf.double(), l_pred, l_det)
    model.float()
    torch.save({"state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()}, "noisy_seg": noisy,
                "time_seg": tf, "l_pred": l_pred, "l_det": l_det, "residual": residual, "residual64": res64,
                "y_hat_first_window": y_first, "torch_version": str(torch.__version__)}, HERE / "residual_TCN.pt")
    print(f"[residual] {tuple(residual.shape)} max|r|={residual.abs().max().item():.4f} "
          f"fp32 vs fp64 {(residual.double() - res64).abs().max().item():.2e}")


def main() -> None:
    ref_detector, ref_utils = _import_reference()
    if "--residual-only" in sys.argv:
        make_residual_golden(ref_utils)
        return
    sensors = yaml.safe_load((REF / "configs/sim_LTA.yaml").read_text())["sensors"]["pressure_node_ids"]
    info = make_graph_goldens(ref_utils, sensors)
    lta_pipes = info["LTA"][1]
    lt_pipes = info["LT"][1]
    # README example classes: default_rng(198).choice(764, 2) under the build image's numpy
    # (leak_generation.py:90-99 pick_pipes) -> two class pipes
    idx = np.sort(np.random.default_rng(198).choice(len(lta_pipes), 2, replace=False))
    readme_pipes = [lta_pipes[i] for i in idx]
    make_detector_golden(ref_detector, "LTA_P2", "LTA", readme_pipes, sensors, batch=4, l_det=36, node_hidden=64,
                         sensor_hidden=64, gnn_layers=2, seed=42)
    make_detector_golden(ref_detector, "LTA_Pall", "LTA", lta_pipes, sensors, batch=3, l_det=36, node_hidden=64,
                         sensor_hidden=64, gnn_layers=2, seed=42)
    make_detector_golden(ref_detector, "LT_Pall", "LT", lt_pipes, sensors, batch=2, l_det=36, node_hidden=64,
                         sensor_hidden=64, gnn_layers=2, seed=43)
    make_detector_golden(ref_detector, "LTA_D128_L3", "LTA", lta_pipes[:40], sensors, batch=2, l_det=12,
                         node_hidden=128, sensor_hidden=32, gnn_layers=3, seed=44)
    make_residual_golden(ref_utils)


if __name__ == "__main__":
    main()
