#!/usr/bin/env python
"""bench.py -- egonets/s (fwd+bwd) of the PGAT+WMR(+LBM) hot path on MAG-CS-shaped synthetic batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = TaxoExpan.forward (propagate + readout + match) + InfoNCE loss + backward on one batch of
256 queries x (1 positive + 31 negatives) = 8192 egonets per GPU (BASELINE.json configs[1]; configs[2] is the same
per-GPU batch sharded by query group over N GPUs with one NCCL all-reduce of the flat gradient -> weak scaling).
Prints ONE JSON line (rank 0). `value` times the step with inputs resident in HBM; `e2e` times it through the public
API from pinned HOST buffers (H2D of features/queries/egonet counts + structure build + D2H of the loss inside the
timed region).  `--impl reference` times the CPU port of the reference path (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "egonets/s (fwd+bwd) PGAT d=250 on MAG-CS-shaped batches"
UNIT = "egonets/s"
MAGCS = dict(in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1],
             feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)       # config_files/config.mag.json:11-20
NEGATIVE_SIZE = 31                                                              # config.mag.json:30


def csrc_sha16():
    """hash of the CUDA sources the library is built from (ties profiles/traffic.json to the kernels it was measured on)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "taxoexpan_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def load_peaks():
    """(HBM GB/s, bf16 TFLOP/s, source).  MEASURED_PEAKS.json is written by the driver; its key names are not part of any contract
    here, so any numeric entry whose key mentions hbm / bandwidth (resp. bf16 / tflop) is accepted, nested dicts included."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = tf = None
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
        except (OSError, ValueError):
            d = {}

        def walk(obj, prefix=""):
            if isinstance(obj, dict):
                for k, v in obj.items():
                    yield from walk(v, f"{prefix}.{k}".lower())
            elif isinstance(obj, (int, float)) and not isinstance(obj, bool):
                yield prefix, float(obj)

        items = list(walk(d))
        for key, val in items:
            if hbm is None and ("hbm" in key or "bandwidth" in key or "copy" in key) and 500.0 < val < 20000.0:
                hbm = val
            if ("bf16" in key or "tflop" in key or "tensor" in key) and 100.0 < val < 5000.0:
                if tf is None or "sustain" in key:
                    tf = val
        for key, val in items:          # bandwidth given in TB/s
            if hbm is None and ("hbm" in key or "bandwidth" in key) and 0.5 < val < 20.0:
                hbm = val * 1000.0
    if hbm is not None:
        return hbm, (tf if tf is not None else 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# algorithmic bytes, per GAT layer launch, W = H * D'
# ------------------------------------------------------------------------------------------------
def gaa_fwd_bytes(n, e, heads, width, training=True):
    return 4 * (n * width + n * width + 2 * n * heads + (e * heads if training else 0)) + 4 * (n + 1 + e)


def gaa_bwd_bytes(n, e, heads, width):
    return 4 * (n * width + n * width + n * width + e * heads + 2 * n * heads) + 8 * (n + 1 + e)


def gemm_flops_per_node(cfg):
    k0 = cfg["in_dim"] + cfg["pos_dim"]
    f0 = cfg["hidden_dim"] * cfg["heads"][0]
    k1 = f0 + cfg["pos_dim"]
    f1 = cfg["out_dim"] * cfg["heads"][1]
    return 2 * (k0 * f0 + k1 * f1)


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Default: in-process NVML queries
    (nvidia_ml_py) issued by the enqueue thread itself, twice inside every timed loop (`sample_once`; the loop is GPU-bound, the
    ~0.1 ms query hides behind the queued launches).  TAXO_SAMPLER=thread polls from a thread every 20 ms; without NVML a polling
    `nvidia-smi -lms` child is the fallback.  (Both background variants disturbed the FIRST timed loop on some boxes: the host's
    enqueue time went from 1.6 to 3-5 ms/step, i.e. the headline dropped by up to 60 % while the later loops were normal.)"""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.samples = []            # (sm_mhz, sm_max_mhz, reasons bitmask) from NVML
        self.nvml = None
        self.thread = None
        self._stop = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = torch.cuda.get_device_properties(self.index).uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if self.index < len(ids) and ids[self.index].strip().isdigit():
                    idx = int(ids[self.index])
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self, threaded=True):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            if threaded:
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
            else:
                self.sample_once()
            return
        except Exception:
            self.nvml = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def sample_once(self):
        n = self.nvml
        if not n:
            return
        try:
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                why = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                why = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((sm, self.max_mhz, why))
        except Exception:
            pass

    def _poll(self):
        while not self._stop:
            self.sample_once()
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """Blocks until the first sample is in: the sampler's start-up (NVML initialisation takes driver locks) must not overlap the
        first timed loop."""
        t0 = time.perf_counter()
        while (self.nvml or self.proc) and not (self.samples or self.lines) and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if self.nvml:
            self._stop = True
            if self.thread:
                self.thread.join(timeout=1)
            sm = [s[0] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= s[2]
            n = self.nvml
            names = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz if sm else None,
                    "reasons": sorted(k for k, b in names if bits & b), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference path on the host cores.  kind "reference" = the UNMODIFIED reference model code (model/model.py, model_zoo.py,
# loss.py byte-compiled into oracle/_ref by oracle/build_ref.py) through oracle/dgl_shim (DGL 0.4.0 is not installable offline);
# kind "port" = the oracle restatement (oracle/taxo_oracle.py), used when oracle/_ref has not been built.
# ------------------------------------------------------------------------------------------------
WORKLOAD = "configs[1]: MAG-CS PGAT+WMR+LBM, d=250, 2-hop egonets, batch=256 queries x 32 = 8192 egonets per GPU"


WORDNET = dict(in_dim=300, hidden_dim=600, out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1],
               feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)     # config_files/config.wordnet.json:11-20
ARCHS = {"mag-cs": ("PGAT", "WMR", "LBM", MAGCS), "wordnet": ("PGCN", "MR", "BIM", WORDNET)}     # BASELINE configs[1] / configs[0]


def _cpu_batches(n_queries, n_batches, rank=0, shape_model="mag-cs"):
    from taxoexpan_b200 import synth
    d = ARCHS[shape_model][3]["in_dim"]
    out = []
    for b in range(n_batches):
        shapes = synth.sample_shapes(n_queries, NEGATIVE_SIZE, shape_model, seed=20200420 + 1000 * rank + b)   # = the B200 arm's batches
        x = torch.from_numpy(synth.unit_rows(shapes.total_nodes, d, seed=11 + 1000 * rank + b))
        qf = torch.from_numpy(synth.unit_rows(shapes.num_graphs, d, seed=13 + 1000 * rank + b))
        out.append((shapes, x, qf))
    return out


def cpu_reference_run(n_queries, steps, warmup, n_batches=1, kind=None, arch="mag-cs"):
    """fwd + InfoNCE + bwd (dropout 0.1 active, no optimizer, no batch construction in the timed region: BASELINE.md section 4) of the
    MAG-CS config on all host cores.  Returns egonets/s over `steps` steps rotating over `n_batches` seeded batches."""
    from oracle import build_ref
    from oracle import taxo_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if kind is None:
        kind = "reference" if build_ref.available() else "port"
    data = _cpu_batches(n_queries, n_batches, shape_model=arch)
    pm, rm, mm, dims = ARCHS[arch]
    torch.manual_seed(0)
    if kind == "reference":
        TaxoExpan, info_nce_loss, dgl = build_ref.import_reference()
        model = TaxoExpan(pm, rm, mm, **dims)
        model.train()
        graphs = []
        for shapes, x, qf in data:
            gs, off = [], 0
            for a, s_ in zip(shapes.n_gp.tolist(), shapes.n_sib.tolist()):       # data_loader/dataset.py:429-435, one DGLGraph per egonet
                n = a + 1 + s_
                gr = dgl.DGLGraph()
                gr.add_nodes(n, {"x": x[off:off + n], "_id": torch.arange(off, off + n), "pos": torch.tensor([0] * a + [1] + [2] * s_)})
                gr.add_edges(list(range(a)), a)
                gr.add_edges(a, list(range(a + 1, n)))
                gr.add_edges(gr.nodes(), gr.nodes())
                gs.append(gr)
                off += n
            bg = dgl.batch(gs)                                                    # data_loader/data_loaders.py:25
            graphs.append((bg, bg.ndata.pop("x"), bg.ndata["pos"].clone(), qf))
        target = torch.zeros(n_queries, dtype=torch.long)

        def step(i):
            bg, h, pos, qf = graphs[i % len(graphs)]
            bg.ndata["pos"] = pos                                                 # PGAT.forward pops it (model_zoo.py:212)
            model.zero_grad()
            scores = model(bg, h, qf)                                             # trainer/trainer.py:51
            loss = info_nce_loss(scores.reshape(n_queries, -1), target)           # trainer.py:52-56, model/loss.py:52-57
            loss.backward()                                                       # trainer.py:60
            return float(loss.detach())
        sizes = [(int(bg.number_of_nodes()), int(bg._src.numel())) for bg, _, _, _ in graphs]
    else:
        cfg = orc.OracleConfig(propagation_method=pm, readout_method=rm, matching_method=mm, **{k: v for k, v in dims.items()})
        ogs = [(orc.batch_star_egonets(sh.n_gp, sh.n_sib), x, qf) for sh, x, qf in data]
        params = {k: v.clone().requires_grad_(True) for k, v in orc.init_model_params(cfg, seed=3).items()}

        def step(i):
            og, x, qf = ogs[i % len(ogs)]
            for p in params.values():
                p.grad = None
            masks = orc.random_keep_masks(cfg, og, seed=i)           # nn.Dropout's bernoulli is part of the reference step
            scores, _, _ = orc.taxoexpan_forward(cfg, og, x, qf, params, masks=masks, training=True)
            loss = orc.info_nce_step_loss(scores, n_queries)
            loss.backward()
            return float(loss.detach())
        sizes = [(og.n, int(og.src.numel())) for og, _, _ in ogs]

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    g = data[0][0].num_graphs
    return {"value": g / dt, "ms_per_step": dt * 1e3, "cores": cores, "egonets": g, "nodes": int(np.mean([s_[0] for s_ in sizes])),
            "edges": int(np.mean([s_[1] for s_ in sizes])), "kind": kind}


def _cpu_sample_text(r, nq, n_batches):
    what = ("the UNMODIFIED reference model/model.py + model_zoo.py + loss.py (byte-compiled, oracle/_ref) on torch-CPU through oracle/dgl_shim"
            if r["kind"] == "reference" else "torch-CPU port (oracle/taxo_oracle.py) of reference model_zoo.py PGAT/WMR/LBM + InfoNCE")
    return (f"{nq} queries x {1 + NEGATIVE_SIZE} = {r['egonets']} egonets ({r['nodes']} nodes) per step, {n_batches} rotating batch(es) of the "
            f"MAG-CS config, TaxoExpan.forward + info_nce_loss + backward, dropout 0.1 active, {r['cores']} host threads: {what} "
            "(real DGL 0.4.0 is not installable offline)")


def run_reference(args, rank):
    if rank != 0:
        return
    nq, nb = args.queries, args.batches
    r = cpu_reference_run(nq, args.steps, args.warmup, n_batches=nb)
    port = cpu_reference_run(32, 4, 1, kind="port") if r["kind"] == "reference" else None
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "egonets_per_gpu_step": r["egonets"], "nodes_per_gpu_step": r["nodes"],
                       "edges_per_gpu_step": r["edges"], "dropout": 0.1, "parallelism": "host cores (one process)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": _cpu_sample_text(r, nq, nb)},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if port is not None:
        line["cpu_port"] = {"value": port["value"], "unit": UNIT, "cores": port["cores"], "kind": "port",
                            "sample": _cpu_sample_text(port, 32, 1)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# Side legs (never the headline): the other BASELINE.json configs on one GPU, each a few seconds.
# ------------------------------------------------------------------------------------------------
def _event_ms(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def leg_pgcn_wordnet(dev, hbm_peak, with_cpu):
    """configs[0]: SemEval-Noun PGCN+MR(+BIM), d=300, batch=32 queries x 32 = 1024 egonets, fwd + InfoNCE + bwd; CPU arm beside it."""
    import taxoexpan_b200 as tx
    from taxoexpan_b200 import synth
    from taxoexpan_b200._lib import Stats
    pm, rm, mm, dims = ARCHS["wordnet"]
    nq = 32
    torch.manual_seed(0)
    model = tx.TaxoExpan(pm, rm, mm, **dims).to(dev).train()
    data = []
    for b in range(4):
        sh = synth.sample_shapes(nq, NEGATIVE_SIZE, "wordnet", seed=20200420 + b)
        data.append((sh, torch.from_numpy(synth.unit_rows(sh.total_nodes, dims["in_dim"], seed=11 + b)).to(dev),
                     torch.from_numpy(synth.unit_rows(sh.num_graphs, dims["in_dim"], seed=13 + b)).to(dev)))

    def step(i):
        sh, x, qf = data[i % 4]
        model.zero_grad(set_to_none=True)
        g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
        loss = tx.info_nce_loss(model(g, x, qf).reshape(nq, -1), None)
        loss.backward()
    ms = _event_ms(step, 30, 5)
    Stats.reset()
    Stats.profiling = True
    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    Stats.profiling = False
    prof = Stats.timings_ms()
    n = float(np.mean([d[0].total_nodes for d in data]))
    e = float(np.mean([d[0].total_edges for d in data]))
    kern, rl = {}, []
    for (name, tag), v in sorted(prof.items()):
        kern[f"{name}[{tag}]"] = round(float(np.sum(v)) / 8, 4)
    # GCN aggregate (model_zoo.py:39-47): fwd reads y once, writes out once (+ norm, CSR); bwd reads g once, writes dy once
    for tag, w in (("L0", dims["hidden_dim"]), ("L1", dims["out_dim"])):
        for nm in ("tx_gcn_aggregate_fwd", "tx_gcn_aggregate_bwd"):
            t = float(np.mean(prof.get((nm, tag), [0.0])))
            if t > 0:
                by = 4 * (2 * n * w + n) + 4 * (n + 1 + e)
                rl.append({"kernel": f"{nm}[{tag}]", "ms": round(t, 4), "bytes": int(by), "achieved": round(by / t / 1e6, 1),
                           "frac": round(by / t / 1e6 / hbm_peak, 4)})
    out = {"workload": "configs[0]: SemEval-Noun PGCN+MR+BIM, d=300, 2-hop egonets, batch=32 queries x 32 = 1024 egonets",
           "egonets_per_step": nq * 32, "nodes_per_step": int(n), "ms_per_step": round(ms, 4), "egonets_per_s": round(nq * 32 / ms * 1e3, 1),
           "kernel_ms_per_step": kern, "roofline_gcn_aggregate": rl,
           "note": "4.5 k nodes per step: every kernel of this config is launch-latency bound (a few us each), not bandwidth bound"}
    if with_cpu:
        r = cpu_reference_run(nq, 10, 2, n_batches=1, arch="wordnet")
        out["cpu_baseline"] = {"value": round(r["value"], 1), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                               "sample": f"the same config, {r['egonets']} egonets ({r['nodes']} nodes) per step x 10 steps on the host cores"}
    return out


def leg_inference(dev):
    """configs[3]: MAG-Full-shaped inference (test_fast.py:149-225, --batch_size 30000): forward-only encode of 30 000-egonet chunks,
    then all-pairs scoring + ranking of 2048 queries against the encoded positions."""
    import taxoexpan_b200 as tx
    from taxoexpan_b200 import synth
    torch.manual_seed(0)
    model = tx.TaxoExpan("PGAT", "WMR", "LBM", **MAGCS).to(dev).eval()
    name = "mag-full" if "mag-full" in synth.SHAPE_MODELS else "mag-cs"
    chunks = []
    for c in range(3):
        sh = synth.sample_shapes(30000, 0, name, seed=100 + c, positives=False)
        chunks.append((sh, torch.from_numpy(synth.unit_rows(sh.total_nodes, 250, seed=200 + c)).to(dev)))
    G = sum(sh.num_graphs for sh, _ in chunks)
    N = sum(sh.total_nodes for sh, _ in chunks)
    hold = {}

    def encode(i):
        hold["hg"] = tx.inference.encode_positions(model, ((tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib), x) for sh, x in chunks))
    ms = _event_ms(encode, 5, 2)
    hg = hold["hg"]
    Q = 2048
    queries = torch.from_numpy(synth.unit_rows(Q, 250, seed=7)).to(dev)
    rng = np.random.default_rng(0)
    positives = [rng.choice(hg.shape[0], size=2, replace=False).tolist() for _ in range(Q)]
    tx.inference.score_and_rank(model, hg, queries[:256], positives[:256])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = tx.inference.score_and_rank(model, hg, queries, positives)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"workload": f"configs[3]: {name}-shaped inference, PGAT+WMR+LBM forward-only, chunks of 30000 egonets (test_fast.py:160)",
            "encode": {"egonets": G, "nodes": N, "ms": round(ms, 3), "egonets_per_s": round(G / ms * 1e3, 1),
                       "includes": "host construction of each EgonetBatch + its structure kernel + propagate + readout"},
            "score_and_rank": {"queries": Q, "positions": int(hg.shape[0]), "ms": round(dt * 1e3, 2),
                               "pairs_per_s": round(Q * hg.shape[0] / dt, 1), "macro_mr": round(tx.inference.macro_mr(res["ranks"]), 1),
                               "includes": "U = hg W, one NT GEMM per 256-query chunk, on-GPU ranks + top-5, D2H of the results (wall clock)"}}


def leg_d512_sweep(dev, hbm_peak):
    """configs[4] on one GPU: d = 512, three propagation layers (num_layers = 2, heads [4, 4, 1], pos_dim 64; SURVEY 8d "config 5"),
    fwd + bwd over G = 2^13 .. 2^16 egonets per step (TAXO_SWEEP_MAX_LOG2 raises the cap), plus the general-CSR stress variant: one
    power-law graph (in-degrees up to 10^4) through a 2-layer GAT on the general kernels."""
    import taxoexpan_b200 as tx
    from taxoexpan_b200 import synth
    dims = dict(in_dim=512, hidden_dim=512, out_dim=512, pos_dim=64, num_layers=2, heads=[4, 4, 1], feat_drop=0.1, attn_drop=0.1,
                hidden_drop=0.1, out_drop=0.1)
    torch.manual_seed(0)
    model = tx.TaxoExpan("PGAT", "WMR", "LBM", **dims).to(dev).train()
    rows = []
    top = int(os.environ.get("TAXO_SWEEP_MAX_LOG2", "16"))
    for lg in range(13, top + 1):
        G = 1 << lg
        nq = G // 32
        sh = synth.sample_shapes(nq, NEGATIVE_SIZE, "mag-cs", seed=500 + lg)
        x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 512, seed=600 + lg)).to(dev)
        qf = torch.from_numpy(synth.unit_rows(sh.num_graphs, 512, seed=700 + lg)).to(dev)

        def step(i):
            model.zero_grad(set_to_none=True)
            g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
            tx.info_nce_loss(model(g, x, qf).reshape(nq, -1), None).backward()
        ms = _event_ms(step, 6 if lg >= 15 else 12, 3)
        n, e = sh.total_nodes, sh.total_edges
        gaa = sum(gaa_fwd_bytes(n, e, h, 512 * h) + gaa_bwd_bytes(n, e, h, 512 * h) for h in (4, 4, 1))
        rows.append({"egonets": G, "nodes": n, "ms_per_step": round(ms, 3), "egonets_per_s": round(G / ms * 1e3, 1),
                     "gaa_algorithmic_gb": round(gaa / 1e9, 3), "gaa_only_roofline_ms": round(gaa / hbm_peak / 1e6, 3)})
        del x, qf
        torch.cuda.empty_cache()
    # general-CSR stress: N = 2^17 nodes, Zipf in-degrees capped at 10^4
    rng = np.random.default_rng(3)
    n = 1 << 17
    deg = np.minimum(rng.zipf(1.6, n), 10_000).astype(np.int64)
    deg[rng.choice(n, 8, replace=False)] = 10_000
    dst = np.repeat(np.arange(n), deg)
    src = rng.integers(0, n, dst.shape[0])
    g = tx.DGLGraph()
    g.add_nodes(n)
    g.add_edges(torch.from_numpy(src), torch.from_numpy(dst))
    gat = tx.GAT(256, 128, 128, 1, [4, 1], torch.nn.functional.leaky_relu, 0.1, 0.1).to(dev).train()
    xg = torch.from_numpy(synth.unit_rows(n, 256, seed=9)).to(dev)
    w = torch.full((n, 128), 1e-3, device=dev)

    def gstep(i):
        gat.zero_grad(set_to_none=True)
        (gat(g, xg) * w).sum().backward()
    gms = _event_ms(gstep, 6, 2)
    e = int(dst.shape[0])
    by = sum(gaa_fwd_bytes(n, e, h, 128 * h) + gaa_bwd_bytes(n, e, h, 128 * h) for h in (4, 1))
    return {"workload": "configs[4] (one GPU): PGAT+WMR+LBM d=512, 3 propagation layers, heads [4,4,1], fwd+bwd, sweep over egonets per step",
            "sweep": rows,
            "general_csr_stress": {"nodes": n, "edges": e, "max_in_degree": int(deg.max()), "model": "GAT 256 -> 4x128 -> 128, fwd+bwd, general-CSR kernels",
                                   "ms_per_step": round(gms, 3), "gaa_algorithmic_gb": round(by / 1e9, 3),
                                   "edges_per_s": round(e / gms * 1e3, 1)}}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    import taxoexpan_b200 as tx
    from taxoexpan_b200 import _lib, synth
    from taxoexpan_b200 import functional as txf
    from taxoexpan_b200._lib import Stats

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False          # parity bar is fp32 1e-5
    # one process per GPU: give every rank its own slice of the host cores (the enqueue thread of a rank, its autograd thread and
    # NCCL's proxy threads otherwise migrate over / pile up on the same cores; a step is host-bound at ~1.6 ms of enqueue)
    host_info = {"cpus_visible": None, "cpus_pinned": None}
    try:
        cpus = sorted(os.sched_getaffinity(0))
        host_info["cpus_visible"] = len(cpus)
        lw = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        if world > 1 and args.pin_cores and len(cpus) >= 2 * lw:
            per = len(cpus) // lw
            mine = cpus[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            host_info["cpus_pinned"] = len(mine)
    except (AttributeError, OSError):
        pass
    _lib.load()
    nq = args.queries
    nb = args.batches

    torch.manual_seed(0)
    model = tx.TaxoExpan("PGAT", "WMR", "LBM", **MAGCS).to(dev)
    if world > 1:   # identical replicas
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    model.train()
    from taxoexpan_b200.dist import FlatGradBucket
    # gradients live in one flat fp32 bucket -> ONE all-reduce per step, issued as soon as the last gradient is final.  Two segments with
    # the first one travelling during layer 0's backward (TAXO_BUCKET_SEGMENTS=2) measured no better on 4 and 8 ranks when reduced on the
    # spot (1.901 vs 1.872 ms, 1.929 vs 1.922 ms: NCCL's CTAs take SMs the persistent one-CTA-per-SM kernels count on - the layer-0
    # star backward ends 0.04-0.07 ms later, which is what the overlap saved) and within the noise when gated behind that kernel
    # (FlatGradBucket(gate=True): 1.870 vs 1.889 ms resident, 2.002 vs 1.984 ms end to end on 4 ranks); profiles/r2_nccl_overlap.log.
    bucket = FlatGradBucket(model.parameters(), segments=int(os.environ.get("TAXO_BUCKET_SEGMENTS", "1")),
                            overlap=os.environ.get("TAXO_BUCKET_OVERLAP", "1") not in ("", "0"),
                            gate=os.environ.get("TAXO_BUCKET_GATE", "1") not in ("", "0"))
    flat = bucket.flat

    # rotating seeded batches (different shapes/features per rank and per slot), host copies pinned.  The same rows are available two
    # ways: as per-batch feature matrices (x_host / qf_host: what a DataLoader collating rows on the host would hand over) and as node
    # ids into a node-embedding table resident in HBM (ids_host / qids_host; table = the batch's rows, every row referenced once)
    batches = []
    for b in range(nb):
        shapes = synth.sample_shapes(nq, NEGATIVE_SIZE, "mag-cs", seed=20200420 + 1000 * rank + b)
        n = shapes.total_nodes
        x = torch.from_numpy(synth.unit_rows(n, MAGCS["in_dim"], seed=11 + 1000 * rank + b)).pin_memory()
        qf = torch.from_numpy(synth.unit_rows(shapes.num_graphs, MAGCS["in_dim"], seed=13 + 1000 * rank + b)).pin_memory()
        g = tx.EgonetBatch.from_counts(shapes.n_gp, shapes.n_sib).pin_memory()
        rng_b = np.random.default_rng(77 + 1000 * rank + b)
        perm = rng_b.permutation(n + shapes.num_graphs).astype(np.int32)        # rows of the table in shuffled order: a real gather
        table = torch.empty((n + shapes.num_graphs, MAGCS["in_dim"]), dtype=torch.float32)
        table[torch.from_numpy(perm[:n].astype(np.int64))] = x
        table[torch.from_numpy(perm[n:].astype(np.int64))] = qf
        batches.append(dict(shapes=shapes, x_host=x, qf_host=qf, graph=g, x=x.to(dev), qf=qf.to(dev), table=table.to(dev),
                            ids_host=torch.from_numpy(perm[:n].copy()).pin_memory(), qids_host=torch.from_numpy(perm[n:].copy()).pin_memory()))
        g.structure(dev)
    torch.cuda.synchronize()

    def fwd_bwd(g, x, qf):
        bucket.zero_()
        scores = model(g, x, qf)                                              # trainer.py:51
        loss = tx.info_nce_loss(scores.reshape(nq, -1), None)                 # trainer.py:52-56, loss.py:52-57 (target = zeros)
        loss.backward()                                                       # trainer.py:60
        bucket.all_reduce()                                                   # the only exchange of the path (no-op at N = 1)
        return loss

    def step_resident(i):
        b = batches[i % nb]
        g = b["graph"]
        g.ndata["pos"] = tx.graph._LazyPos(g)        # PGAT.forward pops 'pos' (model_zoo.py:212)
        return fwd_bwd(g, b["x"], b["qf"])

    # ---- end to end through the public API from HOST buffers: a prefetching loader (copy stream) overlaps the H2D of step
    # i+1 (features, queries, egonet counts: 45.6 MB) with the compute of step i, as a DataLoader(pin_memory=True) feeding
    # trainer.py:44-48 would; every copy and every loss read-back happens inside the timed region ----
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    e2e_mode = {"ids": True}       # True: ship node ids, gather rows from the resident table on the GPU; False: ship the feature rows

    def prefetch(i):
        b = batches[i % nb]
        sh = b["shapes"]
        with torch.cuda.stream(copy_stream):
            g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)    # fresh batch object: structure is rebuilt from host counts
            g._packed = b["graph"]._packed                       # reuse the pinned staging buffer
            g.stage(dev)
            if e2e_mode["ids"]:
                x = b["ids_host"].to(dev, non_blocking=True)
                qf = b["qids_host"].to(dev, non_blocking=True)
            else:
                x = b["x_host"].to(dev, non_blocking=True)
                qf = b["qf_host"].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return g, x, qf, ev, b["table"]

    e2e_state = {"next": None, "pending": []}
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(4)]     # pinned landing slots of the per-step loss

    host_t = {"prefetch": 0.0, "fwd_bwd": 0.0, "item": 0.0, "n": 0}

    def step_e2e(i):
        t0 = time.perf_counter()
        if e2e_state["next"] is None:
            e2e_state["next"] = prefetch(i)
        g, x, qf, ev, table = e2e_state["next"]
        e2e_state["next"] = prefetch(i + 1)                      # overlaps with this step's compute
        main_stream.wait_event(ev)
        for t in (x, qf, g._staged):
            t.record_stream(main_stream)
        t1 = time.perf_counter()
        if e2e_mode["ids"]:                                      # rows from the table resident in HBM (tx_gather_rows)
            x, qf = txf.gather_rows(table, x), txf.gather_rows(table, qf)
        loss = fwd_bwd(g, x, qf)
        slot = loss_host[i % 4]
        slot.copy_(loss.detach(), non_blocking=True)             # D2H of this step's result, every step, asynchronously ...
        ev_l = torch.cuda.Event()
        ev_l.record(main_stream)
        e2e_state["pending"].append((slot, ev_l))
        t2 = time.perf_counter()
        if len(e2e_state["pending"]) > 2:                        # ... and read on the host two steps later (a logging loop that blocks on the
            s_old, e_old = e2e_state["pending"].pop(0)           # previous step's loss serialises host enqueue and GPU work: measured +0.5 ms/step)
            e_old.synchronize()
            float(s_old)
        t3 = time.perf_counter()
        host_t["prefetch"] += t1 - t0; host_t["fwd_bwd"] += t2 - t1; host_t["item"] += t3 - t2; host_t["n"] += 1

    def e2e_flush():
        for s_old, e_old in e2e_state["pending"]:
            e_old.synchronize()
            float(s_old)
        e2e_state["pending"] = []
        e2e_state["next"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, profile=False, flush=None):
        barrier()
        Stats.reset()
        Stats.profiling = profile
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = time.perf_counter()
        e0.record()
        for i in range(steps):
            step_fn(i)
            if sampler_mode == "inline" and rank == 0 and (i + 1) % max(steps // 3, 1) == 0 and i + 1 < steps:
                sampler.sample_once()      # clocks / throttle reasons from the enqueue thread itself, twice per loop
        if flush is not None:
            flush()
        e1.record()
        timed.host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / max(steps, 1)
        barrier()
        Stats.profiling = False
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), Stats.total_launches(), Stats.timings_ms() if profile else {}

    sampler_mode = os.environ.get("TAXO_SAMPLER", "inline")     # inline (default) | thread | off
    sampler = ClockSampler(local_rank)
    if rank == 0 and sampler_mode != "off":
        sampler.start(threaded=sampler_mode == "thread")          # sampled across the value, per-kernel and e2e loops (a 30-step loop alone is ~0.1 s)
    for i in range(nb):          # setup: touch every rotating batch once (its tile table, the allocator's block sizes for its shape);
        step_resident(i)         # a batch first seen inside the timed region costs a cudaMalloc storm of ~100 ms
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    if rank == 0 and sampler.thread is not None:
        sampler.wait_ready()     # nvidia-smi's start-up stays outside the timed loops; from here on it only polls every 100 ms
    for i in range(max(args.warmup, 3)):      # second warm-up pass with the sampler polling: on some boxes the first loop after its
        step_resident(i)                      # start enqueued at ~3 ms/step instead of 1.6 (host-bound), whatever came next was normal
    import gc
    gc.collect()
    gc.freeze()                               # setup objects out of the collector's way: no generation-2 pause inside a 60 ms loop
    total_ms, launches, _ = timed(step_resident, args.steps)
    host_enqueue_ms = timed.host_enqueue_ms

    # per-kernel CUDA-event timings (separate pass so the headline loop carries no event overhead)
    _, _, prof = timed(step_resident, args.steps, profile=True)
    for i in range(max(3, nb)):
        step_e2e(i)
    e2e_flush()
    host_t.update(prefetch=0.0, fwd_bwd=0.0, item=0.0, n=0)
    e2e_ms, _, _ = timed(step_e2e, args.steps, flush=e2e_flush)
    e2e_host = {k: round(v / max(host_t["n"], 1) * 1e3, 4) for k, v in host_t.items() if k != "n"}   # host ms per step by phase
    # the same loop shipping the feature ROWS (45.6 MB per step) instead of node ids: what round 1 reported as e2e
    e2e_mode["ids"] = False
    for i in range(max(3, nb)):
        step_e2e(i)
    e2e_flush()
    e2e_rows_ms, _, _ = timed(step_e2e, args.steps, flush=e2e_flush)
    e2e_mode["ids"] = True
    clocks = sampler.stop() if rank == 0 else None

    # propagate + readout alone (graph_propagate -> readout, fwd + bwd against a fixed upstream gradient; no matching, no
    # loss, no all-reduce).  A side measurement: it never touches the headline, and a failure is reported instead of raised.
    prop_ro = None
    try:
        gout = {}

        def step_prop_readout(i):
            b = batches[i % nb]
            g = b["graph"]
            g.ndata["pos"] = tx.graph._LazyPos(g)
            bucket.zero_()
            pos = g.ndata["pos"].to(dev)
            g.ndata["h"] = model.graph_propagate(g, b["x"])
            hg = model.readout(g, pos)
            go = gout.get(tuple(hg.shape))
            if go is None:
                go = gout[tuple(hg.shape)] = torch.full_like(hg, 1e-3)
            hg.backward(go)

        bucket.active = False                # this leg exchanges nothing
        for i in range(max(3, nb)):
            step_prop_readout(i)
        pr_ms, _, _ = timed(step_prop_readout, args.steps)
        bucket.active = True
        prop_ro = {"ms_per_step": round(pr_ms / args.steps, 4)}
    except Exception as e:      # noqa: BLE001 - diagnostic leg only
        prop_ro = {"error": f"{type(e).__name__}: {e}"[:200]}

    # isolated host->device bandwidth of the pinned feature buffer (explains e2e: the 45.6 MB/step copy runs on its own stream one
    # step ahead, so e2e = max(GPU step, H2D time) whenever the host keeps up)
    torch.cuda.synchronize()
    hx = batches[0]["x_host"]
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        hx.to(dev, non_blocking=True)
        c0.record(copy_stream)
        for _ in range(5):
            hx.to(dev, non_blocking=True)
        c1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_gbs = 5 * hx.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9

    # totals over ranks
    egonets = torch.tensor([sum(batches[i % nb]["shapes"].num_graphs for i in range(args.steps))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(egonets)
    total_egonets = float(egonets.item())
    if rank != 0:
        return

    value = total_egonets / (total_ms * 1e-3)
    if prop_ro and "ms_per_step" in prop_ro:
        prop_ro["egonets_per_s"] = round(total_egonets / (prop_ro["ms_per_step"] * args.steps * 1e-3), 1)
    e2e_value = total_egonets / (e2e_ms * 1e-3)
    b0 = batches[0]
    sh = b0["shapes"]
    n_avg = float(np.mean([b["shapes"].total_nodes for b in batches]))
    e_avg = float(np.mean([b["shapes"].total_edges for b in batches]))
    hbm_peak, tf_peak, peak_src = load_peaks()
    H0, H1 = MAGCS["heads"]
    W0, W1 = MAGCS["hidden_dim"] * H0, MAGCS["out_dim"] * H1

    def avg(name, tag):
        v = prof.get((name, tag), [])
        return float(np.mean(v)) if v else 0.0

    kern = {}
    for (name, tag), v in sorted(prof.items()):
        kern[f"{name}[{tag}]"] = round(float(np.sum(v)) / args.steps, 4)
    # ABI calls recorded INSIDE the gemm_* / split_dy regions (not added twice to the per-step sum)
    nested = ("tx_gemm_nt_tf32x3", "tx_gemm_tn_tf32x3", "tx_split_tf32", "tx_gemm_nt_f16x3", "tx_gemm_tn_f16x3", "tx_split_f16", "tx_absmax")
    in_region = {"tx_reduce_partials"} if txf.GEMM_BACKEND in ("tf32x3", "f16x3") else set()
    step_prof_ms = sum(v for k, v in kern.items() if not k.startswith(nested) or txf.GEMM_BACKEND == "cublas")
    rl = []
    for tag, H, W in (("L0", H0, W0), ("L1", H1, W1)):
        for fwd_names, label in ((("tx_gat_star_fwd",), "tx_gat_star_fwd"), (("tx_gat_fused_fwd_staged",), "tx_gat_fused_fwd_staged"),
                                 (("tx_gat_fused_fwd", "tx_gat_fused_fwd_f16"), "tx_gat_fused_fwd"),
                                 (("tx_gat_node_logits", "tx_gat_aggregate_fwd"), "tx_gat_node_logits+aggregate_fwd")):
            t_f = sum(avg(nm, tag) for nm in fwd_names)
            if t_f > 0:
                by = gaa_fwd_bytes(n_avg, e_avg, H, W)
                rl.append({"kernel": f"{label}[{tag}]", "ms": t_f, "bytes": by, "achieved": by / t_f / 1e6})
        for bwd_names, label in ((("tx_gat_star_bwd",), "tx_gat_star_bwd"), (("tx_gat_fused_bwd_staged",), "tx_gat_fused_bwd_staged"),
                                 (("tx_gat_fused_bwd",), "tx_gat_fused_bwd"),
                                 (("tx_epilogue_bwd", "tx_gat_aggregate_bwd_dst", "tx_gat_aggregate_bwd_src", "tx_gat_attn_grad_partials"),
                                  "tx_epilogue_bwd+aggregate_bwd_dst+src+attn_grad")):
            if not avg(bwd_names[-1] if len(bwd_names) == 1 else "tx_gat_aggregate_bwd_dst", tag):
                continue
            t_b = sum(avg(nm, tag) for nm in bwd_names)
            by = gaa_bwd_bytes(n_avg, e_avg, H, W)
            rl.append({"kernel": f"{label}[{tag}]", "ms": t_b, "bytes": by, "achieved": by / t_b / 1e6})
    for r in rl:
        r["frac"] = r["achieved"] / hbm_peak
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        # the capture is only quoted for the kernel sources it was taken on (scripts/make_profiles.py records their hash): a kernel
        # that changed since reports traffic = null instead of a stale number
        if tj.get("csrc_sha16") == csrc_sha16():
            traffic = tj.get("per_launch_bytes", {})
    for r in rl:
        r["traffic"] = traffic.get(r["kernel"])
    # the dominant HBM kernel = the fused gather-attend-aggregate launch that moves the most algorithmic bytes (the layer-0 backward: it
    # is also the longest of the four; roofline_all lists every launch)
    dom = max(rl, key=lambda r: r["bytes"]) if rl else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": round(dom["achieved"], 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(dom["frac"], 4), "traffic": dom.get("traffic"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(dom["bytes"]), "ms_per_launch": round(dom["ms"], 4)}
    gemm_ms = sum(v for k, v in kern.items() if k.startswith("gemm") or k.startswith("split_dy"))
    gemm_flops = 3 * gemm_flops_per_node(MAGCS) * n_avg - 2 * n_avg * MAGCS["in_dim"] * W0   # dz0 only for the 50 pos columns
    rows_bytes = int(b0["x_host"].numel() * 4 + b0["qf_host"].numel() * 4 + b0["graph"]._packed.numel() * 4)
    x_bytes = int(b0["ids_host"].numel() * 4 + b0["qids_host"].numel() * 4 + b0["graph"]._packed.numel() * 4)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(64, 6, 1)        # bounded sample: 2048 egonets per step, ~10-20 s of host work
        cpu = {"value": round(r["value"], 1), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": _cpu_sample_text(r, 64, 1)}

    other = None
    if world == 1 and not args.no_side_legs:
        other = {}
        del batches
        torch.cuda.empty_cache()
        for key, fn in (("configs[0]", lambda: leg_pgcn_wordnet(dev, hbm_peak, not args.no_cpu_baseline)),
                        ("configs[3]", lambda: leg_inference(dev)), ("configs[4]", lambda: leg_d512_sweep(dev, hbm_peak))):
            try:
                other[key] = fn()
            except Exception as e_:      # noqa: BLE001 - side measurements never break the headline
                other[key] = {"error": f"{type(e_).__name__}: {e_}"[:300]}
            torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD
                               + (f" (configs[2] sharding: {world} x 256 queries, one NCCL all-reduce of {flat.numel()} fp32 grads)" if world > 1 else ""),
                   "egonets_per_gpu_step": sh.num_graphs, "nodes_per_gpu_step": int(n_avg), "edges_per_gpu_step": int(e_avg),
                   "dropout": 0.1, "parallelism": f"dp{world} (egonet shards by query group)",
                   "l2": f"inputs larger than L2: per-step intermediates ~{(n_avg * (W0 * 3 + 2052 * 2 + W1 * 3) * 4) / 1e9:.2f} GB; {nb} rotating batches",
                   "dense": {"f16x3": "tcgen05 kind::f16 on fp16 hi/lo operand pairs, 3 MMAs per product (tx_gemm.cu), fp32-faithful",
                             "tf32x3": "tcgen05 3xTF32 (tx_gemm.cu), fp32-faithful"}.get(txf.GEMM_BACKEND, "torch.mm (cuBLAS fp32, TF32 off)")},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": x_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": round(e2e_ms / args.steps, 4), "host_ms_per_step": e2e_host,
                "inputs": "per step from pinned host memory: egonet counts + int32 node ids + int32 query ids; the node-embedding table is "
                          "resident in HBM (like the model's parameters) and the rows are gathered by tx_gather_rows inside the timed region",
                "feature_rows_variant": {"value": round(total_egonets / (e2e_rows_ms * 1e-3), 1), "ms_per_step": round(e2e_rows_ms / args.steps, 4),
                                         "h2d_bytes_per_step": rows_bytes, "h2d_gb_per_s_isolated": round(h2d_gbs, 2),
                                         "note": "the same loop shipping the fp32 feature ROWS of every batch instead of ids"}},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 4),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_all": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in rl],
        "gemm": {"ms_per_step": round(gemm_ms, 4), "tflops": round(gemm_flops / (gemm_ms * 1e-3) / 1e12, 2) if gemm_ms else None,
                 "flops_per_step": int(gemm_flops)},
        "kernel_ms_per_step": kern,
        "kernel_ms_sum": round(step_prof_ms, 4),
        "cpu_baseline": cpu,
        "host": host_info,
        "propagate_readout": prop_ro,
        "star_bwd_reruns": int(txf.star_bwd_reruns(dev).item()),
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=256, help="queries per GPU per step (x32 egonets)")
    ap.add_argument("--batches", type=int, default=4, help="distinct rotating synthetic batches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the configs[0] / [3] / [4] side measurements (N = 1 only)")
    ap.add_argument("--no-pin-cores", dest="pin_cores", action="store_false", help="N > 1: do not partition the host cores among the ranks")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
