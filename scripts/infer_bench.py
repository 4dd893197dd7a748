"""BASELINE.json configs[3] (MAG-Full inference, --batch_size 30000, test_fast.py:149-225): forward-only encode throughput of
30 000-egonet chunks and all-pairs scoring/ranking, on synthetic MAG-Full-shaped candidate positions."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import synth
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tx.TaxoExpan("PGAT", "WMR", "LBM", **bench.MAGCS).to(dev).eval()
name = "mag-full" if "mag-full" in synth.SHAPE_MODELS else "mag-cs"
chunks = []
for c in range(3):
    sh = synth.sample_shapes(30000, 0, name, seed=100 + c, positives=False)
    x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 250, seed=200 + c)).to(dev)
    chunks.append((sh, x))
def batches():
    for sh, x in chunks:
        yield tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib), x
hg = tx.inference.encode_positions(model, batches())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    hg = tx.inference.encode_positions(model, batches())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
G = sum(sh.num_graphs for sh, _ in chunks); N = sum(sh.total_nodes for sh, _ in chunks)
print(f"encode ({name}-shaped): {G} egonets / {N} nodes in {ms:.2f} ms  -> {G / ms * 1e3 / 1e6:.2f} M egonets/s forward-only "
      f"(EgonetBatch built from host counts inside the timed region; star forward {'on' if tx.functional.STAR_FWD else 'off'})")
t0 = time.perf_counter()
for _ in range(3):
    built = [(tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib), x) for sh, x in chunks]
t_build = (time.perf_counter() - t0) / 3 * 1e3
for g, _ in built:
    g.structure(dev)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    for g, _ in built:
        g.ndata["pos"] = tx.graph._LazyPos(g)
    hg = tx.inference.encode_positions(model, iter(built))
e1.record(); torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / 3
print(f"  host construction of the 3 batches: {t_build:.2f} ms; encode with resident structures: {ms2:.2f} ms -> {G / ms2 * 1e3 / 1e6:.2f} M egonets/s")
Q = 2048
queries = torch.from_numpy(synth.unit_rows(Q, 250, seed=7)).to(dev)
rng = np.random.default_rng(0)
positives = [rng.choice(hg.shape[0], size=2, replace=False).tolist() for _ in range(Q)]
tx.inference.score_and_rank(model, hg, queries[:256], positives[:256])
torch.cuda.synchronize(); t0 = time.perf_counter()
res = tx.inference.score_and_rank(model, hg, queries, positives)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"score+rank: {Q} queries x {hg.shape[0]} positions in {dt * 1e3:.1f} ms -> {Q * hg.shape[0] / dt / 1e9:.2f} G pairs/s; macro MR {tx.inference.macro_mr(res['ranks']):.1f}")
