"""Debug: per-phase cycle counters of the staged backward kernel (needs a build with TAXO_NVCC_FLAGS=-DTX_BWD_PROFILE)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import synth, _lib
import torch.nn.functional as F
import bench
dev = torch.device("cuda", 0)
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
torch.manual_seed(0)
model = tx.TaxoExpan("PGAT", "WMR", "LBM", **bench.MAGCS).to(dev).train()
nq = 256
sh = synth.sample_shapes(nq, 31, "mag-cs", seed=20200420)
x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 250, seed=11)).to(dev)
qf = torch.from_numpy(synth.unit_rows(sh.num_graphs, 250, seed=13)).to(dev)
g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
target = torch.zeros(nq, dtype=torch.long, device=dev)
for _ in range(3):
    g.ndata["pos"] = tx.graph._LazyPos(g)
    model.zero_grad()
    loss = F.cross_entropy(model(g, x, qf).reshape(nq, -1), target, reduction="sum")
    loss.backward()
torch.cuda.synchronize()
buf = np.zeros(148 * 4 * 3 * 6, dtype=np.int64)
raw.tx_debug_bwd_prof(buf.ctypes.data_as(ctypes.c_void_p))
# the last launch was the L0 backward (H = 4, 37 x 4 CTAs)
a = buf.reshape(-1, 3, 6)[:148]
tot = a.sum(0)
print("mode: tiles rows wait dots softmax phaseB   (cycles summed over 148 CTAs; per-CTA mean in parentheses)")
for m in range(3):
    print(m, tot[m].tolist(), [int(v / 148) for v in tot[m]])
print("per-CTA total cycles (mean, max):", a[:, :, 2:].sum((1, 2)).mean(), a[:, :, 2:].sum((1, 2)).max())
for m in range(3):
    if tot[m][0]:
        print(f"mode {m}: per tile: wait {tot[m][2]/tot[m][0]:.0f} dots {tot[m][3]/tot[m][0]:.0f} softmax {tot[m][4]/tot[m][0]:.0f} B {tot[m][5]/tot[m][0]:.0f}; rows/tile {tot[m][1]/tot[m][0]:.1f}")
