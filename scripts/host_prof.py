"""cProfile of the host side of one resident training step (where do the ~2 ms of Python / ctypes / torch overhead go?)."""
import cProfile, pstats, os, sys, io
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import synth
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tx.TaxoExpan("PGAT", "WMR", "LBM", **bench.MAGCS).to(dev).train()
nq = 256
sh = synth.sample_shapes(nq, 31, "mag-cs", seed=20200420)
x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 250, seed=11)).to(dev)
qf = torch.from_numpy(synth.unit_rows(sh.num_graphs, 250, seed=13)).to(dev)
g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
target = torch.zeros(nq, dtype=torch.long, device=dev)
def step():
    g.ndata["pos"] = tx.graph._LazyPos(g)
    model.zero_grad()
    loss = F.cross_entropy(model(g, x, qf).reshape(nq, -1), target, reduction="sum")
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
