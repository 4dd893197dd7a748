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
from taxoexpan_b200.dist import FlatGradBucket
bucket = FlatGradBucket(model.parameters())
def step():
    g.ndata["pos"] = tx.graph._LazyPos(g)
    bucket.zero_()
    loss = tx.info_nce_loss(model(g, x, qf).reshape(nq, -1), None)
    loss.backward()
    bucket.all_reduce()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:9000])
import time
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host enqueue per step (no profiler): %.3f ms" % ((t1 - t0) / 50 * 1e3))
t0 = time.perf_counter()
for _ in range(50): tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
print("EgonetBatch.from_counts: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
