"""Why is the end-to-end step slower than the resident step?  Times variants of the e2e loop (CUDA events, 30 steps)."""
import os, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import synth
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tx.TaxoExpan("PGAT", "WMR", "LBM", **bench.MAGCS).to(dev).train()
nq, nb = 256, 4
B = []
for b in range(nb):
    sh = synth.sample_shapes(nq, 31, "mag-cs", seed=20200420 + b)
    x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 250, seed=11 + b)).pin_memory()
    qf = torch.from_numpy(synth.unit_rows(sh.num_graphs, 250, seed=13 + b)).pin_memory()
    g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib).pin_memory()
    g.structure(dev)
    B.append(dict(sh=sh, xh=x, qh=qf, g=g, x=x.to(dev), q=qf.to(dev)))
target = torch.zeros(nq, dtype=torch.long, device=dev)
copy_stream = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream(dev)
def fwd_bwd(g, x, qf):
    model.zero_grad(set_to_none=False) if False else None
    for p in model.parameters():
        p.grad = None
    loss = F.cross_entropy(model(g, x, qf).reshape(nq, -1), target, reduction="sum")
    loss.backward()
    return loss
def run(name, fresh_graph, h2d, sync_item):
    state = {"next": None, "loss": None}
    def prefetch(i):
        b = B[i % nb]
        with torch.cuda.stream(copy_stream):
            if fresh_graph:
                g = tx.EgonetBatch.from_counts(b["sh"].n_gp, b["sh"].n_sib); g._packed = b["g"]._packed; g.stage(dev)
            else:
                g = b["g"]; g.ndata["pos"] = tx.graph._LazyPos(g)
            if h2d:
                x = b["xh"].to(dev, non_blocking=True); q = b["qh"].to(dev, non_blocking=True)
            else:
                x, q = b["x"], b["q"]
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return g, x, q, ev
    def step(i):
        if state["next"] is None: state["next"] = prefetch(i)
        g, x, q, ev = state["next"]
        state["next"] = prefetch(i + 1)
        main.wait_event(ev)
        if h2d:
            x.record_stream(main); q.record_stream(main)
        if fresh_graph: g._staged.record_stream(main)
        prev = state["loss"]
        state["loss"] = fwd_bwd(g, x, q)
        if sync_item and prev is not None: prev.item()
    for i in range(4): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(30): step(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:48s} gpu {e0.elapsed_time(e1)/30:.3f} ms/step   host {(time.perf_counter()-t0)/30*1e3:.3f} ms/step")
run("resident (no fresh graph, no H2D, no item)", False, False, False)
run("resident + item() of previous loss", False, False, True)
run("fresh graph", True, False, True)
run("H2D only", False, True, True)
run("fresh graph + H2D (= e2e)", True, True, True)
run("fresh graph + H2D, no item", True, True, False)
