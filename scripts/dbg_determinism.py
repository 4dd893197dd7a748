"""Run-to-run determinism of the default hot path: the same full-size step repeated N times in one process must give bit-identical
scores, node states and gradients (dropout off).  usage: python scripts/dbg_determinism.py [reps] ; TAXO_PDL=0 etc. select variants"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import _lib
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dims = dict(in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1], feat_drop=0.0, attn_drop=0.0,
            hidden_drop=0.0, out_drop=0.0)
torch.manual_seed(3)
model = tx.TaxoExpan("PGAT", "WMR", "LBM", **dims).to(dev).train()
nq = 256
shapes = [tx.synth.sample_shapes(nq, 31, "mag-cs", seed=100 + i) for i in range(2)]
data = []
for i, sh in enumerate(shapes):
    data.append((sh, torch.from_numpy(tx.synth.unit_rows(sh.total_nodes, 250, seed=i)).to(dev),
                 torch.from_numpy(tx.synth.unit_rows(sh.num_graphs, 250, seed=50 + i)).to(dev)))
# noise between the measured steps: other shapes / values so that recycled workspaces hold different scalars
noise = [(tx.synth.sample_shapes(64, 31, "mag-cs", seed=7), 37.0), (tx.synth.sample_shapes(16, 31, "mag-cs", seed=8), 0.01)]
noise = [(sh, torch.from_numpy(tx.synth.unit_rows(sh.total_nodes, 250, seed=3)).to(dev) * s,
          torch.from_numpy(tx.synth.unit_rows(sh.num_graphs, 250, seed=4)).to(dev)) for sh, s in noise]

def step(sh, x, qf, n):
    g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
    model.zero_grad(set_to_none=True)
    scores = model(g, x, qf)
    loss = tx.info_nce_loss(scores.reshape(n, -1), None)
    loss.backward()
    return [scores.detach().clone(), g.ndata['h'].detach().clone()] + [p.grad.detach().clone() for p in model.parameters()]

ref = [step(sh, x, qf, nq) for sh, x, qf in data]
torch.cuda.synchronize()
bad = 0
ref_file = os.environ.get("TAXO_DET_REF")          # cold-process mode: the FIRST step of a fresh process against a saved reference
if ref_file:
    cpu = [[t.cpu() for t in r] for r in ref]
    if not os.path.exists(ref_file):
        torch.save(cpu, ref_file)
        print("saved", ref_file)
    else:
        old = torch.load(ref_file)
        for k in range(len(cpu)):
            for j, (a, b) in enumerate(zip(old[k], cpu[k])):
                if not torch.equal(a, b):
                    bad += 1
                    print(f"COLD batch {k} tensor {j} shape {tuple(a.shape)}: {int((a != b).sum())} entries differ, max |diff| {float((a - b).abs().max()):.3e} (max |ref| {float(a.abs().max()):.3e})")
        print("cold check: mismatching tensors", bad)
    sys.exit(0)
for r in range(reps):
    for k, (sh, x, qf) in enumerate(data):
        if r % 3 == 0:
            nsh, nx, nqf = noise[(r // 3) % len(noise)]
            step(nsh, nx, nqf, nsh.num_graphs // 32)
        out = step(sh, x, qf, nq)
        for j, (a, b) in enumerate(zip(ref[k], out)):
            if not torch.equal(a, b):
                bad += 1
                d = (a - b).abs()
                print(f"rep {r} batch {k} tensor {j} shape {tuple(a.shape)}: {int((a != b).sum())} entries differ, max |diff| {float(d.max()):.3e} "
                      f"(max |ref| {float(a.abs().max()):.3e})")
torch.cuda.synchronize()
print("reps", reps, "mismatching tensors", bad, "PDL", os.environ.get("TAXO_PDL", "1"))
