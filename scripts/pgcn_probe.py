import os, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import taxoexpan_b200 as tx, bench
from taxoexpan_b200 import synth, functional as txf
dev = torch.device('cuda', 0)
for nq in (32, 256):
    for fused in (True, False):
        txf.GCN_FUSED = fused
        pm, rm, mm, dims = bench.ARCHS['wordnet']
        torch.manual_seed(0)
        model = tx.TaxoExpan(pm, rm, mm, **dims).to(dev).train()
        sh = synth.sample_shapes(nq, 31, 'wordnet', seed=20200420)
        x = torch.from_numpy(synth.unit_rows(sh.total_nodes, 300, seed=11)).to(dev)
        qf = torch.from_numpy(synth.unit_rows(sh.num_graphs, 300, seed=13)).to(dev)
        def step(i):
            model.zero_grad(set_to_none=True)
            g = tx.EgonetBatch.from_counts(sh.n_gp, sh.n_sib)
            tx.info_nce_loss(model(g, x, qf).reshape(nq, -1), None).backward()
        ms = bench._event_ms(step, 30, 5)
        print(f'PGCN wordnet dims, {nq} queries ({sh.total_nodes} nodes), fused={fused}: {ms:.3f} ms/step = {nq*32/ms*1e3:.0f} egonets/s')
