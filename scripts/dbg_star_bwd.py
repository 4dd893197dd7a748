import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import functional as txf
from oracle import taxo_oracle as orc
from tests.test_gpu_parity import build_model, MAGCS, dev

n_gp, n_sib = [40, 0, 33], [3, 170, 0]
cfg = orc.OracleConfig(**MAGCS)
params = orc.init_model_params(cfg, seed=5)
og = orc.batch_star_egonets(n_gp, n_sib)
x = torch.from_numpy(tx.synth.unit_rows(og.n, cfg.in_dim, seed=1))
qf = torch.from_numpy(tx.synth.unit_rows(og.num_graphs, cfg.in_dim, seed=2))
# fp64 oracle
p = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
h = x.double().clone().requires_grad_(True)
s, _, _ = orc.taxoexpan_forward(cfg, og, h, qf.double(), p)
s.sum().backward()
ref = {k: v.grad for k, v in p.items()}
for star in (True, False):
    for opt in (4.0, 1e-6):
        txf.STAR_BWD = star; txf.DFT_OPTIMISM = opt
        model = build_model(cfg, params).train()
        g = tx.EgonetBatch.from_counts(n_gp, n_sib)
        hh = x.to(dev()).requires_grad_(True)
        model(g, hh, qf.to(dev())).sum().backward()
        torch.cuda.synchronize()
        gs = max(float(v.abs().max()) for v in ref.values())
        print(f"star={star} optimism={opt}: dh rel {float((hh.grad.cpu().double()-h.grad).abs().max()/h.grad.abs().max()):.2e}")
        for k, v in model.named_parameters():
            r = ref[k]
            print(f"   {k:55s} rel-to-max {float((v.grad.cpu().double()-r).abs().max()/r.abs().max()):.2e}  (max {float(r.abs().max()):.3e}, gscale {gs:.3e})")
