"""Debug: where does d(features) of the full-size MAG-CS golden case deviate?  (GPU)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxoexpan_b200 as tx
from taxoexpan_b200 import functional as txf
from tests._golden import load_case
from tests.test_gpu_parity import build_model, run_cuda, run_oracle

cfg, og, x, qf, params, fx = load_case("pgat_wmr_lbm_magcs_full")
nq = int(fx["n_queries"][0])
ref = run_oracle(cfg, og, x, qf, params, nq, dtype=torch.float64)
dh_ref = ref[5]
print("ref dh max", np.abs(dh_ref).max(), "rms", np.sqrt((dh_ref ** 2).mean()))
rown = np.abs(dh_ref).max(1)
print("row max quantiles", np.quantile(rown, [0, .01, .1, .5, .9, .99, 1]))
for backend in ("f16x3", "tf32x3", "cublas"):
    for fused in (True, False):
        txf.GEMM_BACKEND = backend
        txf.FUSED_ENABLED = fused
        model = build_model(cfg, params).train()
        g = tx.EgonetBatch.from_counts(fx["n_gp"], fx["n_sib"])
        got = run_cuda(model, g, x, qf, nq)
        d = np.abs(got[5].astype(np.float64) - dh_ref)
        i, j = np.unravel_index(d.argmax(), d.shape)
        gerr = {k: float(np.abs(got[4][k].astype(np.float64) - ref[4][k]).max() / np.abs(ref[4][k]).max()) for k in ref[4]}
        print(f"{backend} fused={fused}: dh max err {d.max():.3e} at row {i} col {j} (ref {dh_ref[i, j]:.3e}, row max {rown[i]:.3e}); "
              f"rel-to-max {d.max() / np.abs(dh_ref).max():.2e}; node_h err {np.abs(got[2] - ref[2]).max():.2e}; worst param grad rel {max(gerr.values()):.2e} ({max(gerr, key=gerr.get)})")
        # per-row relative error distribution
        rr = d.max(1) / np.maximum(rown, 1e-30)
        print("   per-row rel err quantiles", np.quantile(rr, [.5, .9, .99, 1]))
