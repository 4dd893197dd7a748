"""Generate tests/golden/*.npz by running the UNMODIFIED reference model code.

TEST INFRASTRUCTURE ONLY. Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz and checks the oracle restatement

What runs: /root/reference/model/model.py + model/model_zoo.py, byte-for-byte unmodified, on CPU torch,
with oracle/dgl_shim standing in for DGL 0.4.0 (not installable offline) and for the unused `ipdb` import.
Egonets are built through the same DGLGraph calls as data_loader/dataset.py:429-435 and batched with
dgl.batch as data_loader/data_loaders.py:25 does.

Each fixture stores: the egonet shapes, seeds, checksums of the seeded inputs/parameters (parameters and
features are regenerated from the seeds by the tests - storing 1.75 M floats per case would not be a small
fixture), and the reference's fp32 outputs (node states, readout, scores, loss) and gradients, plus the same
run in fp64 so tests know the reference's own fp32 noise.  The script finally asserts that
oracle/taxo_oracle.py reproduces the reference on every case (fp64: <= 1e-12).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TAXO_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "dgl_shim"), REF, ROOT]

import dgl  # noqa: E402  (the shim)
from model.model import TaxoExpan  # noqa: E402  (the reference, unmodified)

from oracle import taxo_oracle as orc  # noqa: E402
from taxoexpan_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (cfg kwargs, shapes spec, n_queries for InfoNCE)
    "pgat_wmr_lbm_small": dict(
        cfg=dict(propagation_method="PGAT", readout_method="WMR", matching_method="LBM", in_dim=12, hidden_dim=8,
                 out_dim=8, pos_dim=4, num_layers=1, heads=[4, 1]),
        n_gp=[1, 0, 2, 1, 3, 0, 1, 2], n_sib=[3, 0, 5, 0, 1, 4, 50, 7], n_queries=2),
    "pgat_wmr_lbm_3layer_small": dict(
        cfg=dict(propagation_method="PGAT", readout_method="WMR", matching_method="BIM", in_dim=16, hidden_dim=12,
                 out_dim=20, pos_dim=4, num_layers=2, heads=[2, 3, 2]),
        n_gp=[1, 0, 2, 1], n_sib=[3, 0, 5, 9], n_queries=2),
    "gat_mr_mlp_small": dict(
        cfg=dict(propagation_method="GAT", readout_method="MR", matching_method="MLP", in_dim=12, hidden_dim=8,
                 out_dim=8, pos_dim=4, num_layers=1, heads=[4, 1]),
        n_gp=[1, 0, 2, 1], n_sib=[3, 0, 5, 0], n_queries=2),
    "pgcn_mr_bim_small": dict(
        cfg=dict(propagation_method="PGCN", readout_method="MR", matching_method="BIM", in_dim=12, hidden_dim=8,
                 out_dim=8, pos_dim=4, num_layers=1, heads=[4, 1]),
        n_gp=[1, 0, 2, 1, 3, 0], n_sib=[3, 0, 5, 0, 1, 4], n_queries=2),
    "gcn_cr_lbm_small": dict(
        cfg=dict(propagation_method="GCN", readout_method="CR", matching_method="LBM", in_dim=12, hidden_dim=8,
                 out_dim=8, pos_dim=4, num_layers=2, heads=[4, 1]),
        n_gp=[1, 0, 2, 1], n_sib=[3, 0, 5, 0], n_queries=2),
    # BASELINE config 2 dims (config_files/config.mag.json:11-17), 2 queries x 32 egonets
    "pgat_wmr_lbm_magcs": dict(
        cfg=dict(propagation_method="PGAT", readout_method="WMR", matching_method="LBM", in_dim=250, hidden_dim=500,
                 out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1]),
        synth=dict(n_queries=2, negative_size=31, model="mag-cs", seed=7), n_queries=2),
    # BASELINE config 1 dims (config_files/config.wordnet.json:11-17), 2 queries x 32 egonets
    "pgcn_mr_bim_wordnet": dict(
        cfg=dict(propagation_method="PGCN", readout_method="MR", matching_method="BIM", in_dim=300, hidden_dim=600,
                 out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1]),
        synth=dict(n_queries=2, negative_size=31, model="wordnet", seed=11), n_queries=2),
    # BASELINE configs[1] at FULL size: 256 queries x 32 = 8192 egonets (the batch bench.py times), all gradients; large tensors are
    # stored sub-sampled (every 997th element of node_h / dh / hg, every 97th of the large parameter gradients)
    "pgat_wmr_lbm_magcs_full": dict(
        cfg=dict(propagation_method="PGAT", readout_method="WMR", matching_method="LBM", in_dim=250, hidden_dim=500,
                 out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1]),
        synth=dict(n_queries=256, negative_size=31, model="mag-cs", seed=20200420), n_queries=256, big_step=997),
    # BASELINE configs[0] at FULL size: SemEval-Noun / WordNet dims PGCN+MR(+BIM), 32 queries x 32 = 1024 egonets
    "pgcn_mr_bim_wordnet_full": dict(
        cfg=dict(propagation_method="PGCN", readout_method="MR", matching_method="BIM", in_dim=300, hidden_dim=600,
                 out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1]),
        synth=dict(n_queries=32, negative_size=31, model="wordnet", seed=20200420), n_queries=32, big_step=97),
}


def build_ref_graph(n_gp, n_sib, x, dtype):
    """One DGLGraph per egonet through the calls of dataset.py:429-435, then dgl.batch (data_loaders.py:25)."""
    graphs, off = [], 0
    for g, s in zip(n_gp, n_sib):
        g, s = int(g), int(s)
        n = g + 1 + s
        nodes_pos = [0] * g + [1] + [2] * s
        gr = dgl.DGLGraph()
        gr.add_nodes(n, {"x": x[off:off + n].to(dtype), "_id": torch.arange(off, off + n), "pos": torch.tensor(nodes_pos)})
        gr.add_edges(list(range(g)), g)
        gr.add_edges(g, list(range(g + 1, n)))
        gr.add_edges(gr.nodes(), gr.nodes())
        graphs.append(gr)
        off += n
    return dgl.batch(graphs)


def checksum(t):
    t = np.asarray(t, dtype=np.float64)
    return np.array([t.sum(), np.abs(t).sum()])


def subsample(a, step=97):
    return np.ascontiguousarray(np.asarray(a).reshape(-1)[::step])


def sub_rows(a, step):
    """every step-th ROW (scores / hg of the full-size cases)"""
    return np.ascontiguousarray(np.asarray(a)[::step])


def run_case(name, spec):
    cfg_kw = dict(spec["cfg"])
    cfg = orc.OracleConfig(**cfg_kw)
    if "synth" in spec:
        shapes = synth.sample_shapes(**spec["synth"])
        n_gp, n_sib = shapes.n_gp, shapes.n_sib
    else:
        n_gp, n_sib = np.asarray(spec["n_gp"]), np.asarray(spec["n_sib"])
    og = orc.batch_star_egonets(n_gp, n_sib)
    N, G = og.n, og.num_graphs
    x = torch.from_numpy(synth.unit_rows(N, cfg.in_dim, seed=101))
    qf = torch.from_numpy(synth.unit_rows(G, cfg.in_dim, seed=202))
    params = orc.init_model_params(cfg, seed=5)
    big = cfg.in_dim >= 100
    out = {"n_gp": n_gp, "n_sib": n_sib, "x_checksum": checksum(x), "qf_checksum": checksum(qf),
           "param_checksum": np.stack([checksum(v) for _, v in sorted(params.items())]),
           "feature_seed": np.array([101, 202]), "param_seed": np.array([5]), "n_queries": np.array([spec["n_queries"]])}
    res = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        model = TaxoExpan(cfg.propagation_method, cfg.readout_method, cfg.matching_method,
                          in_dim=cfg.in_dim, hidden_dim=cfg.hidden_dim, out_dim=cfg.out_dim, pos_dim=cfg.pos_dim,
                          num_layers=cfg.num_layers, heads=list(cfg.heads), feat_drop=0.0, attn_drop=0.0,
                          hidden_drop=0.0, out_drop=0.0)
        missing = model.load_state_dict(params, strict=True)
        model = model.to(dtype)
        model.train()  # dropout rates are 0: train == eval arithmetic, exercises the training graph
        bg = build_ref_graph(n_gp, n_sib, x, dtype)
        # bit-exact indexing pin: the reference-built edge list must equal the oracle's closed form
        assert torch.equal(bg._src, og.src) and torch.equal(bg._dst, og.dst)
        assert torch.equal(bg.ndata["pos"], og.pos) and bg.batch_num_nodes == og.batch_num_nodes
        h = bg.ndata.pop("x").clone().requires_grad_(True)          # trainer.py:48
        scores = model(bg, h, qf.to(dtype))                          # trainer.py:51
        node_h = bg.ndata["h"]
        pos = og.pos
        hg = model.readout(bg, pos)
        n_q = spec["n_queries"]
        pred = scores.reshape(n_q, -1)                               # trainer.py:54
        loss = F.cross_entropy(pred, torch.zeros(n_q, dtype=torch.long), reduction="sum")   # loss.py:57
        loss.backward()
        grads = {k: p.grad.detach().numpy() for k, p in model.named_parameters()}
        res[tag] = dict(scores=scores.detach().numpy(), node_h=node_h.detach().numpy(), hg=hg.detach().numpy(),
                        loss=loss.detach().numpy(), grads=grads, dh=h.grad.numpy())

        # ---- check the restatement against the reference ----
        p_o = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
        h_o = x.to(dtype).clone().requires_grad_(True)
        s_o, hg_o, nh_o = orc.taxoexpan_forward(cfg, og, h_o, qf.to(dtype), p_o)
        l_o = orc.info_nce_step_loss(s_o, n_q)
        l_o.backward()
        tol = 1e-12 if dtype == torch.float64 else 2e-5
        gscale = max(float(p.grad.abs().max()) for p in model.parameters())

        def rel(a, b, floor=0.0):
            a = torch.as_tensor(a); b = torch.as_tensor(b)
            return float((a - b).abs().max() / max(float(b.abs().max()), floor, 1e-30))
        errs = {"scores": rel(s_o.detach(), scores.detach()), "hg": rel(hg_o.detach(), hg.detach()),
                "node_h": rel(nh_o.detach(), node_h.detach()), "dh": rel(h_o.grad, h.grad)}
        for k, p in model.named_parameters():
            errs["d" + k] = rel(p_o[k].grad, p.grad, floor=1e-3 * gscale)  # near-cancelling grads: scale by the largest grad
        worst = max(errs.values())
        print(f"  [{name}/{tag}] oracle-vs-reference worst rel err {worst:.3e}")
        assert worst < tol, (name, tag, errs)

    f32, f64 = res["f32"], res["f64"]
    big_step = int(spec.get("big_step", 97))
    out["big_step"] = np.array([big_step])
    hg_rows = 1 if f32["hg"].size <= 200000 else 64          # full-size cases keep every 64th row of hg
    out["hg_row_step"] = np.array([hg_rows])
    for key in ("scores", "loss"):
        out[key] = f32[key]
        out[key + "_f64"] = f64[key]
    out["hg"], out["hg_f64"] = sub_rows(f32["hg"], hg_rows), sub_rows(f64["hg"], hg_rows)
    if big:
        out["node_h_sub"] = subsample(f32["node_h"], big_step)
        out["node_h_sub_f64"] = subsample(f64["node_h"], big_step)
        out["dh_sub"] = subsample(f32["dh"], big_step)
        out["dh_sub_f64"] = subsample(f64["dh"], big_step)
    else:
        out["node_h"], out["node_h_f64"] = f32["node_h"], f64["node_h"]
        out["dh"], out["dh_f64"] = f32["dh"], f64["dh"]
    for k in f32["grads"]:
        g32, g64 = f32["grads"][k], f64["grads"][k]
        if g32.size > 20000:
            out["grad_sub." + k], out["grad_sub_f64." + k] = subsample(g32), subsample(g64)
        else:
            out["grad." + k], out["grad_f64." + k] = g32, g64
        out["gradnorm." + k] = np.array([np.sqrt((g64.astype(np.float64) ** 2).sum())])
    noise = max(float(np.abs(f32[k] - f64[k]).max()) for k in ("scores", "hg", "node_h"))
    out["ref_fp32_noise"] = np.array([noise])
    print(f"  [{name}] N={N} E={og.src.numel()} G={G}; reference fp32-vs-fp64 max-abs noise {noise:.3e}")
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        print(name)
        run_case(name, spec)
    print("golden fixtures written to", GOLD)
