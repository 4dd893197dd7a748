"""Drop-in surface against the UNMODIFIED reference (runs only where /root/reference exists, i.e. in the build container): constructor
and forward signatures of every module on the path, `info_nce_loss`, and the state_dict keys / shapes of the headline configurations.
The reference is imported in a subprocess through oracle/dgl_shim (its `model` package name would shadow nothing here, but its
`import dgl` must not leak into this process)."""
import inspect
import json
import os
import subprocess
import sys

import pytest

import taxoexpan_b200 as tx
from taxoexpan_b200 import loss as tx_loss
from taxoexpan_b200 import model_zoo as tx_zoo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TAXO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="reference checkout not present")

CONFIGS = {
    "mag": ("PGAT", "WMR", "LBM", dict(in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1], feat_drop=0.1,
                                      attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)),
    "wordnet": ("PGCN", "MR", "MLP", dict(in_dim=300, hidden_dim=600, out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1], feat_drop=0.1,
                                          attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)),
    "gat_cr_bim": ("GAT", "CR", "BIM", dict(in_dim=16, hidden_dim=8, out_dim=8, pos_dim=4, num_layers=2, heads=[2, 2, 1], feat_drop=0.0,
                                            attn_drop=0.0, hidden_drop=0.0, out_drop=0.0)),
}
CLASSES = ["GCNLayer", "GATLayer", "GCN", "GAT", "PGCN", "PGAT", "MeanReadout", "WeightedMeanReadout", "ConcatReadout", "MLP", "BIM", "LBM"]

PROBE = r'''
import inspect, json, sys
import model.model_zoo as zoo, model.model as mm, model.loss as loss
def sig(f):
    return [(p.name, repr(p.default) if p.default is not inspect.Parameter.empty else None) for p in inspect.signature(f).parameters.values()]
out = {"init": {c: sig(getattr(zoo, c).__init__) for c in CLASSES}, "forward": {c: sig(getattr(zoo, c).forward) for c in CLASSES},
       "taxoexpan_init": sig(mm.TaxoExpan.__init__), "taxoexpan_forward": sig(mm.TaxoExpan.forward), "info_nce": sig(loss.info_nce_loss), "state": {}}
for name, (pm, rm, mmeth, kw) in CONFIGS.items():
    m = mm.TaxoExpan(pm, rm, mmeth, **kw)
    out["state"][name] = {k: list(v.shape) for k, v in m.state_dict().items()}
print(json.dumps(out))
'''


def _reference_surface():
    code = f"CLASSES = {CLASSES!r}\nCONFIGS = {CONFIGS!r}\n" + PROBE
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _sig(f):
    return [[p.name, repr(p.default) if p.default is not inspect.Parameter.empty else None] for p in inspect.signature(f).parameters.values()]


def test_module_signatures_and_state_dicts_match_the_reference():
    ref = _reference_surface()
    for c in CLASSES:
        ours = getattr(tx_zoo, c)
        assert _sig(ours.__init__) == ref["init"][c], f"{c}.__init__"
        # forward: same positional parameters (names may carry the reference's typos; position and count are the contract)
        assert len(_sig(ours.forward)) == len(ref["forward"][c]), f"{c}.forward arity"
    assert [p[0] for p in _sig(tx.TaxoExpan.__init__)] == [p[0] for p in ref["taxoexpan_init"]]
    assert [p[0] for p in _sig(tx.TaxoExpan.forward)] == [p[0] for p in ref["taxoexpan_forward"]]
    assert [p[0] for p in _sig(tx_loss.info_nce_loss)] == [p[0] for p in ref["info_nce"]]
    for name, (pm, rm, mmeth, kw) in CONFIGS.items():
        m = tx.TaxoExpan(pm, rm, mmeth, **kw)
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == ref["state"][name], name


DATASET_PROBE = r'''
import json, random, sys
import networkx as nx, numpy as np, torch
import dgl
def to_networkx(self):                       # DGL 0.4.0: nodes 0..n-1, edges in edge-id order
    G = nx.MultiDiGraph()
    G.add_nodes_from(range(self.number_of_nodes()))
    for eid, (u, v) in enumerate(zip(self._src.tolist(), self._dst.tolist())):
        G.add_edge(u, v, id=eid)
    return G
dgl.DGLGraph.to_networkx = to_networkx
from data_loader.dataset import MaskedGraphDataset   # the reference, unmodified
spec = json.loads(sys.stdin.read())
g = dgl.DGLGraph()
g.add_nodes(spec["n"], {"x": torch.arange(spec["n"], dtype=torch.float32)[:, None].repeat(1, 3)})
g.add_edges(spec["par"], spec["chi"])
class GD: pass
gd = GD(); gd.g_full = g; gd.vocab = [str(i) for i in range(spec["n"])]
gd.train_node_ids = spec["train"]; gd.validation_node_ids = []; gd.test_node_ids = []
random.seed(spec["seed"])
ds = MaskedGraphDataset(gd, mode="train", sampling_mode=1, negative_size=spec["neg"], expand_factor=spec["ef"], cache_refresh_time=spec["refresh"])
out = {"node_list": [int(v) for v in ds.node_list], "items": []}
for idx in spec["indices"]:
    item = ds[idx]
    out["items"].append([[t[0].ndata["_id"].tolist(), t[0].ndata["pos"].tolist(), t[0]._src.tolist(), t[0]._dst.tolist(), float(t[1][0]), int(t[2])] for t in item])
print(json.dumps(out))
'''


def test_train_batcher_matches_the_unmodified_reference_dataset():
    """The reference's own MaskedGraphDataset (data_loader/dataset.py, byte-for-byte, through the DGL / gensim shims and real
    networkx) against NegativeSampler + taxonomy_masks + TrainBatcher for the same `random` seed: node lists, positions, edge lists,
    query features and labels of every (query, anchor) egonet over three passes.  expand_factor exceeds every out-degree, so the only
    random numbers consumed are the queue shuffles (sibling sub-sampling is counter-based here by design and checked separately)."""
    import random

    import numpy as np
    import torch

    from taxoexpan_b200 import sampler
    rng = np.random.default_rng(17)
    n, neg, seed = 90, 9, 4242
    par = rng.integers(0, n, 260)
    chi = rng.integers(0, n, 260)
    keep = par < chi                                       # a DAG without self loops
    edges = sorted(set(zip(par[keep].tolist(), chi[keep].tolist())), key=lambda e: (e[0], e[1]))
    par, chi = [e[0] for e in edges], [e[1] for e in edges]
    train = list(range(n))
    indices = list(range(40)) * 3
    spec = dict(n=n, par=par, chi=chi, train=train, seed=seed, neg=neg, ef=1000, refresh=3, indices=indices)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", DATASET_PROBE], input=json.dumps(spec), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])

    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    has_parent = set(chi)
    roots = [v for v in range(n) if v not in has_parent]
    node_list = ref["node_list"]                            # the reference's list(set(...)) order is an implementation detail: take it
    assert sorted(node_list) == sorted(set(train) - set(roots))
    feats = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    masks = sampler.taxonomy_masks(tax, node_list, roots)
    batcher = sampler.TrainBatcher(tax, feats, node_list, sampler.NegativeSampler(train, masks, random.Random(seed)), neg,
                                   expand_factor=1000, cache=sampler.EgonetCache(n, 3, seed=1))
    for idx, item in zip(indices, ref["items"]):
        bg, x, qf, labels = batcher.batch([idx])
        assert labels.tolist() == [t[5] for t in item]
        assert qf[:, 0].tolist() == [t[4] for t in item]
        assert bg.ndata["_id"].tolist() == [v for t in item for v in t[0]]
        assert bg.host_pos().tolist() == [v for t in item for v in t[1]]
        src, dst = bg.edges()
        off = np.concatenate([[0], np.cumsum([len(t[0]) for t in item])])
        assert src.tolist() == [int(off[k]) + v for k, t in enumerate(item) for v in t[2]]
        assert dst.tolist() == [int(off[k]) + v for k, t in enumerate(item) for v in t[3]]


def test_replay_batcher_follows_the_reference_through_sub_sampled_anchors():
    """ReplayTrainBatcher against the unmodified MaskedGraphDataset with expand_factor BELOW most out-degrees: every large anchor
    costs the reference one random.choices from the generator that also shuffles the negative queue (dataset.py:416-424), and the
    per-anchor cache (dataset.py:383-402) decides when a negative is redrawn - node ids, order, labels and the queue walk must still
    agree item for item over three passes (the case TrainBatcher's counter-based draws cannot replay)."""
    import random

    import numpy as np
    import torch

    from taxoexpan_b200 import sampler
    rng = np.random.default_rng(23)
    n, neg, seed, ef, refresh = 70, 7, 99, 2, 2
    par = rng.integers(0, 12, 400)                         # few parents: out-degrees far above expand_factor
    chi = rng.integers(0, n, 400)
    keep = par < chi
    edges = sorted(set(zip(par[keep].tolist(), chi[keep].tolist())), key=lambda e: (e[0], e[1]))
    par, chi = [e[0] for e in edges], [e[1] for e in edges]
    assert np.bincount(par).max() > 4 * ef
    train = list(range(n))
    indices = list(range(30)) * 3
    spec = dict(n=n, par=par, chi=chi, train=train, seed=seed, neg=neg, ef=ef, refresh=refresh, indices=indices)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", DATASET_PROBE], input=json.dumps(spec), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])

    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    has_parent = set(chi)
    roots = [v for v in range(n) if v not in has_parent]
    node_list = ref["node_list"]
    feats = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    masks = sampler.taxonomy_masks(tax, node_list, roots)
    batcher = sampler.ReplayTrainBatcher(tax, feats, node_list, sampler.NegativeSampler(train, masks, random.Random(seed)), neg,
                                         expand_factor=ef, cache_refresh_time=refresh)
    sampled = 0
    for idx, item in zip(indices, ref["items"]):
        bg, x, qf, labels = batcher.batch([idx])
        assert labels.tolist() == [t[5] for t in item]
        assert qf[:, 0].tolist() == [t[4] for t in item]
        assert bg.ndata["_id"].tolist() == [v for t in item for v in t[0]]
        assert bg.host_pos().tolist() == [v for t in item for v in t[1]]
        assert x[:, 0].tolist() == [float(v) for t in item for v in t[0]]
        sampled += sum(1 for t in item if t[1].count(2) == ef)
    assert sampled > 100                                    # the sub-sampled branch was the common case


RAW_PROBE = r'''
import json, sys
import dgl
from data_loader.dataset import MAGDataset            # the reference, unmodified
d = MAGDataset(name="toy", path=sys.argv[1], raw=True)
src, dst = d.g_full.edges()
print(json.dumps({"vocab": d.vocab, "src": src.tolist(), "dst": dst.tolist(), "x": d.g_full.ndata["x"].tolist(),
                  "train": list(map(int, d.train_node_ids)), "validation": list(map(int, d.validation_node_ids)),
                  "test": list(map(int, d.test_node_ids))}))
'''


def test_raw_dataset_loader_matches_the_unmodified_reference(tmp_path):
    """dataset_io.load_raw against the reference's own MAGDataset._load_dataset_raw (data_loader/dataset.py:92-194) on the same
    .terms / .taxo / .embed files: vocabulary, node numbering, edge order, features and the random 10 % / 10 % leaf split."""
    import numpy as np

    from taxoexpan_b200 import dataset_io
    from tests.test_dataset_io_cpu import _write_dataset
    _write_dataset(tmp_path, np.random.default_rng(2))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", RAW_PROBE, str(tmp_path)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])
    ds = dataset_io.load_raw(str(tmp_path), "toy")
    assert ds.vocab == ref["vocab"]
    assert ds.parents.tolist() == ref["src"] and ds.children.tolist() == ref["dst"]
    assert np.array_equal(ds.features, np.asarray(ref["x"], dtype=np.float32))
    assert ds.train_node_ids.tolist() == ref["train"]
    assert ds.validation_node_ids.tolist() == ref["validation"] and ds.test_node_ids.tolist() == ref["test"]


METRIC_PROBE = r'''
import json, sys
import numpy as np
import model.metric as M                                   # the reference, unmodified
spec = json.loads(sys.stdin.read())
out = {"ranks": [[int(r) for r in M.calculate_ranks_from_similarities(np.asarray(s, dtype=np.float32), p)] for s, p in zip(spec["sims"], spec["pos"])]}
R = spec["all_ranks"]
out.update(macro_mr=float(M.macro_mr(R)), micro_mr=float(M.micro_mr(R)), hit1=float(M.hit_at_1(R)), hit3=float(M.hit_at_3(R)),
           hit5=float(M.hit_at_5(R)), mrr=float(M.mrr_scaled_10(R)))
print(json.dumps(out))
'''


def test_rank_and_metric_restatements_match_the_unmodified_reference():
    """oracle.ranks_from_similarities (the checker of the GPU all-pairs ranking) and inference.macro_mr / micro_mr / hit_at_k /
    mrr_scaled_10 against model/metric.py:7-19,62-95, ties included."""
    import numpy as np

    from oracle import taxo_oracle as orc
    from taxoexpan_b200 import inference
    rng = np.random.default_rng(9)
    sims, pos = [], []
    for _ in range(40):
        n = int(rng.integers(3, 60))
        s = np.round(rng.standard_normal(n), 1).astype(np.float32)           # coarse values: many ties
        sims.append(s.tolist())
        pos.append(sorted(rng.choice(n, size=int(rng.integers(1, 4)), replace=False).tolist()))
    all_ranks = [[int(v) for v in rng.integers(1, 200, int(rng.integers(1, 5)))] for _ in range(50)]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", METRIC_PROBE], input=json.dumps(dict(sims=sims, pos=pos, all_ranks=all_ranks)), env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])
    assert [orc.ranks_from_similarities(np.asarray(s, dtype=np.float32), p) for s, p in zip(sims, pos)] == ref["ranks"]
    assert abs(inference.macro_mr(all_ranks) - ref["macro_mr"]) < 1e-9
    assert abs(inference.micro_mr(all_ranks) - ref["micro_mr"]) < 1e-9
    assert abs(inference.hit_at_k(all_ranks, 1) - ref["hit1"]) < 1e-12
    assert abs(inference.hit_at_k(all_ranks, 3) - ref["hit3"]) < 1e-12
    assert abs(inference.hit_at_k(all_ranks, 5) - ref["hit5"]) < 1e-12
    assert abs(inference.mrr_scaled_10(all_ranks) - ref["mrr"]) < 1e-12


CKPT_PROBE = r'''
import sys, json
import torch
import model.model as mm                                  # the reference, unmodified
import parse_config                                       # its ConfigParser is pickled into every checkpoint (base_trainer.py:141)
kw = json.loads(sys.argv[2])
torch.manual_seed(7)
m = mm.TaxoExpan("PGAT", "WMR", "LBM", **kw)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
cfg = object.__new__(parse_config.ConfigParser)           # constructor wants argparse + files; the pickled object is what matters
cfg.__dict__.update({"_ConfigParser__config": {"name": "toy", "arch": {"type": "TaxoExpan", "args": kw}}, "resume": None})
state = {"arch": type(m).__name__, "epoch": 3, "state_dict": m.state_dict(), "optimizer": opt.state_dict(), "monitor_best": 12.5,
         "config": cfg}                                    # base/base_trainer.py:134-142
torch.save(state, sys.argv[1])
print(json.dumps({k: [float(v.double().sum()), float(v.double().abs().max())] for k, v in m.state_dict().items()}))
'''


def test_checkpoint_written_by_the_reference_classes_loads_without_the_reference(tmp_path):
    """A checkpoint saved the way base/base_trainer.py:126-149 does - state_dict of the reference's own TaxoExpan plus its pickled
    parse_config.ConfigParser - loaded here with no reference code on the path, into the drop-in model, tensor for tensor."""
    import torch

    from taxoexpan_b200 import dataset_io
    kw = dict(in_dim=12, hidden_dim=8, out_dim=6, pos_dim=4, num_layers=1, heads=[2, 1], feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1,
              out_drop=0.1)
    path = str(tmp_path / "checkpoint-epoch3.pth")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", CKPT_PROBE, path, json.dumps(kw)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])
    assert "parse_config" not in sys.modules and "model.model" not in sys.modules
    model = tx.TaxoExpan("PGAT", "WMR", "LBM", **kw)
    ckpt = dataset_io.load_reference_checkpoint(path, model)
    assert ckpt["arch"] == "TaxoExpan" and ckpt["epoch"] == 3 and ckpt["monitor_best"] == 12.5
    got = model.state_dict()
    assert set(got) == set(ref)
    for k, (s, mx) in ref.items():
        assert abs(float(got[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)) and float(got[k].double().abs().max()) == mx, k
    assert type(ckpt["config"]).__name__ == "ConfigParser"          # a stub carrying the pickled attributes
    assert ckpt["config"].__dict__["_ConfigParser__config"]["name"] == "toy"
