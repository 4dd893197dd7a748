"""DGL-free dataset / checkpoint formats (taxoexpan_b200/dataset_io.py) against the reference's own construction order, restated
with networkx exactly as dataset.py:103-150 and :225 (to_networkx of the DGL graph) do it."""
import random

import networkx as nx
import numpy as np
import torch

import taxoexpan_b200 as tx
from taxoexpan_b200 import dataset_io, sampler


def _write_dataset(tmp_path, rng):
    n = 60
    ids = [f"t{1000 + i}" for i in range(n)]
    order = rng.permutation(n)
    with open(tmp_path / "toy.terms", "w") as f:
        for i in order:
            f.write(f"{ids[i]}\tname {i}\n")
    edges = []
    for _ in range(140):
        p, c = rng.integers(0, n, 2)
        if p != c:
            edges.append((ids[p], ids[c]))
    edges += edges[:7]                                   # duplicate lines collapse in a DiGraph
    with open(tmp_path / "toy.taxo", "w") as f:
        for p, c in edges:
            f.write(f"{p}\t{c}\n")
    emb = rng.standard_normal((n, 8)).astype(np.float32)
    with open(tmp_path / "toy.terms.embed", "w") as f:
        f.write(f"{n} 8\n")
        for i in rng.permutation(n):
            f.write(ids[i] + " " + " ".join(f"{v:.6f}" for v in emb[i]) + "\n")
    return ids, order, edges, emb


def test_raw_files_load_in_the_reference_order(tmp_path):
    rng = np.random.default_rng(2)
    ids, order, edges, emb = _write_dataset(tmp_path, rng)
    ds = dataset_io.load_raw(str(tmp_path), "toy")
    # --- the reference's construction (dataset.py:103-150), restated with networkx ---
    taxonomy = nx.DiGraph()
    for i in order:
        taxonomy.add_node(ids[i])
    for p, c in edges:
        taxonomy.add_edge(p, c)
    tx_id2node_id = {node: idx for idx, node in enumerate(taxonomy.nodes())}
    ref_edges = [(tx_id2node_id[p], tx_id2node_id[c]) for p, c in taxonomy.edges()]
    assert ds.tx_ids == list(taxonomy.nodes())
    assert list(zip(ds.parents.tolist(), ds.children.tolist())) == ref_edges
    assert ds.vocab[3] == f"name {order[3]}@@@3"
    for node_id, tx_id in enumerate(ds.tx_ids):
        assert np.allclose(ds.features[node_id], emb[ids.index(tx_id)], atol=1e-6)
    # split: 10 % / 10 % of the leaves with random.seed(47) (dataset.py:167-180)
    leaf = [tx_id2node_id[nd] for nd in taxonomy.nodes() if taxonomy.out_degree(nd) == 0]
    random.seed(47)
    random.shuffle(leaf)
    k = int(len(leaf) * 0.1)
    assert ds.validation_node_ids.tolist() == leaf[:k] and ds.test_node_ids.tolist() == leaf[k:2 * k]
    assert len(ds.train_node_ids) == len(ids) - 2 * k
    # --- egonets: the graph MaskedGraphDataset walks is to_networkx() of the DGL graph built from ref_edges (dataset.py:225) ---
    g = nx.DiGraph()
    g.add_nodes_from(range(len(ids)))
    g.add_edges_from(ref_edges)
    tax = ds.taxonomy()
    for a in range(len(ids)):
        par = tax.par_idx[tax.par_ptr[a]:tax.par_ptr[a + 1]].tolist()
        chi = tax.chi_idx[tax.chi_ptr[a]:tax.chi_ptr[a + 1]].tolist()
        assert par == [e[0] for e in g.in_edges(a)] and chi == [e[1] for e in g.out_edges(a)]
    # round trip through the DGL-free binary
    ds.save(str(tmp_path / "toy.npz"))
    ds2 = dataset_io.TaxonomyDataset.load(str(tmp_path / "toy.npz"))
    assert ds2.vocab == ds.vocab and np.array_equal(ds2.parents, ds.parents) and np.array_equal(ds2.features, ds.features)
    sub = ds.taxonomy(ds.train_node_ids)                  # training graph = subgraph of the train nodes (dataset.py:234)
    held = set(ds.validation_node_ids.tolist()) | set(ds.test_node_ids.tolist())
    assert not (set(sub.chi_idx.tolist()) & held) and not (set(sub.par_idx.tolist()) & held)


def test_reference_checkpoint_layout_loads_into_the_drop_in_model(tmp_path):
    kw = dict(in_dim=12, hidden_dim=8, out_dim=6, pos_dim=4, num_layers=1, heads=[2, 1], feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)
    src = tx.TaxoExpan("PGAT", "WMR", "LBM", **kw)

    class ConfigParser:                                   # stands for parse_config.ConfigParser pickled by base_trainer.py:141
        def __init__(self):
            self.config = {"arch": {"type": "TaxoExpan"}}
    ConfigParser.__module__ = "parse_config"
    ConfigParser.__qualname__ = "ConfigParser"
    import sys
    import types
    mod = types.ModuleType("parse_config")
    mod.ConfigParser = ConfigParser
    sys.modules["parse_config"] = mod
    try:
        state = {"arch": "TaxoExpan", "epoch": 3, "state_dict": {"module." + k: v for k, v in src.state_dict().items()},
                 "optimizer": {}, "monitor_best": 1.0, "config": ConfigParser()}
        torch.save(state, tmp_path / "ckpt.pth")
    finally:
        del sys.modules["parse_config"]                   # the loader must cope without the reference code base
    dst = tx.TaxoExpan("PGAT", "WMR", "LBM", **kw)
    ck = dataset_io.load_reference_checkpoint(str(tmp_path / "ckpt.pth"), dst)
    assert ck["epoch"] == 3 and type(ck["config"]).__name__ == "ConfigParser"
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v)
