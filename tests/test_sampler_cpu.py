"""Vectorised egonet construction (taxoexpan_b200/sampler.py) against the oracle's restatement of dataset.py:404-426; runs on CPU
tensors (the same tensor program runs on the GPU)."""
import numpy as np
import torch

from oracle import taxo_oracle as orc
from taxoexpan_b200 import sampler


def _random_taxonomy(n, n_edges, rng):
    par = rng.integers(0, n, n_edges)
    chi = rng.integers(0, n, n_edges)
    keep = par != chi
    par, chi = par[keep], chi[keep]
    parents_of, children_of = {}, {}
    for p, c in zip(par.tolist(), chi.tolist()):
        parents_of.setdefault(c, []).append(p)
        children_of.setdefault(p, []).append(c)
    return par, chi, parents_of, children_of


def test_node_order_and_counts_match_dataset_py_when_nothing_is_sampled():
    rng = np.random.default_rng(3)
    n = 400
    par, chi, parents_of, children_of = _random_taxonomy(n, 1500, rng)
    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    G = 300
    anchors = rng.integers(0, n, G)
    modes = rng.integers(0, 2, G)
    queries = np.array([rng.choice(children_of[a]) if (m == 1 and a in children_of) else rng.integers(0, n) for a, m in zip(anchors, modes)])
    ids, n_gp, n_sib = sampler.egonet_node_ids(tax, anchors, queries, modes, expand_factor=10 ** 6)
    want_ids, want_gp, want_sib, want_pos = [], [], [], []
    for a, q, m in zip(anchors.tolist(), queries.tolist(), modes.tolist()):
        nodes, pos = orc.get_subgraph_nodes(parents_of, children_of, q, a, m, expand_factor=10 ** 6)
        want_ids += nodes
        want_pos += pos
        want_gp.append(pos.count(0))
        want_sib.append(pos.count(2))
    assert ids.tolist() == want_ids and n_gp.tolist() == want_gp and n_sib.tolist() == want_sib
    feats = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    bg, x, ids2 = sampler.build_egonet_batch(tax, feats, anchors, queries, modes, expand_factor=10 ** 6)
    assert torch.equal(ids2, ids) and torch.equal(x[:, 0].long(), ids)
    assert bg.host_pos().tolist() == want_pos            # closed-form positions = the reference's nodes_pos
    assert list(bg.batch_num_nodes) == [a + 1 + b for a, b in zip(want_gp, want_sib)]


def test_sampled_children_follow_random_choices_semantics():
    """More children than expand_factor: exactly expand_factor draws with replacement, all of them children; positives lose the
    draws that hit the query (dataset.py:419,424)."""
    n, ef = 200, 5
    par = np.zeros(150, dtype=np.int64)                   # node 0 has 150 children: 1..150
    chi = np.arange(1, 151)
    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    g = torch.Generator().manual_seed(0)
    G = 4000
    ids, n_gp, n_sib = sampler.egonet_node_ids(tax, np.zeros(G, np.int64), np.full(G, 7), np.ones(G, np.int64), expand_factor=ef, generator=g)
    assert int(n_gp.sum()) == 0 and int(n_sib.max()) <= ef and int(n_sib.min()) >= 0
    off = np.concatenate([[0], np.cumsum((n_sib + 1).numpy())])
    ids = ids.numpy()
    hits = 0
    for k in range(G):
        seg = ids[off[k]:off[k + 1]]
        assert seg[0] == 0 and ((seg[1:] >= 1) & (seg[1:] <= 150)).all() and (seg[1:] != 7).all()
        hits += ef - (len(seg) - 1)
    assert abs(hits / (G * ef) - 1.0 / 150) < 3e-3        # each draw hits the query with probability 1 / 150
    neg_ids, _, neg_sib = sampler.egonet_node_ids(tax, np.zeros(8, np.int64), np.full(8, 7), np.zeros(8, np.int64), expand_factor=ef, generator=g)
    assert neg_sib.tolist() == [ef] * 8                   # negatives keep every draw
