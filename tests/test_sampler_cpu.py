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


def test_counter_draws_match_the_oracle_definition():
    rng = np.random.default_rng(11)
    n = 500
    anchor = rng.integers(0, 10 ** 6, n)
    gen = np.where(rng.random(n) < 0.5, rng.integers(0, 50, n), sampler.POSITIVE_GENERATION_BASE + rng.integers(0, 10 ** 6, n))
    slot = rng.integers(0, 50, n)
    deg = rng.integers(1, 5000, n)
    for seed in (0, 12345, (1 << 63) + 7):
        got = sampler.counter_draws(seed, torch.from_numpy(anchor), torch.from_numpy(gen), torch.from_numpy(slot), torch.from_numpy(deg))
        want = [orc.counter_draw(seed, int(a), int(g), int(s), int(d)) for a, g, s, d in zip(anchor, gen, slot, deg)]
        assert got.tolist() == want
        assert int(got.min()) >= 0 and bool((got < torch.from_numpy(deg)).all())


def test_egonet_cache_reproduces_the_reference_cache_semantics():
    """dataset.py:383-402: a negative anchor's egonet is reused until it has been read cache_refresh_time times, positives are
    always rebuilt; repeated anchors inside one batch advance the counter in batch order.  Checked over several batches against
    the sequential dict-cache restatement, node list by node list."""
    rng = np.random.default_rng(5)
    n, ef, refresh = 300, 4, 2
    par, chi, parents_of, children_of = _random_taxonomy(n, 2500, rng)       # ~8 children per node: most anchors exceed ef = 4
    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    cache = sampler.EgonetCache(n, refresh, seed=99)
    ref = orc.EgonetCacheOracle(parents_of, children_of, ef, refresh, seed=99)
    changed = 0
    last = {}
    for step in range(6):
        G = 200
        anchors = rng.integers(0, 40, G)                 # few distinct anchors: every one is hit many times per batch
        modes = (rng.random(G) < 0.2).astype(np.int64)
        queries = np.array([rng.choice(children_of[a]) if (m == 1 and a in children_of) else rng.integers(0, n)
                            for a, m in zip(anchors, modes)])
        feats = torch.arange(n, dtype=torch.float32)[:, None]
        bg, x, ids = sampler.build_egonet_batch(tax, feats, anchors, queries, modes, expand_factor=ef, cache=cache)
        off = np.concatenate([[0], np.cumsum(bg.batch_num_nodes)])
        ids = ids.tolist()
        for k, (a, q, m) in enumerate(zip(anchors.tolist(), queries.tolist(), modes.tolist())):
            want = ref.get(q, a, m)
            assert ids[off[k]:off[k + 1]] == want, (step, k)
            if m == 0:
                changed += int(a in last and last[a] != want)
                last[a] = want
    assert changed > 20                                   # the cache did refresh (the test would be vacuous otherwise)
    assert cache.positives == ref.positives
    assert {a: int(c) for a, c in enumerate(cache.uses.tolist()) if c} == \
        {a: (ref.created[a] - 1) * (refresh + 1) + ref.cache_counter[a] + 1 for a in ref.created}


def test_negative_sampler_matches_the_reference_queue_walk():
    import random
    rng = np.random.default_rng(8)
    n = 120
    par, chi, parents_of, children_of = _random_taxonomy(n, 200, rng)
    keep = par < chi                                      # a DAG
    par, chi = par[keep], chi[keep]
    parents_of, children_of = {}, {}
    for p, c in zip(par.tolist(), chi.tolist()):
        parents_of.setdefault(c, []).append(p)
        children_of.setdefault(p, []).append(c)
    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    roots = [v for v in range(n) if v not in parents_of]
    train = [v for v in range(n) if v % 5 != 0]
    queries = [v for v in train if v not in roots]
    want_masks = orc.node_masks(parents_of, children_of, queries, roots)
    masks = sampler.taxonomy_masks(tax, queries, roots)
    assert {k: sorted(v) for k, v in want_masks.items()} == {k: v.tolist() for k, v in masks.items()}
    got_s = sampler.NegativeSampler(train, masks, random.Random(42))
    ref_s = orc.NegativeQueueOracle(train, want_masks, random.Random(42))
    for rep in range(40):                                 # > one pass over the queue: exercises the reshuffle
        for q in queries[:20]:
            assert got_s.exactly_k(q, 31) == ref_s.exactly_k(q, 31)
        assert got_s.pointer == ref_s.pointer
    for q in queries[:30]:
        assert got_s.at_most_k(q, 7) == ref_s.at_most_k(q, 7)
    # a query whose mask covers every candidate: the reference gives up after 10 tries and pads from the head of the queue
    full = {queries[0]: np.arange(n)}
    a = sampler.NegativeSampler(train, full, random.Random(1)).exactly_k(queries[0], 5)
    b = orc.NegativeQueueOracle(train, {queries[0]: set(range(n))}, random.Random(1)).exactly_k(queries[0], 5)
    assert a == b and len(a) >= 5
    anchors, qs, modes = got_s.batch(queries[:4], [parents_of[q][0] for q in queries[:4]], 31)
    assert anchors.shape == (4 * 32,) and modes.reshape(4, 32)[:, 0].tolist() == [1] * 4 and int(modes.sum()) == 4
    assert qs.reshape(4, 32)[:, 0].tolist() == queries[:4]


def test_train_batcher_reproduces_getitem_plus_collate():
    """dataset.py:290-332 (`__getitem__`, sampling_mode 1) + data_loaders.py:9-28 (collate) for whole batches: positive pointer
    cycling over several parents, negatives from the shared queue, cached negative egonets, labels and query features."""
    import random
    rng = np.random.default_rng(21)
    n, ef, neg, refresh = 150, 3, 7, 2
    par, chi, _, _ = _random_taxonomy(n, 420, rng)
    keep = par < chi
    par, chi = par[keep], chi[keep]
    parents_of, children_of = {}, {}
    for p, c in zip(par.tolist(), chi.tolist()):
        parents_of.setdefault(c, []).append(p)
        children_of.setdefault(p, []).append(c)
    tax = sampler.TaxonomyCSR.from_edges(par, chi, n)
    roots = [v for v in range(n) if v not in parents_of]
    node_list = [v for v in range(n) if v in parents_of]
    feats = torch.from_numpy(rng.standard_normal((n, 5)).astype(np.float32))
    masks = sampler.taxonomy_masks(tax, node_list, roots)
    got = sampler.TrainBatcher(tax, feats, node_list, sampler.NegativeSampler(list(range(n)), masks, random.Random(3)), neg,
                               expand_factor=ef, cache=sampler.EgonetCache(n, refresh, seed=5))
    ref = orc.TrainItemOracle(node_list, parents_of, orc.NegativeQueueOracle(list(range(n)), orc.node_masks(parents_of, children_of, node_list, roots),
                                                                           random.Random(3)),
                              orc.EgonetCacheOracle(parents_of, children_of, ef, refresh, seed=5), neg)
    order = list(range(len(node_list)))
    for epoch in range(3):                                    # pointers cycle through multi-parent nodes
        random.Random(epoch).shuffle(order)
        for b0 in range(0, len(order) - 15, 16):
            idx = order[b0:b0 + 16]
            bg, x, qf, labels = got.batch(idx)
            items = [it for i in idx for it in ref[i]]
            want_ids = [v for nodes, _, _ in items for v in nodes]
            assert bg.ndata["_id"].tolist() == want_ids
            assert list(bg.batch_num_nodes) == [len(nodes) for nodes, _, _ in items]
            assert labels.tolist() == [lab for _, _, lab in items]
            assert torch.equal(qf, feats[[q for _, q, _ in items]])
            assert torch.equal(x, feats[want_ids])
            assert labels.reshape(16, 1 + neg)[:, 0].tolist() == [1] * 16 and int(labels.sum()) == 16


def test_chunk_by_node_limit_matches_the_test_stage_collate():
    rng = np.random.default_rng(2)
    cases = [rng.integers(1, 58, 5000).tolist(), [150000, 3, 4], [5, 150000, 3, 200000, 200000, 1], [100000], [100000, 1], [1] * 10, [],
             [99999, 1, 1], [50000, 50000, 1, 99999, 2]]
    for nodes in cases:
        for limit in (100000, 1000, 57):
            assert sampler.chunk_by_node_limit(nodes, limit) == orc.large_batch_chunks(nodes, limit), (nodes[:8], limit)


def test_negative_queue_diverges_from_the_reference_after_a_large_anchor():
    """Documents a deliberate difference: the reference draws the sub-sampled siblings of an anchor with more than
    `expand_factor` children with random.choices from the SAME module-level generator that shuffles the negative queue
    (dataset.py:416-424 interleaved with :340-381), so every such draw moves its queue order on; the sampler's sibling draws are
    counter-based and leave the generator alone.  Identical queue walks are therefore guaranteed only while no anchor is sub-sampled."""
    import random
    train = list(range(40))
    masks = {0: np.array([0])}
    sets = {0: {0}}
    a = sampler.NegativeSampler(train, masks, random.Random(7))
    b = orc.NegativeQueueOracle(train, sets, random.Random(7))
    for _ in range(3):
        assert a.exactly_k(0, 31) == b.exactly_k(0, 31)              # no sub-sampled anchor so far: same walk
    b.rng.choices(range(100), k=50)                                   # the reference samples a large anchor's siblings here
    diverged = False
    for _ in range(12):                                               # crosses the next reshuffle of the 200-entry queue
        diverged |= a.exactly_k(0, 31) != b.exactly_k(0, 31)
    assert diverged
