"""Vectorised egonet batch construction: the per-egonet Python / networkx / DGLGraph loop of
`data_loader/dataset.py:404-437` (`_get_subgraph`) + `data_loaders.py:9-28` (`dgl.batch`) restated as a handful of tensor
operations over a CSR taxonomy, on whatever device the taxonomy lives on (the GPU in production; the same code runs on CPU
tensors, which is how the parity tests exercise it without a GPU).

    tax = TaxonomyCSR.from_edges(parents, children, num_nodes).to("cuda")
    bg, x, ids = build_egonet_batch(tax, features, anchors, queries, modes, expand_factor=50)
    scores = model(bg, x, features[queries])

Node order per egonet is the reference's: [parents of the anchor (pos 0, in in-edge order), anchor (pos 1), children of the
anchor (pos 2, in out-edge order)]; a positive instance (mode 1) drops the query node from the children; an anchor with more
than `expand_factor` children gets `expand_factor` children drawn WITH replacement (`random.choices`, dataset.py:419,424 - so
duplicates can occur, and for positives the draws that hit the query are dropped afterwards).  The edge list itself is never
materialised: `EgonetBatch.from_counts` + `tx_star_batch_structure` produce positions and both CSRs in closed form.
The negative-anchor sampler (`NegativeSampler`, dataset.py:248-258,285-287,334-381) and the egonet cache (`EgonetCache`,
dataset.py:383-402) are here too.  The cache does not store subgraphs: sibling draws are counter-based (a pure function of
(seed, anchor, generation, slot)), so "reuse the cached egonet of this anchor until it has been read cache_refresh_time times" is
one integer per node - the number of negative uses so far - and generation = uses // (cache_refresh_time + 1).
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .graph import EgonetBatch


class TaxonomyCSR:
    """Parents (in-edges) and children (out-edges) of every node, each list in edge-insertion order like networkx / DGL."""

    def __init__(self, par_ptr, par_idx, chi_ptr, chi_idx):
        self.par_ptr, self.par_idx, self.chi_ptr, self.chi_idx = par_ptr, par_idx, chi_ptr, chi_idx

    @classmethod
    def from_edges(cls, parents, children, num_nodes: int) -> "TaxonomyCSR":
        p = torch.as_tensor(parents, dtype=torch.int64).reshape(-1)
        c = torch.as_tensor(children, dtype=torch.int64).reshape(-1)

        def csr(key, val):
            order = torch.sort(key, stable=True).indices                 # stable: lists keep the edge order
            ptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
            ptr[1:] = torch.cumsum(torch.bincount(key, minlength=num_nodes), 0)
            return ptr, val[order]

        par_ptr, par_idx = csr(c, p)     # in-edges of a node: its parents
        chi_ptr, chi_idx = csr(p, c)     # out-edges: its children
        return cls(par_ptr, par_idx, chi_ptr, chi_idx)

    def to(self, device) -> "TaxonomyCSR":
        return TaxonomyCSR(*(t.to(device) for t in (self.par_ptr, self.par_idx, self.chi_ptr, self.chi_idx)))

    @property
    def device(self):
        return self.par_ptr.device


POSITIVE_GENERATION_BASE = 1 << 40


def _i64(c: int) -> int:
    """a 64-bit constant as the signed value torch.int64 holds for the same bit pattern"""
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """logical right shift of int64 bit patterns"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def counter_draws(seed: int, anchor: torch.Tensor, generation: torch.Tensor, slot: torch.Tensor, degree: torch.Tensor) -> torch.Tensor:
    """floor(u * degree) with u a 53-bit uniform from a splitmix64 finaliser of (seed, anchor, generation, slot): the counter-based
    replacement of `random.choices(out_edges, k=expand_factor)` (dataset.py:419,424); bit-for-bit `oracle.counter_draw`.  int64
    tensors wrap modulo 2^64 like the unsigned arithmetic of the definition."""
    z = (anchor * _i64(0x9E3779B97F4A7C15)) ^ (generation * _i64(0xC2B2AE3D27D4EB4F)) ^ (slot * _i64(0x165667B19E3779F9)) ^ _i64(seed & ((1 << 64) - 1))
    z = z + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    pick = (_lsr(z, 11).to(torch.float64) * (degree.to(torch.float64) / 2.0 ** 53)).to(torch.int64)
    return torch.minimum(pick, torch.clamp(degree - 1, min=0))


def egonet_node_ids(tax: TaxonomyCSR, anchors, queries, modes, expand_factor: int = 50,
                    generator: Optional[torch.Generator] = None, draw_seed: Optional[int] = None,
                    generation: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(ids [N] int64 in batched egonet order, n_gp [G], n_sib [G]) on tax.device - dataset.py:404-426 for G egonets at once.
    Sibling draws of anchors with more than expand_factor children: torch.rand (`generator`) by default; with `draw_seed` and a
    per-egonet `generation` [G] they are `counter_draws(draw_seed, anchor, generation, slot, degree)` (see EgonetCache)."""
    dev = tax.device
    a = torch.as_tensor(anchors, dtype=torch.int64, device=dev).reshape(-1)
    q = torch.as_tensor(queries, dtype=torch.int64, device=dev).reshape(-1)
    m = torch.as_tensor(modes, dtype=torch.int64, device=dev).reshape(-1)
    G = a.numel()
    n_gp = tax.par_ptr[a + 1] - tax.par_ptr[a]
    deg = tax.chi_ptr[a + 1] - tax.chi_ptr[a]
    n_sib_raw = torch.clamp(deg, max=expand_factor)
    cnt = n_gp + 1 + n_sib_raw
    off = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(cnt, 0)
    total = int(off[-1])                                               # the one host sync of the construction
    gid = torch.repeat_interleave(torch.arange(G, device=dev), cnt, output_size=total)
    local = torch.arange(total, device=dev) - off[gid]
    a_g, ngp_g, deg_g = a[gid], n_gp[gid], deg[gid]
    is_gp = local < ngp_g
    is_sib = local > ngp_g
    t_sib = local - ngp_g - 1
    # children: all of them in out-edge order, or expand_factor uniform draws with replacement
    if draw_seed is not None:
        gen = torch.as_tensor(generation, dtype=torch.int64, device=dev).reshape(-1)
        drawn = counter_draws(int(draw_seed), a_g, gen[gid], torch.clamp(t_sib, min=0), deg_g)
    else:
        draw = torch.rand(total, device=dev, generator=generator)
        drawn = torch.clamp((draw * deg_g).to(torch.int64), max=torch.clamp(deg_g - 1, min=0))
    pick = torch.where(deg_g <= expand_factor, t_sib, drawn)
    gp_src = tax.par_ptr[a_g] + local
    sib_src = tax.chi_ptr[a_g] + pick
    n_par, n_chi = tax.par_idx.numel(), tax.chi_idx.numel()
    ids = torch.where(is_gp, tax.par_idx[torch.clamp(gp_src, 0, max(n_par - 1, 0))] if n_par else a_g,
                      torch.where(is_sib, tax.chi_idx[torch.clamp(sib_src, 0, max(n_chi - 1, 0))] if n_chi else a_g, a_g))
    # positives: the query itself is not a sibling (dataset.py:422,424)
    drop = is_sib & (m[gid] == 1) & (ids == q[gid])
    keep = ~drop
    n_sib = n_sib_raw - torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, gid, drop.to(torch.int64))
    return ids[keep], n_gp, n_sib


class EgonetCache:
    """dataset.py:383-402 without stored subgraphs.  The reference keeps, per negative anchor, the last egonet it built and hands
    it out again until it has been read `cache_refresh_time` times (positives are rebuilt every time and never cached).  With
    counter-based draws the egonet of an anchor is a pure function of its generation number, so the whole cache is `uses[node]` =
    how often the node has served as a negative anchor: use number u belongs to generation u // (cache_refresh_time + 1).
    `generations()` numbers the egonets of one batch exactly as the reference's sequential loop would (duplicates of an anchor
    inside a batch advance its counter one by one, in batch order)."""

    def __init__(self, num_nodes: int, cache_refresh_time: int, seed: int = 0, device="cpu"):
        self.uses = torch.zeros(num_nodes, dtype=torch.int64, device=device)
        self.period = int(cache_refresh_time) + 1
        self.seed = int(seed)
        self.positives = 0

    def generations(self, anchors, modes) -> torch.Tensor:
        dev = self.uses.device
        a = torch.as_tensor(anchors, dtype=torch.int64, device=dev).reshape(-1)
        m = torch.as_tensor(modes, dtype=torch.int64, device=dev).reshape(-1)
        gen = torch.zeros_like(a)
        pos_idx = torch.nonzero(m == 1).reshape(-1)
        gen[pos_idx] = POSITIVE_GENERATION_BASE + self.positives + torch.arange(pos_idx.numel(), device=dev)
        self.positives += int(pos_idx.numel())
        neg_idx = torch.nonzero(m != 1).reshape(-1)
        if neg_idx.numel():
            an = a[neg_idx]
            order = torch.sort(an, stable=True).indices              # occurrences of one anchor stay in batch order
            sa = an[order]
            first = torch.ones_like(sa, dtype=torch.bool)
            first[1:] = sa[1:] != sa[:-1]
            start = torch.cummax(torch.where(first, torch.arange(sa.numel(), device=dev), torch.zeros_like(sa)), 0).values
            rank = torch.arange(sa.numel(), device=dev) - start      # 0, 1, 2, ... within each anchor
            g_sorted = (self.uses[sa] + rank) // self.period
            gn = torch.empty_like(g_sorted)
            gn[order] = g_sorted
            gen[neg_idx] = gn
            self.uses.index_add_(0, an, torch.ones_like(an))
        return gen


def taxonomy_masks(tax: TaxonomyCSR, nodes: Sequence[int], roots: Sequence[int]) -> Dict[int, np.ndarray]:
    """node2masks of dataset.py:248-258 as sorted arrays: descendants(n) + parents(n) + [n] + roots for every n in `nodes`
    (frontier expansion over the children CSR; the graph is a DAG but cycles are tolerated)."""
    chi_ptr, chi_idx = tax.chi_ptr.cpu().numpy(), tax.chi_idx.cpu().numpy()
    par_ptr, par_idx = tax.par_ptr.cpu().numpy(), tax.par_idx.cpu().numpy()
    roots = np.asarray(list(roots), dtype=np.int64)
    out = {}
    for n in nodes:
        seen = np.zeros(0, dtype=np.int64)
        frontier = chi_idx[chi_ptr[n]:chi_ptr[n + 1]]
        while frontier.size:
            frontier = np.setdiff1d(np.unique(frontier), seen, assume_unique=True)
            frontier = frontier[frontier != n]
            if not frontier.size:
                break
            seen = np.union1d(seen, frontier)
            cnt = chi_ptr[frontier + 1] - chi_ptr[frontier]
            if not cnt.sum():
                break
            base = np.repeat(chi_ptr[frontier], cnt)
            within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            frontier = chi_idx[base + within]
        out[int(n)] = np.unique(np.concatenate([seen, par_idx[par_ptr[n]:par_ptr[n + 1]], [n], roots]).astype(np.int64))
    return out


class NegativeSampler:
    """dataset.py:285-287,334-381: negative anchors come from a queue (train ids x 5) walked by a pointer shared by all queries and
    reshuffled with `random.shuffle` whenever it is exhausted; positions in the query's mask are skipped.  The queue walk gives the same
    sequence as the reference for the same `random.Random` state AS LONG AS no anchor of the run has more than `expand_factor`
    children: the reference draws the sub-sampled siblings of such an anchor with `random.choices` from the SAME module-level generator
    (`_get_subgraph`, dataset.py:416-424), interleaved with these shuffles, while the sibling draws here are counter-based
    (`counter_draws`) and leave the generator alone - after the first large anchor the reference's queue order moves on, this one's does
    not (tests/test_sampler_cpu.py::test_negative_queue_diverges_from_the_reference_after_a_large_anchor documents it;
    `ReplayTrainBatcher` below is the mode that follows the reference through such anchors too).  The membership test runs on sorted
    mask arrays instead of Python sets."""

    def __init__(self, train_node_ids: Sequence[int], node2masks: Dict[int, np.ndarray], rng: Optional[random.Random] = None):
        self.queue = list(train_node_ids) * 5
        self.pointer = 0
        self.node2masks = node2masks
        self.rng = rng if rng is not None else random
        self._arr = None                      # numpy view of the queue, rebuilt after every shuffle

    def _shuffle(self):
        self.rng.shuffle(self.queue)
        self._arr = None

    def _window(self, lo: int, n: int, mask: np.ndarray) -> list:
        if self._arr is None:
            self._arr = np.asarray(self.queue, dtype=np.int64)
        w = self._arr[lo:lo + n]
        if not mask.size or not w.size:
            return w.tolist()
        at = np.searchsorted(mask, w)
        hit = mask[np.minimum(at, mask.size - 1)] == w
        return w[~hit].tolist()

    def at_most_k(self, query_node: int, negative_size: int) -> list:
        """dataset.py:340-356 (sampling_mode 0)"""
        if self.pointer == 0:
            self._shuffle()
        mask = self.node2masks[query_node]
        while True:
            negatives = self._window(self.pointer, negative_size, mask)
            if len(negatives) > 0:
                break
        self.pointer += negative_size
        if self.pointer >= len(self.queue):
            self.pointer = 0
        return negatives

    def exactly_k(self, query_node: int, negative_size: int) -> list:
        """dataset.py:358-381 (sampling_mode 1: the InfoNCE layout of one positive followed by exactly negative_size negatives)"""
        if self.pointer == 0:
            self._shuffle()
        mask = self.node2masks[query_node]
        negatives = []
        max_try = 0
        while len(negatives) != negative_size:
            n_lack = negative_size - len(negatives)
            negatives.extend(self._window(self.pointer, n_lack, mask))
            self.pointer += n_lack
            if self.pointer >= len(self.queue):
                self.pointer = 0
                self._shuffle()
            max_try += 1
            if max_try > 10:                  # corner case of the reference: trim / pad from the head of the queue, mask ignored
                if len(negatives) > negative_size:
                    negatives = negatives[:negative_size]
                else:
                    negatives.extend(self.queue[:negative_size - len(negatives)])
        return negatives

    def batch(self, query_nodes: Sequence[int], positive_parents: Sequence[int], negative_size: int):
        """(anchors, queries, modes) of a training batch in the reference's order (dataset.py:308-332 + data_loaders.py:9-28):
        per query its positive parent first, then exactly `negative_size` negatives."""
        anchors, queries, modes = [], [], []
        for q, p in zip(query_nodes, positive_parents):
            neg = self.exactly_k(int(q), negative_size)
            anchors += [int(p)] + neg
            queries += [int(q)] * (1 + len(neg))
            modes += [1] + [0] * len(neg)
        return np.asarray(anchors, np.int64), np.asarray(queries, np.int64), np.asarray(modes, np.int64)


def build_egonet_batch(tax: TaxonomyCSR, features: torch.Tensor, anchors, queries, modes, expand_factor: int = 50,
                       generator: Optional[torch.Generator] = None, cache: Optional[EgonetCache] = None):
    """(EgonetBatch, x [N, d], ids [N]): the batched graph of `collate_graph_and_node_small_batch` (data_loaders.py:9-28) with
    ndata 'x' / '_id' / 'pos' semantics - x and ids are returned as device tensors, positions come from the closed-form structure."""
    if cache is not None:           # dataset.py:383-402: negatives reuse their anchor's egonet until it has been read cache_refresh_time times
        ids, n_gp, n_sib = egonet_node_ids(tax, anchors, queries, modes, expand_factor, draw_seed=cache.seed,
                                           generation=cache.generations(anchors, modes).to(tax.device))
    else:
        ids, n_gp, n_sib = egonet_node_ids(tax, anchors, queries, modes, expand_factor, generator)
    bg = EgonetBatch.from_counts(n_gp.cpu().numpy().astype(np.int32), n_sib.cpu().numpy().astype(np.int32))
    x = features.index_select(0, ids.to(features.device))
    bg.ndata["_id"] = ids
    return bg, x, ids


class TrainBatcher:
    """One training batch = `collate_graph_and_node_small_batch` (data_loaders.py:9-28) of `MaskedGraphDataset.__getitem__`
    (dataset.py:290-332, sampling_mode 1) for a list of dataset indices: per query node its next true parent (the cyclic positive
    pointer of dataset.py:316-321), then exactly `negative_size` negative anchors from the shared queue, every (anchor, query) pair
    turned into an egonet - all egonets of the batch built at once on the taxonomy's device instead of one DGLGraph at a time in 20
    worker processes.  Returns what the trainer consumes (trainer.py:44-51): (batched graph, node features x, query features, labels)."""

    def __init__(self, tax: TaxonomyCSR, features: torch.Tensor, node_list: Sequence[int], negatives: NegativeSampler,
                 negative_size: int, expand_factor: int = 50, cache: Optional[EgonetCache] = None):
        self.tax, self.features = tax, features
        self.node_list = [int(v) for v in node_list]
        self.negatives, self.negative_size, self.expand_factor, self.cache = negatives, int(negative_size), int(expand_factor), cache
        self._par_ptr = tax.par_ptr.cpu().numpy()
        self._par_idx = tax.par_idx.cpu().numpy()
        self.positive_pointer: Dict[int, int] = {}          # node2positive_pointer (dataset.py:252,316-321)

    def __len__(self):
        return len(self.node_list)

    def _next_parent(self, q: int) -> int:
        lo, hi = int(self._par_ptr[q]), int(self._par_ptr[q + 1])
        if hi == lo:
            raise ValueError(f"query node {q} has no parent in the training graph (roots are excluded, dataset.py:241-244)")
        ptr = self.positive_pointer.get(q, 0)
        self.positive_pointer[q] = (ptr + 1) % (hi - lo)
        return int(self._par_idx[lo + ptr])

    def batch(self, indices: Sequence[int]):
        queries = [self.node_list[int(i)] for i in indices]
        parents = [self._next_parent(q) for q in queries]
        anchors, qs, modes = self.negatives.batch(queries, parents, self.negative_size)
        bg, x, _ = build_egonet_batch(self.tax, self.features, anchors, qs, modes, self.expand_factor, cache=self.cache)
        dev = self.features.device
        qf = self.features.index_select(0, torch.as_tensor(qs, dtype=torch.int64, device=dev))
        labels = torch.as_tensor(modes, dtype=torch.int64, device=dev)
        return bg, x, qf, labels


class ReplayTrainBatcher(TrainBatcher):
    """TrainBatcher that replays the reference's single-process loader (num_workers=0) DRAW FOR DRAW, sub-sampled anchors included.
    The reference takes the siblings of an anchor with more than `expand_factor` children with `random.choices` from the same
    generator that shuffles the negative queue (dataset.py:416-424), inside the per-item loop of `__getitem__` (dataset.py:290-332):
    positive egonet first, then the queue walk, then the negatives' egonets, each negative read from the per-anchor cache until it has
    been served `cache_refresh_time` times (dataset.py:383-402).  This batcher keeps exactly that order on the host - one
    `rng.choices` per large anchor at the position the reference makes it, the same dict cache and counters - and hands the
    resulting node-id lists to the device as one batch.  It trades the vectorised device construction of `TrainBatcher` for the
    ability to continue a reference run bit for bit; `rng` must be the generator the NegativeSampler uses."""

    def __init__(self, tax: TaxonomyCSR, features: torch.Tensor, node_list: Sequence[int], negatives: NegativeSampler,
                 negative_size: int, expand_factor: int = 50, cache_refresh_time: int = 64):
        super().__init__(tax, features, node_list, negatives, negative_size, expand_factor, cache=None)
        self._chi_ptr = tax.chi_ptr.cpu().numpy()
        self._chi_idx = tax.chi_idx.cpu().numpy()
        self.rng = negatives.rng
        self.cache_refresh_time = int(cache_refresh_time)
        self._cached: Dict[int, tuple] = {}                 # anchor -> (grand-parents, siblings) of its last negative egonet
        self._served: Dict[int, int] = {}

    def _egonet(self, query: int, anchor: int, positive: bool):
        """dataset.py:404-427 without the DGL object: (grand-parent ids, sibling ids)"""
        gps = self._par_idx[self._par_ptr[anchor]:self._par_ptr[anchor + 1]].tolist()
        children = self._chi_idx[self._chi_ptr[anchor]:self._chi_ptr[anchor + 1]].tolist()
        if len(children) > self.expand_factor:
            children = self.rng.choices(children, k=self.expand_factor)
        if positive:
            children = [c for c in children if c != query]
        return gps, children

    def _negative_egonet(self, query: int, anchor: int):
        if anchor in self._cached and self._served[anchor] < self.cache_refresh_time:
            self._served[anchor] += 1
            return self._cached[anchor]
        ego = self._egonet(query, anchor, False)
        self._cached[anchor] = ego
        self._served[anchor] = 0
        return ego

    def batch(self, indices: Sequence[int]):
        ids, n_gp, n_sib, qs, modes = [], [], [], [], []

        def push(q, a, ego, mode):
            gps, sibs = ego
            ids.extend(gps); ids.append(a); ids.extend(sibs)
            n_gp.append(len(gps)); n_sib.append(len(sibs)); qs.append(q); modes.append(mode)

        for i in indices:
            q = self.node_list[int(i)]
            p = self._next_parent(q)
            push(q, p, self._egonet(q, p, True), 1)
            for a in self.negatives.exactly_k(q, self.negative_size):
                push(q, int(a), self._negative_egonet(q, int(a)), 0)
        dev = self.features.device
        bg = EgonetBatch.from_counts(np.asarray(n_gp, np.int32), np.asarray(n_sib, np.int32))
        idt = torch.as_tensor(ids, dtype=torch.int64, device=dev)
        bg.ndata["_id"] = idt
        x = self.features.index_select(0, idt)
        qf = self.features.index_select(0, torch.as_tensor(qs, dtype=torch.int64, device=dev))
        return bg, x, qf, torch.as_tensor(modes, dtype=torch.int64, device=dev)


BATCH_GRAPH_NODE_LIMIT = 100000      # data_loaders.py:7


def chunk_by_node_limit(nodes_per_egonet: Sequence[int], limit: int = BATCH_GRAPH_NODE_LIMIT):
    """Chunk boundaries of `collate_graph_and_node_large_batch` (data_loaders.py:31-72, the test-stage collate): egonets are appended
    to the current chunk and the chunk is closed right AFTER the egonet that pushes its node count above `limit` (if it holds more
    than one egonet).  Returns [(first, last + 1), ...] over the egonet list; found with one searchsorted per chunk instead of a
    Python loop over all egonets."""
    n = np.asarray(nodes_per_egonet, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(n)])
    out, lo, total = [], 0, n.shape[0]
    while lo < total:
        # first egonet index hi >= lo with csum[hi + 1] - csum[lo] > limit
        hi = int(np.searchsorted(csum, csum[lo] + limit, side="right")) - 1
        if hi >= total:
            out.append((lo, total))
            break
        if hi == lo:                      # a single egonet above the limit never closes a chunk on its own (len(gs) > 1 rule)
            hi = lo + 1
            if hi >= total:
                out.append((lo, total))
                break
        out.append((lo, hi + 1))
        lo = hi + 1
    return out
