"""Seeded synthetic egonet batches with the shape statistics of the reference datasets.

There is no network for the real MAG-CS / MAG-Full / SemEval pickles, so benchmarks and
parity tests draw egonet SHAPES from a model fitted to the statistics printed in the
reference's preprocessing notebooks and lay every egonet out exactly
as `data_loader/dataset.py:404-437` does: nodes [grand-parents..., anchor, siblings...],
edges [gp->anchor..., anchor->sib..., self-loops in node order].

numpy only: the same arrays feed the CUDA path, the CPU oracle and the golden fixtures.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

# name -> (poisson mean of extra grand-parents, leaf probability, power-law exponent, k_max)
SHAPE_MODELS = {
    # MAG-CS: 29,654 nodes / 46,248 edges / 24,508 leaves (data_preprocessing/mag-cs-fos.ipynb)
    "mag-cs": (0.56, 0.8265, 1.8, 2000),
    # MAG-Full: 431,416 nodes / 698,743 edges / 378,044 leaves (data_preprocessing/mag-all-fos.ipynb)
    "mag-full": (0.62, 0.876, 1.8, 2000),
    # WordNet noun / SemEval: tree-like (data_preprocessing/semeval-task14.ipynb)
    "wordnet": (0.02, 0.79, 1.8, 2000),
}


@dataclass
class EgonetShapes:
    """Per-egonet counts. n = n_gp + 1 + n_sib nodes, e = 2n - 1 edges (dataset.py:431-435)."""
    n_gp: np.ndarray       # int64 [G]
    n_sib: np.ndarray      # int64 [G]

    @property
    def num_graphs(self) -> int:
        return int(self.n_gp.shape[0])

    @property
    def num_nodes(self) -> np.ndarray:
        return self.n_gp + 1 + self.n_sib

    @property
    def total_nodes(self) -> int:
        return int(self.num_nodes.sum())

    @property
    def total_edges(self) -> int:
        return int((2 * self.num_nodes - 1).sum())


def _power_law(kmax: int, expo: float) -> np.ndarray:
    k = np.arange(1, kmax + 1, dtype=np.float64)
    p = k ** (-expo)
    return p / p.sum()


def sample_shapes(n_queries: int, negative_size: int = 31, model: str = "mag-cs", expand_factor: int = 50,
                  seed: int = 20200420, positives: bool = True) -> EgonetShapes:
    """Shapes of n_queries * (1 + negative_size) egonets, one positive then `negative_size` negatives per query
    (sampling_mode=1, dataset.py:308-313). With positives=False every egonet is a negative-style (node-uniform)
    anchor -- the inference layout of test_fast.py:93-97 (one egonet per candidate position).
    """
    lam, leaf_p, expo, kmax = SHAPE_MODELS[model]
    rng = np.random.default_rng(seed)
    per = (1 + negative_size) if positives else 1
    g = n_queries * per
    n_gp = 1 + rng.poisson(lam, size=g)
    pk = _power_law(kmax, expo)
    ks = np.arange(1, kmax + 1)
    # node-uniform out-degree law (negatives): 0 w.p. leaf_p else power law
    k_neg = np.where(rng.random(g) < leaf_p, 0, rng.choice(ks, size=g, p=pk))
    n_sib = np.minimum(k_neg, expand_factor)
    if positives:
        # the true parent is reached through an edge: size-biased law ~ k P(k), and the query itself is
        # removed from the sibling list (dataset.py:421-424)
        pb = pk * ks
        pb /= pb.sum()
        k_pos = rng.choice(ks, size=n_queries, p=pb)
        sib_pos = np.where(k_pos <= expand_factor, k_pos - 1, expand_factor)
        n_sib = n_sib.reshape(n_queries, per)
        n_sib[:, 0] = sib_pos
        n_sib = n_sib.reshape(-1)
    return EgonetShapes(n_gp.astype(np.int64), n_sib.astype(np.int64))


def star_batch_arrays(shapes: EgonetShapes) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised `dgl.batch` of star egonets: returns (src, dst, pos, node_offsets[G+1], edge_offsets[G+1]).

    Edge-id order per egonet follows dataset.py:431-435 (gp->anchor, anchor->sib, self-loops), graphs are
    concatenated with node/edge id offsets (data_loaders.py:25).
    """
    n_gp, n_sib = shapes.n_gp, shapes.n_sib
    n = n_gp + 1 + n_sib
    e = 2 * n - 1
    g = n.shape[0]
    noff = np.zeros(g + 1, dtype=np.int64)
    np.cumsum(n, out=noff[1:])
    eoff = np.zeros(g + 1, dtype=np.int64)
    np.cumsum(e, out=eoff[1:])
    N, E = int(noff[-1]), int(eoff[-1])
    gid_n = np.repeat(np.arange(g), n)
    local = np.arange(N) - noff[gid_n]
    anchor_local = n_gp[gid_n]
    pos = np.where(local < anchor_local, 0, np.where(local == anchor_local, 1, 2)).astype(np.int64)
    gid_e = np.repeat(np.arange(g), e)
    le = np.arange(E) - eoff[gid_e]                      # local edge id
    ngp_e, nsib_e, base = n_gp[gid_e], n_sib[gid_e], noff[gid_e]
    is_gp = le < ngp_e
    is_sib = (~is_gp) & (le < ngp_e + nsib_e)
    self_i = le - (ngp_e + nsib_e)
    src_l = np.where(is_gp, le, np.where(is_sib, ngp_e, self_i))
    dst_l = np.where(is_gp, ngp_e, np.where(is_sib, le + 1, self_i))
    return (src_l + base).astype(np.int64), (dst_l + base).astype(np.int64), pos, noff, eoff


def unit_rows(n: int, d: int, seed: int) -> np.ndarray:
    """Row-L2-normalised N(0,1) features (all configs use normalize_embed=true, dataset.py:222-223)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x
