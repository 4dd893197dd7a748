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
Not covered (still "next"): the negative-anchor sampler and the egonet cache of dataset.py:334-402.
"""
from __future__ import annotations

from typing import Optional, Tuple

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


def egonet_node_ids(tax: TaxonomyCSR, anchors, queries, modes, expand_factor: int = 50,
                    generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(ids [N] int64 in batched egonet order, n_gp [G], n_sib [G]) on tax.device - dataset.py:404-426 for G egonets at once."""
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
    draw = torch.rand(total, device=dev, generator=generator)
    pick = torch.where(deg_g <= expand_factor, t_sib, torch.clamp((draw * deg_g).to(torch.int64), max=torch.clamp(deg_g - 1, min=0)))
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


def build_egonet_batch(tax: TaxonomyCSR, features: torch.Tensor, anchors, queries, modes, expand_factor: int = 50,
                       generator: Optional[torch.Generator] = None):
    """(EgonetBatch, x [N, d], ids [N]): the batched graph of `collate_graph_and_node_small_batch` (data_loaders.py:9-28) with
    ndata 'x' / '_id' / 'pos' semantics - x and ids are returned as device tensors, positions come from the closed-form structure."""
    ids, n_gp, n_sib = egonet_node_ids(tax, anchors, queries, modes, expand_factor, generator)
    bg = EgonetBatch.from_counts(n_gp.cpu().numpy().astype(np.int32), n_sib.cpu().numpy().astype(np.int32))
    x = features.index_select(0, ids.to(features.device))
    bg.ndata["_id"] = ids
    return bg, x, ids
