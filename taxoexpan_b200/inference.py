"""All-pairs inference scoring: the evaluation loops of the reference's `test_fast.py:93-218` /
`infer.py:139-159` - encode every candidate position's egonet once, then score every query against every position and rank the
true positions - with the per-query Python loop (`#queries x #positions` bilinear forms, one `model.match` call per chunk and
query) replaced by two GEMMs on the library's fp32-faithful tensor-core kernels and on-GPU rank / top-k extraction.

    hg      = encode_positions(model, batches)                       # test_fast.py:25-28,156-179 under no_grad
    result  = score_and_rank(model, hg, query_features, positives)   # test_fast.py:187-218 + metric.py:7-60

Ranks follow `model/metric.py:7-19` (similarity mode, used with the InfoNCE loss, `test_fast.py:70-73`): the rank of a true
position is 1 + the number of NON-true positions that score strictly higher.  For the bilinear matchers (`BIM`, `LBM`,
model_zoo.py:301-328) the score matrix is `(hg W) Q^T` (`LBM`'s exp is monotone: ranks are taken on the bilinear form and exp is
applied only to the reported top-k scores); any other matcher falls back to `model.match` on expanded chunks exactly like the
reference.  No host round trip per query: one D2H copy of the ranks and the top-k ids per query chunk.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np
import torch

from . import functional as txf
from .model_zoo import BIM, LBM


@torch.no_grad()
def encode_graph(model, bg, h, pos):
    """reference `test_fast.py:25-28`."""
    bg.ndata['h'] = model.graph_propagate(bg, h)
    return model.readout(bg, pos)


@torch.no_grad()
def encode_positions(model, batches: Iterable) -> torch.Tensor:
    """`batches` yields (batched graph, features [N, in_dim] on the device); returns hg [n_positions, l_dim] on the device
    (`test_fast.py:156-179`; the reference parks the chunks on the CPU, 180 GB of HBM make that unnecessary)."""
    was_training = model.training
    model.eval()
    out = []
    for bg, h in batches:
        pos = bg.ndata['pos'].to(h.device)
        out.append(encode_graph(model, bg, h, pos))
    model.train(was_training)
    return torch.cat(out, 0) if out else torch.zeros(0)


def _bilinear_left(model, hg: torch.Tensor) -> torch.Tensor:
    """U = hg W for the bilinear matchers: score(p, q) = <U_p, q>."""
    w = model.match.W.weight[0]                      # [l, r]
    if txf.GEMM_BACKEND == "f16x3" and hg.shape[0] > 0:
        wp, wt = txf.split_f16_weight(w)
        return txf.gemm_nt_f16(txf.split_f16(hg, w.shape[0]), w.shape[0], wt, w.shape[1]).contiguous()
    return hg @ w


@torch.no_grad()
def score_and_rank(model, hg: torch.Tensor, queries: torch.Tensor, positives: Sequence[Sequence[int]], topk: int = 5,
                   query_chunk: int = 256) -> dict:
    """hg [P, l_dim]: encodings of all candidate positions; queries [Q, r_dim]; positives[j]: indices (into the P positions) of
    query j's true positions.  Returns {'ranks': list (per query) of int64 arrays (one rank per true position, in the order given),
    'topk_idx' [Q, k] int64, 'topk_score' [Q, k] float32}."""
    dev = hg.device
    P, Q = hg.shape[0], queries.shape[0]
    k = min(topk, P)
    bilinear = isinstance(model.match, BIM)
    U = _bilinear_left(model, hg) if bilinear else None
    ranks: List[np.ndarray] = []
    top_idx = torch.empty((Q, k), dtype=torch.int64)
    top_val = torch.empty((Q, k), dtype=torch.float32)
    for q0 in range(0, Q, query_chunk):
        q1 = min(Q, q0 + query_chunk)
        qc = q1 - q0
        qf = queries[q0:q1].to(dev, torch.float32).contiguous()
        if bilinear:                                                    # S[p, j] = <U_p, q_j>: one NT GEMM per chunk
            if txf.GEMM_BACKEND == "f16x3" and P > 0:
                S = txf.gemm_nt_f16(txf.split_f16(U, U.shape[1]), U.shape[1], txf.split_f16(qf, qf.shape[1]), qc)
            else:
                S = U @ qf.t()
        else:                                                           # generic matcher: expanded chunks like test_fast.py:192-198
            S = torch.stack([model.match(hg, qf[j:j + 1].expand(P, -1)).reshape(P) for j in range(qc)], 1)
        # flat (query, true position) pairs of the chunk
        cnt = [len(positives[q0 + j]) for j in range(qc)]
        qi = torch.from_numpy(np.repeat(np.arange(qc), cnt)).to(dev)
        pi = torch.from_numpy(np.concatenate([np.asarray(positives[q0 + j], dtype=np.int64) for j in range(qc)]) if sum(cnt) else
                              np.zeros(0, np.int64)).to(dev)
        s_pos = S[pi, qi]                                               # [K]
        greater = torch.zeros(pi.numel(), dtype=torch.int64, device=dev)
        for p0 in range(0, P, 1 << 16):                                 # positions in slabs: the [slab, K] comparison stays small
            greater += (S[p0:p0 + (1 << 16)][:, qi] > s_pos[None, :]).sum(0)
        # true positions that score higher are not counted (metric.py:15-17 masks them)
        same_q = qi[:, None] == qi[None, :]
        greater -= (same_q & (s_pos[None, :] > s_pos[:, None])).sum(1)
        r = (greater + 1).cpu().numpy()
        off = np.concatenate([[0], np.cumsum(cnt)])
        ranks.extend(r[off[j]:off[j + 1]] for j in range(qc))
        tv, ti = torch.topk(S, k, dim=0)                                # larger similarity preferred (info_nce, test_fast.py:204)
        if isinstance(model.match, LBM):
            tv = torch.exp(tv)
        top_idx[q0:q1] = ti.t().cpu()
        top_val[q0:q1] = tv.t().float().cpu()
    return {"ranks": ranks, "topk_idx": top_idx, "topk_score": top_val}


# ---- metric.py:62-97 on the rank lists (pure numpy; kept here so an evaluation script needs nothing else) ----
def macro_mr(all_ranks):
    return float(np.mean([np.mean(r) for r in all_ranks]))


def micro_mr(all_ranks):
    return float(np.mean(np.concatenate(all_ranks)))


def hit_at_k(all_ranks, k):
    r = np.concatenate(all_ranks)
    return float((r <= k).sum() / len(r))


def mrr_scaled_10(all_ranks):
    r = np.concatenate(all_ranks)
    return float((1.0 / np.ceil(r / 10)).mean())
