"""CPU oracle: a plain-torch restatement of TaxoExpan's propagation + readout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `taxoexpan_b200/` imports this file; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and only as the checker / the CPU arm -- never as the product path.

What is restated (reference file:line, mickeysjm/TaxoExpan @ 3e38336):
  egonet node/edge layout          data_loader/dataset.py:404-437
  batching (disjoint union)        data_loader/data_loaders.py:24-26
  GCNLayer / GCN / PGCN            model/model_zoo.py:13-50, 116-137, 139-167
  GATLayer / GAT / PGAT            model/model_zoo.py:52-114, 169-190, 192-220
  MeanReadout / WMR / ConcatReadout model/model_zoo.py:227-258
  MLP / BIM / LBM matching         model/model_zoo.py:281-328
  TaxoExpan.forward glue           model/model.py:70-87
  info_nce_loss + step reshape     model/loss.py:52-57, trainer/trainer.py:52-56

Parity status: PINNED against the reference's own unmodified model/model.py +
model/model_zoo.py executed in the build container through the DGL-semantics shim
in oracle/dgl_shim (see oracle/make_golden.py, tests/golden/*.npz and
tests/test_oracle_golden.py).  DGL 0.4.0 itself (un-vendored third-party dependency,
reference README.md:7-12) is not installable offline, so the semantics of
update_all / edge_softmax / mean_nodes / batch are restated from DGL's published
behaviour: for that layer parity is UNPINNED (the reference ships no tests or golden
vectors).

All functions are dtype-generic (fp32 oracle, fp64 to measure the oracle's own noise)
and differentiable, so torch autograd provides the gradient oracle.  Dropout takes
EXPLICIT keep-masks so a CUDA run with a known mask can be checked exactly:
drop(x) = x * keep / (1 - p)  (torch.nn.Dropout semantics, model_zoo.py:57-64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import os

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# Graph layout (integer work; bit-exact)
# --------------------------------------------------------------------------------------
@dataclass
class OracleGraph:
    """Batched egonet graph: what `dgl.batch` of `_get_subgraph` outputs would hold."""
    n: int
    src: torch.Tensor                 # int64 [E], edge-id order
    dst: torch.Tensor                 # int64 [E]
    pos: torch.Tensor                 # int64 [N] in {0 grand-parent, 1 anchor, 2 sibling}
    batch_num_nodes: List[int]
    batch_num_edges: List[int] = field(default_factory=list)

    @property
    def num_graphs(self) -> int:
        return len(self.batch_num_nodes)

    def graph_ids(self) -> torch.Tensor:
        return torch.repeat_interleave(torch.arange(self.num_graphs), torch.tensor(self.batch_num_nodes, dtype=torch.int64))

    def in_degrees(self) -> torch.Tensor:
        return torch.bincount(self.dst, minlength=self.n)


def star_egonet(n_gp: int, n_sib: int):
    """Edges / positions of ONE egonet exactly as dataset.py:404-437 builds it.

    nodes : [grand-parents (pos 0) x n_gp, anchor (pos 1), siblings (pos 2) x n_sib]   :406-426
    edges : gp_k -> anchor (k = 0..n_gp-1)                                            :431
            anchor -> sib_k                                                          :432
            i -> i for every node i in order                                         :435
    """
    n = n_gp + 1 + n_sib
    a = n_gp
    src = list(range(n_gp)) + [a] * n_sib + list(range(n))
    dst = [a] * n_gp + list(range(a + 1, n)) + list(range(n))
    pos = [0] * n_gp + [1] + [2] * n_sib
    return (np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), np.asarray(pos, dtype=np.int64), n)


def batch_star_egonets(n_gp: Sequence[int], n_sib: Sequence[int]) -> OracleGraph:
    """`dgl.batch` (data_loaders.py:25) of star egonets: ids of graph k shifted by the totals before it."""
    srcs, dsts, poss, nn, ne, off = [], [], [], [], [], 0
    for g, s in zip(n_gp, n_sib):
        u, v, p, n = star_egonet(int(g), int(s))
        srcs.append(u + off)
        dsts.append(v + off)
        poss.append(p)
        nn.append(n)
        ne.append(len(u))
        off += n
    cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64))
    return OracleGraph(off, torch.from_numpy(cat(srcs)), torch.from_numpy(cat(dsts)), torch.from_numpy(cat(poss)), nn, ne)


def csr_by_dst(n: int, src: np.ndarray, dst: np.ndarray):
    """Destination-sorted CSR with in-edges kept in edge-id order (stable). Returns indptr, src_sorted, eid."""
    eid = np.argsort(dst, kind="stable")
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
    return indptr, src[eid], eid


# --------------------------------------------------------------------------------------
# Message passing primitives (DGL 0.4.0 semantics, restated)
# --------------------------------------------------------------------------------------
def _drop(x, keep, p):
    if keep is None or not p:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


def _sum_in(g: OracleGraph, msg: torch.Tensor) -> torch.Tensor:
    """update_all(..., fn.sum): out[i] = sum over in-edges e:(j->i) of msg[e]; 0 with no in-edge."""
    out = torch.zeros((g.n, *msg.shape[1:]), dtype=msg.dtype)
    return out.index_add(0, g.dst, msg)


def edge_softmax(g: OracleGraph, logits: torch.Tensor) -> torch.Tensor:
    """dgl.nn.pytorch.edge_softmax (model_zoo.py:112): per destination, per trailing index."""
    tail = logits.shape[1:]
    idx = g.dst.view(-1, *([1] * len(tail))).expand_as(logits)
    mx = torch.full((g.n, *tail), float("-inf"), dtype=logits.dtype)
    mx = mx.scatter_reduce(0, idx, logits.detach(), reduce="amax", include_self=True)
    e = torch.exp(logits - mx[g.dst])
    return e / _sum_in(g, e)[g.dst]


def gcn_norm(g: OracleGraph, dtype) -> torch.Tensor:
    """model_zoo.py:130-134 / 157-161: in-degree ** -0.5, inf -> 0, shape [N,1].

    The reference computes it as `g.in_degrees().float()` -> always fp32, whatever the model dtype; an fp64
    model promotes the fp32 value.  Restated literally so the fp64 run matches the reference's fp64 run.
    """
    norm = torch.pow(g.in_degrees().float(), -0.5)
    norm[torch.isinf(norm)] = 0
    return norm.unsqueeze(1).to(dtype)


def gcn_layer(g, h, weight, bias, norm, activation, keep=None, p=0.0):
    """GCNLayer.forward, model_zoo.py:34-50."""
    h = _drop(h, keep, p)                       # :35-36
    h = torch.mm(h, weight)                     # :37
    h = h * norm                                # :39
    h = _sum_in(g, h[g.src])                    # :41  copy_src / sum
    h = h * norm                                # :44
    if bias is not None:
        h = h + bias                            # :47
    if activation:
        h = activation(h)                       # :49
    return h


def gat_layer(g, feature, fc_weight, attn_l, attn_r, num_heads, negative_slope=0.2,
              feat_keep=None, p_feat=0.0, attn_keep=None, p_attn=0.0, return_attention=False):
    """GATLayer.forward, model_zoo.py:80-114 (residual branch is dead for every config)."""
    h = _drop(feature, feat_keep, p_feat)                                # :82
    ft = F.linear(h, fc_weight).reshape(h.shape[0], num_heads, -1)       # :83
    a1 = (ft * attn_l).sum(dim=-1).unsqueeze(-1)                         # :84
    a2 = (ft * attn_r).sum(dim=-1).unsqueeze(-1)                         # :85
    a = F.leaky_relu(a1[g.src] + a2[g.dst], negative_slope)              # :108
    attention = edge_softmax(g, a)                                       # :112
    a_drop = _drop(attention, attn_keep, p_attn)                         # :114
    ret = _sum_in(g, ft[g.src] * a_drop)                                 # :95
    if return_attention:
        return ret, attention
    return ret


# --------------------------------------------------------------------------------------
# Parameter containers following the reference state_dict names
# --------------------------------------------------------------------------------------
def _xavier_normal(shape, gain, gen):
    # torch.nn.init.xavier_normal_: fan_in = size(1)*rf, fan_out = size(0)*rf, rf = prod(shape[2:])
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    std = gain * math.sqrt(2.0 / (fan_in + fan_out))
    return torch.randn(shape, generator=gen) * std


def init_pgat_params(in_dim, hidden_dim, out_dim, pos_dim, num_layers, heads, seed=0, positional=True,
                     position_vocab_size=3) -> Dict[str, torch.Tensor]:
    """Same shapes/names/init laws as PGAT.__init__ / GATLayer.__init__ (model_zoo.py:192-208, 52-69).

    (Values are drawn from a private generator, not bit-identical to constructing the reference module.)
    """
    gen = torch.Generator().manual_seed(seed)
    pd = pos_dim if positional else 0
    dims = [(in_dim + pd, hidden_dim, heads[0])]
    for l in range(1, num_layers):
        dims.append((hidden_dim * heads[l - 1] + pd, hidden_dim, heads[l]))
    dims.append((hidden_dim * heads[-2] + pd, out_dim, heads[-1]))
    p = {}
    for i, (k, d, hds) in enumerate(dims):
        p[f"gat_layers.{i}.fc.weight"] = _xavier_normal((hds * d, k), 1.414, gen)
        p[f"gat_layers.{i}.attn_l"] = _xavier_normal((1, hds, d), 1.414, gen)
        p[f"gat_layers.{i}.attn_r"] = _xavier_normal((1, hds, d), 1.414, gen)
        if positional:
            p[f"prop_position_embeddings.{i}.weight"] = torch.randn((position_vocab_size, pos_dim), generator=gen)
    return p


def init_pgcn_params(in_dim, hidden_dim, out_dim, pos_dim, num_layers, seed=0, positional=True,
                     position_vocab_size=3) -> Dict[str, torch.Tensor]:
    """Shapes/names/init laws of PGCN.__init__ / GCNLayer.reset_parameters (model_zoo.py:139-153, 28-32)."""
    gen = torch.Generator().manual_seed(seed)
    pd = pos_dim if positional else 0
    dims = [(in_dim + pd, hidden_dim)] + [(hidden_dim + pd, hidden_dim)] * (num_layers - 1) + [(hidden_dim + pd, out_dim)]
    p = {}
    for i, (k, d) in enumerate(dims):
        stdv = 1.0 / math.sqrt(d)
        p[f"layers.{i}.weight"] = (torch.rand((k, d), generator=gen) * 2 - 1) * stdv
        p[f"layers.{i}.bias"] = (torch.rand((d,), generator=gen) * 2 - 1) * stdv
        if positional:
            p[f"prop_position_embeddings.{i}.weight"] = torch.randn((position_vocab_size, pos_dim), generator=gen)
    return p


# --------------------------------------------------------------------------------------
# Propagation stacks
# --------------------------------------------------------------------------------------
def _activate(h, activation, branch, l, tol=2e-6):
    """Hidden-layer activation.  `branch` (optional {"act.<l>": bool [N, F]}) pins the leaky-relu branch (True = positive
    side) taken by the implementation under test: the derivative is discontinuous at 0, so two fp32-accurate
    implementations may legitimately pick different branches for pre-activations within rounding noise of 0 and their
    GRADIENTS then differ by O(1e-3).  A pinned branch may only disagree with the oracle's own sign where
    |pre-activation| <= tol (asserted) or where the element is dropped afterwards ("keep.<l>" mask)."""
    key = f"act.{l}"
    if branch is None or key not in branch:
        return activation(h)
    assert activation is F.leaky_relu, "branch replay is implemented for leaky_relu(0.01)"
    pos = branch[key]
    if os.environ.get("TAXO_DEBUG_PINS") == "2":      # cross-process repeatability probe of the oracle itself
        hv = h.detach().contiguous()
        bits = hv.view(torch.int32 if hv.dtype == torch.float32 else torch.int64)
        print(f"DEBUG_PINS oracle pre-activation l={l} dtype={hv.dtype} checksum={int(bits.to(torch.int64).sum())} threads={torch.get_num_threads()}")
    disagree = (h.detach() > 0) != pos
    keep = branch.get(f"keep.{l}")
    if keep is not None:
        disagree = disagree & keep
    worst = float(h.detach().abs()[disagree].max()) if bool(disagree.any()) else 0.0
    assert worst <= tol, f"activation branch of layer {l} differs at |pre-activation| = {worst:.3e} > {tol:.1e}"
    return torch.where(pos, h, 0.01 * h)


def pgat_forward(g: OracleGraph, features, params, num_layers, heads, activation=F.leaky_relu,
                 negative_slope=0.2, p_feat=0.0, p_attn=0.0, masks: Optional[dict] = None, positional=True):
    """PGAT.forward (model_zoo.py:210-220); positional=False gives GAT.forward (:183-190).

    masks: optional {"feat.<l>": bool [N, K_l], "attn.<l>": bool [E, H_l, 1]} keep-masks (edge-id order), and
           optional {"act.<l>": bool [N, F_l]} leaky-relu branch pins (see _activate).
    """
    masks = masks or {}
    h = features
    n_total = num_layers + 1
    for l in range(n_total):
        if positional:
            p = params[f"prop_position_embeddings.{l}.weight"][g.pos]             # :214,218
            z = torch.cat((h, p), 1)                                               # :215,219
        else:
            z = h
        out = gat_layer(g, z, params[f"gat_layers.{l}.fc.weight"], params[f"gat_layers.{l}.attn_l"],
                        params[f"gat_layers.{l}.attn_r"], heads[l], negative_slope,
                        masks.get(f"feat.{l}"), p_feat, masks.get(f"attn.{l}"), p_attn)
        if l < num_layers:
            h = _activate(out.flatten(1), activation, masks, l)                    # :215-216
        else:
            h = out.mean(1)                                                        # :219
    return h


def pgcn_forward(g: OracleGraph, features, params, num_layers, activation=F.leaky_relu,
                 p_in=0.0, p_hidden=0.0, p_out=0.0, masks: Optional[dict] = None, positional=True):
    """PGCN.forward (model_zoo.py:155-167); positional=False gives GCN.forward (:128-137)."""
    masks = masks or {}
    h = features
    norm = gcn_norm(g, features.dtype)                                             # :157-161
    n_total = num_layers + 1
    for l in range(n_total):
        if positional:
            z = torch.cat((h, params[f"prop_position_embeddings.{l}.weight"][g.pos]), 1)   # :165-166
        else:
            z = h
        p = p_in if l == 0 else (p_out if l == n_total - 1 else p_hidden)          # :145-152
        act = (lambda v, _l=l: _activate(v, activation, masks, _l)) if l < n_total - 1 else None
        h = gcn_layer(g, z, params[f"layers.{l}.weight"], params[f"layers.{l}.bias"], norm, act,
                      masks.get(f"feat.{l}"), p)
    return h


# --------------------------------------------------------------------------------------
# Readouts (dgl.mean_nodes / sum_nodes semantics restated)
# --------------------------------------------------------------------------------------
def _segment_sum(g: OracleGraph, x):
    out = torch.zeros((g.num_graphs, *x.shape[1:]), dtype=x.dtype)
    return out.index_add(0, g.graph_ids(), x)


def mean_readout(g: OracleGraph, h):
    """MeanReadout, model_zoo.py:231-232: per-graph arithmetic mean."""
    return _segment_sum(g, h) / torch.tensor(g.batch_num_nodes, dtype=h.dtype).unsqueeze(1)


def weighted_mean_readout(g: OracleGraph, h, position_weights):
    """WeightedMeanReadout, model_zoo.py:240-242: a = softplus(w[pos]); sum(a h) / sum(a) per graph."""
    a = F.softplus(position_weights[g.pos])            # [N,1]
    return _segment_sum(g, h * a) / _segment_sum(g, a)


def concat_readout(g: OracleGraph, h):
    """ConcatReadout, model_zoo.py:248-258: [sum_gp/n, mean over anchors (=anchor row), sum_sib/n]."""
    n = torch.tensor(g.batch_num_nodes, dtype=h.dtype).unsqueeze(1)
    a_gp = (g.pos == 0).to(h.dtype).unsqueeze(1)
    a_p = (g.pos == 1).to(h.dtype).unsqueeze(1)
    a_sib = (g.pos == 2).to(h.dtype).unsqueeze(1)
    gp = _segment_sum(g, h * a_gp) / n
    pe = _segment_sum(g, h * a_p) / _segment_sum(g, a_p)
    sib = _segment_sum(g, h * a_sib) / n
    return torch.cat((gp, pe, sib), 1)


# --------------------------------------------------------------------------------------
# Matching + loss (the step right after the hot path; used for end-to-end gradients)
# --------------------------------------------------------------------------------------
def match_bim(hg, qf, W):
    """BIM, model_zoo.py:301-313: nn.Bilinear(l, r, 1, bias=False); W is [1, l, r]."""
    return torch.einsum("gl,olr,gr->go", hg, W, qf)


def match_lbm(hg, qf, W):
    """LBM, model_zoo.py:316-328."""
    return torch.exp(match_bim(hg, qf, W))


def match_mlp(hg, qf, w0, b0, w1, b1):
    """MLP, model_zoo.py:281-298."""
    return F.linear(F.relu(F.linear(torch.cat((hg, qf), 1), w0, b0)), w1, b1)


def info_nce_step_loss(scores, n_queries):
    """trainer.py:52-56 + loss.py:52-57: reshape to [n_queries, 1+neg], CE(sum) against class 0."""
    pred = scores.reshape(n_queries, -1)
    return F.cross_entropy(pred, torch.zeros(n_queries, dtype=torch.long), reduction="sum")


# --------------------------------------------------------------------------------------
# Whole model (TaxoExpan.forward, model.py:70-87) for the CPU arm of the bench
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    propagation_method: str = "PGAT"
    readout_method: str = "WMR"
    matching_method: str = "LBM"
    in_dim: int = 250
    hidden_dim: int = 500
    out_dim: int = 500
    pos_dim: int = 50
    num_layers: int = 1
    heads: Sequence[int] = (4, 1)
    feat_drop: float = 0.0
    attn_drop: float = 0.0
    hidden_drop: float = 0.0
    out_drop: float = 0.0


def init_model_params(cfg: OracleConfig, seed=0) -> Dict[str, torch.Tensor]:
    pm = cfg.propagation_method
    if pm in ("PGAT", "GAT"):
        pp = init_pgat_params(cfg.in_dim, cfg.hidden_dim, cfg.out_dim, cfg.pos_dim, cfg.num_layers, list(cfg.heads),
                              seed, positional=(pm == "PGAT"))
    else:
        pp = init_pgcn_params(cfg.in_dim, cfg.hidden_dim, cfg.out_dim, cfg.pos_dim, cfg.num_layers, seed,
                              positional=(pm == "PGCN"))
    params = {f"graph_propagate.{k}": v for k, v in pp.items()}
    gen = torch.Generator().manual_seed(seed + 1)
    if cfg.readout_method == "WMR":
        params["readout.position_weights.weight"] = torch.randn((3, 1), generator=gen)
    l_dim = cfg.out_dim * (3 if cfg.readout_method == "CR" else 1)
    r_dim = cfg.in_dim
    if cfg.matching_method in ("LBM", "BIM"):
        bound = 1.0 / math.sqrt(l_dim)
        params["match.W.weight"] = (torch.rand((1, l_dim, r_dim), generator=gen) * 2 - 1) * bound
    else:
        k0 = 1.0 / math.sqrt(l_dim + r_dim)
        k1 = 1.0 / math.sqrt(cfg.hidden_dim)
        params["match.ffn.0.weight"] = (torch.rand((cfg.hidden_dim, l_dim + r_dim), generator=gen) * 2 - 1) * k0
        params["match.ffn.0.bias"] = (torch.rand((cfg.hidden_dim,), generator=gen) * 2 - 1) * k0
        params["match.ffn.2.weight"] = (torch.rand((1, cfg.hidden_dim), generator=gen) * 2 - 1) * k1
        params["match.ffn.2.bias"] = (torch.rand((1,), generator=gen) * 2 - 1) * k1
    return params


def propagate(cfg: OracleConfig, g: OracleGraph, h, params, masks=None, training=False):
    sub = {k[len("graph_propagate."):]: v for k, v in params.items() if k.startswith("graph_propagate.")}
    pm = cfg.propagation_method
    tr = 1.0 if training else 0.0
    if pm in ("PGAT", "GAT"):
        return pgat_forward(g, h, sub, cfg.num_layers, list(cfg.heads), p_feat=cfg.feat_drop * tr,
                            p_attn=cfg.attn_drop * tr, masks=masks, positional=(pm == "PGAT"))
    return pgcn_forward(g, h, sub, cfg.num_layers, p_in=cfg.feat_drop * tr, p_hidden=cfg.hidden_drop * tr,
                        p_out=cfg.out_drop * tr, masks=masks, positional=(pm == "PGCN"))


def readout(cfg: OracleConfig, g: OracleGraph, h, params):
    if cfg.readout_method == "MR":
        return mean_readout(g, h)
    if cfg.readout_method == "WMR":
        return weighted_mean_readout(g, h, params["readout.position_weights.weight"])
    return concat_readout(g, h)


def match(cfg: OracleConfig, hg, qf, params):
    if cfg.matching_method == "LBM":
        return match_lbm(hg, qf, params["match.W.weight"])
    if cfg.matching_method == "BIM":
        return match_bim(hg, qf, params["match.W.weight"])
    return match_mlp(hg, qf, params["match.ffn.0.weight"], params["match.ffn.0.bias"],
                     params["match.ffn.2.weight"], params["match.ffn.2.bias"])


def taxoexpan_forward(cfg: OracleConfig, g: OracleGraph, h, qf, params, masks=None, training=False):
    """TaxoExpan.forward, model.py:83-87. Returns (scores [G,1], hg [G,l], node_h [N,out])."""
    node_h = propagate(cfg, g, h, params, masks, training)
    hg = readout(cfg, g, node_h, params)
    return match(cfg, hg, qf, params), hg, node_h


def random_keep_masks(cfg: OracleConfig, g: OracleGraph, seed=0):
    """torch-CPU Bernoulli keep-masks for the CPU arm in train mode (timing only; not bit-parity)."""
    gen = torch.Generator().manual_seed(seed)
    masks = {}
    n_total = cfg.num_layers + 1
    pd = cfg.pos_dim if cfg.propagation_method in ("PGAT", "PGCN") else 0
    if cfg.propagation_method in ("PGAT", "GAT"):
        heads = list(cfg.heads)
        k = cfg.in_dim + pd
        for l in range(n_total):
            masks[f"feat.{l}"] = torch.rand((g.n, k), generator=gen) >= cfg.feat_drop
            masks[f"attn.{l}"] = torch.rand((g.src.numel(), heads[l], 1), generator=gen) >= cfg.attn_drop
            k = cfg.hidden_dim * heads[l] + pd
    else:
        k = cfg.in_dim + pd
        for l in range(n_total):
            p = cfg.feat_drop if l == 0 else (cfg.out_drop if l == n_total - 1 else cfg.hidden_drop)
            masks[f"feat.{l}"] = torch.rand((g.n, k), generator=gen) >= p
            k = cfg.hidden_dim + pd
    return masks


# ---------------------------------------------------------------------------------------------------------------------
# evaluation: rank of every true position of a query among all candidate positions
# ---------------------------------------------------------------------------------------------------------------------
def ranks_from_similarities(all_similarities: np.ndarray, positive_relations) -> List[int]:
    """reference model/metric.py:7-19: rank of each true position = 1 + number of NON-true positions with a strictly larger
    similarity (the true positions themselves are masked out)."""
    all_similarities = np.asarray(all_similarities)
    pos = np.asarray(list(positive_relations), dtype=np.int64)
    neg_mask = np.ones(all_similarities.shape[0], dtype=bool)
    neg_mask[pos] = False
    neg = all_similarities[neg_mask]
    return [int((neg > all_similarities[p]).sum()) + 1 for p in pos]


def all_pairs_ranks(cfg: "OracleConfig", hg, queries, params, positives) -> List[List[int]]:
    """reference test_fast.py:187-218: for every query, score it against every position with model.match on the expanded query
    feature, then rank its true positions (similarity mode, used with the InfoNCE loss)."""
    out = []
    for j in range(queries.shape[0]):
        scores = match(cfg, hg, queries[j:j + 1].expand(hg.shape[0], -1), params).reshape(-1).detach().numpy()
        out.append(ranks_from_similarities(scores, positives[j]))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# egonet construction (the data layer's format spec for the kernels' input)
# ---------------------------------------------------------------------------------------------------------------------
def get_subgraph_nodes(parents_of, children_of, query_node, anchor_node, instance_mode, expand_factor=50, rng=None):
    """reference data_loader/dataset.py:404-426 (`_get_subgraph`) for ONE egonet: (node ids, positions) in the reference's node
    order.  parents_of / children_of: dict node -> list in edge order (networkx in_edges / out_edges)."""
    import random
    rng = rng or random
    nodes = list(parents_of.get(anchor_node, []))
    pos = [0] * len(nodes)
    nodes.append(anchor_node)
    pos.append(1)
    kids = list(children_of.get(anchor_node, []))
    if len(kids) > expand_factor:
        kids = rng.choices(kids, k=expand_factor)
    if instance_mode == 1:
        kids = [k for k in kids if k != query_node]
    nodes.extend(kids)
    pos.extend([2] * len(kids))
    return nodes, pos


# ---------------------------------------------------------------------------------------------------------------------
# negative-anchor sampling and the egonet cache (the step before egonet construction)
# ---------------------------------------------------------------------------------------------------------------------
def node_masks(parents_of, children_of, nodes, roots):
    """reference data_loader/dataset.py:248-258: node2masks[n] = descendants(n) + parents(n) + [n] + roots, for every n in `nodes`
    (positions that must not be drawn as negative anchors of query n)."""
    out = {}
    for n in nodes:
        seen, stack = set(), list(children_of.get(n, []))
        while stack:                                   # nx.descendants: everything reachable through out-edges, n itself excluded
            v = stack.pop()
            if v in seen or v == n:
                continue
            seen.add(v)
            stack.extend(children_of.get(v, []))
        out[n] = set(list(seen) + list(parents_of.get(n, [])) + [n] + list(roots))
    return out


class NegativeQueueOracle:
    """reference data_loader/dataset.py:285-287,334-381 restated line by line: a queue of candidate positions (train ids x 5)
    walked by a pointer and reshuffled when exhausted; `rng` is a random.Random (the reference uses the module-level one)."""

    def __init__(self, train_node_ids, node2masks, rng):
        self.queue = (list(train_node_ids) * 5).copy()
        self.pointer = 0
        self.node2masks = node2masks
        self.rng = rng

    def at_most_k(self, query_node, negative_size):          # dataset.py:340-356
        if self.pointer == 0:
            self.rng.shuffle(self.queue)
        while True:
            negatives = [e for e in self.queue[self.pointer: self.pointer + negative_size] if e not in self.node2masks[query_node]]
            if len(negatives) > 0:
                break
        self.pointer += negative_size
        if self.pointer >= len(self.queue):
            self.pointer = 0
        return negatives

    def exactly_k(self, query_node, negative_size):          # dataset.py:358-381
        if self.pointer == 0:
            self.rng.shuffle(self.queue)
        masks = self.node2masks[query_node]
        negatives = []
        max_try = 0
        while len(negatives) != negative_size:
            n_lack = negative_size - len(negatives)
            negatives.extend([e for e in self.queue[self.pointer: self.pointer + n_lack] if e not in masks])
            self.pointer += n_lack
            if self.pointer >= len(self.queue):
                self.pointer = 0
                self.rng.shuffle(self.queue)
            max_try += 1
            if max_try > 10:
                if len(negatives) > negative_size:
                    negatives = negatives[:negative_size]
                else:
                    negatives.extend([e for e in self.queue[: (negative_size - len(negatives))]])
        return negatives


_M64 = (1 << 64) - 1
POSITIVE_GENERATION_BASE = 1 << 40


def counter_draw(seed, anchor, generation, slot, degree):
    """The counter-based replacement of `random.choices(out_edges, k=expand_factor)` (dataset.py:419,424) shared by the oracle and
    taxoexpan_b200.sampler: draw number `slot` of generation `generation` of `anchor` = floor(u * degree), u a 53-bit uniform from a
    splitmix64 finaliser of the four integers.  Pure function of its arguments, so a 'cached' egonet is reproducible from its
    generation number alone."""
    z = (seed ^ (anchor * 0x9E3779B97F4A7C15) ^ (generation * 0xC2B2AE3D27D4EB4F) ^ (slot * 0x165667B19E3779F9)) & _M64
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return min(int(float(z >> 11) * (degree / 2.0 ** 53)), degree - 1)


class EgonetCacheOracle:
    """reference data_loader/dataset.py:383-402 (`_get_subgraph_and_node_pair`) restated: negative egonets are cached per anchor
    and reused until they have been read `cache_refresh_time` times; positives are always rebuilt and never cached.  The random
    sibling draws of dataset.py:419,424 come from `counter_draw` (generation = number of earlier creations for that anchor)."""

    def __init__(self, parents_of, children_of, expand_factor, cache_refresh_time, seed):
        self.parents_of, self.children_of = parents_of, children_of
        self.expand_factor, self.cache_refresh_time, self.seed = expand_factor, cache_refresh_time, seed
        self.cache, self.cache_counter = {}, {}
        self.created, self.positives = {}, 0

    def _get_subgraph(self, query_node, anchor_node, instance_mode):
        nodes = list(self.parents_of.get(anchor_node, []))
        nodes.append(anchor_node)
        kids = list(self.children_of.get(anchor_node, []))
        if instance_mode == 0:
            generation = self.created.get(anchor_node, 0)
            self.created[anchor_node] = generation + 1
        else:
            generation = POSITIVE_GENERATION_BASE + self.positives
            self.positives += 1
        if len(kids) > self.expand_factor:
            kids = [kids[counter_draw(self.seed, anchor_node, generation, t, len(kids))] for t in range(self.expand_factor)]
        if instance_mode == 1:
            kids = [k for k in kids if k != query_node]
        return nodes + kids

    def get(self, query_node, anchor_node, instance_mode):
        if instance_mode == 0 and (anchor_node in self.cache) and (self.cache_counter[anchor_node] < self.cache_refresh_time):
            g = self.cache[anchor_node]
            self.cache_counter[anchor_node] += 1
        else:
            g = self._get_subgraph(query_node, anchor_node, instance_mode)
            if instance_mode == 0:
                self.cache[anchor_node] = g
                self.cache_counter[anchor_node] = 0
        return g


class TrainItemOracle:
    """reference MaskedGraphDataset.__getitem__ for sampling_mode 1 in train mode (dataset.py:290-332) restated on top of the
    restated queue walk and egonet cache: [(egonet node list, query node, label)], the positive first."""

    def __init__(self, node_list, node2parents, queue: NegativeQueueOracle, cache: EgonetCacheOracle, negative_size):
        self.node_list, self.node2parents, self.queue, self.cache, self.negative_size = node_list, node2parents, queue, cache, negative_size
        self.node2positive_pointer = {n: 0 for n in node2parents}

    def __getitem__(self, idx):
        res = []
        query_node = self.node_list[idx]
        positive_pointer = self.node2positive_pointer[query_node]
        parent_node = self.node2parents[query_node][positive_pointer]
        res.append((self.cache.get(query_node, parent_node, 1), query_node, 1))
        self.node2positive_pointer[query_node] = (positive_pointer + 1) % len(self.node2parents[query_node])
        for negative_parent in self.queue.exactly_k(query_node, self.negative_size):
            res.append((self.cache.get(query_node, negative_parent, 0), query_node, 0))
        return res


def large_batch_chunks(nodes_per_graph, limit=100000):
    """reference data_loaders.py:31-72 (`collate_graph_and_node_large_batch`) restated: [(first, last + 1)] of every emitted batch."""
    out, start, count, size = [], 0, 0, 0
    for i, n in enumerate(nodes_per_graph):
        size += 1
        count += n
        if count > limit and size > 1:
            out.append((start, i + 1))
            start, count, size = i + 1, 0, 0
    if size != 0:
        out.append((start, len(nodes_per_graph)))
    return out
