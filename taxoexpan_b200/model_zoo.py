"""Drop-in replacements for the reference's propagation / readout / matching modules.

Same class names, constructor arguments, forward signatures, parameter names/shapes/initialisers and side effects
as reference `model/model_zoo.py`, so `model/model.py`-style glue (and released checkpoints: state_dict keys
`gat_layers.{i}.fc.weight|attn_l|attn_r`, `layers.{i}.weight|bias`, `prop_position_embeddings.{i}.weight`,
`position_weights.weight`, `W.weight`) work unchanged.  The message passing itself runs in hand-written sm_100a CUDA
kernels behind the C ABI (taxoexpan_b200/functional.py); there is no DGL and no CPU fallback.

Graph argument `g`: a `taxoexpan_b200.graph.DGLGraph` / `EgonetBatch` (dict-like `ndata` with 'pos', `batch_num_nodes`).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import functional as txf
from .graph import as_int32_pos


def _act_slope(activation) -> float:
    """Negative-side slope of a leaky-relu-family activation; 1.0 = identity."""
    if activation is None:
        return 1.0
    if activation is F.leaky_relu:
        return 0.01                      # F.leaky_relu default negative_slope (reference model.py:25,30,35,40)
    if activation is F.relu or activation is torch.relu:
        return 0.0
    if isinstance(activation, nn.LeakyReLU):
        return float(activation.negative_slope)
    if isinstance(activation, nn.ReLU):
        return 0.0
    raise NotImplementedError(f"activation {activation!r}: the fused epilogue supports leaky_relu / relu / None")


def _rate(p) -> float:
    return float(p) if p else 0.0


# =====================================================================================================
# Graph propagation modules: GCN, GAT, PGCN, PGAT
# =====================================================================================================
class GCNLayer(nn.Module):
    """reference model_zoo.py:13-50"""

    def __init__(self, in_feats, out_feats, activation, dropout, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        self.bias = nn.Parameter(torch.Tensor(out_feats)) if bias else None
        self.activation = activation
        self.dropout = _rate(dropout)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, g, h):
        """Stand-alone layer call (needs g.ndata['norm'] semantics -> taken from the graph structure)."""
        st = g.structure(h.device)
        p = self.dropout if self.training else 0.0
        z = txf.ConcatPosDropout.apply(h, None, None, p, txf.new_seed() if p else 0, 0)
        cfg = txf.GcnLayerCfg(k=h.shape[1], dim=self.weight.shape[1], hidden=True, act_slope=_act_slope(self.activation))
        out = txf.GcnLayer.apply(z, self.weight, self.bias, None, st, None, cfg)
        return out[:, :self.weight.shape[1]]


class GATLayer(nn.Module):
    """reference model_zoo.py:52-114 (the residual branch is never enabled by model/model.py and is not built)"""

    def __init__(self, in_dim, out_dim, num_heads=1, feat_drop=0.5, attn_drop=0.5, leaky_relu_alpha=0.2, residual=False):
        super().__init__()
        if residual:
            raise NotImplementedError("residual GAT layers are dead code in the reference (model.py never sets residual)")
        self.num_heads = num_heads
        self.out_dim = out_dim
        self.fc = nn.Linear(in_dim, num_heads * out_dim, bias=False)
        self.feat_drop = _rate(feat_drop)
        self.attn_drop = _rate(attn_drop)
        self.attn_l = nn.Parameter(torch.Tensor(size=(1, num_heads, out_dim)))
        self.attn_r = nn.Parameter(torch.Tensor(size=(1, num_heads, out_dim)))
        nn.init.xavier_normal_(self.fc.weight.data, gain=1.414)
        nn.init.xavier_normal_(self.attn_l.data, gain=1.414)
        nn.init.xavier_normal_(self.attn_r.data, gain=1.414)
        self.negative_slope = float(leaky_relu_alpha)
        self.residual = False

    def forward(self, g, feature):
        """Stand-alone layer call: returns [N, H, D'] like the reference."""
        st = g.structure(feature.device)
        tr = self.training
        p_f = self.feat_drop if tr else 0.0
        p_a = self.attn_drop if tr else 0.0
        z = txf.ConcatPosDropout.apply(feature, None, None, p_f, txf.new_seed() if p_f else 0, 0)
        cfg = txf.GatLayerCfg(k=feature.shape[1], heads=self.num_heads, dim=self.out_dim, neg_slope=self.negative_slope,
                              p_attn=p_a, attn_seed=txf.new_seed() if p_a else 0, attn_stream=1, hidden=True, act_slope=1.0)
        out = txf.GatLayer.apply(z, self.fc.weight, self.attn_l, self.attn_r, None, st, None, cfg)
        return out[:, :self.num_heads * self.out_dim].reshape(feature.shape[0], self.num_heads, self.out_dim)


def _gat_stack_forward(mod, g, features, positions, pos_tables):
    """Shared by GAT (pos_tables=None) and PGAT: model_zoo.py:183-190 / 210-220."""
    dev = features.device
    st = g.structure(dev)
    pos32 = as_int32_pos(positions, dev) if positions is not None else None
    tr = mod.training
    layers = mod.gat_layers
    n_total = len(layers)
    seed = txf.new_seed() if tr else 0
    slope = _act_slope(mod.activation)
    tab = (lambda l: pos_tables[l].weight) if pos_tables is not None else (lambda l: None)
    pd = pos_tables[0].weight.shape[1] if pos_tables is not None else 0
    p0 = layers[0].feat_drop if tr else 0.0
    link0 = txf.MaskLink() if txf.concat_publishes_f16() else None      # layer-0 input handed over as an fp16 operand pair
    z = txf.ConcatPosDropout.apply(features, tab(0), pos32, p0, seed, 0, link0)
    if link0 is not None and link0.c_state is None:
        link0 = None
    k = features.shape[1] + pd
    links = [txf.MaskLink() for _ in range(n_total - 1)]     # layer l's epilogue mask -> layer l+1's d(z) GEMM epilogue
    for l, layer in enumerate(layers):
        hidden = l < n_total - 1
        cfg = txf.GatLayerCfg(
            in_link=links[l - 1] if l > 0 else link0, out_link=links[l] if hidden else None,
            k=k, heads=layer.num_heads, dim=layer.out_dim, neg_slope=layer.negative_slope,
            p_attn=layer.attn_drop if tr else 0.0, attn_seed=seed, attn_stream=2 * l + 1, hidden=hidden,
            act_slope=slope if hidden else 1.0,
            p_next=(layers[l + 1].feat_drop if tr else 0.0) if hidden else 0.0, next_seed=seed, next_stream=2 * (l + 1),
            dz_from=features.shape[1] if (l == 0 and not features.requires_grad) else 0, tag=f"L{l}")
        z = txf.GatLayer.apply(z, layer.fc.weight, layer.attn_l, layer.attn_r, tab(l + 1) if hidden else None, st, pos32, cfg)
        k = layer.num_heads * layer.out_dim + pd
    return z


class GAT(nn.Module):
    """reference model_zoo.py:169-190"""

    def __init__(self, in_dim, hidden_dim, out_dim, num_layers, heads, activation, feat_drop=0.5, attn_drop=0.5,
                 leaky_relu_alpha=0.2, residual=False):
        super().__init__()
        self.num_layers = num_layers
        self.gat_layers = nn.ModuleList()
        self.activation = activation
        self.gat_layers.append(GATLayer(in_dim, hidden_dim, heads[0], feat_drop, attn_drop, leaky_relu_alpha, False))
        for l in range(1, num_layers):
            self.gat_layers.append(GATLayer(hidden_dim * heads[l - 1], hidden_dim, heads[l], feat_drop, attn_drop, leaky_relu_alpha, residual))
        self.gat_layers.append(GATLayer(hidden_dim * heads[-2], out_dim, heads[-1], feat_drop, attn_drop, leaky_relu_alpha, residual))

    def forward(self, g, features):
        return _gat_stack_forward(self, g, features, None, None)


class PGAT(nn.Module):
    """reference model_zoo.py:192-220"""

    def __init__(self, in_dim, hidden_dim, out_dim, pos_dim, num_layers, heads, activation, feat_drop=0.5, attn_drop=0.5,
                 leaky_relu_alpha=0.2, residual=False, position_vocab_size=3):
        super().__init__()
        self.num_layers = num_layers
        self.gat_layers = nn.ModuleList()
        self.prop_position_embeddings = nn.ModuleList()
        self.activation = activation
        self.gat_layers.append(GATLayer(in_dim + pos_dim, hidden_dim, heads[0], feat_drop, attn_drop, leaky_relu_alpha, False))
        self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))
        for l in range(1, num_layers):
            self.gat_layers.append(GATLayer(hidden_dim * heads[l - 1] + pos_dim, hidden_dim, heads[l], feat_drop, attn_drop, leaky_relu_alpha, residual))
            self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))
        self.gat_layers.append(GATLayer(hidden_dim * heads[-2] + pos_dim, out_dim, heads[-1], feat_drop, attn_drop, leaky_relu_alpha, residual))
        self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))

    def forward(self, g, features):
        positions = g.ndata.pop('pos')          # same side effect as the reference (model_zoo.py:212)
        return _gat_stack_forward(self, g, features, positions, self.prop_position_embeddings)


def _gcn_stack_forward(mod, g, features, positions, pos_tables):
    """Shared by GCN and PGCN: model_zoo.py:128-137 / 155-167."""
    dev = features.device
    st = g.structure(dev)
    pos32 = as_int32_pos(positions, dev) if positions is not None else None
    tr = mod.training
    layers = mod.layers
    n_total = len(layers)
    seed = txf.new_seed() if tr else 0
    tab = (lambda l: pos_tables[l].weight) if pos_tables is not None else (lambda l: None)
    pd = pos_tables[0].weight.shape[1] if pos_tables is not None else 0
    g.ndata['norm'] = st.gcn_norm().unsqueeze(1)      # reference side effect (model_zoo.py:134,161)
    p0 = layers[0].dropout if tr else 0.0
    link0 = txf.MaskLink() if (txf.concat_publishes_f16() and txf.GCN_FUSED) else None
    z = txf.ConcatPosDropout.apply(features, tab(0), pos32, p0, seed, 0, link0)
    if link0 is not None and link0.c_state is None:
        link0 = None
    k = features.shape[1] + pd
    links = [txf.MaskLink() for _ in range(n_total - 1)]     # layer l's epilogue mask -> layer l+1's d(z) GEMM epilogue
    for l, layer in enumerate(layers):
        hidden = l < n_total - 1
        cfg = txf.GcnLayerCfg(
            in_link=links[l - 1] if l > 0 else link0, out_link=links[l] if hidden else None,
            k=k, dim=layer.weight.shape[1], hidden=hidden, act_slope=_act_slope(layer.activation) if hidden else 1.0,
            p_next=(layers[l + 1].dropout if tr else 0.0) if hidden else 0.0, next_seed=seed, next_stream=2 * (l + 1),
            dz_from=features.shape[1] if (l == 0 and not features.requires_grad) else 0, tag=f"L{l}")
        if not hidden and layer.activation is not None:
            raise NotImplementedError("the output GCN layer has no activation in the reference (model_zoo.py:126,152)")
        z = txf.GcnLayer.apply(z, layer.weight, layer.bias, tab(l + 1) if hidden else None, st, pos32, cfg)
        k = layer.weight.shape[1] + pd
    return z


class GCN(nn.Module):
    """reference model_zoo.py:116-137"""

    def __init__(self, in_dim, hidden_dim, out_dim, num_layers, activation, in_dropout=0.1, hidden_dropout=0.1, output_dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(GCNLayer(in_dim, hidden_dim, activation, in_dropout))
        for l in range(num_layers - 1):
            self.layers.append(GCNLayer(hidden_dim, hidden_dim, activation, hidden_dropout))
        self.layers.append(GCNLayer(hidden_dim, out_dim, None, output_dropout))

    def forward(self, g, features):
        return _gcn_stack_forward(self, g, features, None, None)


class PGCN(nn.Module):
    """reference model_zoo.py:139-167"""

    def __init__(self, in_dim, hidden_dim, out_dim, pos_dim, num_layers, activation, in_dropout=0.1, hidden_dropout=0.1,
                 output_dropout=0.0, position_vocab_size=3):
        super().__init__()
        self.layers = nn.ModuleList()
        self.prop_position_embeddings = nn.ModuleList()
        self.layers.append(GCNLayer(in_dim + pos_dim, hidden_dim, activation, in_dropout))
        self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))
        for l in range(num_layers - 1):
            self.layers.append(GCNLayer(hidden_dim + pos_dim, hidden_dim, activation, hidden_dropout))
            self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))
        self.layers.append(GCNLayer(hidden_dim + pos_dim, out_dim, None, output_dropout))
        self.prop_position_embeddings.append(nn.Embedding(position_vocab_size, pos_dim))

    def forward(self, g, features):
        positions = g.ndata.pop('pos')          # model_zoo.py:163
        return _gcn_stack_forward(self, g, features, positions, self.prop_position_embeddings)


# =====================================================================================================
# Graph readout modules: MR, WMR, CR
# =====================================================================================================
class MeanReadout(nn.Module):
    """reference model_zoo.py:227-232"""
    kind = _lib.TX_READOUT_MEAN

    def __init__(self):
        super().__init__()

    def forward(self, g, pos=None):
        h = g.ndata['h']
        return txf.Readout.apply(h, None, g.structure(h.device), None, _lib.TX_READOUT_MEAN)


class WeightedMeanReadout(nn.Module):
    """reference model_zoo.py:234-242"""
    kind = _lib.TX_READOUT_WMEAN

    def __init__(self):
        super().__init__()
        self.position_weights = nn.Embedding(3, 1)
        self.nonlinear = F.softplus

    def forward(self, g, pos):
        h = g.ndata['h']
        pos32 = as_int32_pos(pos, h.device)
        return txf.Readout.apply(h, self.position_weights.weight, g.structure(h.device), pos32, _lib.TX_READOUT_WMEAN)


class ConcatReadout(nn.Module):
    """reference model_zoo.py:244-258"""

    def __init__(self):
        super().__init__()

    def forward(self, g, pos):
        h = g.ndata['h']
        pos32 = as_int32_pos(pos, h.device)
        return txf.Readout.apply(h, None, g.structure(h.device), pos32, _lib.TX_READOUT_CONCAT)


# =====================================================================================================
# Graph matching modules: the bilinear forms' projection e1 W and its autograd GEMMs run on the library's
# fp32-faithful tensor-core kernels (they were 180 us of cuBLAS SIMT sgemm per step), the row-dot with the query features (+ LBM's exp)
# on tx_match_rowdot_fwd/bwd; the InfoNCE loss that follows is taxoexpan_b200.loss.info_nce_loss.  MLP stays torch glue.
# =====================================================================================================
class MLP(nn.Module):
    """reference model_zoo.py:281-298"""

    def __init__(self, l_dim, r_dim, hidden_dim):
        super().__init__()
        self.ffn = nn.Sequential(nn.Linear(l_dim + r_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, 1))

    def forward(self, e1, e2):
        return self.ffn(torch.cat((e1, e2), 1))


class BIM(nn.Module):
    """reference model_zoo.py:301-313; same parameter (`W.weight` [1, l, r]) evaluated as (e1 W) . e2"""

    def __init__(self, l_dim, r_dim):
        super().__init__()
        self.W = nn.Bilinear(l_dim, r_dim, 1, bias=False)

    apply_exp = False

    def forward(self, e1, e2):
        return txf.match_rowdot(txf.dense_right(e1, self.W.weight[0]), e2, self.apply_exp)     # tx_match_rowdot_fwd/bwd


class LBM(BIM):
    """reference model_zoo.py:316-328: exp of the bilinear form (fused into the row-dot kernel)"""

    apply_exp = True
