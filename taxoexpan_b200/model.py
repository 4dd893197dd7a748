"""`TaxoExpan`: the reference's string-dispatched assembly of propagate / readout / match (model/model.py:13-87),
built from the B200 modules of taxoexpan_b200.model_zoo.  Same constructor, same attributes
(`graph_propagate`, `readout`, `match`, `readout_method`, ...), same forward signature and side effects."""
import numpy as np
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import functional as txf
from .graph import as_int32_pos
from .model_zoo import GCN, GAT, PGCN, PGAT, MeanReadout, WeightedMeanReadout, ConcatReadout, MLP, BIM, LBM


class BaseModel(nn.Module):
    """reference base/base_model.py:6-25 (prints the trainable-parameter count)"""

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + '\nTrainable parameters: {}'.format(params)


class TaxoExpan(BaseModel):
    def __init__(self, propagation_method, readout_method, matching_method, **options):
        super().__init__()
        self.propagation_method = propagation_method
        self.readout_method = readout_method
        self.matching_method = matching_method
        self.options = options
        o = options
        if propagation_method == "GCN":
            self.graph_propagate = GCN(o["in_dim"], o["hidden_dim"], o["out_dim"], num_layers=o["num_layers"],
                                       activation=F.leaky_relu, in_dropout=o["feat_drop"], hidden_dropout=o["hidden_drop"],
                                       output_dropout=o["out_drop"])
        elif propagation_method == "PGCN":
            self.graph_propagate = PGCN(o["in_dim"], o["hidden_dim"], o["out_dim"], o["pos_dim"], num_layers=o["num_layers"],
                                        activation=F.leaky_relu, in_dropout=o["feat_drop"], hidden_dropout=o["hidden_drop"],
                                        output_dropout=o["out_drop"])
        elif propagation_method == "GAT":
            self.graph_propagate = GAT(o["in_dim"], o["hidden_dim"], o["out_dim"], num_layers=o["num_layers"], heads=o["heads"],
                                       activation=F.leaky_relu, feat_drop=o["feat_drop"], attn_drop=o["attn_drop"])
        elif propagation_method == "PGAT":
            self.graph_propagate = PGAT(o["in_dim"], o["hidden_dim"], o["out_dim"], o["pos_dim"], num_layers=o["num_layers"],
                                        heads=o["heads"], activation=F.leaky_relu, feat_drop=o["feat_drop"],
                                        attn_drop=o["attn_drop"])
        else:
            # the reference's `assert f"..."` never fires (model.py:43); an unknown name is a configuration error
            raise ValueError(f"Unacceptable Graph Propagation Method: {propagation_method}")

        if readout_method == "MR":
            self.readout = MeanReadout()
            l_dim, r_dim = o["out_dim"], o["in_dim"]
        elif readout_method == "WMR":
            self.readout = WeightedMeanReadout()
            l_dim, r_dim = o["out_dim"], o["in_dim"]
        elif readout_method == "CR":
            self.readout = ConcatReadout()
            l_dim, r_dim = o["out_dim"] * 3, o["in_dim"]
        else:
            raise ValueError(f"Unacceptable Readout Method: {readout_method}")

        if matching_method == "MLP":
            self.match = MLP(l_dim, r_dim, o["hidden_dim"])
        elif matching_method == "LBM":
            self.match = LBM(l_dim, r_dim)
        elif matching_method == "BIM":
            self.match = BIM(l_dim, r_dim)
        else:
            raise ValueError(f"Unacceptable Matching Method: {matching_method}")

    def forward(self, g, h, qf):
        """model/model.py:70-87: scores[G,1] of each egonet in batched graph g against its query feature."""
        if h.shape[0] == 0:          # empty batch (the reference cannot build one: dgl.batch([]) fails): nothing to score
            return h.new_zeros((0, 1))
        pos = g.ndata['pos'].to(h.device)
        g.ndata['h'] = self.graph_propagate(g, h)
        kind = getattr(self.readout, "kind", None)
        w_match = None
        if type(self.match) in (BIM, LBM):     # nn.Bilinear weight [1, l, r] seen as [l, r] - a VIEW (weight[0] would cost a zero-fill + copy in
            w_match = self.match.W.weight      # its select-backward, and hide the parameter from the gradient sinks)
            w_match = w_match.view(w_match.shape[1], w_match.shape[2])
        if w_match is not None and kind is not None and txf.head_native_ok(g.ndata['h'], qf, kind, w_match):
            # model.py:85-86 as ONE native call per direction (readout, projection GEMM, row-dot; tx_head_fwd / tx_head_bwd): the same
            # kernels as self.readout(g, pos) followed by self.match(hg, qf)
            pw = getattr(self.readout, "position_weights", None)
            pos32 = as_int32_pos(pos, h.device) if kind == _lib.TX_READOUT_WMEAN else None
            return txf.ReadoutMatch.apply(g.ndata['h'], None if pw is None else pw.weight, w_match, qf,
                                          g.structure(h.device), pos32, kind, self.match.apply_exp)
        hg = self.readout(g, pos)
        scores = self.match(hg, qf)
        return scores
