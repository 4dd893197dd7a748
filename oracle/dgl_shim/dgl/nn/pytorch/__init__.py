"""`dgl.nn.pytorch.edge_softmax` stand-in (DGL 0.4.0 semantics, restated).

TEST INFRASTRUCTURE ONLY -- see dgl/__init__.py in this directory.
"""
import torch


def edge_softmax(graph, logits):
    """Softmax of `logits` [E, ...] over the in-edges of every destination node.

    DGL 0.4.0 `EdgeSoftmax.forward`: score = exp(x - max_over_in_edges) and
    out = score / sum_over_in_edges(score), independently per trailing index.
    Written with differentiable torch ops so autograd yields the same gradient
    as DGL's hand-written backward (out*g - out*sum_in(out*g)).
    """
    dst = graph._dst
    n = graph.number_of_nodes()
    tail = logits.shape[1:]
    idx = dst.view(-1, *([1] * len(tail))).expand_as(logits)
    mx = torch.full((n, *tail), float("-inf"), dtype=logits.dtype)
    mx = mx.scatter_reduce(0, idx, logits.detach(), reduce="amax", include_self=True)
    e = torch.exp(logits - mx[dst])
    den = torch.zeros((n, *tail), dtype=logits.dtype).index_add_(0, dst, e)
    return e / den[dst]
