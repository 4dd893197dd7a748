"""Pure-torch stand-in for the handful of DGL 0.4.0 calls the reference hot path uses.

TEST INFRASTRUCTURE ONLY.  DGL 0.4.0 (reference README.md:7-12) is an un-vendored
third-party dependency that cannot be installed offline, so its semantics are
RESTATED here from its published behaviour ("parity unpinned" for DGL itself).
This package is put on PYTHONPATH together with /root/reference so that the
reference's model/model.py and model/model_zoo.py run byte-for-byte unmodified on
CPU torch; it is used only by oracle/make_golden.py and tests, never by the
product package.

Call sites covered (reference file:line):
  DGLGraph(), add_nodes, add_edges, nodes        data_loader/dataset.py:429-435
  dgl.batch                                      data_loader/data_loaders.py:25
  g.ndata / g.edata (dict-like, pop)             model/model.py:83-84, model_zoo.py:40-42,86-96,112-114
  g.in_degrees()                                 model_zoo.py:130,157
  g.apply_edges(udf)                             model_zoo.py:90
  g.update_all(copy_src|src_mul_edge, sum)       model_zoo.py:41,95
  dgl.mean_nodes / dgl.sum_nodes                 model_zoo.py:232,242,252-256
  g.batch_num_nodes                              model_zoo.py:249
"""
import torch

from . import function  # noqa: F401
from . import nn  # noqa: F401


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class DGLGraph:
    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.int64)
        self._dst = torch.zeros(0, dtype=torch.int64)
        self.ndata = {}
        self.edata = {}
        self.batch_num_nodes = [0]
        self.batch_num_edges = [0]

    # --- construction -------------------------------------------------
    def add_nodes(self, num, data=None):
        assert self._n == 0, "shim supports a single add_nodes call (dataset.py:430)"
        self._n = int(num)
        self.batch_num_nodes = [self._n]
        if data:
            self.ndata.update(data)

    def add_edges(self, u, v):
        u = torch.as_tensor(u, dtype=torch.int64).reshape(-1)
        v = torch.as_tensor(v, dtype=torch.int64).reshape(-1)
        if u.numel() == 1 and v.numel() != 1:
            u = u.expand(v.numel())
        if v.numel() == 1 and u.numel() != 1:
            v = v.expand(u.numel())
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        self.batch_num_edges = [self._src.numel()]

    # --- queries ------------------------------------------------------
    def nodes(self):
        return torch.arange(self._n, dtype=torch.int64)

    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def edges(self):
        return self._src, self._dst

    # --- message passing ---------------------------------------------
    def apply_edges(self, udf):
        src = {k: v[self._src] for k, v in self.ndata.items() if torch.is_tensor(v) and v.is_floating_point()}
        dst = {k: v[self._dst] for k, v in self.ndata.items() if torch.is_tensor(v) and v.is_floating_point()}
        self.edata.update(udf(_EdgeBatch(src, dst, self.edata)))

    def update_all(self, msg, red):
        assert red.kind == "sum" and red.msg == msg.out
        m = self.ndata[msg.a][self._src]
        if msg.kind == "src_mul_edge":
            m = m * self.edata[msg.b]
        else:
            assert msg.kind == "copy_src"
        out = torch.zeros((self._n, *m.shape[1:]), dtype=m.dtype)
        self.ndata[red.out] = out.index_add(0, self._dst, m)


def batch(graphs):
    """Disjoint union; node and edge ids of graph k are shifted by the totals of graphs < k."""
    bg = DGLGraph()
    bg.batch_num_nodes = [g._n for g in graphs]
    bg.batch_num_edges = [int(g._src.numel()) for g in graphs]
    off, srcs, dsts = 0, [], []
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g._n
    bg._n = off
    bg._src = torch.cat(srcs) if srcs else bg._src
    bg._dst = torch.cat(dsts) if dsts else bg._dst
    keys = graphs[0].ndata.keys() if graphs else []
    for k in keys:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    return bg


def _graph_ids(g):
    return torch.repeat_interleave(torch.arange(len(g.batch_num_nodes)), torch.tensor(g.batch_num_nodes))


def _w(g, weight, feat):
    # DGL 0.4.0 `_sum_on/_mean_on`: the weight is reshaped to (-1, 1, ..., 1) so that it broadcasts over the
    # feature dims (this is what lets ConcatReadout pass a rank-1 mask, model_zoo.py:251-256).
    return g.ndata[weight].reshape((-1,) + (1,) * (feat.dim() - 1))


def sum_nodes(g, feat, weight=None):
    x = g.ndata[feat]
    if weight is not None:
        x = x * _w(g, weight, x)
    out = torch.zeros((len(g.batch_num_nodes), *x.shape[1:]), dtype=x.dtype)
    return out.index_add(0, _graph_ids(g), x)


def mean_nodes(g, feat, weight=None):
    s = sum_nodes(g, feat, weight)
    if weight is None:
        den = torch.tensor(g.batch_num_nodes, dtype=s.dtype).view(-1, *([1] * (s.dim() - 1)))
    else:
        w = _w(g, weight, g.ndata[feat])
        den = torch.zeros((len(g.batch_num_nodes), *w.shape[1:]), dtype=w.dtype).index_add(0, _graph_ids(g), w)
    return s / den
