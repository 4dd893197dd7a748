"""Batched egonet graphs without DGL.

Mirrors the handful of DGL 0.4.0 graph calls the reference uses to build and consume egonet batches
(data_loader/dataset.py:429-435, data_loader/data_loaders.py:25, model/model.py:83-84,
model/model_zoo.py:130,157,163,212,241,249) on top of plain torch tensors, and owns the DEVICE structure the
CUDA kernels need: int32 CSR by destination and by source, graph offsets, positions.

Two ways to get a batch:
  * DGLGraph() + add_nodes/add_edges + batch([...])  -- the reference's per-egonet construction, any edge list;
    CSRs are built on the GPU by tx_build_csr_by_dst / tx_build_csr_by_src.
  * EgonetBatch.from_counts(n_gp, n_sib)             -- star egonets in the dataset.py:404-437 layout given only the
    per-egonet counts; structure is generated on the GPU in closed form (tx_star_batch_structure), so a step
    ships 4 small int32 vectors instead of the edge list.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import os

import numpy as np
import torch

from . import _lib


STAR_CHUNK = int(os.environ.get("TAXO_STAR_CHUNK", "4"))   # siblings per work item of the star-specialised forward kernel
# siblings per work item of the star backward: its chunks of one egonet meet in a partial-row combine (a fence + an atomic per chunk), so
# larger chunks pay there while the forward prefers small ones for its tail (measured L0: backward 0.200 -> 0.192 ms at 8, forward
# 0.182 -> 0.192 ms at 8)
STAR_BWD_CHUNK = int(os.environ.get("TAXO_STAR_BWD_CHUNK", "8"))
STAR_MAX_CHUNKS = 128    # = tx_gat_star_max_chunks()


class GraphStructure:
    """Device-resident structure of one batched graph (all int32)."""

    __slots__ = ("device", "n", "e", "g", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_slot", "node_off",
                 "pos", "src", "dst", "max_nodes", "max_out_deg", "max_in_deg", "_norm", "is_star", "_bwd_tiles", "_dh_bound", "star", "star_bwd", "counts")

    def __init__(self, device):
        self.device = device
        self._norm = None
        self._bwd_tiles = {}
        self._dh_bound = None      # (data_ptr of the readout's d(h), device bound of max|d(h)|): hand-over to the output layer's backward
        self.max_out_deg = 0       # host-side upper bound of the largest out-degree (bounds |dft| for the fp16-split GEMM operands)
        self.max_in_deg = 0        # ... and of the largest in-degree (bounds a GCN layer's output)
        self.pos = None
        self.src = self.dst = None
        self.is_star = False
        self.star = None           # (task records, n_tasks, chunk) of an EgonetBatch on the device (tx_gat_star_fwd)
        self.star_bwd = None       # the same with the backward's chunk size (tx_gat_star_bwd)
        self.counts = None         # (n_gp, n_sib, node_off, edge_off) device vectors of an EgonetBatch (tx_gat_star_bwd)

    def bwd_tiles(self, dim: int) -> torch.Tensor:
        """Tile table of the TMA-staged fused GAT backward (tx_gat_bwd_tiles) for per-head width `dim`; built once per batch."""
        lib = _lib.load()
        rows = int(lib.tx_gat_bwd_tile_rows(dim))
        t = self._bwd_tiles.get(rows)
        if t is None:
            nt = int(lib.tx_gat_bwd_num_tiles(self.n, dim))
            t = torch.empty((nt + 1) * 4, dtype=torch.int32, device=self.device)
            _lib.check(lib.tx_gat_bwd_tiles(_lib.ptr(self.node_off), self.g, self.n, _lib.ptr(self.in_ptr), _lib.ptr(self.out_ptr), dim,
                                            _lib.ptr(t), _lib.current_stream()), "tx_gat_bwd_tiles")
            self._bwd_tiles[rows] = t
        return t

    def gcn_norm(self) -> torch.Tensor:
        """in_degree ** -0.5 with inf -> 0 (model_zoo.py:157-161), fp32 [N]."""
        if self._norm is None:
            lib = _lib.load()
            norm = torch.empty(self.n, dtype=torch.float32, device=self.device)
            _lib.check(lib.tx_gcn_norm(_lib.ptr(self.in_ptr), self.n, _lib.ptr(norm), _lib.current_stream()), "tx_gcn_norm")
            self._norm = norm
        return self._norm


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TaxoLibraryError(
            "taxoexpan_b200 runs the propagation/readout path only on a CUDA device (sm_100a); "
            f"got device '{device}'. There is no CPU fallback.")
    return device


def _build_structure_from_edges(src: torch.Tensor, dst: torch.Tensor, n: int, node_off: torch.Tensor, device) -> GraphStructure:
    lib = _lib.load()
    device = _require_cuda(device)
    e = int(src.numel())
    if n >= 2 ** 31 - 1 or e >= 2 ** 31 - 1:
        raise ValueError("graph too large for int32 indices")
    st = GraphStructure(device)
    st.n, st.e, st.g = n, e, int(node_off.numel()) - 1
    with torch.cuda.device(device):
        s32 = src.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
        d32 = dst.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
        st.src, st.dst = s32, d32
        st.node_off = node_off.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
        i32 = dict(dtype=torch.int32, device=device)
        st.in_ptr, st.in_src, st.in_eid = torch.empty(n + 1, **i32), torch.empty(e, **i32), torch.empty(e, **i32)
        st.out_ptr, st.out_dst, st.out_slot = torch.empty(n + 1, **i32), torch.empty(e, **i32), torch.empty(e, **i32)
        slot_of_eid = torch.empty(e, **i32)
        ws = torch.empty(n + 1 + e, **i32)
        stream = _lib.current_stream()
        _lib.check(lib.tx_build_csr_by_dst(_lib.ptr(s32), _lib.ptr(d32), n, e, _lib.ptr(st.in_ptr), _lib.ptr(st.in_src),
                                           _lib.ptr(st.in_eid), _lib.ptr(slot_of_eid), _lib.ptr(ws), stream),
                   "tx_build_csr_by_dst")
        _lib.check(lib.tx_build_csr_by_src(_lib.ptr(s32), _lib.ptr(d32), _lib.ptr(slot_of_eid), n, e, _lib.ptr(st.out_ptr),
                                           _lib.ptr(st.out_dst), _lib.ptr(st.out_slot), _lib.ptr(ws), stream),
                   "tx_build_csr_by_src")
    return st


class DGLGraph:
    """Drop-in for the subset of `dgl.DGLGraph` / `dgl.BatchedDGLGraph` the reference touches."""

    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.int64)
        self._dst = torch.zeros(0, dtype=torch.int64)
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self.batch_num_nodes: List[int] = [0]
        self.batch_num_edges: List[int] = [0]
        self._structure: Dict[torch.device, GraphStructure] = {}

    # ---- construction (dataset.py:429-435) ----
    def add_nodes(self, num, data=None):
        if len(self.batch_num_nodes) != 1:
            raise RuntimeError("cannot add nodes to a batched graph")
        self._n += int(num)
        self.batch_num_nodes = [self._n]
        if data:
            for k, v in data.items():
                self.ndata[k] = v if k not in self.ndata else torch.cat([self.ndata[k], v], 0)
        self._structure.clear()

    def add_edges(self, u, v):
        u = torch.as_tensor(u, dtype=torch.int64).reshape(-1)
        v = torch.as_tensor(v, dtype=torch.int64).reshape(-1)
        if u.numel() == 1 and v.numel() != 1:
            u = u.expand(v.numel())
        if v.numel() == 1 and u.numel() != 1:
            v = v.expand(u.numel())
        if u.numel() != v.numel():
            raise ValueError("add_edges: u and v must have the same length (or one of them length 1)")
        if u.numel() and (int(u.max()) >= self._n or int(v.max()) >= self._n or int(u.min()) < 0 or int(v.min()) < 0):
            raise ValueError("add_edges: node id out of range")
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        self.batch_num_edges = [int(self._src.numel())]
        self._structure.clear()

    # ---- queries ----
    def nodes(self):
        return torch.arange(self._n, dtype=torch.int64)

    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    @property
    def batch_size(self):
        return len(self.batch_num_nodes)

    def node_offsets(self) -> torch.Tensor:
        off = torch.zeros(len(self.batch_num_nodes) + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(torch.as_tensor(self.batch_num_nodes, dtype=torch.int64), 0)
        return off

    # ---- device structure ----
    def structure(self, device) -> GraphStructure:
        device = _require_cuda(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        st = self._structure.get(device)
        if st is None:
            st = _build_structure_from_edges(self._src, self._dst, self._n, self.node_offsets(), device)
            st.max_nodes = max(self.batch_num_nodes) if self.batch_num_nodes else 0
            st.max_out_deg = int(torch.bincount(self._src.reshape(-1).to(torch.int64)).max()) if self._src.numel() else 0
            st.max_in_deg = int(torch.bincount(self._dst.reshape(-1).to(torch.int64)).max()) if self._dst.numel() else 0
            self._structure[device] = st
        return st


def batch(graphs: Sequence[DGLGraph]) -> DGLGraph:
    """`dgl.batch` (data_loaders.py:25): disjoint union, ids of graph k shifted by the totals of graphs < k,
    node features concatenated in list order."""
    bg = DGLGraph()
    bg.batch_num_nodes = [g._n for g in graphs]
    bg.batch_num_edges = [int(g._src.numel()) for g in graphs]
    offs = np.concatenate([[0], np.cumsum(bg.batch_num_nodes)]).astype(np.int64) if graphs else np.zeros(1, np.int64)
    bg._n = int(offs[-1])
    if graphs:
        bg._src = torch.cat([g._src + int(o) for g, o in zip(graphs, offs[:-1])])
        bg._dst = torch.cat([g._dst + int(o) for g, o in zip(graphs, offs[:-1])])
        for k in graphs[0].ndata:
            bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    return bg


class EgonetBatch(DGLGraph):
    """A batch of star egonets described only by per-egonet (n_gp, n_sib) counts.

    Same surface as a batched DGLGraph (ndata with 'pos', batch_num_nodes, in_degrees, ...). The edge list in the
    reference's edge-id order is materialised on the host lazily (only if somebody asks for .edges());
    the device structure comes from tx_star_batch_structure.
    """

    def __init__(self, n_gp, n_sib, ndata: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.n_gp = np.ascontiguousarray(np.asarray(n_gp, dtype=np.int32))
        self.n_sib = np.ascontiguousarray(np.asarray(n_sib, dtype=np.int32))
        if self.n_gp.shape != self.n_sib.shape or self.n_gp.ndim != 1:
            raise ValueError("n_gp and n_sib must be 1-D arrays of the same length")
        if (self.n_gp < 0).any() or (self.n_sib < 0).any():
            raise ValueError("negative egonet counts")
        # Host work per batch is a handful of reductions over the counts: totals for the allocations and the launch sizes.  Offsets and
        # the work-item tables of the star kernels are built on the GPU by tx_star_batch_plan from the two count vectors (they were
        # two cumsums and two row-repeats per table in numpy: 0.5 ms per 8192-egonet batch on the path of every step).
        g = self.n_gp.shape[0]
        n = self.n_gp.astype(np.int64) + 1 + self.n_sib
        self._n = int(n.sum()) if g else 0
        self._e = 2 * self._n - g
        if self._e >= 2 ** 31 - 1:
            raise ValueError("batch too large for int32 indices")
        self._n_per_graph = n                 # batch_num_nodes / batch_num_edges lists are materialised on first use only
        self._bnn = self._bne = None          # (two 8192-element tolist() calls were 0.2 ms of every freshly built batch)
        self._offs = None                     # host (node_off, edge_off), materialised on first use only
        self._edges_built = False
        self._max_nodes = int(n.max()) if g else 0
        # work items of the star-specialised kernels (tx_gat_star_fwd / tx_gat_star_bwd): one 16-byte record {node_off, edge_off,
        # n_gp | chunk << 24, n_sib} per (egonet, chunk of C siblings), C = STAR_CHUNK for the forward and STAR_BWD_CHUNK for the backward;
        # none when the batch exceeds the encoding (the general fused kernels take over)
        s_max = int(self.n_sib.max()) if g else 0
        ok = g > 0 and int(self.n_gp.max()) < (1 << 24) and (s_max + STAR_CHUNK - 1) // STAR_CHUNK <= STAR_MAX_CHUNKS

        def count(chunk):
            return int(np.maximum((self.n_sib + (chunk - 1)) // chunk, 1).sum(dtype=np.int64))
        self._n_tasks = count(STAR_CHUNK) if ok else 0
        self._n_tasks_bwd = (self._n_tasks if STAR_BWD_CHUNK == STAR_CHUNK else count(STAR_BWD_CHUNK)) if ok else 0
        packed = np.empty(2 * g, dtype=np.int32)          # the staging buffer of a step: [n_gp | n_sib]
        packed[:g] = self.n_gp
        packed[g:] = self.n_sib
        self._packed = torch.from_numpy(packed)
        self._g = g
        if ndata:
            self.ndata.update(ndata)
        if "pos" not in self.ndata:
            self.ndata["pos"] = _LazyPos(self)

    def _host_offsets(self):
        if self._offs is None:
            n = self._n_per_graph
            node_off = np.zeros(n.shape[0] + 1, dtype=np.int64)
            np.cumsum(n, out=node_off[1:])
            edge_off = np.zeros(n.shape[0] + 1, dtype=np.int64)
            np.cumsum(2 * n - 1, out=edge_off[1:])
            self._offs = (node_off, edge_off)
        return self._offs

    @property
    def _node_off(self):
        return self._host_offsets()[0]

    @property
    def _edge_off(self):
        return self._host_offsets()[1]

    @property
    def batch_num_nodes(self):
        if self._bnn is None:
            self._bnn = self._n_per_graph.tolist() if hasattr(self, "_n_per_graph") else [0]
        return self._bnn

    @batch_num_nodes.setter
    def batch_num_nodes(self, v):
        self._bnn = v

    @property
    def batch_num_edges(self):
        if self._bne is None:
            self._bne = (2 * self._n_per_graph - 1).tolist() if hasattr(self, "_n_per_graph") else [0]
        return self._bne

    @batch_num_edges.setter
    def batch_num_edges(self, v):
        self._bne = v

    @classmethod
    def from_counts(cls, n_gp, n_sib, ndata=None) -> "EgonetBatch":
        return cls(n_gp, n_sib, ndata)

    def pin_memory(self):
        if torch.cuda.is_available() and not self._packed.is_pinned():
            self._packed = self._packed.pin_memory()
        return self

    def stage(self, device):
        """Enqueue the host->device copy of the 4 count vectors on the CURRENT stream (e.g. a copy stream of a prefetching
        loader); structure() then only launches the closed-form structure kernel."""
        device = _require_cuda(device)
        self._staged = self._packed.to(device, non_blocking=True)
        return self

    def _build_edges(self):
        if not self._edges_built:
            from .synth import EgonetShapes, star_batch_arrays
            src, dst, _, _, _ = star_batch_arrays(EgonetShapes(self.n_gp.astype(np.int64), self.n_sib.astype(np.int64)))
            self._src, self._dst = torch.from_numpy(src), torch.from_numpy(dst)
            self._edges_built = True

    def edges(self):
        self._build_edges()
        return self._src, self._dst

    def number_of_edges(self):
        return self._e

    def in_degrees(self):
        # gp: 1 (self loop), anchor: n_gp + 1, sibling: 2
        n = self._n_per_graph
        gid = np.repeat(np.arange(self._g), n)
        local = np.arange(self._n) - self._node_off[gid]
        a = self.n_gp.astype(np.int64)[gid]
        return torch.from_numpy(np.where(local < a, 1, np.where(local == a, a + 1, 2)))

    def host_pos(self) -> torch.Tensor:
        n = self._n_per_graph
        gid = np.repeat(np.arange(self._g), n)
        local = np.arange(self._n) - self._node_off[gid]
        a = self.n_gp.astype(np.int64)[gid]
        return torch.from_numpy(np.where(local < a, 0, np.where(local == a, 1, 2)).astype(np.int64))

    def node_offsets(self) -> torch.Tensor:
        return torch.from_numpy(self._node_off)

    def add_nodes(self, *a, **k):
        raise RuntimeError("EgonetBatch is immutable")

    add_edges = add_nodes

    def structure(self, device) -> GraphStructure:
        device = _require_cuda(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        st = self._structure.get(device)
        if st is not None:
            return st
        lib = _lib.load()
        st = GraphStructure(device)
        st.n, st.e, st.g = self._n, self._e, self._g
        st.max_nodes = self._max_nodes
        st.max_out_deg = self._max_nodes          # the anchor: every sibling + its self-loop (<= graph size)
        st.max_in_deg = self._max_nodes           # the anchor: every grand-parent + its self-loop
        st.is_star = True
        g = self._g
        with torch.cuda.device(device):
            packed = getattr(self, "_staged", None)
            if packed is None or packed.device != device:
                packed = self._packed.to(device, non_blocking=True)
            n_gp, n_sib = packed[:g], packed[g:2 * g]
            i32 = dict(dtype=torch.int32, device=device)
            # one allocation for everything tx_star_batch_plan writes: node_off | edge_off | forward records | backward records
            o1 = (g + 1 + 3) // 4 * 4
            o2 = 2 * o1
            two = STAR_BWD_CHUNK != STAR_CHUNK
            o3 = o2 + 4 * self._n_tasks
            plan = torch.empty(o3 + (4 * self._n_tasks_bwd if two else 0), **i32)
            node_off, edge_off = plan[:g + 1], plan[o1:o1 + g + 1]
            t_f = plan[o2:o3] if self._n_tasks else None
            t_b = (plan[o3:o3 + 4 * self._n_tasks_bwd] if two else t_f) if self._n_tasks else None
            if g:
                _lib.check(lib.tx_star_batch_plan(_lib.ptr(n_gp), _lib.ptr(n_sib), g, STAR_CHUNK, STAR_BWD_CHUNK, _lib.ptr(node_off),
                                                  _lib.ptr(edge_off), _lib.ptr(t_f), _lib.ptr(t_b) if two else None,
                                                  _lib.current_stream()), "tx_star_batch_plan")
            else:
                plan.zero_()
            st.node_off = node_off
            st.counts = (n_gp, n_sib, node_off, edge_off)
            if self._n_tasks:
                st.star = (t_f, self._n_tasks, STAR_CHUNK)
                st.star_bwd = (t_b, self._n_tasks_bwd, STAR_BWD_CHUNK)
            st.pos = torch.empty(st.n, **i32)
            st.in_ptr, st.in_src, st.in_eid = torch.empty(st.n + 1, **i32), torch.empty(st.e, **i32), torch.empty(st.e, **i32)
            st.out_ptr, st.out_dst, st.out_slot = torch.empty(st.n + 1, **i32), torch.empty(st.e, **i32), torch.empty(st.e, **i32)
            if g == 0:
                st.in_ptr.zero_()
                st.out_ptr.zero_()
            _lib.check(lib.tx_star_batch_structure(_lib.ptr(n_gp), _lib.ptr(n_sib), _lib.ptr(node_off), _lib.ptr(edge_off), g,
                                                   _lib.ptr(st.pos), None, None, _lib.ptr(st.in_ptr), _lib.ptr(st.in_src),
                                                   _lib.ptr(st.in_eid), _lib.ptr(st.out_ptr), _lib.ptr(st.out_dst),
                                                   _lib.ptr(st.out_slot), _lib.current_stream()),
                       "tx_star_batch_structure")
        self._structure[device] = st
        return st


class _LazyPos:
    """Placeholder for ndata['pos'] of an EgonetBatch: resolves to the device int32 positions written by
    tx_star_batch_structure when moved with .to(device) (as model/model.py:83 and model_zoo.py:212 do), or to
    the host int64 tensor when used on the CPU."""

    def __init__(self, batch: EgonetBatch):
        self._batch = batch

    def to(self, device, *a, **k):
        device = torch.device(device)
        if device.type == "cuda":
            return self._batch.structure(device).pos
        return self._batch.host_pos()

    def cpu(self):
        return self._batch.host_pos()

    def __getattr__(self, name):
        return getattr(self._batch.host_pos(), name)


def as_int32_pos(pos, device) -> torch.Tensor:
    """Positions as a contiguous int32 device tensor (accepts int64 host/device tensors like the reference's)."""
    if isinstance(pos, _LazyPos):
        return pos.to(device)
    if pos.dtype == torch.int32 and pos.device == device and pos.is_contiguous():
        return pos
    return pos.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
