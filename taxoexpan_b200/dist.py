"""Egonet-sharded data parallelism: the path shards by QUERY GROUP, replicas + one gradient all-reduce.

The reference has no working multi-device path (torch.nn.DataParallel cannot split a DGL batched graph,
base/base_trainer.py:18-19; every config sets n_gpu = 1).  Here: one process per GPU, each rank holds a full replica and a
contiguous range of queries (a query's 1 positive + `negative_size` negatives stay together so the per-query InfoNCE
soft-max of trainer/trainer.py:52-56 is rank-local), the loss reduction is a sum (model/loss.py:57), so the exact
single-GPU gradient is the SUM over ranks: the flat gradient buffer is all-reduced once per step, no other exchange.

The all-reduce is issued in (by default two) contiguous SEGMENTS of the flat buffer, each as soon as autograd has finished the
last parameter of the segment (post-accumulate-grad hooks), asynchronously on NCCL's stream: the segment holding the matching,
readout and output-layer gradients (60 % of the bytes) travels while the first layer's backward kernels still run.
"""
from __future__ import annotations

import weakref
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch


def shard_queries(nodes_per_query: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous query ranges [(begin, end)) per rank, balanced by NODE count (egonet sizes vary 1..57), never
    splitting a query group. Every rank gets at least one query when there are enough queries."""
    nodes_per_query = np.asarray(nodes_per_query, dtype=np.int64)
    q = int(nodes_per_query.shape[0])
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if q < world_size:
        raise ValueError(f"cannot shard {q} queries over {world_size} ranks")
    csum = np.concatenate([[0], np.cumsum(nodes_per_query)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        cut = int(np.searchsorted(csum, target, side="left"))
        cut = max(cut, bounds[-1] + 1)                 # at least one query per rank
        cut = min(cut, q - (world_size - r))           # leave one query for every later rank
        bounds.append(cut)
    bounds.append(q)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class FlatGradBucket:
    """All gradients of a module live in ONE flat fp32 buffer (each parameter's .grad is a view into it; 1 755 303 floats = 7.0 MB
    for the MAG-CS PGAT+WMR+LBM model), all-reduced in `segments` contiguous pieces.

    Protocol per step:   bucket.zero_()  ->  forward / loss.backward()  ->  bucket.all_reduce()  ->  optimizer.step()

    * `zero_()` (or `optimizer.zero_grad(set_to_none=False)`) clears the buffer in place.  `optimizer.zero_grad()` with torch's
      default `set_to_none=True` (what the reference trainer calls, trainer/trainer.py:50) drops the views; the bucket notices:
      the post-accumulate-grad hook of every parameter copies a gradient that autograd allocated elsewhere back into the flat
      buffer and re-binds `.grad` to its view, and `all_reduce()` re-checks every parameter (a parameter that received no gradient
      contributes zeros) - so the reduced buffer always equals the sum of the ranks' gradients, whichever way the caller zeroes.
    * with `overlap=True` (default) a segment's all-reduce starts inside backward, as soon as its last gradient is final;
      `all_reduce()` launches whatever has not started and makes the CURRENT stream wait for all of it (no host block).
    * with `write_through=True` (default) every slice is registered as the home of its parameter's gradient
      (`functional.register_grad_sink`): the native backward calls write the gradient straight into the slice and autograd adopts the
      alias as `.grad` - no add / copy kernel per parameter - PROVIDED `.grad` is None when backward runs, so `zero_()` then drops
      the gradients instead of zero-filling the buffer (a parameter that receives no gradient is zero-filled by `all_reduce()`).
      Gradients produced elsewhere (torch ops, the per-kernel paths) are copied in by the hooks as before.
    * with `gate=True` (default; CUDA, more than one segment) a segment that completes inside backward is not reduced on the spot
      - NCCL's CTAs would take SMs the persistent one-CTA-per-SM kernels of the next layer's backward count on (its star backward
      ended 0.04-0.07 ms later on 4 / 8 GPUs, as much as the overlap saved) - but from a side stream that waits for the event
      `tx_gat_layer_bwd` records right after it has launched its star backward (tx_set_after_star_bwd_event): the collective then runs
      beside the weight- and input-gradient GEMMs that follow, which leave SMs idle.  Only segments that were complete BEFORE the
      last such record use the gate (the library counts the records); everything else is reduced from the current stream.
    """

    _gate_owner = None          # the bucket whose event the library records after a layer's star backward (one per process)

    def __init__(self, params: Iterable[torch.nn.Parameter], segments: int = 2, overlap: bool = True, group=None,
                 write_through: bool = True, gate: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.group = group
        self.overlap = overlap
        self.segments = segments
        self.write_through = bool(write_through)
        self._gate_event = self._side = self._txlib = None
        if gate and segments > 1 and self.params[0].is_cuda:
            try:
                from . import _lib
                self._txlib = _lib.load()
                with torch.cuda.device(self.params[0].device):
                    self._gate_event = torch.cuda.Event()
                    self._gate_event.record()                       # creates the underlying cudaEvent_t
                    self._side = torch.cuda.Stream()
                self._txlib.tx_set_after_star_bwd_event(self._gate_event.cuda_event)
                FlatGradBucket._gate_owner = weakref.ref(self)      # the library records ONE event: the newest bucket owns it
            except Exception:                                       # no library: reduce on the spot
                self._gate_event = self._side = self._txlib = None
        self._deferred = []                                         # (segment, record count when it became ready)
        self.gated_launches = 0                                     # collectives started from the side stream so far (tests)
        self.active = True                        # False: hooks only keep the views bound (no exchange is launched)
        self._rebuilt = False
        self._fire_order = []
        self._layout(list(range(len(self.params))), None)
        self._ready = [0] * len(self._seg_range)
        self._launched = [False] * len(self._seg_range)
        self._seen = [False] * len(self.params)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _layout(self, order, old_views):
        """(Re)build the flat buffer with the parameters laid out in `order` (indices into self.params) and cut it into segments of
        about equal size, each boundary at the parameter edge closest to its target."""
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=dt)
        self._views = [None] * len(self.params)
        edges = [0]
        for i in order:
            p = self.params[i]
            v = self.flat[edges[-1]:edges[-1] + p.numel()].view_as(p)
            if old_views is not None:
                v.copy_(old_views[i])
            self._views[i] = v
            p.grad = v
            if self.write_through:
                from . import functional as txf
                txf.register_grad_sink(p, v)
            edges.append(edges[-1] + p.numel())
        n_seg = max(1, min(int(self.segments), len(order)))
        total = self.flat.numel()
        cuts = [0]
        for k in range(1, n_seg):
            target = total * k / n_seg
            j = min(range(cuts[-1] + 1, len(order) - (n_seg - 1 - k)), key=lambda i: abs(edges[i] - target), default=None)
            if j is None:
                break
            cuts.append(j)
        cuts.append(len(order))
        self._seg_of = [0] * len(self.params)
        self._seg_range, self._seg_size = [], []
        for s_, (a, b) in enumerate(zip(cuts, cuts[1:])):
            for i in order[a:b]:
                self._seg_of[i] = s_
            self._seg_range.append((edges[a], edges[b]))
            self._seg_size.append(b - a)

    # ---- internals ----
    def _owns_gate(self) -> bool:
        ref = FlatGradBucket._gate_owner
        return ref is not None and ref() is self

    def _dist_active(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _rebind(self, i: int):
        """Make parameter i's .grad the flat view again: copy a gradient autograd allocated elsewhere (after
        zero_grad(set_to_none=True)); a parameter without any gradient this step contributes zeros."""
        p, v = self.params[i], self._views[i]
        g = p.grad
        if g is None:
            v.zero_()
        elif g.data_ptr() != v.data_ptr() or g.shape != v.shape:
            v.copy_(g)
        p.grad = v

    def _make_hook(self, i: int):
        def hook(param):
            s = self._seg_of[i]
            if not self.active:
                self._rebind(i)
                return
            if self._launched[s]:
                raise RuntimeError("FlatGradBucket: backward ran again after this segment's all-reduce was launched; call "
                                   "all_reduce() / zero_() between backward passes or construct the bucket with overlap=False")
            self._rebind(i)
            if not self._seen[i]:
                self._seen[i] = True
                self._ready[s] += 1
                if not self._rebuilt:
                    self._fire_order.append(i)
            if self.overlap and self._ready[s] == self._seg_size[s]:
                if self._gate_event is not None and self._owns_gate() and self._dist_active():
                    self._launched[s] = True                         # claimed: a late gradient for it is an error, as before
                    self._deferred.append((s, int(self._txlib.tx_after_star_bwd_event_count())))
                else:
                    self._launch(s)
        return hook

    def _launch(self, s: int):
        self._launched[s] = True
        if self._dist_active():
            import torch.distributed as dist
            a, b = self._seg_range[s]
            self._works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _launch_deferred(self):
        """Reduce the segments that completed inside backward: gated on the after-star-backward event when one was recorded after they
        were complete (the collective may then start at that point of the compute stream), else ordered after the current stream."""
        if not self._deferred:
            return
        import torch.distributed as dist
        now = int(self._txlib.tx_after_star_bwd_event_count())
        cur = torch.cuda.current_stream(self.flat.device)
        for s, count in self._deferred:
            a, b = self._seg_range[s]
            if now > count and self._owns_gate():
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._gate_event)
                    self.gated_launches += 1
                    self._works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self._works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._deferred = []
        del cur

    def _reset_step(self):
        self._deferred = []
        self._ready = [0] * len(self._seg_range)
        self._launched = [False] * len(self._seg_range)
        self._seen = [False] * len(self.params)
        self._works = []

    # ---- public ----
    def zero_(self):
        """Clear the flat buffer in place and (re-)bind every .grad view: call before each backward."""
        for w in self._works:                     # a reduction still in flight must not race the clear / the next backward's writes
            w.wait()
        if self.write_through:
            for p in self.params:                 # .grad None: the gradient sinks are armed, autograd adopts what backward returns
                p.grad = None
        else:
            self.flat.zero_()
            for p, v in zip(self.params, self._views):
                p.grad = v
        self._reset_step()

    def all_reduce(self, group=None):
        """Finish the step's gradient exchange: returns the flat buffer holding the SUM over ranks (stream-ordered: the current
        stream waits for NCCL's stream, the host does not block)."""
        if group is not None:
            self.group = group
        for i, (p, v) in enumerate(zip(self.params, self._views)):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if self._launched[self._seg_of[i]]:
                    raise RuntimeError("FlatGradBucket: a gradient was replaced after its segment's all-reduce was launched")
                self._rebind(i)
        self._launch_deferred()
        for s in range(len(self._seg_range) - 1, -1, -1):
            if not self._launched[s]:
                self._launch(s)
        for w in self._works:
            w.wait()
        if not self._rebuilt and len(self._fire_order) == len(self.params):
            # like DDP after its first iteration: lay the buffer out in the order autograd finishes the gradients (last finished first), so
            # that every segment but the first is complete - and on the wire - while the layers below still run their backward
            self._layout(self._fire_order[::-1], self._views)
            self._rebuilt = True
        self._reset_step()
        return self.flat

    def __del__(self):
        # the library records into this bucket's event by raw handle: never let it outlive the torch.cuda.Event that owns the handle
        try:
            if getattr(self, "_gate_event", None) is not None and self._owns_gate():
                self._txlib.tx_set_after_star_bwd_event(None)
                FlatGradBucket._gate_owner = None
        except Exception:
            pass

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._gate_event is not None:
            if self._owns_gate():
                self._txlib.tx_set_after_star_bwd_event(None)
                FlatGradBucket._gate_owner = None
            self._gate_event = None
        if self.write_through:
            from . import functional as txf
            for p in self.params:
                txf.unregister_grad_sink(p)
