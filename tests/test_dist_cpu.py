"""CPU, world_size = 2 over gloo: the N > 1 path of the hot path is 'shard by query group + ONE all-reduce(sum) of the flat
gradient'. Checked with the CPU oracle model: summed shard gradients == unsharded gradients (loss reduction is a sum)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import taxo_oracle as orc
from taxoexpan_b200 import synth
from taxoexpan_b200.dist import FlatGradBucket, shard_queries

CFG = dict(propagation_method="PGAT", readout_method="WMR", matching_method="LBM", in_dim=24, hidden_dim=16, out_dim=16,
           pos_dim=4, num_layers=1, heads=[2, 1])
N_Q, NEG = 8, 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleModule(torch.nn.Module):
    def __init__(self, cfg, params):
        super().__init__()
        self.cfg = cfg
        self.names = sorted(params)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(params[k].clone()) for k in self.names])

    def loss(self, og, x, qf, n_q):
        p = dict(zip(self.names, self.ps))
        scores, _, _ = orc.taxoexpan_forward(self.cfg, og, x, qf, p)
        return orc.info_nce_step_loss(scores, n_q)


def _data():
    shapes = synth.sample_shapes(N_Q, NEG, "mag-cs", seed=21)
    per = 1 + NEG
    nodes = shapes.num_nodes.reshape(N_Q, per)
    x = torch.from_numpy(synth.unit_rows(shapes.total_nodes, CFG["in_dim"], seed=5))
    qf = torch.from_numpy(synth.unit_rows(shapes.num_graphs, CFG["in_dim"], seed=6))
    return shapes, nodes, x, qf


def _shard(shapes, nodes, x, qf, q0, q1):
    per = 1 + NEG
    g0, g1 = q0 * per, q1 * per
    n0 = int(shapes.num_nodes[:g0].sum())
    n1 = int(shapes.num_nodes[:g1].sum())
    og = orc.batch_star_egonets(shapes.n_gp[g0:g1], shapes.n_sib[g0:g1])
    return og, x[n0:n1], qf[g0:g1], q1 - q0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = orc.OracleConfig(**CFG)
    model = OracleModule(cfg, orc.init_model_params(cfg, seed=9))
    bucket = FlatGradBucket(model.parameters())
    shapes, nodes, x, qf = _data()
    q0, q1 = shard_queries(nodes.sum(1), world)[rank]
    og, xs, qs, nq = _shard(shapes, nodes, x, qf, q0, q1)
    bucket.zero_()
    model.loss(og, xs, qs, nq).backward()
    bucket.all_reduce()                       # the single exchange of the path
    if rank == 0:
        torch.save(torch.cat([p.grad.reshape(-1) for p in model.parameters()]), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_unsharded(tmp_path):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = orc.OracleConfig(**CFG)
    model = OracleModule(cfg, orc.init_model_params(cfg, seed=9))
    bucket = FlatGradBucket(model.parameters())
    shapes, nodes, x, qf = _data()
    og, xs, qs, nq = _shard(shapes, nodes, x, qf, 0, N_Q)
    model.loss(og, xs, qs, nq).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))


def test_shard_queries_balances_nodes_and_keeps_groups_whole():
    rng = np.random.default_rng(0)
    nodes = rng.integers(32, 400, size=64)
    for world in (1, 2, 4, 8):
        sh = shard_queries(nodes, world)
        assert sh[0][0] == 0 and sh[-1][1] == 64
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:])) and all(e > b for b, e in sh)
        loads = [int(nodes[b:e].sum()) for b, e in sh]
        assert max(loads) <= 1.35 * (sum(loads) / world) + 400
    import pytest
    with pytest.raises(ValueError):
        shard_queries(nodes[:3], 4)


def test_flat_bucket_views_alias_the_buffer():
    lin = torch.nn.Linear(3, 2)
    b = FlatGradBucket(lin.parameters())
    assert b.flat.numel() == 8
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.flat[:6].view(2, 3), lin.weight.grad) and float(b.flat.abs().sum()) > 0
    b.zero_()        # write-through buckets drop the gradients (the sinks are armed; all_reduce() zero-fills what receives none) ...
    assert lin.weight.grad is None and lin.bias.grad is None
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.flat[:6].view(2, 3), lin.weight.grad) and torch.equal(lin.weight.grad, torch.ones(2, 3))
    assert lin.weight.grad.data_ptr() == b.flat.data_ptr()
    b.zero_()
    lin.weight.sum().backward()                      # the bias receives no gradient this step
    assert torch.equal(b.all_reduce(), torch.tensor([1.0] * 6 + [0.0] * 2))
    b.close()
    b2 = FlatGradBucket(lin.parameters(), write_through=False)     # ... the classic protocol zero-fills in place
    lin(torch.ones(1, 3)).sum().backward()
    b2.zero_()
    assert float(lin.weight.grad.abs().sum()) == 0.0 and lin.weight.grad.data_ptr() == b2.flat.data_ptr()
    b2.close()


def test_flat_bucket_survives_zero_grad_set_to_none():
    """The reference trainer calls optimizer.zero_grad() (trainer.py:50; torch's default set_to_none=True drops the views): the bucket
    must still end up holding the real gradient (it used to all-reduce zeros)."""
    lin = torch.nn.Linear(3, 2)
    b = FlatGradBucket(lin.parameters())
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    for _ in range(2):
        opt.zero_grad()                              # set_to_none=True
        assert lin.weight.grad is None
        lin(torch.ones(1, 3)).sum().backward()
        flat = b.all_reduce()
        assert torch.equal(flat, torch.ones(8))
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        assert all(lo <= p.grad.data_ptr() < hi for p in lin.parameters())       # .grad is a view of the buffer again
        assert torch.equal(lin.weight.grad, torch.ones(2, 3)) and torch.equal(lin.bias.grad, torch.ones(2))
    # a parameter that received no gradient contributes zeros, not last step's values
    opt.zero_grad()
    (lin.weight.sum()).backward()
    flat = b.all_reduce()
    assert float(flat.sum()) == 6.0 and torch.equal(lin.weight.grad, torch.ones(2, 3)) and float(lin.bias.grad.abs().sum()) == 0.0
    # in-place zeroing keeps working
    b.zero_()
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.all_reduce(), torch.ones(8))


def test_flat_bucket_relayouts_in_backward_order_after_the_first_step():
    """Like DDP: after the first backward the buffer is laid out in the order the gradients become final (last finished first), so the
    segment holding the LATE layers completes - and is all-reduced - while the early layers still run their backward."""
    net = torch.nn.Sequential(torch.nn.Linear(4, 300), torch.nn.Linear(300, 200), torch.nn.Linear(200, 1))
    b = FlatGradBucket(net.parameters(), segments=2)
    x = torch.ones(2, 4)
    net(x).sum().backward()
    ref = [p.grad.clone() for p in net.parameters()]
    b.all_reduce()
    assert b._rebuilt
    # the first segment now holds the first layer (its gradients are final last), the last segment the last layers
    first_layer = {id(p) for p in net[0].parameters()}
    seg0 = {id(p) for i, p in enumerate(b.params) if b._seg_of[i] == 0}
    assert first_layer <= seg0 and id(net[2].weight) not in seg0
    assert all(torch.equal(p.grad, r) for p, r in zip(net.parameters(), ref))    # values survive the re-layout
    b.zero_()
    net(x).sum().backward()
    b.all_reduce()
    assert all(torch.equal(p.grad, r) for p, r in zip(net.parameters(), ref))


def test_flat_bucket_segments_cover_the_buffer():
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (600000, 200, 50, 1025000, 500, 150, 3, 125000)]
    b = FlatGradBucket(ps, segments=2)
    assert b._seg_range[0][0] == 0 and b._seg_range[-1][1] == b.flat.numel()
    assert all(x[1] == y[0] for x, y in zip(b._seg_range, b._seg_range[1:])) and sum(b._seg_size) == len(ps)
    assert len(b._seg_range) == 2


def test_gradient_sink_registry_rules():
    """functional.grad_sink (what the native backward calls consult before they write a parameter gradient): a fresh 1-D alias of the
    registered slice, found through ANY tensor that starts at the parameter's data pointer (e.g. nn.Bilinear's weight viewed as
    [l, r]); none when the parameter already holds a gradient (autograd has to accumulate), when the size does not match, after
    unregistering, or once the parameter is gone."""
    import gc
    from taxoexpan_b200 import functional as txf
    lin = torch.nn.Bilinear(3, 2, 1, bias=False)              # weight [1, 3, 2]
    w = lin.weight
    flat = torch.zeros(10)
    txf.register_grad_sink(w, flat[2:8].view_as(w))
    try:
        a = txf.grad_sink(w.view(3, 2), 6)
        assert a is not None and a.shape == (6,) and a.data_ptr() == flat[2:].data_ptr()
        b = txf.grad_sink(w, 6)
        assert b is not a and b.data_ptr() == a.data_ptr()    # a fresh alias every time: autograd only adopts unshared tensors
        assert txf.grad_sink(w, 5) is None                     # wrong size (a row view like weight[0][0] shares the data pointer)
        assert txf.grad_sink(torch.zeros(6), 6) is None        # unrelated tensor
        assert txf.grad_sink(None, 6) is None
        w.grad = torch.ones_like(w)
        assert txf.grad_sink(w, 6) is None                     # an existing gradient must be accumulated into, not overwritten
        w.grad = None
        txf.unregister_grad_sink(w)
        assert txf.grad_sink(w, 6) is None
        txf.register_grad_sink(w, flat[2:8].view_as(w))
        ptr_key = w.data_ptr()
        probe = torch.zeros(1)
        del lin, w, a, b
        gc.collect()
        assert ptr_key not in txf._grad_sinks or txf._grad_sinks[ptr_key][0]() is None
    finally:
        txf._grad_sinks.clear()

