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
        torch.save(bucket.flat.clone(), out)
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
    ref = bucket.flat
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
    b.zero_()
    assert float(lin.weight.grad.abs().sum()) == 0.0


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
        assert torch.equal(flat[:6].view(2, 3), torch.ones(2, 3)) and torch.equal(flat[6:], torch.ones(2))
        assert lin.weight.grad.data_ptr() == flat.data_ptr()       # .grad is the view again
    # a parameter that received no gradient contributes zeros, not last step's values
    opt.zero_grad()
    (lin.weight.sum()).backward()
    flat = b.all_reduce()
    assert torch.equal(flat[:6], torch.ones(6)) and float(flat[6:].abs().sum()) == 0.0
    # in-place zeroing keeps working
    b.zero_()
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.all_reduce()[:6], torch.ones(6))


def test_flat_bucket_segments_cover_the_buffer():
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (600000, 200, 50, 1025000, 500, 150, 3, 125000)]
    b = FlatGradBucket(ps, segments=2)
    assert b._seg_range[0][0] == 0 and b._seg_range[-1][1] == b.flat.numel()
    assert all(x[1] == y[0] for x, y in zip(b._seg_range, b._seg_range[1:])) and sum(b._seg_size) == len(ps)
    assert len(b._seg_range) == 2
