"""Worker of tests/test_dist_gpu.py (launched by torch.distributed.run, one rank per GPU, NCCL).

Every rank: full replica of the MAG-CS PGAT+WMR+LBM model, its shard of the queries (taxoexpan_b200.dist.shard_queries), forward +
InfoNCE + backward on the CUDA kernels, FlatGradBucket.all_reduce().  Rank 0 then runs the UNSHARDED batch on its own GPU and
compares: the loss reduction is a sum (reference model/loss.py:57 on the reshape of trainer/trainer.py:52-56), so the summed shard
gradients must equal the single-GPU gradient up to fp32 summation order."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import taxoexpan_b200 as tx  # noqa: E402
from taxoexpan_b200.dist import FlatGradBucket, shard_queries  # noqa: E402

CFG = dict(in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1])
N_Q, NEG = 32, 31


def build(dev, drop=0.0):
    torch.manual_seed(1234)                     # same initial replica on every rank
    m = tx.TaxoExpan("PGAT", "WMR", "LBM", feat_drop=drop, attn_drop=drop, hidden_drop=drop, out_drop=drop, **CFG)
    return m.to(dev).train()


def grads_of(model):
    """all gradients in registration order (the bucket re-lays its buffer out in backward order after the first step)"""
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()])


def step(model, bucket, shapes, x, qf, q0, q1, dev):
    per = 1 + NEG
    g0, g1 = q0 * per, q1 * per
    n0, n1 = int(shapes.num_nodes[:g0].sum()), int(shapes.num_nodes[:g1].sum())
    g = tx.EgonetBatch.from_counts(shapes.n_gp[g0:g1], shapes.n_sib[g0:g1])
    bucket.zero_()
    scores = model(g, x[n0:n1].to(dev), qf[g0:g1].to(dev))
    loss = tx.info_nce_loss(scores.reshape(q1 - q0, -1), torch.zeros(q1 - q0, dtype=torch.long, device=dev))
    loss.backward()
    return loss


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    shapes = tx.synth.sample_shapes(N_Q, NEG, "mag-cs", seed=77)
    x = torch.from_numpy(tx.synth.unit_rows(shapes.total_nodes, CFG["in_dim"], seed=5))
    qf = torch.from_numpy(tx.synth.unit_rows(shapes.num_graphs, CFG["in_dim"], seed=6))
    nodes_per_query = shapes.num_nodes.reshape(N_Q, 1 + NEG).sum(1)
    q0, q1 = shard_queries(nodes_per_query, world)[rank]
    res = {}
    for overlap in (True, False):
        model = build(dev)
        bucket = FlatGradBucket(model.parameters(), overlap=overlap)
        loss = step(model, bucket, shapes, x, qf, q0, q1, dev)
        bucket.all_reduce()
        flat = grads_of(model)
        loss_sum = loss.detach().clone()
        dist.all_reduce(loss_sum)
        # the reference trainer's zeroing (optimizer.zero_grad(), set_to_none=True) must give the same reduced buffer
        for p in model.parameters():
            p.grad = None
        scores_loss = None
        per = 1 + NEG
        g0, g1 = q0 * per, q1 * per
        n0, n1 = int(shapes.num_nodes[:g0].sum()), int(shapes.num_nodes[:g1].sum())
        g = tx.EgonetBatch.from_counts(shapes.n_gp[g0:g1], shapes.n_sib[g0:g1])
        bucket._reset_step()
        s = model(g, x[n0:n1].to(dev), qf[g0:g1].to(dev))
        tx.info_nce_loss(s.reshape(q1 - q0, -1), torch.zeros(q1 - q0, dtype=torch.long, device=dev)).backward()
        bucket.all_reduce()
        flat_none = grads_of(model)
        torch.cuda.synchronize()
        if rank == 0:
            ref_model = build(dev)
            ref_bucket = FlatGradBucket(ref_model.parameters(), overlap=False, group=None, gate=False)
            ref_bucket._dist_active = lambda: False                       # single-GPU run: no exchange
            ref_loss = step(ref_model, ref_bucket, shapes, x, qf, 0, N_Q, dev)
            ref = grads_of(ref_model)
            torch.cuda.synchronize()
            scale = max(1.0, float(ref.abs().max()))
            res[f"overlap={overlap}"] = {
                "max_abs_diff": float((flat - ref).abs().max()), "max_abs_diff_set_to_none": float((flat_none - ref).abs().max()),
                "scale": scale, "grad_max": float(ref.abs().max()), "loss_sharded": float(loss_sum), "loss_single": float(ref_loss),
                "numel": int(ref.numel()), "world": world, "segments": [list(r) for r in bucket._seg_range],
                "gated_launches": int(bucket.gated_launches)}
        bucket.close()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
