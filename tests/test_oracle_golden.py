"""CPU: the oracle restatement must reproduce the golden vectors frozen from the UNMODIFIED reference model code
(oracle/make_golden.py). fp32 against fp32 golden, and fp64 oracle against the reference's fp64 run."""
import numpy as np
import pytest
import torch

from oracle import taxo_oracle as orc
from tests._golden import CASES, compare_to_fixture, load_case, sub


def _run(cfg, og, x, qf, params, dtype, n_q):
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    h = x.to(dtype).clone().requires_grad_(True)
    scores, hg, node_h = orc.taxoexpan_forward(cfg, og, h, qf.to(dtype), p)
    loss = orc.info_nce_step_loss(scores, n_q)
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in p.items()}
    return scores.detach().numpy(), hg.detach().numpy(), node_h.detach().numpy(), loss.detach().numpy(), grads, h.grad.numpy()


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp32_matches_reference_golden(name):
    cfg, og, x, qf, params, fx = load_case(name)
    out = _run(cfg, og, x, qf, params, torch.float32, int(fx["n_queries"][0]))
    compare_to_fixture(fx, *out, tol=1e-5, gtol=2e-5)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_reference_fp64(name):
    cfg, og, x, qf, params, fx = load_case(name)
    scores, hg, node_h, loss, grads, dh = _run(cfg, og, x, qf, params, torch.float64, int(fx["n_queries"][0]))
    np.testing.assert_allclose(scores, fx["scores_f64"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(hg, fx["hg_f64"], rtol=1e-11, atol=1e-12)
    if "node_h_f64" in fx:
        np.testing.assert_allclose(node_h, fx["node_h_f64"], rtol=1e-11, atol=1e-12)
    else:
        np.testing.assert_allclose(sub(node_h), fx["node_h_sub_f64"], rtol=1e-11, atol=1e-12)
    for k, g in grads.items():
        ref = fx["grad_f64." + k] if "grad_f64." + k in fx else fx["grad_sub_f64." + k]
        got = g if "grad_f64." + k in fx else sub(g)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * max(1.0, float(fx["gradnorm." + k][0])))


def test_star_layout_matches_dataset_py():
    """dataset.py:404-437 for one egonet with 2 grand-parents and 3 siblings."""
    src, dst, pos, n = orc.star_egonet(2, 3)
    assert n == 6
    assert pos.tolist() == [0, 0, 1, 2, 2, 2]
    assert src.tolist() == [0, 1, 2, 2, 2, 0, 1, 2, 3, 4, 5]
    assert dst.tolist() == [2, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5]
    # minimum egonet: a root anchor that is also a leaf -> 1 node, 1 self-loop
    src, dst, pos, n = orc.star_egonet(0, 0)
    assert (n, src.tolist(), dst.tolist(), pos.tolist()) == (1, [0], [0], [1])


def test_vectorised_batch_equals_per_graph_construction():
    from taxoexpan_b200 import synth
    shapes = synth.sample_shapes(8, 31, "mag-cs", seed=3)
    og = orc.batch_star_egonets(shapes.n_gp, shapes.n_sib)
    src, dst, pos, noff, eoff = synth.star_batch_arrays(shapes)
    assert np.array_equal(src, og.src.numpy()) and np.array_equal(dst, og.dst.numpy())
    assert np.array_equal(pos, og.pos.numpy())
    assert np.array_equal(np.diff(noff), np.asarray(og.batch_num_nodes))
    assert np.array_equal(np.diff(eoff), np.asarray(og.batch_num_edges))
    assert shapes.total_nodes == og.n and shapes.total_edges == og.src.numel()


def test_rank_restatement_matches_its_definition():
    """oracle.ranks_from_similarities (metric.py:7-19) against the definition, ties and multiple true positions included."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(2, 40))
        s = rng.integers(0, 6, n).astype(np.float64)          # many ties
        pos = rng.choice(n, size=int(rng.integers(1, min(4, n))), replace=False).tolist()
        got = orc.ranks_from_similarities(s, pos)
        want = [1 + sum(1 for i in range(n) if i not in pos and s[i] > s[p]) for p in pos]
        assert got == want
