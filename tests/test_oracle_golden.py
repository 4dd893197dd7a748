"""CPU: the oracle restatement must reproduce the golden vectors frozen from the UNMODIFIED reference model code
(oracle/make_golden.py). fp32 against fp32 golden, and fp64 oracle against the reference's fp64 run."""
import numpy as np
import pytest
import torch

from oracle import taxo_oracle as orc
from tests._golden import CASES, FULL_CASES, compare_to_fixture, load_case, sub


def _run(cfg, og, x, qf, params, dtype, n_q):
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    h = x.to(dtype).clone().requires_grad_(True)
    scores, hg, node_h = orc.taxoexpan_forward(cfg, og, h, qf.to(dtype), p)
    loss = orc.info_nce_step_loss(scores, n_q)
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in p.items()}
    return scores.detach().numpy(), hg.detach().numpy(), node_h.detach().numpy(), loss.detach().numpy(), grads, h.grad.numpy()


@pytest.mark.parametrize("name", CASES + FULL_CASES)
def test_oracle_fp32_matches_reference_golden(name):
    cfg, og, x, qf, params, fx = load_case(name)
    out = _run(cfg, og, x, qf, params, torch.float32, int(fx["n_queries"][0]))
    compare_to_fixture(fx, *out, tol=1e-5, gtol=2e-5)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_reference_fp64(name):
    cfg, og, x, qf, params, fx = load_case(name)
    scores, hg, node_h, loss, grads, dh = _run(cfg, og, x, qf, params, torch.float64, int(fx["n_queries"][0]))
    np.testing.assert_allclose(scores, fx["scores_f64"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(hg[::int(fx["hg_row_step"][0])], fx["hg_f64"], rtol=1e-11, atol=1e-12)
    if "node_h_f64" in fx:
        np.testing.assert_allclose(node_h, fx["node_h_f64"], rtol=1e-11, atol=1e-12)
    else:
        np.testing.assert_allclose(sub(node_h, int(fx["big_step"][0])), fx["node_h_sub_f64"], rtol=1e-11, atol=1e-12)
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


def test_closed_form_star_backward_matches_autograd():
    """oracle/star_backward.py (the per-egonet closed form a star-specialised backward kernel implements) against torch autograd of
    the restated GAT layer, fp64, with attention dropout masks and egonets that cover roots, leaves and many grand-parents."""
    import numpy as np
    import torch
    from oracle import star_backward
    n_gp, n_sib = [1, 0, 3, 2, 0, 5], [2, 0, 0, 7, 3, 1]
    og = orc.batch_star_egonets(n_gp, n_sib)
    H, D, K = 3, 8, 6
    gen = torch.Generator().manual_seed(4)
    z = torch.randn(og.n, K, generator=gen, dtype=torch.float64)
    W = (torch.randn(H * D, K, generator=gen, dtype=torch.float64) * 0.5).requires_grad_(True)
    al = torch.randn(1, H, D, generator=gen, dtype=torch.float64).requires_grad_(True)
    ar = torch.randn(1, H, D, generator=gen, dtype=torch.float64).requires_grad_(True)
    p = 0.3
    keep = (torch.rand(og.src.shape[0], H, 1, generator=gen) > p)
    ft_holder = {}
    orig_linear = torch.nn.functional.linear

    def linear(x, w):                         # capture ft with its gradient
        y = orig_linear(x, w)
        y.retain_grad()
        ft_holder["ft"] = y
        return y
    torch.nn.functional.linear = linear
    try:
        out = orc.gat_layer(og, z, W, al, ar, H, negative_slope=0.2, attn_keep=keep, p_attn=p)
    finally:
        torch.nn.functional.linear = orig_linear
    gout = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    out.backward(gout)
    ft = ft_holder["ft"]
    keepw = (keep.double() / (1.0 - p)).reshape(-1, H).numpy()
    dft, dal, dar = star_backward.star_gat_backward(n_gp, n_sib, ft.detach().reshape(og.n, H, D).numpy(), gout.numpy(),
                                                   al.detach()[0].numpy(), ar.detach()[0].numpy(), keepw, 0.2)
    assert np.abs(dft.reshape(og.n, -1) - ft.grad.numpy()).max() <= 1e-12
    assert np.abs(dal - al.grad[0].numpy()).max() <= 1e-12
    assert np.abs(dar - ar.grad[0].numpy()).max() <= 1e-12
