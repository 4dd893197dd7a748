"""CPU: the C-ABI library loads and exports every symbol include/taxo_b200.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import taxoexpan_b200 as tx
from taxoexpan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "taxo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from taxoexpan_b200 import build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/taxo_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes binding and header disagree"
    l = _lib.load()
    assert l.tx_abi_version() == 1
    assert l.tx_target_arch() == b"sm_100a"
    assert l.tx_row_blocks(65) == 2


def test_no_kernel_touches_global_memory_before_it_waits_for_its_predecessor():
    """Programmatic dependent launch: a kernel may be scheduled while its predecessor still runs, so nothing that reads or writes global
    memory may sit in front of its griddepcontrol.wait (SASS: ACQBULK).  ld.global.nc loads are NOT ordered by inline-asm memory clobbers -
    the build once placed `LDG.E.CONSTANT` of a predecessor-written scalar before the wait - hence the check on the disassembly of the
    library that ships (scripts/check_pdl_sass.py)."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    from taxoexpan_b200 import build
    build.build()
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import check_pdl_sass
    res = check_pdl_sass.check(_lib.LIB_PATH)
    assert len(res) >= 40, "the hot-path kernels are expected to wait on their predecessors"
    bad = {k: v for k, v in res.items() if v}
    assert not bad, bad
    names = " ".join(res)
    for must in ("gat_star_fwd_kernel", "gat_star_bwd_kernel", "gemm_tf32x3_pair_kernel", "readout_fwd_fast_kernel", "absmax_kernel"):
        assert must in names, must


def test_the_sass_checker_flags_a_load_hoisted_over_the_wait():
    """The checker itself: the hoisted non-coherent load that broke the GCN forward is reported, the GEMM prologue's shared-memory
    traffic and loads behind the wait are not."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import check_pdl_sass
    listing = '''
	Function : bad_kernel
        /*0000*/                   LDC R1, c[0x0][0x37c] ;
        /*0060*/                   LDG.E.CONSTANT R2, desc[UR8][R2.64] ;
        /*0080*/                   PREEXIT ;
        /*0090*/                   ACQBULK ;
        /*00a0*/                   LDG.E R4, desc[UR8][R4.64] ;
	Function : good_kernel
        /*0000*/                   LDC R1, c[0x0][0x37c] ;
        /*0010*/                   SYNCS.ARRIVE.TRANS64.RED RZ, [R7+URZ], R8 ;
        /*0020*/                   CCTL.IVALL ;
        /*0030*/                   LD.E.STRONG.SYS R3, desc[UR20][R2.64+0x30040] ;
        /*0080*/                   PREEXIT ;
        /*0090*/                   ACQBULK ;
        /*00a0*/              @!P0 EXIT ;
        /*00b0*/                   LDG.E.CONSTANT R2, desc[UR8][R2.64] ;
        /*00c0*/              @P1  STG.E desc[UR8][R4.64], R2 ;
	Function : predicated_store_before
        /*0010*/              @P0  STG.E desc[UR8][R4.64], R2 ;
        /*0090*/                   ACQBULK ;
	Function : no_wait_kernel
        /*0010*/                   LDG.E R4, desc[UR8][R4.64] ;
'''
    res = check_pdl_sass.check_text(listing)
    assert set(res) == {"bad_kernel", "good_kernel", "predicated_store_before"}
    assert res["bad_kernel"] == ["LDG.E.CONSTANT R2, desc[UR8][R2.64]"] and res["good_kernel"] == []
    assert len(res["predicated_store_before"]) == 1


def test_argument_validation_without_gpu():
    l = _lib.load()
    # invalid dropout rate is rejected before any CUDA call
    rc = l.tx_dropout_keep_mask(1, 0, 0, 16, 1.5, None, None)
    assert rc == -1 and b"p_drop" in l.tx_last_error()
    rc = l.tx_readout_fwd(7, None, 4, None, None, None, 1, 4, None, 4, None)
    assert rc == -1 and b"unknown kind" in l.tx_last_error()


def test_product_path_refuses_cpu_tensors():
    m = tx.TaxoExpan("PGAT", "WMR", "LBM", in_dim=8, hidden_dim=4, out_dim=4, pos_dim=4, num_layers=1, heads=[2, 1],
                     feat_drop=0.0, attn_drop=0.0, hidden_drop=0.0, out_drop=0.0)
    g = tx.EgonetBatch.from_counts([1, 0], [2, 0])
    with pytest.raises(tx.TaxoLibraryError):
        m(g, torch.randn(5, 8), torch.randn(2, 8))


def test_unknown_method_names_raise():
    with pytest.raises(ValueError):
        tx.TaxoExpan("XGAT", "WMR", "LBM", in_dim=8, hidden_dim=4, out_dim=4, pos_dim=4, num_layers=1, heads=[2, 1],
                     feat_drop=0.0, attn_drop=0.0, hidden_drop=0.0, out_drop=0.0)


def test_state_dict_keys_match_reference_layout():
    m = tx.TaxoExpan("PGAT", "WMR", "LBM", in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1],
                     feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "graph_propagate.gat_layers.0.attn_l": (1, 4, 500), "graph_propagate.gat_layers.0.attn_r": (1, 4, 500),
        "graph_propagate.gat_layers.0.fc.weight": (2000, 300),
        "graph_propagate.gat_layers.1.attn_l": (1, 1, 500), "graph_propagate.gat_layers.1.attn_r": (1, 1, 500),
        "graph_propagate.gat_layers.1.fc.weight": (500, 2050),
        "graph_propagate.prop_position_embeddings.0.weight": (3, 50),
        "graph_propagate.prop_position_embeddings.1.weight": (3, 50),
        "readout.position_weights.weight": (3, 1), "match.W.weight": (1, 500, 250)}
    assert sum(v.numel() for v in sd.values()) == 1755303
    m2 = tx.TaxoExpan("PGCN", "MR", "BIM", in_dim=300, hidden_dim=600, out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1],
                      feat_drop=0.1, attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)
    assert {k: tuple(v.shape) for k, v in m2.state_dict().items()} == {
        "graph_propagate.layers.0.weight": (350, 600), "graph_propagate.layers.0.bias": (600,),
        "graph_propagate.layers.1.weight": (650, 300), "graph_propagate.layers.1.bias": (300,),
        "graph_propagate.prop_position_embeddings.0.weight": (3, 50),
        "graph_propagate.prop_position_embeddings.1.weight": (3, 50), "match.W.weight": (1, 300, 300)}


def test_graph_host_api_follows_dgl_calls_of_dataset_py():
    """dataset.py:429-435 + data_loaders.py:25 through the DGL-free graph class."""
    from oracle import taxo_oracle as orc
    gs = []
    for n_gp, n_sib in [(2, 3), (0, 0), (1, 4)]:
        n = n_gp + 1 + n_sib
        g = tx.DGLGraph()
        g.add_nodes(n, {"x": torch.randn(n, 4), "_id": torch.arange(n), "pos": torch.tensor([0] * n_gp + [1] + [2] * n_sib)})
        g.add_edges(list(range(n_gp)), n_gp)
        g.add_edges(n_gp, list(range(n_gp + 1, n)))
        g.add_edges(g.nodes(), g.nodes())
        gs.append(g)
    bg = tx.batch(gs)
    og = orc.batch_star_egonets([2, 0, 1], [3, 0, 4])
    assert torch.equal(bg.edges()[0], og.src) and torch.equal(bg.edges()[1], og.dst)
    assert bg.batch_num_nodes == og.batch_num_nodes and torch.equal(bg.ndata["pos"], og.pos)
    assert torch.equal(bg.in_degrees(), og.in_degrees())
    eb = tx.EgonetBatch.from_counts([2, 0, 1], [3, 0, 4])
    assert torch.equal(eb.edges()[0], og.src) and torch.equal(eb.edges()[1], og.dst)
    assert torch.equal(eb.in_degrees(), og.in_degrees()) and torch.equal(eb.ndata["pos"].cpu(), og.pos)
    assert eb.batch_num_nodes == og.batch_num_nodes and eb.number_of_edges() == og.src.numel()
    assert bg.ndata.pop("pos") is not None and "pos" not in bg.ndata


def test_synthetic_shapes_statistics():
    s = tx.synth.sample_shapes(256, 31, "mag-cs")
    assert s.num_graphs == 8192
    assert s.total_edges == 2 * s.total_nodes - s.num_graphs
    assert 3.5 < s.total_nodes / s.num_graphs < 6.0 and int(s.num_nodes.max()) <= 1 + 50 + 8
    assert (s.n_sib <= 50).all() and (s.n_gp >= 1).all()


def test_star_task_counts_and_the_encoding_limit():
    """EgonetBatch sizes the work-item tables of the star kernels on the host - one record per (egonet, chunk of STAR_CHUNK resp.
    STAR_BWD_CHUNK siblings), chunk 0 for every egonet - and ships only the two count vectors; the records themselves are written on
    the GPU by tx_star_batch_plan (checked bit for bit in tests/test_gpu_parity.py::test_star_batch_plan_tables_are_bit_exact)."""
    import numpy as np
    from taxoexpan_b200 import graph as txg
    rng = np.random.default_rng(0)
    n_gp = rng.integers(0, 5, 200)
    n_sib = np.concatenate([rng.integers(0, 60, 190), [0, 1, txg.STAR_CHUNK, txg.STAR_CHUNK + 1, 4 * txg.STAR_CHUNK, 0, 0, 7, 50, 50]])
    g = tx.EgonetBatch.from_counts(n_gp, n_sib)
    assert g._n_tasks == int(np.maximum(1, -(-n_sib // txg.STAR_CHUNK)).sum())
    assert g._n_tasks_bwd == int(np.maximum(1, -(-n_sib // txg.STAR_BWD_CHUNK)).sum())
    assert g._packed.numel() == 2 * 200 and np.array_equal(g._packed.numpy()[:200], n_gp) and np.array_equal(g._packed.numpy()[200:], n_sib)
    assert g.number_of_nodes() == int((n_gp + 1 + n_sib).sum()) and g.number_of_edges() == 2 * g.number_of_nodes() - 200
    assert np.array_equal(g.node_offsets().numpy(), np.concatenate([[0], np.cumsum(n_gp + 1 + n_sib)]))
    # a batch the encoding cannot hold (more chunks than the 7-bit field) falls back to the general kernel: no records
    big = tx.EgonetBatch.from_counts([1], [txg.STAR_CHUNK * txg.STAR_MAX_CHUNKS + 1])
    assert big._n_tasks == 0 and big._n_tasks_bwd == 0


def test_matching_and_loss_have_no_cpu_path():
    """SURVEY 8 f1 surface: `info_nce_loss(output, target)` (model/loss.py:52-57) and the BIM / LBM row-dot run on the CUDA library
    only - CPU tensors raise instead of silently taking a torch fallback."""
    import torch
    out = torch.randn(4, 32)
    with pytest.raises(tx.TaxoLibraryError):
        tx.info_nce_loss(out, torch.zeros(4, dtype=torch.long))
    with pytest.raises(ValueError):
        tx.info_nce_loss(out.reshape(-1), None)
    from taxoexpan_b200 import functional as txf
    with pytest.raises(tx.TaxoLibraryError):
        txf.match_rowdot(torch.randn(4, 8), torch.randn(4, 8), True)
