"""Drop-in surface against the UNMODIFIED reference (runs only where /root/reference exists, i.e. in the build container): constructor
and forward signatures of every module on the path, `info_nce_loss`, and the state_dict keys / shapes of the headline configurations.
The reference is imported in a subprocess through oracle/dgl_shim (its `model` package name would shadow nothing here, but its
`import dgl` must not leak into this process)."""
import inspect
import json
import os
import subprocess
import sys

import pytest

import taxoexpan_b200 as tx
from taxoexpan_b200 import loss as tx_loss
from taxoexpan_b200 import model_zoo as tx_zoo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TAXO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="reference checkout not present")

CONFIGS = {
    "mag": ("PGAT", "WMR", "LBM", dict(in_dim=250, hidden_dim=500, out_dim=500, pos_dim=50, num_layers=1, heads=[4, 1], feat_drop=0.1,
                                      attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)),
    "wordnet": ("PGCN", "MR", "MLP", dict(in_dim=300, hidden_dim=600, out_dim=300, pos_dim=50, num_layers=1, heads=[4, 1], feat_drop=0.1,
                                          attn_drop=0.1, hidden_drop=0.1, out_drop=0.1)),
    "gat_cr_bim": ("GAT", "CR", "BIM", dict(in_dim=16, hidden_dim=8, out_dim=8, pos_dim=4, num_layers=2, heads=[2, 2, 1], feat_drop=0.0,
                                            attn_drop=0.0, hidden_drop=0.0, out_drop=0.0)),
}
CLASSES = ["GCNLayer", "GATLayer", "GCN", "GAT", "PGCN", "PGAT", "MeanReadout", "WeightedMeanReadout", "ConcatReadout", "MLP", "BIM", "LBM"]

PROBE = r'''
import inspect, json, sys
import model.model_zoo as zoo, model.model as mm, model.loss as loss
def sig(f):
    return [(p.name, repr(p.default) if p.default is not inspect.Parameter.empty else None) for p in inspect.signature(f).parameters.values()]
out = {"init": {c: sig(getattr(zoo, c).__init__) for c in CLASSES}, "forward": {c: sig(getattr(zoo, c).forward) for c in CLASSES},
       "taxoexpan_init": sig(mm.TaxoExpan.__init__), "taxoexpan_forward": sig(mm.TaxoExpan.forward), "info_nce": sig(loss.info_nce_loss), "state": {}}
for name, (pm, rm, mmeth, kw) in CONFIGS.items():
    m = mm.TaxoExpan(pm, rm, mmeth, **kw)
    out["state"][name] = {k: list(v.shape) for k, v in m.state_dict().items()}
print(json.dumps(out))
'''


def _reference_surface():
    code = f"CLASSES = {CLASSES!r}\nCONFIGS = {CONFIGS!r}\n" + PROBE
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "dgl_shim"), REF]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _sig(f):
    return [[p.name, repr(p.default) if p.default is not inspect.Parameter.empty else None] for p in inspect.signature(f).parameters.values()]


def test_module_signatures_and_state_dicts_match_the_reference():
    ref = _reference_surface()
    for c in CLASSES:
        ours = getattr(tx_zoo, c)
        assert _sig(ours.__init__) == ref["init"][c], f"{c}.__init__"
        # forward: same positional parameters (names may carry the reference's typos; position and count are the contract)
        assert len(_sig(ours.forward)) == len(ref["forward"][c]), f"{c}.forward arity"
    assert [p[0] for p in _sig(tx.TaxoExpan.__init__)] == [p[0] for p in ref["taxoexpan_init"]]
    assert [p[0] for p in _sig(tx.TaxoExpan.forward)] == [p[0] for p in ref["taxoexpan_forward"]]
    assert [p[0] for p in _sig(tx_loss.info_nce_loss)] == [p[0] for p in ref["info_nce"]]
    for name, (pm, rm, mmeth, kw) in CONFIGS.items():
        m = tx.TaxoExpan(pm, rm, mmeth, **kw)
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == ref["state"][name], name
