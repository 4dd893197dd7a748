"""Build recipe for oracle/_ref: the UNMODIFIED reference model code, compiled where it lies.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py --impl reference and the cpu_baseline leg; never the product path).

    python oracle/build_ref.py [--reference /root/reference]

The reference is pure Python: "compiling" it means byte-compiling the packages its hot path imports - model/ (model.py, model_zoo.py,
loss.py, metric.py) and what `from base import BaseModel` (model/model.py:7) drags in (base/, logger/, utils/) - straight from the
read-only checkout into oracle/_ref/<package>/<module>.refpyc (byte code only: no source file of the reference is copied into this
repository; the extension is not `.pyc` because snapshot tools drop those - import_reference() installs a finder for them).  oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot, where
/root/reference does not exist; there the reference's own TaxoExpan.forward + info_nce_loss + backward run on the host cores through
oracle/dgl_shim (DGL 0.4.0 cannot be installed offline).  MANIFEST.json records the sha256 of every source file compiled.
"""
import argparse
import hashlib
import importlib.abc
import importlib.machinery
import importlib.util
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("model", "base", "logger", "utils")
EXT = ".refpyc"


def build(reference="/root/reference", quiet=False):
    if not os.path.isdir(os.path.join(reference, "model")):
        return None
    manifest = {"reference": reference, "python": sys.version.split()[0], "files": {}}
    for pkg in PACKAGES:
        src_dir = os.path.join(reference, pkg)
        dst_dir = os.path.join(OUT, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(".py"):
                continue
            src = os.path.join(src_dir, name)
            py_compile.compile(src, cfile=os.path.join(dst_dir, name[:-3] + EXT), dfile=f"<reference>/{pkg}/{name}", doraise=True, optimize=0)
            with open(src, "rb") as f:
                manifest["files"][f"{pkg}/{name}"] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print(f"oracle/_ref: {len(manifest['files'])} modules byte-compiled from {reference}")
    return OUT


def available():
    return os.path.exists(os.path.join(OUT, "model", "model" + EXT)) and os.path.exists(os.path.join(OUT, "MANIFEST.json"))


class _RefFinder(importlib.abc.MetaPathFinder):
    """Imports the reference's packages from their byte code under oracle/_ref."""

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] not in PACKAGES:
            return None
        base = os.path.join(OUT, *parts)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__" + EXT)
            if os.path.exists(init):
                return importlib.util.spec_from_file_location(fullname, init, loader=importlib.machinery.SourcelessFileLoader(fullname, init),
                                                              submodule_search_locations=[base])
        f = base + EXT
        if os.path.exists(f):
            return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f))
        return None


def import_reference():
    """(TaxoExpan, info_nce_loss, dgl shim module) of the compiled reference; oracle/dgl_shim joins sys.path, the byte-code finder
    sys.meta_path."""
    shim = os.path.join(HERE, "dgl_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    import dgl  # noqa: F401  (the shim)
    from model.loss import info_nce_loss
    from model.model import TaxoExpan
    return TaxoExpan, info_nce_loss, dgl


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("TAXO_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    if build(a.reference) is None:
        raise SystemExit(f"{a.reference} not found")
