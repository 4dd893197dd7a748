"""DGL-free dataset and checkpoint formats.

The reference keeps its datasets as a pickle that embeds a `dgl.DGLGraph` (`data_loader/dataset.py:183-194`; unreadable without
DGL 0.4) built from three text files (`README.md:23-51`, `dataset.py:92-160`): `<name>.terms` (id \\t surface name), `<name>.taxo`
(parent id \\t child id) and `<name>.terms[.<suffix>].embed` (word2vec text format).  This module reads the text files directly
into plain arrays - node ids, edge order and the train / validation / test split exactly as `MAGDataset._load_dataset_raw`
assigns them - and stores them as one `.npz`; `TaxonomyDataset.taxonomy()` hands the `sampler.TaxonomyCSR` to the vectorised
egonet construction.  `load_reference_checkpoint` loads a reference `.pth` (`base/base_trainer.py:126-149`) into the drop-in
`TaxoExpan` without importing the reference's `parse_config` (the pickled ConfigParser is replaced by a stub while unpickling).
"""
from __future__ import annotations

import io
import os
import pickle
import random
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .sampler import TaxonomyCSR


@dataclass
class TaxonomyDataset:
    name: str
    vocab: List[str]                 # "<surface name>@@@<node id>" (dataset.py:143)
    tx_ids: List[str]                # original taxon ids, by node id
    parents: np.ndarray              # edge list in the reference's DGL edge-id order (grouped by parent node, dataset.py:146-150)
    children: np.ndarray
    features: np.ndarray             # fp32 [V, d]
    train_node_ids: np.ndarray
    validation_node_ids: np.ndarray
    test_node_ids: np.ndarray

    def taxonomy(self, node_subset: Optional[np.ndarray] = None) -> TaxonomyCSR:
        """CSR of the full taxonomy or of the subgraph induced by `node_subset` (node ids are kept: the training graph of
        dataset.py:234 is the subgraph of the train nodes)."""
        p, c = self.parents, self.children
        if node_subset is not None:
            keep = np.zeros(len(self.vocab), dtype=bool)
            keep[np.asarray(node_subset, dtype=np.int64)] = True
            m = keep[p] & keep[c]
            p, c = p[m], c[m]
        return TaxonomyCSR.from_edges(p, c, len(self.vocab))

    def save(self, path: str):
        np.savez_compressed(path, name=np.array(self.name), vocab=np.array(self.vocab), tx_ids=np.array(self.tx_ids), parents=self.parents,
                            children=self.children, features=self.features, train_node_ids=self.train_node_ids,
                            validation_node_ids=self.validation_node_ids, test_node_ids=self.test_node_ids)

    @classmethod
    def load(cls, path: str) -> "TaxonomyDataset":
        d = np.load(path, allow_pickle=False)
        return cls(str(d["name"]), d["vocab"].tolist(), d["tx_ids"].tolist(), d["parents"], d["children"], d["features"],
                   d["train_node_ids"], d["validation_node_ids"], d["test_node_ids"])


def _read_pairs(path):
    out = []
    with open(path, "r") as fin:
        for line in fin:
            line = line.strip()
            if line:
                segs = line.split("\t")
                if len(segs) != 2:
                    raise ValueError(f"Wrong number of segmentations {line}")      # dataset.py:113,124
                out.append((segs[0], segs[1]))
    return out


def _read_word2vec_text(path):
    with open(path, "r") as fin:
        v, d = (int(t) for t in fin.readline().split())
        keys, rows = [], np.zeros((v, d), dtype=np.float32)
        for i in range(v):
            segs = fin.readline().rstrip().split(" ")
            keys.append(segs[0])
            rows[i] = np.asarray(segs[1:1 + d], dtype=np.float32)
    return {k: rows[i] for i, k in enumerate(keys)}, d


def load_raw(dir_path: str, name: str, embed_suffix: str = "", existing_partition: bool = False,
             normalize_embed: bool = False) -> TaxonomyDataset:
    """`MAGDataset._load_dataset_raw` (dataset.py:82-194) without networkx / gensim / DGL."""
    terms = _read_pairs(os.path.join(dir_path, f"{name}.terms"))
    tx_id2node = {}
    names = []
    for tx_id, surface in terms:
        if tx_id in tx_id2node:
            # The reference would add a SECOND graph node here (its Taxon class defines neither __eq__ nor __hash__, so
            # nx.DiGraph.add_node(taxon) at dataset.py:114-116 never merges two lines with the same id) and node ids / edge order
            # would depend on that accident.  No released data set repeats an id; refuse instead of guessing.
            raise ValueError(f"{name}.terms: taxon id {tx_id!r} appears twice")
        tx_id2node[tx_id] = len(names)
        names.append(surface)
    tx_ids = [None] * len(names)
    for k, v in tx_id2node.items():
        tx_ids[v] = k
    # edges: a DiGraph keeps one edge per (parent, child); taxonomy.edges() iterates parents in node order, children in insertion order
    seen, per_parent = set(), {}
    for p, c in _read_pairs(os.path.join(dir_path, f"{name}.taxo")):
        e = (tx_id2node[p], tx_id2node[c])
        if e not in seen:
            seen.add(e)
            per_parent.setdefault(e[0], []).append(e[1])
    # a node first seen as the parent of an edge keeps its place in the adjacency dict: all nodes were added from .terms first, so
    # the iteration order is simply the node order
    parents, children = [], []
    for p in range(len(names)):
        for c in per_parent.get(p, ()):
            parents.append(p)
            children.append(c)
    suffix = f".{embed_suffix}" if embed_suffix else ""
    emb, dim = _read_word2vec_text(os.path.join(dir_path, f"{name}.terms{suffix}.embed"))
    feats = np.zeros((len(emb), dim), dtype=np.float32)          # dataset.py:153: shaped like the embedding matrix
    for node_id, tx_id in enumerate(tx_ids):
        feats[node_id] = emb[tx_id]
    if normalize_embed:                                           # dataset.py:222-223
        feats = feats / np.maximum(np.linalg.norm(feats, axis=1, keepdims=True), 1e-12)
    vocab = [f"{names[i]}@@@{i}" for i in range(len(names))]
    if existing_partition:
        def ids(suffix2):
            with open(os.path.join(dir_path, f"{name}.terms.{suffix2}")) as fin:
                return [tx_id2node[line.strip()] for line in fin if line.strip()]
        train, val, test = ids("train"), ids("validation"), ids("test")
    else:                                                         # dataset.py:167-180: 10 % / 10 % of the leaves, seed 47
        has_child = set(parents)
        leaf = [i for i in range(len(names)) if i not in has_child]
        random.seed(47)
        random.shuffle(leaf)
        n_val = int(len(leaf) * 0.1)
        n_test = int(len(leaf) * 0.1)
        val, test = leaf[:n_val], leaf[n_val:n_val + n_test]
        held = set(val) | set(test)
        train = [i for i in range(len(names)) if i not in held]
    return TaxonomyDataset(name, vocab, tx_ids, np.asarray(parents, np.int64), np.asarray(children, np.int64), feats,
                           np.asarray(train, np.int64), np.asarray(val, np.int64), np.asarray(test, np.int64))


class _Stub:
    """Stands in for classes of the reference code base (parse_config.ConfigParser, ...) pickled inside a checkpoint."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


# globals a reference checkpoint legitimately needs: tensor / storage rebuilders, OrderedDict, numpy scalars.  Everything else in the
# pickle (the reference's ConfigParser, loggers, pathlib objects ...) is replaced by an inert stub: loading a downloaded checkpoint must
# not import - i.e. execute - arbitrary modules, and only state_dict / epoch / monitor_best are used anyway.
_SAFE_GLOBALS = {
    ("collections", "OrderedDict"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_tensor"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch.serialization", "_get_layout"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "str"), ("builtins", "bool"), ("builtins", "bytes"),
}


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS or (module == "torch" and name.endswith("Storage")) or \
                (module == "torch" and name in ("float32", "float64", "float16", "bfloat16", "int64", "int32", "int16", "int8", "uint8", "bool")):
            return super().find_class(module, name)
        return type(name, (_Stub,), {})


class _TolerantPickle:
    Unpickler = _TolerantUnpickler
    __name__ = "pickle"

    @staticmethod
    def load(f, **kw):
        return _TolerantUnpickler(f, **kw).load()


def load_reference_checkpoint(path_or_file, model=None, map_location="cpu"):
    """Reads a reference checkpoint ({'arch', 'epoch', 'state_dict', 'optimizer', 'monitor_best', 'config'},
    base_trainer.py:134-142) and, if `model` is given, loads its state_dict (DataParallel's 'module.' prefix removed)."""
    ckpt = torch.load(path_or_file, map_location=map_location, pickle_module=_TolerantPickle, weights_only=False)
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ckpt["state_dict"].items()}
    if model is not None:
        model.load_state_dict(state, strict=True)
    ckpt["state_dict"] = state
    return ckpt
