"""Opaque message/reduce tokens standing in for `dgl.function` (DGL 0.4.0).

TEST INFRASTRUCTURE ONLY -- see dgl/__init__.py in this directory.
Only the four builtins used by reference model/model_zoo.py:41,95 exist.
"""
from collections import namedtuple

_Msg = namedtuple("_Msg", "kind a b out")
_Red = namedtuple("_Red", "kind msg out")


def copy_src(src, out):            # model_zoo.py:41
    return _Msg("copy_src", src, None, out)


def src_mul_edge(src, edge, out):  # model_zoo.py:95
    return _Msg("src_mul_edge", src, edge, out)


def sum(msg, out):                 # model_zoo.py:41,95
    return _Red("sum", msg, out)
