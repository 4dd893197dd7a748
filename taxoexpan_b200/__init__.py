"""taxoexpan_b200: TaxoExpan's position-enhanced graph propagation + readout hot path, B200-native.

    from taxoexpan_b200 import TaxoExpan, EgonetBatch, DGLGraph, batch

The CUDA library (libtaxo_sm100.so, C ABI in include/taxo_b200.h) is loaded on first use; build it with
`python -m taxoexpan_b200.build`.
"""
from . import dataset_io, inference, loss, sampler, synth  # noqa: F401
from ._lib import TaxoLibraryError  # noqa: F401
from .graph import DGLGraph, EgonetBatch, batch  # noqa: F401
from .loss import info_nce_loss  # noqa: F401
from .model import TaxoExpan  # noqa: F401
from .model_zoo import (BIM, GAT, GCN, LBM, MLP, PGAT, PGCN, ConcatReadout, GATLayer, GCNLayer, MeanReadout,  # noqa: F401
                        WeightedMeanReadout)

__version__ = "0.1.0"
