"""Importable placeholders for `dgl.nn.pytorch.glob` (dead code in the reference,
model_zoo.py:7,260-276). TEST INFRASTRUCTURE ONLY."""
import torch.nn as nn


class SumPooling(nn.Module):
    pass


class MaxPooling(nn.Module):
    pass
