"""dgl_kgat_b200 -- B200-native KGAT attention + aggregation hot path.

A drop-in for the hot path of jennyzhang0215/DGL-KGAT: the reference's Python surface
(``models.Model`` / ``KGATConv``, ``metric.calc_recall_ndcg``, ``dataset.DataLoader``,
``kgat`` flags) on top of hand-written sm_100a CUDA kernels reached through the C ABI of
``include/kgat_b200.h`` (``csrc/libkgat_b200.so``).  No DGL, no CPU fallback.
"""
from . import _lib  # noqa: F401
from .graph import CKGraph  # noqa: F401

__version__ = "0.1.0"
