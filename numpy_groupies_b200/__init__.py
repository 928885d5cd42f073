"""numpy_groupies_b200 -- B200-native drop-in for numpy-groupies' ``aggregate`` hot path.

Public surface mirrors what the reference's facade exports for this path
(numpy_groupies/__init__.py:1-43): ``aggregate`` (here always the CUDA implementation --
no backend selection), the implementation module ``aggregate_cuda`` and ``unpack``.
"""
from . import aggregate_cuda
from .aggregate_cuda import aggregate, unpack

__all__ = ["aggregate", "aggregate_cuda", "unpack"]
__version__ = "0.1.0"
