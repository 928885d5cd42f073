"""numpy_groupies_b200 -- B200-native drop-in for numpy-groupies' ``aggregate`` hot path.

Public surface mirrors what the reference's facade exports for this path
(numpy_groupies/__init__.py:1-43): ``aggregate`` (here always the CUDA implementation --
no backend selection), the implementation module ``aggregate_cuda`` and ``unpack``.
"""
from . import aggregate_cuda
from .aggregate_cuda import (aggregate, aggregate_multi, group_csr, is_duck_array, label_contiguous_1d, multi_arange,
                             relabel_groups_masked, relabel_groups_unique, step_count, step_indices, uaggregate, unpack)

# the names the reference's facade exports for this path (numpy_groupies/__init__.py:10-43), all served by the CUDA
# implementation; `aggregate_cuda` sits where the facade keeps aggregate_np / aggregate_nb
__all__ = ["aggregate", "aggregate_cuda", "aggregate_multi", "group_csr", "is_duck_array", "label_contiguous_1d",
           "multi_arange", "relabel_groups_masked", "relabel_groups_unique", "step_count", "step_indices", "uaggregate",
           "unpack"]
__version__ = "0.1.0"
