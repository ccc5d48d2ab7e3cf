"""folddisco_b200 -- B200 (sm_100a) accelerator for the folddisco index/query hot path.

The compute lives in libfolddisco_b200.so (hand-written CUDA behind the C ABI of
include/folddisco_b200.h).  This package is the thin host-side mirror of the reference's
interface for that path; it has no CPU fallback and raises if the library is missing.
"""
from .capi import (  # noqa: F401
    FdError,
    Context,
    StructBatch,
    HashParams,
    PrefilterParams,
    IndexBuffers,
    lib,
    library_path,
)

__all__ = ["FdError", "Context", "StructBatch", "HashParams", "PrefilterParams", "IndexBuffers", "lib",
           "library_path"]
