"""schemanet_b200 -- B200-native (sm_100a) schema-inference head of SchemaNet.

`native`  : ctypes binding of libschemahead.so (hand-written CUDA kernels behind a C ABI)
`head`    : the fused, sync-free tensor path (discretize -> instance graphs -> match) used by the drop-in modules
`dist`    : batch / class sharding helpers for one-process-per-GPU runs
The drop-in modules that mirror the reference's names live next to this package:
`cpp_extension`, `discretization`, `schema_inference.graph`, `schema_inference.utils`.
"""
from . import native  # noqa: F401

__all__ = ["native"]
