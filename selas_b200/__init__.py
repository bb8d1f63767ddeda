"""selas_b200 -- B200-native (sm_100a) path-tracing hot path of schuttejoe/Selas behind a C ABI.

The product is selas_b200/lib/libselas_cuda.so (sources in selas_b200/csrc, interface in include/selas_cuda.h);
`capi` is the ctypes binding, `scenes` the synthetic-input generators for the BASELINE.json configs.
"""
from . import capi, scenes  # noqa: F401

__all__ = ["capi", "scenes"]
