"""mmannot_b200 -- B200-native read-annotation hot path of mmannot behind a C ABI.

host    : config / GTF / SAM+BAM front-end (libmmannot_host.so, C++)
device  : the CUDA path (libmmannot_b200.so, sm_100a) -- no CPU fallback
"""
from . import host  # noqa: F401
from . import device  # noqa: F401

__all__ = ["host", "device"]
