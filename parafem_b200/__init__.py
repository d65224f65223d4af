"""parafem_b200 -- B200-native EBE-PCG hot path of ParaFEM p121 / p123.

csrc/      CUDA kernels (kernels.cuh), the C-ABI (device.cu), host helpers (host.cpp)
_lib.py    ctypes binding of include/parafem_b200.h
host.py    host-side mirror of the ParaFEM library calls around the path
solver.py  Python mirror of the device API
driver.py  p121 / p123 program flow on top of the two
"""
from ._lib import PfError  # noqa: F401
