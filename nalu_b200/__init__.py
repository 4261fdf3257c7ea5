"""nalu_b200 -- B200-native Hex8 CVFEM element assembly (the one hot path of NaluCFD/Nalu).

Layout:
  csrc/      CUDA kernels (sm_100a) + the C ABI of include/nalu_b200.h  -> libnalu_b200.so
  host/      C++ mirror of the reference's Kernel / ElemDataRequests / AssembleElemSolverAlgorithm /
             LinearSystem registration API on top of the C ABI           -> libnalu_b200_host.so
  lib.py     ctypes door to the C ABI
  linsys.py  DeviceCsrLinearSystem + assemble helpers used by tests / bench
  mesh.py, fields.py   flat-array mesh substrate and analytic test fields
"""
from . import fields, mesh  # noqa: F401

__all__ = ["mesh", "fields"]
