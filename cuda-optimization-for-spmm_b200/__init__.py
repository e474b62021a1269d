"""cuda-optimization-for-spmm_b200 -- B200-native SpMM engine behind the
Cuda-Optimization-for-SpMM format/engine API.

The product is ``libcuspmm_b200.so`` (C ABI: include/cuspmm_b200.h, sources: csrc/) and the
C++ host layer in ``host/`` (storage classes, Engine<FMT>, runEngine, the ``cuspmm`` CLI).
This Python package is plumbing for tests and bench.py only: ``binding`` is a ctypes view of
the C ABI over torch CUDA tensors (torch is used for device memory and streams, nothing
else).  The directory name contains '-' so it is imported by path:
``__graft_entry__.load_package()``.
"""
from . import binding  # noqa: F401
from .binding import lib, CuspmmError  # noqa: F401
