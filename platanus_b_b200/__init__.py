"""platanus_b_b200 -- B200-native (sm_100a) k-mer occurrence counting for Platanus_B.

The product is libpbk.so (platanus_b_b200/csrc, C ABI in include/pbk.h) plus the C++ host shim in
platanus_b_b200/host/.  This Python package is plumbing: it builds and loads the library (ctypes),
mirrors the reference's Counter interface for tests/benchmarks, and generates synthetic reads.
"""
from .capi import KmerCounter, KmerGroup, PbkError, load_library, library_path, microbench_atomics  # noqa: F401

__all__ = ["KmerCounter", "KmerGroup", "PbkError", "load_library", "library_path", "microbench_atomics"]
