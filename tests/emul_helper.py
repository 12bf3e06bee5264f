"""TEST ONLY: loads tests/cpu_emul (the product kernels compiled for the host, one sequential
thread) so the kernel logic can be checked against the oracle in the GPU-less build container."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_emul", "emul.cpp")
OUT = os.path.join(HERE, "cpu_emul", "_build", "libpbk_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "platanus_b_b200", "csrc")

_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC, os.path.join(HERE, "cpu_emul", "cuda_shim.h")] + [
            os.path.join(CSRC, f) for f in ("pbk_device.cuh", "pbk_kernels_impl.cuh")]
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", OUT, SRC],
                           check=True)
        _lib = C.CDLL(OUT)
    return _lib


def emul_count(bases: np.ndarray, offsets: np.ndarray, k: int, encoding: int = 0, n_pos=None, n_pos_off=None,
               table_slots: int | None = None, n_shards: int = 1, rank: int = 0, min_count: int = 1,
               partition: bool = False):
    W = (k + 31) // 32
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    pad = np.zeros(64, np.uint8)
    bases_p = np.concatenate([bases, pad])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_reads = len(offsets) - 1
    n_bases = int(offsets[-1])
    windows = max(1024, n_bases)
    slots = table_slots or 2 * windows
    keys = np.zeros((windows, W), np.uint64)
    counts = np.zeros(windows, np.uint16)
    occ = np.zeros(65535, np.uint64)
    lh = np.zeros(500001, np.uint64)
    n_out, n_inst, n_remote = C.c_uint64(), C.c_uint64(), C.c_uint64()
    err = C.c_uint32()
    remote = np.zeros((windows, W + 1), np.uint64)
    if n_pos is None:
        n_pos = np.zeros(1, np.int32)
        n_pos_off = np.zeros(n_reads + 1, np.uint64)
    n_pos = np.ascontiguousarray(n_pos, dtype=np.int32)
    n_pos_off = np.ascontiguousarray(n_pos_off, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().emul_set_partition(int(partition))
    rc = lib().emul_count(p(bases_p), p(offsets), C.c_uint64(n_reads), k, encoding, p(n_pos), p(n_pos_off),
                          C.c_uint64(slots), n_shards, rank, min_count, p(keys), p(counts), C.c_uint64(windows),
                          C.byref(n_out), p(occ), p(lh), C.byref(n_inst), C.byref(err), p(remote), C.byref(n_remote))
    assert rc == 0
    n = n_out.value
    keys, counts = keys[:n], counts[:n]
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    return dict(keys=keys[order], counts=counts[order], occ_hist=occ, len_hist=lh, n_instances=n_inst.value,
                err=err.value, remote=remote[:n_remote.value])


def emul_insert_records(records: np.ndarray, k: int):
    W = (k + 31) // 32
    records = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, W + 1)
    n = len(records)
    keys = np.zeros((max(n, 1), W), np.uint64)
    counts = np.zeros(max(n, 1), np.uint16)
    n_out = C.c_uint64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_insert_records(p(records), C.c_uint64(n), k, C.c_uint64(max(1024, 2 * n)), p(keys), p(counts),
                                   C.c_uint64(max(n, 1)), C.byref(n_out))
    assert rc == 0
    m = n_out.value
    keys, counts = keys[:m], counts[:m]
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    return keys[order], counts[order]
