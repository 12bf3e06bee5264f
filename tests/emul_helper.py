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


def emul_keyx_partition(bases: np.ndarray, offsets: np.ndarray, k: int, n_dest: int, n_regions: int, seg_cap: int,
                        bin_cap: int = 3):
    """Pass A of the key exchange: (send [n_dest, n_regions, seg_cap] u64, cursors [n_dest, n_regions] u64,
    n_instances, spilled (key, weight) records)."""
    bases = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(64, np.uint8)])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_reads = len(offsets) - 1
    send = np.zeros((n_dest, n_regions, seg_cap), np.uint64)
    cursors = np.zeros((n_dest, n_regions), np.uint64)
    spill_cap = max(1024, int(offsets[-1]))
    spill = np.zeros((spill_cap, 2), np.uint64)
    n_inst, n_spill, err = C.c_uint64(), C.c_uint64(), C.c_uint32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_keyx_partition(p(bases), p(offsets), C.c_uint64(n_reads), k, n_dest, n_regions, C.c_uint64(seg_cap), bin_cap,
                                   p(send), p(cursors), C.byref(n_inst), p(spill), C.c_uint64(spill_cap), C.byref(n_spill),
                                   C.byref(err))
    assert rc == 0 and err.value == 0 and n_spill.value <= spill_cap
    return send, cursors, n_inst.value, spill[:n_spill.value]


def emul_keyx_insert(recv: np.ndarray, recv_cursors: np.ndarray, seg_cap: int, k: int, extra=None):
    """Pass B of the key exchange over recv [n_src, n_regions, seg_cap] / recv_cursors [n_src, n_regions]; `extra` =
    (key, weight) records received by the record route.  Returns the sorted (keys, counts) of this rank's table."""
    recv = np.ascontiguousarray(recv, dtype=np.uint64)
    recv_cursors = np.ascontiguousarray(recv_cursors, dtype=np.uint64)
    n_src, n_regions = recv_cursors.shape
    extra = np.zeros((0, 2), np.uint64) if extra is None else np.ascontiguousarray(extra, dtype=np.uint64).reshape(-1, 2)
    cap = int(np.minimum(recv_cursors, seg_cap).sum()) + len(extra) + 1
    keys = np.zeros((cap, 1), np.uint64)
    counts = np.zeros(cap, np.uint16)
    n_out = C.c_uint64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_keyx_insert(p(recv), p(recv_cursors), n_src, n_regions, C.c_uint64(seg_cap), k, p(extra),
                                C.c_uint64(len(extra)), p(keys), p(counts), C.c_uint64(cap), C.byref(n_out))
    assert rc == 0
    m = n_out.value
    keys, counts = keys[:m], counts[:m]
    order = np.argsort(keys[:, 0], kind="stable")
    return keys[order], counts[order]


def emul_lookup(bases: np.ndarray, offsets: np.ndarray, k: int, keys: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """lookup_kernel over a table built from (keys, counts): u16 per base, indexed by window start (pbk_lookup's output)."""
    bases = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(64, np.uint8)])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    out = np.zeros(int(offsets[-1]) + 1, np.uint16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_lookup(p(bases), p(offsets), C.c_uint64(len(offsets) - 1), k, p(keys), p(counts), C.c_uint64(len(counts)), p(out))
    assert rc == 0, rc
    return out[:int(offsets[-1])]


def emul_seeded_count(bases, offsets, k: int, seed_keys, seed_counts):
    """pbk_push_reads + pbk_seed_entries + pbk_finalize: sorted (keys, counts) of the table, n_instances counted."""
    W = (k + 31) // 32
    bases = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(64, np.uint8)])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    sk = np.ascontiguousarray(seed_keys, dtype=np.uint64).reshape(-1, W)
    sc = np.ascontiguousarray(seed_counts, dtype=np.uint16)
    cap = int(offsets[-1]) + len(sc) + 1
    keys = np.zeros((cap, W), np.uint64)
    counts = np.zeros(cap, np.uint16)
    n_out, n_inst = C.c_uint64(), C.c_uint64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_seeded(p(bases), p(offsets), C.c_uint64(len(offsets) - 1), k, p(sk), p(sc), C.c_uint64(len(sc)), p(keys), p(counts),
                           C.c_uint64(cap), C.byref(n_out), C.byref(n_inst), None)
    assert rc == 0, rc
    keys, counts = keys[:n_out.value], counts[:n_out.value]
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    return keys[order], counts[order], n_inst.value


def emul_match_reads(bases, offsets, k: int, keys, counts) -> np.ndarray:
    """pbk_match_reads over a table holding (keys, counts): bool per read."""
    W = (k + 31) // 32
    bases = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(64, np.uint8)])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    sk = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, W)
    sc = np.ascontiguousarray(counts, dtype=np.uint16)
    out = np.zeros(len(offsets), np.uint8)
    n_out, n_inst = C.c_uint64(), C.c_uint64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_seeded(p(bases), p(offsets), C.c_uint64(len(offsets) - 1), k, p(sk), p(sc), C.c_uint64(len(sc)), None, None,
                           C.c_uint64(0), C.byref(n_out), C.byref(n_inst), p(out))
    assert rc == 0, rc
    return out[:len(offsets) - 1].astype(bool)


# ---- the WHOLE C ABI compiled for the host (cuda_rt_shim.h): pbk_api.cu + the launch wrappers of pbk_kernels.cu ------------
ABI_OUT = os.path.join(HERE, "cpu_emul", "_build", "libpbk_emul_abi.so")


def abi_lib_path() -> str:
    """Build (if stale) and return the emulated libpbk: same sources as the product, kernels run as one host thread.
    pbk_kernels.cu is compiled from a copy whose `<<<grid, block, smem, stream>>>` launch configurations are removed."""
    import re
    emul_dir = os.path.join(HERE, "cpu_emul")
    srcs = [os.path.join(CSRC, f) for f in ("pbk_kernels.cu", "pbk_api.cu", "pbk_host.cpp", "pbk_device.cuh", "pbk_kernels.cuh",
                                            "pbk_kernels_impl.cuh", "pbk_group.cu")]
    deps = srcs + [os.path.join(emul_dir, f) for f in ("cuda_shim.h", "cuda_rt_shim.h", "abi_extra.cpp")] + [
        os.path.join(os.path.dirname(HERE), "include", "pbk.h")]
    if os.path.exists(ABI_OUT) and os.path.getmtime(ABI_OUT) >= max(os.path.getmtime(d) for d in deps):
        return ABI_OUT
    build = os.path.dirname(ABI_OUT)
    os.makedirs(build, exist_ok=True)
    text = open(srcs[0]).read()
    text, n = re.subn(r"<<<[^;]*?>>>", "", text)
    assert n >= 25, n
    text = text.replace('#include "pbk_kernels.cuh"', f'#include "{CSRC}/pbk_kernels.cuh"').replace(
        '#include "pbk_kernels_impl.cuh"', f'#include "{CSRC}/pbk_kernels_impl.cuh"')
    gen = os.path.join(build, "pbk_kernels_emul.cpp")
    open(gen, "w").write("// GENERATED from platanus_b_b200/csrc/pbk_kernels.cu (launch configurations removed)\n" + text)
    flags = ["-O1", "-std=c++17", "-fPIC", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-DPBK_CPU_EMUL=1",
             "-include", os.path.join(emul_dir, "cuda_shim.h"), "-include", os.path.join(emul_dir, "cuda_rt_shim.h")]
    objs = []
    for src in (gen, srcs[1], os.path.join(CSRC, "pbk_group.cu"), os.path.join(emul_dir, "abi_extra.cpp"), srcs[2]):
        obj = os.path.join(build, os.path.basename(src).split(".")[0] + "_abi.o")
        subprocess.run(["g++", *flags, "-x", "c++", "-c", src, "-o", obj], check=True)
        objs.append(obj)
    subprocess.run(["g++", "-shared", "-o", ABI_OUT, *objs, "-lpthread"], check=True)
    return ABI_OUT


def abi_cli_path() -> str:
    """pbk_assemble (platanus_b_b200/host) linked against the emulated libpbk: the C++ host side end to end on CPU."""
    lib = abi_lib_path()
    build = os.path.join(os.path.dirname(lib), "cli")
    os.makedirs(build, exist_ok=True)
    link = os.path.join(build, "libpbk.so")
    if not os.path.exists(link) or os.path.getmtime(link) < os.path.getmtime(lib):
        import shutil
        shutil.copyfile(lib, link)
    host = os.path.join(os.path.dirname(HERE), "platanus_b_b200", "host")
    srcs = [os.path.join(host, f) for f in ("pbk_assemble.cpp", "pbk_counter.hpp", "pbk_ingest.hpp")]
    exe = os.path.join(build, "pbk_assemble")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(x) for x in srcs + [link]):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-pthread", "-o", exe, srcs[0], "-L" + build, "-lpbk",
                        "-Wl,-rpath,$ORIGIN"], check=True, env=env)
    return exe


def emul_contigs(bases, offsets, k: int, coverage, min_occurrence: int = 1):
    """contig_max_kernel (pbk_push_contigs): sorted (keys, values) of the table built from contigs and their coverages."""
    W = (k + 31) // 32
    bases = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(64, np.uint8)])
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cov = np.ascontiguousarray(coverage, dtype=np.uint16)
    cap = int(offsets[-1]) + 1
    keys = np.zeros((cap, W), np.uint64)
    counts = np.zeros(cap, np.uint16)
    n_out = C.c_uint64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().emul_contigs(p(bases), p(offsets), C.c_uint64(len(offsets) - 1), k, p(cov), C.c_uint64(min_occurrence), p(keys), p(counts),
                            C.c_uint64(cap), C.byref(n_out))
    assert rc == 0, rc
    keys, counts = keys[:n_out.value], counts[:n_out.value]
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    return keys[order], counts[order]
