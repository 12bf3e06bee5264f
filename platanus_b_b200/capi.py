"""ctypes binding of include/pbk.h (libpbk.so) -- plumbing only, no compute happens in Python.

`KmerCounter` mirrors the part of the reference's ``Counter<KMER>`` interface that the
`assemble -kmer_occ_only` path uses (counter.h:36-201): makeKmerReadDistributionMT,
getLeftLocalMinimalValue, calcOccurrenceDistributionAverage, calcLengthDistributionAverage,
getMaxOccurrence, outputOccurrenceDistribution, sortedKeyFromKmerFile, loadKmer,
outputOccurrenceTableBinary -- same names (snake_case), argument meaning and error behaviour.
The library has no CPU fallback: without a CUDA device `KmerCounter()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OCC_BINS = 65535
LEN_BINS = 500001
COUNT_SAT = 65534
MAX_K = 256
ENC_ASCII, ENC_PLATANUS = 0, 1
F_TIMING = 1
F_NO_PARTITION = 2
F_FORCE_PARTITION = 4
F_NO_PIPELINE = 8
F_UNKNOWN_AS_N = 16

STATUS = {0: "PBK_OK", -1: "PBK_E_ARG", -2: "PBK_E_NO_DEVICE", -3: "PBK_E_CUDA", -4: "PBK_E_NOMEM",
          -5: "PBK_E_READ_TOO_LONG", -6: "PBK_E_BAD_BASE", -7: "PBK_E_KMER_DIST", -8: "PBK_E_STATE",
          -9: "PBK_E_IO", -10: "PBK_E_UNSUPPORTED_K"}

# every symbol include/pbk.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pbk_strerror", "pbk_last_error", "pbk_abi_version", "pbk_create", "pbk_destroy", "pbk_host_alloc",
    "pbk_host_free", "pbk_push_reads", "pbk_push_reads_device", "pbk_finalize", "pbk_export", "pbk_get_stats",
    "pbk_reset", "pbk_shard_record_bytes", "pbk_shard_send_counts", "pbk_shard_pack_device",
    "pbk_shard_insert_device", "pbk_shard_of_key", "pbk_left_local_min", "pbk_coverage_cutoff",
    "pbk_distribution_average", "pbk_double_hash_size", "pbk_write_frq_tsv", "pbk_write_kmer_occ_bin",
    "pbk_microbench_atomics", "pbk_timer_mark", "pbk_timer_elapsed_ms", "pbk_set_timing",
    "pbk_keyx_plan", "pbk_keyx_partition", "pbk_keyx_partition_device", "pbk_keyx_insert_device",
    "pbk_lookup", "pbk_lookup_device", "pbk_load_entries", "pbk_read_kmer_occ_bin", "pbk_free",
    "pbk_match_reads", "pbk_seed_entries", "pbk_stream_signal", "pbk_stream_wait", "pbk_keyx_partition_device_async",
    "pbk_push_contigs", "pbk_keyx_pull_setup", "pbk_keyx_pull_handle", "pbk_keyx_pull_connect_ipc", "pbk_keyx_pull_connect_local",
    "pbk_keyx_pull_partition", "pbk_keyx_pull_partition_device", "pbk_keyx_pull_insert", "pbk_keyx_pull_release", "pbk_keyx_staged_count_device",
    "pbk_device_count", "pbk_group_create", "pbk_group_destroy", "pbk_group_size", "pbk_group_member", "pbk_group_last_error",
    "pbk_group_reset", "pbk_group_push_reads", "pbk_group_finalize", "pbk_group_export",
    "pbk_pack_reads", "pbk_push_reads_packed", "pbk_neighbor_flags", "pbk_group_neighbor_flags",
]


class PbkConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("k", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32),
                ("n_shards", C.c_uint32), ("shard_rank", C.c_uint32), ("table_slots_hint", C.c_uint64),
                ("hbm_budget_bytes", C.c_uint64), ("n_passes", C.c_uint32), ("pass_index", C.c_uint32)]


class PbkStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_instances", C.c_uint64),
                ("n_distinct", C.c_uint64), ("table_slots", C.c_uint64), ("table_bytes", C.c_uint64),
                ("n_grow", C.c_uint64), ("launches_pack", C.c_uint64), ("launches_count", C.c_uint64),
                ("launches_other", C.c_uint64), ("ms_pack", C.c_double), ("ms_count", C.c_double),
                ("ms_other", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches_partition", C.c_uint64), ("launches_insert", C.c_uint64),
                ("ms_partition", C.c_double), ("ms_insert", C.c_double), ("ms_count_elapsed", C.c_double),
                ("n_pipelined_batches", C.c_uint64), ("n_split_build", C.c_uint64)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PbkKeyxLayout(C.Structure):
    """pbk_keyx_layout (include/pbk.h): the all-to-all layout of the key exchange"""
    _fields_ = [("n_dest", C.c_uint32), ("n_regions", C.c_uint32), ("seg_cap", C.c_uint64), ("entry_bytes", C.c_uint64),
                ("bytes_per_dest", C.c_uint64), ("cursors_per_dest", C.c_uint64)]


class PbkError(RuntimeError):
    """A libpbk call failed.  `.status` is the pbk_status code; the reference throws the platanus::
    exception named in include/pbk.h for the same condition."""

    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        msg = f"{what}: {STATUS.get(status, status)}"
        if detail:
            msg += f" ({detail})"
        super().__init__(msg)


_lib = None


def library_path() -> str:
    return _build.LIB


def load_library(build_if_missing: bool = True):
    """dlopen libpbk.so (built in-tree).  Loading needs no GPU; creating a context does."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(f"{path} is missing: run `python -m platanus_b_b200.build`")
        _build.build_lib()
    L = C.CDLL(path)
    vp, u64p = C.c_void_p, C.c_void_p
    L.pbk_strerror.argtypes = [C.c_int]; L.pbk_strerror.restype = C.c_char_p
    L.pbk_last_error.argtypes = [vp]; L.pbk_last_error.restype = C.c_char_p
    L.pbk_abi_version.restype = C.c_int
    L.pbk_create.argtypes = [C.POINTER(vp), C.POINTER(PbkConfig)]
    L.pbk_destroy.argtypes = [vp]; L.pbk_destroy.restype = None
    L.pbk_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.pbk_host_free.argtypes = [vp]; L.pbk_host_free.restype = None
    L.pbk_push_reads.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p]
    L.pbk_push_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64]
    L.pbk_pack_reads.argtypes = [vp, C.c_uint64, u64p, u64p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.pbk_push_reads_packed.argtypes = [vp, u64p, u64p, C.c_uint64, u64p, C.c_uint64]
    L.pbk_neighbor_flags.argtypes = [vp, C.c_uint32, u64p, C.c_uint64, vp]
    L.pbk_group_neighbor_flags.argtypes = [vp, C.c_uint32, u64p, C.c_uint64, vp]
    L.pbk_finalize.argtypes = [vp, u64p, u64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pbk_export.argtypes = [vp, C.c_uint32, C.c_int, u64p, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.pbk_get_stats.argtypes = [vp, C.POINTER(PbkStats)]
    L.pbk_reset.argtypes = [vp, C.c_uint32]
    L.pbk_set_timing.argtypes = [vp, C.c_int]
    L.pbk_timer_mark.argtypes = [vp, C.c_int]
    L.pbk_timer_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.pbk_shard_record_bytes.argtypes = [vp]; L.pbk_shard_record_bytes.restype = C.c_uint32
    L.pbk_shard_send_counts.argtypes = [vp, u64p]
    L.pbk_shard_pack_device.argtypes = [vp, vp, C.c_uint64]
    L.pbk_shard_insert_device.argtypes = [vp, vp, C.c_uint64]
    L.pbk_shard_of_key.argtypes = [u64p, C.c_uint32, C.c_uint32]; L.pbk_shard_of_key.restype = C.c_uint32
    L.pbk_lookup.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p, vp]
    L.pbk_lookup_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, vp]
    L.pbk_load_entries.argtypes = [vp, u64p, vp, C.c_uint64]
    L.pbk_seed_entries.argtypes = [vp, u64p, vp, C.c_uint64]
    L.pbk_push_contigs.argtypes = [vp, vp, u64p, C.c_uint64, vp, C.c_uint64]
    L.pbk_match_reads.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p, vp]
    L.pbk_read_kmer_occ_bin.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp),
                                        C.POINTER(C.c_uint64)]
    L.pbk_free.argtypes = [vp]; L.pbk_free.restype = None
    L.pbk_keyx_plan.argtypes = [vp, C.c_uint64, C.POINTER(PbkKeyxLayout)]
    L.pbk_keyx_partition.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p, vp, vp]
    L.pbk_keyx_partition_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp]
    L.pbk_keyx_insert_device.argtypes = [vp, vp, vp]
    L.pbk_keyx_partition_device_async.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp]
    L.pbk_keyx_pull_setup.argtypes = [vp, C.c_uint64, C.POINTER(PbkKeyxLayout)]
    L.pbk_keyx_pull_handle.argtypes = [vp, vp]
    L.pbk_keyx_pull_connect_ipc.argtypes = [vp, C.c_uint32, vp]
    L.pbk_keyx_pull_connect_local.argtypes = [vp, C.c_uint32, vp]
    L.pbk_keyx_pull_partition.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p]
    L.pbk_keyx_pull_partition_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int]
    L.pbk_keyx_pull_insert.argtypes = [vp]
    L.pbk_keyx_pull_release.argtypes = [vp]
    L.pbk_keyx_staged_count_device.argtypes = [vp, vp]
    L.pbk_device_count.restype = C.c_int
    L.pbk_group_create.argtypes = [C.POINTER(vp), C.POINTER(PbkConfig), vp, C.c_uint32]
    L.pbk_group_destroy.argtypes = [vp]; L.pbk_group_destroy.restype = None
    L.pbk_group_size.argtypes = [vp]; L.pbk_group_size.restype = C.c_uint32
    L.pbk_group_member.argtypes = [vp, C.c_uint32]; L.pbk_group_member.restype = vp
    L.pbk_group_last_error.argtypes = [vp]; L.pbk_group_last_error.restype = C.c_char_p
    L.pbk_group_reset.argtypes = [vp, C.c_uint32]
    L.pbk_group_push_reads.argtypes = [vp, vp, u64p, C.c_uint64, C.c_int, vp, u64p]
    L.pbk_group_finalize.argtypes = [vp, u64p, u64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pbk_group_export.argtypes = [vp, C.c_uint32, C.c_int, u64p, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.pbk_stream_signal.argtypes = [vp, vp]
    L.pbk_stream_wait.argtypes = [vp, vp]
    L.pbk_left_local_min.argtypes = [u64p, C.c_uint64, C.c_uint64]; L.pbk_left_local_min.restype = C.c_uint64
    L.pbk_coverage_cutoff.argtypes = [u64p, C.c_uint64, C.c_int, C.c_int]; L.pbk_coverage_cutoff.restype = C.c_uint64
    L.pbk_distribution_average.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
    L.pbk_double_hash_size.argtypes = [C.c_uint64, C.c_uint32]; L.pbk_double_hash_size.restype = C.c_uint64
    L.pbk_write_frq_tsv.argtypes = [C.c_char_p, u64p, C.c_uint64]
    L.pbk_write_kmer_occ_bin.argtypes = [C.c_char_p, C.c_uint32, u64p, vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.pbk_microbench_atomics.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack_reads(bases: np.ndarray):
    """pbk_pack_reads: ASCII bases (all reads concatenated) -> (2-bit stream words uint64, absolute N positions uint64) --
    what a host parser that packs as it copies hands to push_reads_packed."""
    L = load_library()
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    n = len(bases)
    words = np.zeros((n + 31) // 32, np.uint64)
    n_n = C.c_uint64()
    rc = L.pbk_pack_reads(_ptr(bases), n, _ptr(words), None, 0, C.byref(n_n))
    if rc:
        raise PbkError(rc, "pbk_pack_reads")
    npos = np.zeros(n_n.value, np.uint64)
    if n_n.value:
        rc = L.pbk_pack_reads(_ptr(bases), n, _ptr(words), _ptr(npos), n_n.value, C.byref(n_n))
        if rc:
            raise PbkError(rc, "pbk_pack_reads")
    return words, npos


class KmerGroup:
    """Several GPUs of one box behind one handle (a `pbk_group`): the library cuts every batch into one slice per device and
    exchanges keys between the devices itself (pull exchange over peer-mapped HBM for k <= 32, records for k > 32)."""

    def __init__(self, k: int, devices, partition: bool | str = True):
        self._L = load_library()
        self._g = C.c_void_p()
        self.k, self.words = int(k), (int(k) + 31) // 32
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        cfg = PbkConfig(C.sizeof(PbkConfig), self.k, -1, F_FORCE_PARTITION if partition == "force" else 0 if partition else F_NO_PARTITION, 0, 0, 0, 0, 0, 0)
        rc = self._L.pbk_group_create(C.byref(self._g), C.byref(cfg), _ptr(devs), len(devs))
        if rc:
            self._g = C.c_void_p()
            raise PbkError(rc, "pbk_group_create", self._L.pbk_strerror(rc).decode())
        self.occ_hist = self.len_hist = None
        self.n_distinct = self.n_instances = self.max_occurrence = 0

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            self._L.pbk_group_destroy(self._g)
            self._g = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc:
            raise PbkError(rc, what, self._L.pbk_group_last_error(self._g).decode())

    def push_reads(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._L.pbk_group_push_reads(self._g, _ptr(bases), _ptr(offsets), len(offsets) - 1, ENC_ASCII, None, None), "pbk_group_push_reads")

    def finalize(self):
        occ = np.zeros(OCC_BINS, np.uint64)
        lh = np.zeros(LEN_BINS, np.uint64)
        nd, ni, mx = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.pbk_group_finalize(self._g, _ptr(occ), _ptr(lh), C.byref(nd), C.byref(ni), C.byref(mx)), "pbk_group_finalize")
        self.occ_hist, self.len_hist = occ, lh
        self.n_distinct, self.n_instances, self.max_occurrence = nd.value, ni.value, mx.value
        return self

    def export(self, min_count: int = 1, sorted: bool = True):
        n = C.c_uint64()
        self._check(self._L.pbk_group_export(self._g, min_count, int(sorted), None, None, 0, C.byref(n)), "pbk_group_export")
        keys = np.zeros((n.value, self.words), np.uint64)
        counts = np.zeros(n.value, np.uint16)
        if n.value:
            self._check(self._L.pbk_group_export(self._g, min_count, int(sorted), _ptr(keys), _ptr(counts), n.value, C.byref(n)), "pbk_group_export")
        return keys, counts

    def neighbor_flags(self, keys: np.ndarray, min_count: int = 1) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = keys.shape[0] if keys.ndim == 2 else len(keys) // self.words
        out = np.zeros(n, np.uint8)
        self._check(self._L.pbk_group_neighbor_flags(self._g, int(min_count), _ptr(keys), n, _ptr(out)), "pbk_group_neighbor_flags")
        return out

    def reset(self, k: int = 0):
        self._check(self._L.pbk_group_reset(self._g, k), "pbk_group_reset")
        if k:
            self.k, self.words = int(k), (int(k) + 31) // 32

    def member_stats(self, i: int) -> dict:
        st = PbkStats()
        rc = self._L.pbk_get_stats(self._L.pbk_group_member(self._g, i), C.byref(st))
        if rc:
            raise PbkError(rc, "pbk_get_stats")
        return st.asdict()


class KmerCounter:
    """One GPU's k-mer occurrence counter (a `pbk_ctx`)."""

    def __init__(self, k: int, device: int = -1, n_shards: int = 1, shard_rank: int = 0, timing: bool = False,
                 table_slots_hint: int = 0, hbm_budget_bytes: int = 0, partition: bool | str = True,
                 pipeline: bool = True, unknown_as_n: bool = False, n_passes: int = 0, pass_index: int = 0):
        """n_passes >= 2: a hash-range pass on one GPU (pbk_config.n_passes) -- this context counts only the k-mers of range
        `pass_index`; push the same reads into one context per pass and add the results up."""
        self._L = load_library()
        self._ctx = C.c_void_p()
        self.k = int(k)
        self.words = (self.k + 31) // 32
        cfg = PbkConfig(C.sizeof(PbkConfig), self.k, device, (F_TIMING if timing else 0) | (F_FORCE_PARTITION if partition == "force" else 0 if partition else F_NO_PARTITION) | (0 if pipeline else F_NO_PIPELINE) | (F_UNKNOWN_AS_N if unknown_as_n else 0),
                        n_shards, shard_rank,
                        table_slots_hint, hbm_budget_bytes, n_passes, pass_index)
        rc = self._L.pbk_create(C.byref(self._ctx), C.byref(cfg))
        if rc:
            self._ctx = C.c_void_p()
            raise PbkError(rc, "pbk_create", self._L.pbk_strerror(rc).decode())
        self.occ_hist = None
        self.len_hist = None
        self.n_distinct = self.n_instances = self.max_occurrence = 0
        self.double_hash_size = 0

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.pbk_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc:
            raise PbkError(rc, what, self._L.pbk_last_error(self._ctx).decode())

    # -- counting -------------------------------------------------------------------------------
    def push_reads(self, bases: np.ndarray, offsets: np.ndarray, encoding: int = ENC_ASCII, n_pos=None,
                   n_pos_offsets=None):
        """Host buffers (numpy; pinned or pageable).  offsets: uint64[n_reads + 1]."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        np_p = npo_p = None
        if encoding == ENC_PLATANUS:
            n_pos = np.ascontiguousarray(n_pos, dtype=np.int32)
            n_pos_offsets = np.ascontiguousarray(n_pos_offsets, dtype=np.uint64)
            np_p, npo_p = _ptr(n_pos), _ptr(n_pos_offsets)
        self._check(self._L.pbk_push_reads(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, encoding,
                                           np_p, npo_p), "pbk_push_reads")

    def push_reads_ptr(self, bases_ptr: int, offsets_ptr: int, n_reads: int):
        """Host pointers (e.g. torch pinned tensors' data_ptr())."""
        self._check(self._L.pbk_push_reads(self._ctx, C.c_void_p(bases_ptr), C.c_void_p(offsets_ptr), n_reads,
                                           ENC_ASCII, None, None), "pbk_push_reads")

    def push_reads_packed(self, words: np.ndarray, offsets: np.ndarray, n_positions: np.ndarray):
        """Opt-in host form: 2-bit words (pack_reads) + absolute N positions: 0.25 byte per base over PCIe, no pack kernel."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_positions = np.ascontiguousarray(n_positions, dtype=np.uint64)
        self._check(self._L.pbk_push_reads_packed(self._ctx, _ptr(words), _ptr(offsets), len(offsets) - 1,
                                                  _ptr(n_positions) if len(n_positions) else None, len(n_positions)), "pbk_push_reads_packed")

    def push_reads_packed_ptr(self, words_ptr: int, offsets_ptr: int, n_reads: int, npos_ptr: int, n_n: int):
        self._check(self._L.pbk_push_reads_packed(self._ctx, C.c_void_p(words_ptr), C.c_void_p(offsets_ptr), n_reads,
                                                  C.c_void_p(npos_ptr) if n_n else None, n_n), "pbk_push_reads_packed")

    def push_reads_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int):
        """Device pointers (inputs already resident in HBM)."""
        self._check(self._L.pbk_push_reads_device(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr),
                                                  n_reads, n_bases), "pbk_push_reads_device")

    def finalize(self):
        occ = np.zeros(OCC_BINS, np.uint64)
        lh = np.zeros(LEN_BINS, np.uint64)
        nd, ni, mx = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.pbk_finalize(self._ctx, _ptr(occ), _ptr(lh), C.byref(nd), C.byref(ni), C.byref(mx)),
                    "pbk_finalize")
        self.occ_hist, self.len_hist = occ, lh
        self.n_distinct, self.n_instances, self.max_occurrence = nd.value, ni.value, mx.value
        return self

    def finalize_light(self):
        """finalize without copying the 4 MB length histogram back (bench inner loop)."""
        occ = np.zeros(OCC_BINS, np.uint64)
        nd, ni, mx = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.pbk_finalize(self._ctx, _ptr(occ), None, C.byref(nd), C.byref(ni), C.byref(mx)),
                    "pbk_finalize")
        self.occ_hist = occ
        self.n_distinct, self.n_instances, self.max_occurrence = nd.value, ni.value, mx.value
        return self

    def export(self, min_count: int = 1, sorted: bool = True):
        n = C.c_uint64()
        self._check(self._L.pbk_export(self._ctx, min_count, int(sorted), None, None, 0, C.byref(n)), "pbk_export")
        keys = np.zeros((n.value, self.words), np.uint64)
        counts = np.zeros(n.value, np.uint16)
        if n.value:
            self._check(self._L.pbk_export(self._ctx, min_count, int(sorted), _ptr(keys), _ptr(counts), n.value,
                                           C.byref(n)), "pbk_export")
        return keys, counts

    def export_count(self, min_count: int = 1) -> int:
        n = C.c_uint64()
        self._check(self._L.pbk_export(self._ctx, min_count, 0, None, None, 0, C.byref(n)), "pbk_export")
        return n.value

    def export_into(self, min_count: int, sorted: bool, keys_ptr: int, counts_ptr: int, capacity: int) -> int:
        """pbk_export into host buffers of the caller (e.g. pinned torch tensors): returns the number of entries."""
        n = C.c_uint64()
        self._check(self._L.pbk_export(self._ctx, min_count, int(sorted), C.c_void_p(keys_ptr), C.c_void_p(counts_ptr), capacity,
                                       C.byref(n)), "pbk_export")
        return n.value

    def reset(self, k: int = 0):
        self._check(self._L.pbk_reset(self._ctx, k), "pbk_reset")
        if k:
            self.k, self.words = int(k), (int(k) + 31) // 32
        self.occ_hist = self.len_hist = None

    def set_timing(self, on: bool):
        self._check(self._L.pbk_set_timing(self._ctx, int(on)), "pbk_set_timing")

    def timer_mark(self, slot: int):
        self._check(self._L.pbk_timer_mark(self._ctx, slot), "pbk_timer_mark")

    def timer_elapsed_ms(self, start: int, stop: int) -> float:
        out = C.c_double()
        self._check(self._L.pbk_timer_elapsed_ms(self._ctx, start, stop, C.byref(out)), "pbk_timer_elapsed_ms")
        return out.value

    def stats(self) -> dict:
        st = PbkStats()
        self._check(self._L.pbk_get_stats(self._ctx, C.byref(st)), "pbk_get_stats")
        return st.asdict()

    # -- sharding -------------------------------------------------------------------------------
    def shard_record_words(self) -> int:
        return self._L.pbk_shard_record_bytes(self._ctx) // 8

    def shard_send_counts(self, n_shards: int) -> np.ndarray:
        cnt = np.zeros(n_shards, np.uint64)
        self._check(self._L.pbk_shard_send_counts(self._ctx, _ptr(cnt)), "pbk_shard_send_counts")
        return cnt

    def shard_pack_device(self, d_records_ptr: int, capacity_records: int):
        self._check(self._L.pbk_shard_pack_device(self._ctx, C.c_void_p(d_records_ptr), capacity_records),
                    "pbk_shard_pack_device")

    def shard_insert_device(self, d_records_ptr: int, n_records: int):
        self._check(self._L.pbk_shard_insert_device(self._ctx, C.c_void_p(d_records_ptr), n_records),
                    "pbk_shard_insert_device")

    # -- consumers of the table (kmer_divide.cpp:151-197; counter.h:967-993) -------------------------
    def lookup(self, bases: np.ndarray, offsets: np.ndarray, encoding: int = ENC_ASCII, n_pos=None, n_pos_offsets=None) -> np.ndarray:
        """u16 per base: occurrence of the k-mer window starting there (ContigDivider::getOccurrenceArray), 0 if absent / N."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(int(offsets[-1]), np.uint16)
        np_p = npo_p = None
        if encoding == ENC_PLATANUS:
            n_pos = np.ascontiguousarray(n_pos, dtype=np.int32)
            n_pos_offsets = np.ascontiguousarray(n_pos_offsets, dtype=np.uint64)
            np_p, npo_p = _ptr(n_pos), _ptr(n_pos_offsets)
        self._check(self._L.pbk_lookup(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, encoding, np_p, npo_p, _ptr(out)),
                    "pbk_lookup")
        return out

    def neighbor_flags(self, keys: np.ndarray, min_count: int = 1) -> np.ndarray:
        """makeInitialBruijnGraph's eight findValue probes per k-mer (graph.h:337-375): u8 per key, (leftFlags << 4) | rightFlags."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = keys.shape[0] if keys.ndim == 2 else len(keys) // self.words
        out = np.zeros(n, np.uint8)
        self._check(self._L.pbk_neighbor_flags(self._ctx, int(min_count), _ptr(keys), n, _ptr(out)), "pbk_neighbor_flags")
        return out

    def match_reads(self, bases: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """Counter::pickupReadMatchedEdgeKmer (counter.h:870-910): bool per read -- has a k-mer that is in the table."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(len(offsets) - 1, np.uint8)
        self._check(self._L.pbk_match_reads(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, ENC_ASCII, None, None, _ptr(out)),
                    "pbk_match_reads")
        return out.astype(bool)

    def seed_entries(self, keys: np.ndarray, counts: np.ndarray):
        """makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750): these k-mers keep their value."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        self._check(self._L.pbk_seed_entries(self._ctx, _ptr(keys), _ptr(counts), len(counts)), "pbk_seed_entries")

    def push_contigs(self, bases: np.ndarray, offsets: np.ndarray, coverage: np.ndarray, min_occurrence: int = 1):
        """makeKmerReadDistributionFromContig (counter.h:511-593): table[k-mer] = max over its contigs of max(coverage, min_occurrence)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        coverage = np.ascontiguousarray(coverage, dtype=np.uint16)
        self._check(self._L.pbk_push_contigs(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, _ptr(coverage), int(min_occurrence)),
                    "pbk_push_contigs")

    def load_entries(self, keys: np.ndarray, counts: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        self._check(self._L.pbk_load_entries(self._ctx, _ptr(keys), _ptr(counts), len(counts)), "pbk_load_entries")

    def read_occurrence_table_binary(self, path: str):
        """Counter::readOccurrenceTableBinary (counter.h:967-993): load PREFIX_kmer_occ.bin into this counter's table."""
        k, idx, n = C.c_uint32(), C.c_uint64(), C.c_uint64()
        keys, counts = C.c_void_p(), C.c_void_p()
        rc = self._L.pbk_read_kmer_occ_bin(path.encode(), C.byref(k), C.byref(idx), C.byref(keys), C.byref(counts), C.byref(n))
        if rc:
            raise PbkError(rc, "pbk_read_kmer_occ_bin", path)
        try:
            if k.value != self.k:
                self.reset(k.value)
            self._check(self._L.pbk_load_entries(self._ctx, keys, counts, n.value), "pbk_load_entries")
        finally:
            self._L.pbk_free(keys); self._L.pbk_free(counts)
        return n.value

    # -- sharding, second form: keys exchanged before counting (k <= 32) ---------------------------
    def keyx_plan(self, max_windows_any_rank: int) -> PbkKeyxLayout:
        lay = PbkKeyxLayout()
        self._check(self._L.pbk_keyx_plan(self._ctx, int(max_windows_any_rank), C.byref(lay)), "pbk_keyx_plan")
        return lay

    def keyx_partition(self, bases: np.ndarray, offsets: np.ndarray, d_send_ptr: int, d_cursors_ptr: int):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._L.pbk_keyx_partition(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, ENC_ASCII, None, None,
                                               C.c_void_p(d_send_ptr), C.c_void_p(d_cursors_ptr)), "pbk_keyx_partition")

    def keyx_partition_ptr(self, bases_ptr: int, offsets_ptr: int, n_reads: int, d_send_ptr: int, d_cursors_ptr: int):
        self._check(self._L.pbk_keyx_partition(self._ctx, C.c_void_p(bases_ptr), C.c_void_p(offsets_ptr), n_reads, ENC_ASCII,
                                               None, None, C.c_void_p(d_send_ptr), C.c_void_p(d_cursors_ptr)), "pbk_keyx_partition")

    def keyx_partition_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, d_send_ptr: int,
                              d_cursors_ptr: int):
        self._check(self._L.pbk_keyx_partition_device(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads,
                                                      n_bases, C.c_void_p(d_send_ptr), C.c_void_p(d_cursors_ptr)),
                    "pbk_keyx_partition_device")

    def keyx_partition_device_async(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, d_send_ptr: int,
                                    d_cursors_ptr: int):
        self._check(self._L.pbk_keyx_partition_device_async(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads,
                                                            n_bases, C.c_void_p(d_send_ptr), C.c_void_p(d_cursors_ptr)),
                    "pbk_keyx_partition_device_async")

    def stream_signal(self, cuda_stream: int):
        """the caller's stream (a raw cudaStream_t) waits for what this context has queued"""
        self._check(self._L.pbk_stream_signal(self._ctx, C.c_void_p(cuda_stream)), "pbk_stream_signal")

    def stream_wait(self, cuda_stream: int):
        """this context waits for what has been queued on the caller's stream"""
        self._check(self._L.pbk_stream_wait(self._ctx, C.c_void_p(cuda_stream)), "pbk_stream_wait")

    def keyx_insert_device(self, d_recv_ptr: int, d_recv_cursors_ptr: int):
        self._check(self._L.pbk_keyx_insert_device(self._ctx, C.c_void_p(d_recv_ptr), C.c_void_p(d_recv_cursors_ptr)),
                    "pbk_keyx_insert_device")

    # -- pull form of the key exchange: Pass B reads the peers' bucket stores in place over NVLink (pbk_keyx_pull_*) -------
    KEYX_HANDLE_BYTES = 64

    def keyx_pull_setup(self, max_windows_any_rank: int) -> PbkKeyxLayout:
        lay = PbkKeyxLayout()
        self._check(self._L.pbk_keyx_pull_setup(self._ctx, int(max_windows_any_rank), C.byref(lay)), "pbk_keyx_pull_setup")
        return lay

    def keyx_pull_handle(self) -> bytes:
        buf = C.create_string_buffer(self.KEYX_HANDLE_BYTES)
        self._check(self._L.pbk_keyx_pull_handle(self._ctx, buf), "pbk_keyx_pull_handle")
        return buf.raw

    def keyx_pull_connect_ipc(self, src_rank: int, handle: bytes):
        buf = C.create_string_buffer(bytes(handle), self.KEYX_HANDLE_BYTES)
        self._check(self._L.pbk_keyx_pull_connect_ipc(self._ctx, int(src_rank), buf), "pbk_keyx_pull_connect_ipc")

    def keyx_pull_connect_local(self, src_rank: int, peer: "KmerCounter"):
        self._check(self._L.pbk_keyx_pull_connect_local(self._ctx, int(src_rank), peer._ctx), "pbk_keyx_pull_connect_local")

    def keyx_pull_partition(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._L.pbk_keyx_pull_partition(self._ctx, _ptr(bases), _ptr(offsets), len(offsets) - 1, ENC_ASCII, None, None),
                    "pbk_keyx_pull_partition")

    def keyx_pull_partition_ptr(self, bases_ptr: int, offsets_ptr: int, n_reads: int):
        self._check(self._L.pbk_keyx_pull_partition(self._ctx, C.c_void_p(bases_ptr), C.c_void_p(offsets_ptr), n_reads, ENC_ASCII, None, None),
                    "pbk_keyx_pull_partition")

    def keyx_pull_partition_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, asynchronous: bool = False):
        self._check(self._L.pbk_keyx_pull_partition_device(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads, n_bases,
                                                           int(asynchronous)), "pbk_keyx_pull_partition_device")

    def keyx_staged_count_device(self, d_count_ptr: int):
        self._check(self._L.pbk_keyx_staged_count_device(self._ctx, C.c_void_p(d_count_ptr)), "pbk_keyx_staged_count_device")

    def keyx_pull_insert(self):
        self._check(self._L.pbk_keyx_pull_insert(self._ctx), "pbk_keyx_pull_insert")

    # -- Counter<KMER> mirror (reference counter.h) ------------------------------------------------
    def make_kmer_read_distribution(self, bases, offsets, memory_bytes: int) -> int:
        """Counter::makeKmerReadDistributionMT (counter.h:276-383): returns doubleHashSize."""
        self.push_reads(bases, offsets)
        self.finalize()
        self.double_hash_size = int(self._L.pbk_double_hash_size(memory_bytes, self.k))
        return self.double_hash_size

    def get_max_occurrence(self) -> int:
        return self.max_occurrence

    def get_left_local_minimal_value(self, window: int = 1) -> int:
        """Counter::getLeftLocalMinimalValue (counter.h:245-267)."""
        return int(self._L.pbk_left_local_min(_ptr(self.occ_hist), self.max_occurrence, window))

    def coverage_cutoff(self, n_opt: int = 0, repeat: bool = False) -> int:
        """cutoff rule of Assemble::initialKmerAssemble (assemble.cpp:318-321)."""
        return int(self._L.pbk_coverage_cutoff(_ptr(self.occ_hist), self.max_occurrence, n_opt, int(repeat)))

    def _average(self, dist, start, end, what):
        out = C.c_double()
        rc = self._L.pbk_distribution_average(_ptr(dist), len(dist), start, end, C.byref(out))
        if rc:
            raise PbkError(rc, what, "platanus::KmerDistError")
        return out.value

    def calc_occurrence_distribution_average(self, start: int, end: int) -> float:
        """Counter::calcOccurrenceDistributionAverage (counter.h:149-150, 221-238)."""
        return self._average(self.occ_hist, start, end, "calc_occurrence_distribution_average")

    def calc_length_distribution_average(self, start: int = 0, end: int = LEN_BINS - 1) -> float:
        """Counter::calcLengthDistributionAverage (counter.h:146-147)."""
        return self._average(self.len_hist, start, end, "calc_length_distribution_average")

    def output_occurrence_distribution(self, path: str):
        """Counter::outputOccurrenceDistribution (counter.h:1000-1007): PREFIX_<k>merFrq.tsv."""
        rc = self._L.pbk_write_frq_tsv(path.encode(), _ptr(self.occ_hist), self.max_occurrence)
        if rc:
            raise PbkError(rc, "pbk_write_frq_tsv", path)

    def sorted_key_from_kmer_file(self, min_occurrence: int):
        """Counter::sortedKeyFromKmerFile (counter.h:917-951): ascending keys with count >= min."""
        return self.export(min_occurrence, sorted=True)

    def output_occurrence_table_binary(self, path: str, min_occurrence: int, double_hash_size: int = 0) -> int:
        """loadKmer + outputOccurrenceTableBinary (counter.h:600-640, 955-963): PREFIX_kmer_occ.bin.
        Returns loadKmer's return value (the new doubleHashSize)."""
        keys, counts = self.export(min_occurrence, sorted=True)
        out = C.c_uint64()
        rc = self._L.pbk_write_kmer_occ_bin(path.encode(), self.k, _ptr(keys), _ptr(counts), len(counts),
                                            double_hash_size or self.double_hash_size, C.byref(out))
        if rc:
            raise PbkError(rc, "pbk_write_kmer_occ_bin", path)
        return out.value


def microbench_atomics(table_bytes: int, n_ops: int, mode: int, device: int = -1) -> float:
    out = C.c_double()
    rc = load_library().pbk_microbench_atomics(device, table_bytes, n_ops, mode, C.byref(out))
    if rc:
        raise PbkError(rc, "pbk_microbench_atomics")
    return out.value
