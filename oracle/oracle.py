"""ctypes wrapper around oracle/kmer_oracle.c -- TEST INFRASTRUCTURE ONLY.

Nothing in platanus_b_b200/ imports this module.  It is used by tests/, by bench.py's
cpu_baseline / ``--impl reference`` leg and by __graft_entry__.smoke() as the checker.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libpbk_oracle.so")
REF_BINARY = os.path.join(HERE, "_ref", "platanus_b")

OCC_BINS = 65535
MAX_READ_LEN = 500000
COUNT_SAT = 65534


class _Reads(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("offsets", C.c_void_p), ("n_reads", C.c_uint64),
                ("cap_bases", C.c_uint64), ("cap_reads", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("k", C.c_uint), ("words", C.c_uint), ("n_distinct", C.c_uint64),
                ("n_instances", C.c_uint64), ("keys", C.POINTER(C.c_uint64)),
                ("counts", C.POINTER(C.c_uint16)), ("occ_hist", C.c_uint64 * OCC_BINS),
                ("len_hist", C.POINTER(C.c_uint64)), ("max_occ", C.c_uint64)]


class _Bin(C.Structure):
    _fields_ = [("k", C.c_uint64), ("index_size", C.c_uint64), ("words", C.c_uint),
                ("n", C.c_uint64), ("slots", C.POINTER(C.c_uint64)),
                ("keys", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint16))]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) if needed; returns the .so path."""
    src = os.path.join(HERE, "kmer_oracle.c")
    hdr = os.path.join(HERE, "kmer_oracle.h")
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "liboracle"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.pbo_reads_new.restype = C.POINTER(_Reads)
        L.pbo_reads_free.argtypes = [C.POINTER(_Reads)]
        L.pbo_reads_add.argtypes = [C.POINTER(_Reads), C.c_char_p, C.c_uint64]
        L.pbo_reads_add_file.argtypes = [C.POINTER(_Reads), C.c_char_p]
        L.pbo_check_file_format.argtypes = [C.c_char_p]
        L.pbo_char2bin.argtypes = [C.c_char]
        L.pbo_char2bin.restype = C.c_ubyte
        L.pbo_count.argtypes = [C.POINTER(_Reads), C.c_uint, C.POINTER(_Result)]
        L.pbo_result_release.argtypes = [C.POINTER(_Result)]
        L.pbo_count_seeded.argtypes = [C.POINTER(_Reads), C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_Result)]
        L.pbo_count_contigs.argtypes = [C.POINTER(_Reads), C.c_uint, C.c_void_p, C.c_uint64, C.POINTER(_Result)]
        L.pbo_match_reads.argtypes = [C.POINTER(_Reads), C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.pbo_occurrence_array.argtypes = [C.POINTER(_Reads), C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.pbo_neighbor_flags.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.pbo_left_local_min.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.pbo_left_local_min.restype = C.c_uint64
        L.pbo_dist_average.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
        L.pbo_coverage_cutoff.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.pbo_coverage_cutoff.restype = C.c_uint64
        for f in (L.pbo_pair_size, L.pbo_key_raw_size):
            f.argtypes = [C.c_uint]
            f.restype = C.c_uint64
        L.pbo_double_hash_size.argtypes = [C.c_uint64, C.c_uint]
        L.pbo_double_hash_size.restype = C.c_uint64
        L.pbo_load_size.argtypes = [C.c_uint64]
        L.pbo_load_size.restype = C.c_uint64
        L.pbo_write_bin.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.pbo_read_bin.argtypes = [C.c_char_p, C.POINTER(_Bin)]
        L.pbo_bin_release.argtypes = [C.POINTER(_Bin)]
        L.pbo_bin_check_reachable.argtypes = [C.POINTER(_Bin)]
        L.pbo_write_tsv.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"oracle: {what} failed with code {code}")
        self.code = code


@dataclass
class CountResult:
    k: int
    words: int
    n_instances: int
    keys: np.ndarray        # (n_distinct, words) uint64, word 0 first, sorted in reference order
    counts: np.ndarray      # (n_distinct,) uint16
    occ_hist: np.ndarray    # (65535,) uint64
    len_hist: np.ndarray    # (500001,) uint64
    max_occ: int

    @property
    def n_distinct(self) -> int:
        return int(self.keys.shape[0])


class Reads:
    """Raw read strings as the reference parser hands them to SEQ::convertFromString."""

    def __init__(self):
        self._p = lib().pbo_reads_new()

    def __del__(self):
        if getattr(self, "_p", None):
            lib().pbo_reads_free(self._p)
            self._p = None

    def add(self, seq: bytes) -> None:
        rc = lib().pbo_reads_add(self._p, seq, len(seq))
        if rc:
            raise OracleError(rc, "pbo_reads_add")

    def add_file(self, path: str) -> None:
        rc = lib().pbo_reads_add_file(self._p, path.encode())
        if rc:
            raise OracleError(rc, "pbo_reads_add_file")

    def add_array(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        b = np.ascontiguousarray(bases, dtype=np.uint8).tobytes()
        for i in range(len(offsets) - 1):
            self.add(b[int(offsets[i]):int(offsets[i + 1])])

    @property
    def n_reads(self) -> int:
        return int(self._p.contents.n_reads)

    def arrays(self):
        """(bases uint8[total], offsets uint64[n_reads + 1]) copies."""
        r = self._p.contents
        n = int(r.n_reads)
        offs = np.ctypeslib.as_array(C.cast(r.offsets, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        total = int(offs[-1])
        if total:
            bases = np.ctypeslib.as_array(C.cast(r.bases, C.POINTER(C.c_uint8)), shape=(total,)).copy()
        else:
            bases = np.zeros(0, np.uint8)
        return bases, offs


def check_file_format(path: str) -> int:
    return lib().pbo_check_file_format(path.encode())


def char2bin(c: int) -> int:
    return lib().pbo_char2bin(bytes([c & 0xFF]))


def count(reads: Reads, k: int, seed_keys=None, seed_counts=None) -> CountResult:
    """makeKmerReadDistributionMT; with seeds (sorted (key, value) dump of the table the counter already holds):
    makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750)."""
    res = _Result()
    if seed_keys is None:
        rc = lib().pbo_count(reads._p, k, C.byref(res))
    else:
        sk = np.ascontiguousarray(seed_keys, dtype=np.uint64)
        sc = np.ascontiguousarray(seed_counts, dtype=np.uint16)
        rc = lib().pbo_count_seeded(reads._p, k, sk.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), len(sc), C.byref(res))
    if rc:
        raise OracleError(rc, "pbo_count")
    try:
        nd, w = int(res.n_distinct), int(res.words)
        keys = (np.ctypeslib.as_array(res.keys, shape=(nd * w,)).copy().reshape(nd, w)
                if nd else np.zeros((0, w), np.uint64))
        counts = np.ctypeslib.as_array(res.counts, shape=(nd,)).copy() if nd else np.zeros(0, np.uint16)
        occ = np.frombuffer(bytes(res.occ_hist), dtype=np.uint64).copy()
        lh = np.ctypeslib.as_array(res.len_hist, shape=(MAX_READ_LEN + 1,)).copy()
        return CountResult(int(res.k), w, int(res.n_instances), keys, counts, occ, lh, int(res.max_occ))
    finally:
        lib().pbo_result_release(C.byref(res))


def occurrence_array(seqs: Reads, k: int, keys: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """ContigDivider::getOccurrenceArray (kmer_divide.cpp:151-197): u16 per base of `seqs` (window start indexed), looked
    up in the table given as its sorted dump (keys [n, words] ascending in reference order, counts [n])."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    _, offs = seqs.arrays()
    out = np.zeros(int(offs[-1]), np.uint16)
    rc = lib().pbo_occurrence_array(seqs._p, k, keys.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                    len(counts), out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc, "pbo_occurrence_array")
    return out


def neighbor_flags(k: int, keys: np.ndarray, counts: np.ndarray, min_count: int = 1) -> np.ndarray:
    """The eight findValue probes of makeInitialBruijnGraph (graph.h:337-375) per key of a sorted table: u8 per key,
    (leftFlags << 4) | rightFlags; keys below min_count are not in the table."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    out = np.zeros(len(counts), np.uint8)
    rc = lib().pbo_neighbor_flags(k, keys.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), len(counts), int(min_count),
                                  out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc, "pbo_neighbor_flags")
    return out


def count_contigs(seqs: Reads, k: int, coverage: np.ndarray, min_occurrence: int = 1) -> CountResult:
    """Counter::makeKmerReadDistributionFromContig (counter.h:511-593)."""
    cov = np.ascontiguousarray(coverage, dtype=np.uint16)
    res = _Result()
    rc = lib().pbo_count_contigs(seqs._p, k, cov.ctypes.data_as(C.c_void_p), int(min_occurrence), C.byref(res))
    if rc:
        raise OracleError(rc, "pbo_count_contigs")
    try:
        nd, w = int(res.n_distinct), int(res.words)
        keys = (np.ctypeslib.as_array(res.keys, shape=(nd * w,)).copy().reshape(nd, w) if nd else np.zeros((0, w), np.uint64))
        counts = np.ctypeslib.as_array(res.counts, shape=(nd,)).copy() if nd else np.zeros(0, np.uint16)
        occ = np.frombuffer(bytes(res.occ_hist), dtype=np.uint64).copy()
        lh = np.ctypeslib.as_array(res.len_hist, shape=(MAX_READ_LEN + 1,)).copy()
        return CountResult(int(res.k), w, 0, keys, counts, occ, lh, int(res.max_occ))
    finally:
        lib().pbo_result_release(C.byref(res))


def run_ref_contig(k: int, contigs_fa: str, min_occurrence: int, workdir: str):
    """oracle/ref_iter_harness.cpp, mode contig: sorted keys, counts, maxOccurrence of makeKmerReadDistributionFromContig."""
    import subprocess
    out = os.path.join(workdir, "contig_out")
    p = subprocess.run([REF_ITER, "contig", str(k), contigs_fa, str(min_occurrence), out], check=True, capture_output=True, text=True, cwd=workdir)
    w = (k + 31) // 32
    raw = np.fromfile(out, dtype=np.uint8).reshape(-1, 8 * w + 2)
    keys = np.ascontiguousarray(raw[:, :8 * w]).view(np.uint64).reshape(-1, w)
    counts = np.ascontiguousarray(raw[:, 8 * w:]).view(np.uint16).reshape(-1)
    order = np.lexsort(tuple(keys[:, j] for j in range(w)))
    return keys[order], counts[order], int(p.stdout.split()[-1])


def match_reads(reads: Reads, k: int, keys: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Counter::pickupReadMatchedEdgeKmer (counter.h:870-910): bool per read."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    out = np.zeros(max(reads.n_reads, 1), np.uint8)
    rc = lib().pbo_match_reads(reads._p, k, keys.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), len(counts),
                               out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc, "pbo_match_reads")
    return out[:reads.n_reads].astype(bool)


REF_ITER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_iter_harness")


def run_ref_iter(mode: str, bin_path: str, reads, n_threads: int, workdir: str, k: int):
    """oracle/ref_iter_harness.cpp: the UNMODIFIED reference's pickupReadMatchedEdgeKmer ('pickup' -> list of kept reads)
    or makeKmerReadDistributionConsideringPreviousGraph ('count' -> sorted keys, counts, maxOccurrence)."""
    import subprocess
    rp = os.path.join(workdir, "iter_reads.txt")
    with open(rp, "w") as fh:
        fh.write("".join(r + "\n" for r in reads))
    out = os.path.join(workdir, "iter_out_" + mode)
    p = subprocess.run([REF_ITER, mode, bin_path, rp, str(n_threads), out], check=True, capture_output=True, text=True, cwd=workdir)
    if mode == "pickup":
        return open(out).read().split("\n")[:-1]
    w = (k + 31) // 32
    raw = np.fromfile(out, dtype=np.uint8).reshape(-1, 8 * w + 2)
    keys = np.ascontiguousarray(raw[:, :8 * w]).view(np.uint64).reshape(-1, w)
    counts = np.ascontiguousarray(raw[:, 8 * w:]).view(np.uint16).reshape(-1)
    order = np.lexsort(tuple(keys[:, j] for j in range(w)))
    return keys[order], counts[order], int(p.stdout.split()[-1])


REF_OCC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_occ_harness")


def run_ref_occurrence(bin_path: str, contigs_fa: str, workdir: str):
    """oracle/ref_occ_harness.cpp: the UNMODIFIED reference's getOccurrenceArray + dumpKmerCoverage.  Returns
    [(name, np.uint64 array)] per contig."""
    import subprocess
    out = os.path.join(workdir, "ref_occ.txt")
    subprocess.run([REF_OCC, bin_path, contigs_fa, out], check=True, capture_output=True)
    res, name, vals = [], None, []
    for ln in open(out):
        ln = ln.strip()
        if ln.startswith(">"):
            if name is not None:
                res.append((name, np.array(vals, np.uint64)))
            name, vals = ln[1:], []
        elif ln:
            vals.append(int(ln))
    if name is not None:
        res.append((name, np.array(vals, np.uint64)))
    return res


def _u64ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def left_local_min(occ_hist: np.ndarray, max_occ: int, window: int = 1) -> int:
    occ = np.ascontiguousarray(occ_hist, dtype=np.uint64)
    return int(lib().pbo_left_local_min(_u64ptr(occ), max_occ, window))


def dist_average(dist: np.ndarray, start: int, end: int) -> float:
    d = np.ascontiguousarray(dist, dtype=np.uint64)
    out = C.c_double()
    rc = lib().pbo_dist_average(_u64ptr(d), len(d), start, end, C.byref(out))
    if rc:
        raise OracleError(rc, "pbo_dist_average")
    return out.value


def coverage_cutoff(occ_hist: np.ndarray, max_occ: int, n_opt: int = 0, repeat: bool = False) -> int:
    occ = np.ascontiguousarray(occ_hist, dtype=np.uint64)
    return int(lib().pbo_coverage_cutoff(_u64ptr(occ), max_occ, n_opt, int(repeat)))


def pair_size(k: int) -> int:
    return int(lib().pbo_pair_size(k))


def key_raw_size(k: int) -> int:
    return int(lib().pbo_key_raw_size(k))


def double_hash_size(memory_bytes: int, k: int) -> int:
    return int(lib().pbo_double_hash_size(memory_bytes, k))


def load_size(total: int) -> int:
    return int(lib().pbo_load_size(total))


def write_bin(path: str, k: int, keys: np.ndarray, counts: np.ndarray, min_count: int, dh_size: int) -> None:
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    rc = lib().pbo_write_bin(path.encode(), k, _u64ptr(keys), _u64ptr(counts), len(counts), min_count, dh_size)
    if rc:
        raise OracleError(rc, "pbo_write_bin")


@dataclass
class BinTable:
    k: int
    index_size: int
    slots: np.ndarray
    keys: np.ndarray     # (n, words), file (slot) order
    counts: np.ndarray
    reachable: bool

    def sorted_dump(self):
        """(keys, counts) in the reference's numeric key order (top word first)."""
        order = np.lexsort(tuple(self.keys[:, w] for w in range(self.keys.shape[1])))
        return self.keys[order], self.counts[order]


def read_bin(path: str) -> BinTable:
    b = _Bin()
    rc = lib().pbo_read_bin(path.encode(), C.byref(b))
    if rc:
        raise OracleError(rc, "pbo_read_bin")
    try:
        n, w = int(b.n), int(b.words)
        slots = np.ctypeslib.as_array(b.slots, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        keys = (np.ctypeslib.as_array(b.keys, shape=(n * w,)).copy().reshape(n, w)
                if n else np.zeros((0, w), np.uint64))
        counts = np.ctypeslib.as_array(b.counts, shape=(n,)).copy() if n else np.zeros(0, np.uint16)
        ok = lib().pbo_bin_check_reachable(C.byref(b)) == 0
        return BinTable(int(b.k), int(b.index_size), slots, keys, counts, ok)
    finally:
        lib().pbo_bin_release(C.byref(b))


def write_tsv(path: str, occ_hist: np.ndarray, max_occ: int) -> None:
    occ = np.ascontiguousarray(occ_hist, dtype=np.uint64)
    rc = lib().pbo_write_tsv(path.encode(), _u64ptr(occ), max_occ)
    if rc:
        raise OracleError(rc, "pbo_write_tsv")


def tsv_text(occ_hist: np.ndarray, max_occ: int) -> str:
    return "".join(f"{i}\t{int(occ_hist[i])}\n" for i in range(1, max_occ + 1))


# ------------------------------------------------------------------------------------------------
# the unmodified reference binary (oracle/_ref/platanus_b), when it has been built
# ------------------------------------------------------------------------------------------------

def have_ref_binary() -> bool:
    return os.path.exists(REF_BINARY) and os.access(REF_BINARY, os.X_OK)


@dataclass
class RefRun:
    returncode: int
    stderr: str
    cutoff: int | None
    ave_read_len: str | None
    kmer_coverage: str | None
    tsv: str | None
    table: BinTable | None
    wall_s: float


def run_reference(files, k: int, workdir: str, threads: int = 1, mem_gb: int = 1, n_opt: int = 0,
                  repeat: bool = False, prefix: str = "ref", parse_bin: bool = True) -> RefRun:
    """`platanus_b assemble -kmer_occ_only` (main.cpp:70 -> assemble.cpp:140) on `files`."""
    import re
    import time
    cmd = [REF_BINARY, "assemble", "-kmer_occ_only", "-k", str(k), "-t", str(threads), "-m", str(mem_gb),
           "-tmp", workdir, "-o", os.path.join(workdir, prefix), "-f", *files]
    if n_opt:
        cmd += ["-n", str(n_opt)]
    if repeat:
        cmd += ["-repeat"]
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=workdir)
    wall = time.time() - t0
    err = p.stderr
    m = re.search(r"^K=%d, KMER_COVERAGE=(\S+) \(>= (\d+)\), COVERAGE_CUTOFF=(\d+)" % k, err, re.M)
    a = re.search(r"^AVE_READ_LEN=(\S+)", err, re.M)
    tsv_path = os.path.join(workdir, f"{prefix}_{k}merFrq.tsv")
    bin_path = os.path.join(workdir, f"{prefix}_kmer_occ.bin")
    tsv = open(tsv_path).read() if os.path.exists(tsv_path) else None
    table = read_bin(bin_path) if (parse_bin and os.path.exists(bin_path)) else None
    return RefRun(p.returncode, err, int(m.group(3)) if m else None, a.group(1) if a else None,
                  m.group(1) if m else None, tsv, table, wall)
