"""Seeded synthetic read sets for the BASELINE.json configs.

One generator feeds the CUDA path, the oracle and the reference binary, so every arm of a
comparison sees byte-identical reads.  Reads are fixed-length, so a read set is a dense
``uint8[n_reads, read_len]`` matrix of ASCII characters (mate-1 block first, then the mate-2 block,
i.e. the order of ``-f r_1.fq r_2.fq``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


@dataclass(frozen=True)
class ReadSetSpec:
    name: str
    genome_lengths: tuple          # one entry per genome
    gc: tuple                      # GC fraction per genome
    genome_seeds: tuple
    abundances: tuple              # relative depth per genome (mean 1.0)
    read_len: int
    coverage: float                # mean depth over all genomes
    insert: int
    sub_rate: float
    n_rate: float
    read_seed: int
    k: int = 32
    repeat_mode: bool = False
    # C5: copies of an operon planted into genome 0: (n_copies, operon_len, identity)
    operon: tuple | None = None

    @property
    def total_genome(self) -> int:
        return int(sum(self.genome_lengths))

    @property
    def n_pairs(self) -> int:
        return int(self.total_genome * self.coverage / (2 * self.read_len))


def config(name: str, scale: float = 1.0) -> ReadSetSpec:
    """The five BASELINE.json configs; ``scale`` < 1 shrinks genome sizes (coverage is kept)."""
    s = lambda n: max(2000, int(n * scale))
    if name in ("C1", "C2"):
        return ReadSetSpec(name, (s(4_600_000),), (0.5,), (42,), (1.0,), 150, 100.0, 400, 0.005, 0.0001, 1042)
    if name == "C3":
        return ReadSetSpec(name, (s(7_000_000),), (0.70,), (43,), (1.0,), 250, 300.0, 600, 0.01, 0.0, 1043)
    if name == "C4":
        rng = np.random.default_rng(4)
        lens = rng.uniform(3e6, 7e6, 20)
        lens = np.round(lens * (100e6 / lens.sum())).astype(np.int64)
        gcs = rng.uniform(0.3, 0.7, 20)
        ab = rng.lognormal(0.0, 1.0, 20)
        ab = ab * (lens.sum() / (ab * lens).sum())          # length-weighted mean depth 1.0
        return ReadSetSpec(name, tuple(s(int(x)) for x in lens), tuple(float(g) for g in gcs),
                           tuple(range(100, 120)), tuple(float(a) for a in ab), 150, 200.0, 400, 0.005, 0.0, 1044)
    if name == "C5":
        return ReadSetSpec(name, (s(1_000_000),), (0.5,), (44,), (1.0,), 150, 1000.0, 400, 0.005, 0.0, 1045,
                           repeat_mode=True, operon=(7, min(5000, s(1_000_000) // 20), 0.995))
    raise ValueError(f"unknown config {name!r}")


def make_genome(length: int, gc: float, seed: int, operon: tuple | None = None) -> np.ndarray:
    """i.i.d. bases as codes 0..3 (A C G T) with the given GC fraction."""
    rng = np.random.default_rng(seed)
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    g = rng.choice(4, size=length, p=p).astype(np.uint8)
    if operon is not None:
        copies, olen, ident = operon
        op = rng.choice(4, size=olen, p=p).astype(np.uint8)
        starts = np.sort(rng.choice(max(1, length // olen - 1), size=copies, replace=False)) * olen
        for st in starts:
            c = op.copy()
            nmut = rng.binomial(olen, 1.0 - ident)
            pos = rng.integers(0, olen, nmut)
            c[pos] = (c[pos] + rng.integers(1, 4, nmut)) % 4
            g[st:st + olen] = c
    return g


@dataclass
class ReadSet:
    spec: ReadSetSpec
    reads: np.ndarray                       # uint8 [n_reads, read_len], ASCII
    meta: dict = field(default_factory=dict)

    @property
    def n_reads(self) -> int:
        return int(self.reads.shape[0])

    @property
    def read_len(self) -> int:
        return int(self.reads.shape[1])

    def flat(self):
        """(bases uint8[n_reads*read_len], offsets uint64[n_reads+1])"""
        n, L = self.reads.shape
        return self.reads.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))

    def n_windows(self, k: int) -> int:
        """upper bound on k-mer instances (windows containing N not subtracted)"""
        return self.n_reads * max(0, self.read_len - k + 1)


def _sample_pairs(genome: np.ndarray, n_pairs: int, L: int, insert: int, rng, chunk: int = 1 << 18):
    G = genome.shape[0]
    insert = min(insert, G)
    Lr = min(L, insert)
    m1 = np.empty((n_pairs, L), dtype=np.uint8)
    m2 = np.empty((n_pairs, L), dtype=np.uint8)
    ar = np.arange(Lr, dtype=np.int64)
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        st = rng.integers(0, G - insert + 1, hi - lo)
        m1[lo:hi, :Lr] = genome[st[:, None] + ar[None, :]]
        # mate 2 = reverse complement of the fragment's far end
        far = genome[(st + insert - 1)[:, None] - ar[None, :]]
        m2[lo:hi, :Lr] = 3 - far
        if Lr < L:                                   # tiny test genomes only
            m1[lo:hi, Lr:] = 0
            m2[lo:hi, Lr:] = 0
    return m1, m2


def make_reads(spec: ReadSetSpec, max_pairs: int | None = None, pair_slice: tuple | None = None) -> ReadSet:
    """Generate the read set.  ``pair_slice=(i, n)`` keeps the i-th of n contiguous slices of the
    pairs of every genome (used to give each rank its share); ``max_pairs`` truncates (bounded CPU
    samples).  Both are deterministic functions of the spec."""
    rng = np.random.default_rng([spec.read_seed, 7] + (list(pair_slice) if pair_slice else []))
    L = spec.read_len
    blocks1, blocks2 = [], []
    weights = np.array(spec.genome_lengths, dtype=np.float64) * np.array(spec.abundances)
    weights /= weights.sum()
    total_pairs = spec.n_pairs
    for gi, glen in enumerate(spec.genome_lengths):
        npairs = int(round(total_pairs * weights[gi]))
        genome = make_genome(glen, spec.gc[gi], spec.genome_seeds[gi], spec.operon if gi == 0 else None)
        grng = np.random.default_rng([spec.read_seed, gi])
        if pair_slice is not None:                   # every rank draws its own share of each genome
            i, n = pair_slice
            grng = np.random.default_rng([spec.read_seed, gi, i, n])
            npairs = npairs // n + (1 if i < npairs % n else 0)
        m1, m2 = _sample_pairs(genome, npairs, L, spec.insert, grng)
        blocks1.append(m1)
        blocks2.append(m2)
    m1 = np.concatenate(blocks1) if len(blocks1) > 1 else blocks1[0]
    m2 = np.concatenate(blocks2) if len(blocks2) > 1 else blocks2[0]
    if max_pairs is not None:
        m1, m2 = m1[:max_pairs], m2[:max_pairs]
    codes = np.concatenate([m1, m2])
    del m1, m2
    flat = codes.reshape(-1)
    total = flat.shape[0]
    # substitutions: exact-count binomial, uniform positions
    nsub = rng.binomial(total, spec.sub_rate) if spec.sub_rate > 0 else 0
    if nsub:
        pos = rng.integers(0, total, nsub)
        flat[pos] = (flat[pos] + rng.integers(1, 4, nsub).astype(np.uint8)) % 4
    ascii_reads = _ACGT[codes]
    nn = rng.binomial(total, spec.n_rate) if spec.n_rate > 0 else 0
    if nn:
        pos = rng.integers(0, total, nn)
        ascii_reads.reshape(-1)[pos] = ord("N")
    # the reference sniffs line 2 of each file: it must be uppercase ACGTN only -- always true here
    return ReadSet(spec, ascii_reads, {"n_sub": int(nsub), "n_N": int(nn)})


def write_fastq(rs: ReadSet, path1: str, path2: str | None = None) -> list:
    """Uncompressed FASTQ with constant quality 'I'.  With ``path2`` the first half of the reads
    (mate 1) goes to path1 and the second half to path2, as `-f r_1.fq r_2.fq`."""
    def dump(block: np.ndarray, path: str, tag: int, first_index: int):
        n, L = block.shape
        hdr = 12
        rec = np.empty((n, hdr + 1 + L + 3 + L + 1), dtype=np.uint8)
        rec[:, 0] = ord("@")
        idx = np.arange(first_index, first_index + n, dtype=np.int64)
        for d in range(9):
            rec[:, 9 - d] = ord("0") + (idx // 10 ** d) % 10
        rec[:, 10] = ord("/")
        rec[:, 11] = ord("0") + tag
        rec[:, hdr] = ord("\n")
        rec[:, hdr + 1:hdr + 1 + L] = block
        rec[:, hdr + 1 + L:hdr + 4 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
        rec[:, hdr + 4 + L:hdr + 4 + 2 * L] = ord("I")
        rec[:, -1] = ord("\n")
        rec.tofile(path)

    n = rs.n_reads
    if path2 is None:
        dump(rs.reads, path1, 1, 0)
        return [path1]
    half = n // 2
    dump(rs.reads[:half], path1, 1, 0)
    dump(rs.reads[half:], path2, 2, 0)
    return [path1, path2]
