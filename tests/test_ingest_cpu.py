"""platanus_b_b200/host/pbk_ingest.hpp -- the range-parallel FASTA/FASTQ parser of pbk_assemble -- against the oracle's
restatement of the reference's serial parsers (assemble.cpp:816-848, 902-942): same reads, in the same order, for any
number of ranges, including the quirks (quality lines starting with '@', empty lines, multi-line records, a header
at the very end, files without any header, the unconditional last flush)."""
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "ingest_check.cpp")
EXE = os.path.join(HERE, "host", "_build", "ingest_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    hdr = os.path.join(ROOT, "platanus_b_b200", "host", "pbk_ingest.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Wextra", "-o", EXE, SRC], check=True)
    return EXE


def oracle_reads(O, path):
    rd = O.Reads()
    rd.add_file(path)
    b, o = rd.arrays()
    return [bytes(b[int(o[i]):int(o[i + 1])]) for i in range(len(o) - 1)]


def ours(exe, path, kind, T):
    out = subprocess.run([exe, path, kind, str(T)], capture_output=True, check=True).stdout
    lines = out.split(b"\n")
    assert lines[-1] == b""
    return lines[:-1]


@pytest.mark.parametrize("name", ["small.fq", "small.fa", "cov.fq", "kat.fa", "tail_header.fa", "empty.fa"])
def test_golden_inputs_any_number_of_ranges(oracle, exe, name):
    path = os.path.join(HERE, "golden", "inputs", name)
    want = oracle_reads(oracle, path)
    for T in (1, 2, 3, 7, 64):
        assert ours(exe, path, "fq" if name.endswith("fq") else "fa", T) == want, (name, T)


@pytest.mark.parametrize("seed", range(12))
def test_random_files_with_quirks(oracle, exe, seed, tmp_path):
    rng = random.Random(seed)
    fastq = seed % 2 == 0
    path = str(tmp_path / ("x.fq" if fastq else "x.fa"))
    seqs = ["".join(rng.choice("ACGTNacgtn") for _ in range(rng.randint(0, 90))) for _ in range(rng.randint(1, 60))]
    seqs[0] = "".join(rng.choice("ACGTN") for _ in range(rng.randint(1, 90)))     # checkFileFormat: line 2 uppercase ACGTN
    with open(path, "w") as fh:
        for i, s in enumerate(seqs):
            if fastq:
                fh.write(f"@r{i}\n")
                w = 200 if i == 0 else rng.choice([200, 30, 11])
                for j in range(0, len(s), w):                      # multi-line sequence
                    fh.write(s[j:j + w] + "\n")
                fh.write("+\n")
                q = "".join(rng.choice("@+I5#") for _ in s)          # quality lines may start with '@' or '+'
                for j in range(0, len(q), w):
                    fh.write(q[j:j + w] + "\n")
                if rng.random() < 0.1:
                    fh.write("\n")
            else:
                fh.write(f">r{i}\n")
                w = 200 if i == 0 else rng.choice([200, 25, 7])
                for j in range(0, len(s), w):
                    fh.write(s[j:j + w] + "\n")
                if rng.random() < 0.1:
                    fh.write("\n")
        if rng.random() < 0.3:
            fh.write("@tail" if fastq else ">tail")                  # header at the very end, no newline
        elif rng.random() < 0.3 and seqs:
            fh.write("ACGT")                                         # last line without newline
    want = oracle_reads(oracle, path)
    for T in (1, 2, 5, 13, 64):
        assert ours(exe, path, "fq" if fastq else "fa", T) == want, (seed, T)


@pytest.mark.parametrize("prog", ["gzip", "bzip2"])
@pytest.mark.parametrize("name", ["small.fq", "kat.fa", "tail_header.fa"])
def test_compressed_inputs(oracle, exe, name, prog, tmp_path):
    """readFastaCompressed / readFastqCompressed (assemble.cpp:851-885, 945-986): the same reads as from the plain file;
    the type is recognised from the magic number, whatever the file is called."""
    import shutil
    if shutil.which(prog) is None:
        pytest.skip(f"{prog} not installed")
    plain = os.path.join(HERE, "golden", "inputs", name)
    packed = str(tmp_path / ("reads_" + name))                          # no .gz/.bz2 suffix on purpose
    with open(packed, "wb") as out:
        subprocess.run([prog, "-c", plain], stdout=out, check=True)
    want = oracle_reads(oracle, plain)
    for T in (1, 3):
        got = subprocess.run([exe, packed, "fq" if name.endswith("fq") else "fa", str(T), str(tmp_path)],
                             capture_output=True, check=True).stdout.split(b"\n")[:-1]
        assert got == want
    assert not [f for f in os.listdir(tmp_path) if not f.startswith("reads_")]      # the temp file is unlinked
