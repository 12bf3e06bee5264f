"""platanus_b_b200/host/pbk_seqlib.hpp -- the paired / tagged read ingest of the reference's other
commands (ReadFastaSingleMT, ReadFastaPairMT and the *Tagged* forms, seqlib.cpp:365-742) -- against the UNMODIFIED reference
readers driven by oracle/ref_seqlib_harness.cpp: byte-identical per-thread SEQ temp files, numPair, totalLength and error
conditions, on generated files with the quirks that control the reference's loops (multi-line records, empty lines, no final
newline, N and lower case, odd read counts, mate-pair reversal, BX:Z: tags) and on gzip input.

Where /root/reference is absent (the GPU box) the harness is the prebuilt oracle/_ref/ref_seqlib_harness; committed golden
digests (tests/golden/seqlib_digests.json) pin a fixed set of cases even without it."""
import hashlib
import json
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "seqlib_check.cpp")
EXE = os.path.join(HERE, "host", "_build", "seqlib_check")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_seqlib_harness")
DIGESTS = os.path.join(HERE, "golden", "seqlib_digests.json")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_seqlib_harness is built where /root/reference exists")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "platanus_b_b200", "host", f) for f in ("pbk_seqlib.hpp", "pbk_ingest.hpp", "pbk_counter.hpp")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++11", "-Wall", "-Wextra", "-pthread", "-o", EXE, SRC], check=True)
    return EXE


def make_file(path, rng, n_reads, fastq, tagged=False, final_newline=True, quirks=True):
    tags = ["AAAC", "GGTA", "TTTT", "ACGTACGT"]
    with open(path, "w") as fh:
        for i in range(n_reads):
            seq = "".join(rng.choice("ACGTNacgtn" if quirks else "ACGT") for _ in range(rng.randint(0 if quirks else 20, 120)))
            hdr = f"r{i}" + (f" BX:Z:{rng.choice(tags)}-1" if tagged and rng.random() < 0.85 else "")
            width = rng.choice([200, 200, 37, 11]) if quirks else 200
            lines = [seq[j:j + width] for j in range(0, len(seq), width)] or [""]
            if fastq:
                fh.write(f"@{hdr}\n" + "\n".join(lines) + "\n+\n" + "\n".join("I" * len(x) for x in lines))
            else:
                fh.write(f">{hdr}\n" + "\n".join(lines))
            if i + 1 < n_reads or final_newline:
                fh.write("\n")


def run(binary, mode, T, mate, fastq, not_pair, prefix, files, cwd):
    p = subprocess.run([binary, mode, str(T), str(int(mate)), str(int(fastq)), str(int(not_pair)), prefix, *files], capture_output=True, text=True, cwd=cwd)
    assert p.returncode == 0, p.stderr
    outs = []
    for i in range(T):
        f = f"{prefix}.{i}"
        outs.append(open(f, "rb").read() if os.path.exists(f) else None)
    return p.stdout.strip(), outs


CASES = [  # mode, n_reads (file 1, file 2), fastq, mate, not_pair, T, final newline
    ("single", (40, 0), True, False, False, 3, True), ("single", (41, 0), True, False, False, 2, True), ("single", (41, 0), False, True, True, 4, False),
    ("single", (0, 0), True, False, False, 2, True), ("pair", (30, 30), True, True, False, 3, True), ("pair", (30, 31), False, False, False, 2, True),
    ("pair", (25, 25), False, False, False, 1, False), ("single_tagged", (40, 0), True, False, False, 3, True),
    ("pair_tagged", (30, 30), False, True, False, 2, True), ("single_tagged", (31, 0), False, False, True, 2, False),
]


def _case_files(tmp, idx, case, seed):
    mode, (n1, n2), fastq, mate, not_pair, T, final_nl = case
    rng = random.Random(1000 * seed + idx)
    ext = "fq" if fastq else "fa"
    files = [os.path.join(tmp, f"c{idx}_1.{ext}")]
    make_file(files[0], rng, n1, fastq, "tagged" in mode, final_nl)
    if mode.startswith("pair"):
        files.append(os.path.join(tmp, f"c{idx}_2.{ext}"))
        make_file(files[1], rng, n2, fastq, "tagged" in mode, final_nl)
    return files


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_readers_match_the_reference_byte_for_byte(exe, seed, tmp_path):
    tmp = str(tmp_path)
    for idx, case in enumerate(CASES):
        mode, _, fastq, mate, not_pair, T, _ = case
        files = _case_files(tmp, idx, case, seed)
        want = run(REF, mode, T, mate, fastq, not_pair, os.path.join(tmp, f"ref{idx}"), files, tmp)
        if "tagged" in mode and os.path.exists(os.path.join(tmp, f"ref{idx}.tags")):
            os.replace(os.path.join(tmp, f"ref{idx}.tags"), os.path.join(tmp, f"our{idx}.tags"))
        got = run(exe, mode, T, mate, fastq, not_pair, os.path.join(tmp, f"our{idx}"), files, tmp)
        assert got[0] == want[0], (case, got[0], want[0])
        assert got[1] == want[1], case


@needs_ref
def test_gzip_input_and_a_larger_file(exe, tmp_path):
    tmp = str(tmp_path)
    rng = random.Random(7)
    f1, f2 = os.path.join(tmp, "big_1.fq"), os.path.join(tmp, "big_2.fq")
    make_file(f1, rng, 20000, True, quirks=False)
    make_file(f2, rng, 20000, True, quirks=False)
    want = run(REF, "pair", 5, False, True, False, os.path.join(tmp, "ref"), [f1, f2], tmp)
    got = run(exe, "pair", 5, False, True, False, os.path.join(tmp, "our"), [f1, f2], tmp)
    assert got == want and want[0].startswith("numPair 20000 ")
    subprocess.run(["gzip", "-k", f1, f2], check=True)            # the reference cannot sniff compression here (no `file` utility): ours only
    gz = run(exe, "pair", 5, False, True, False, os.path.join(tmp, "gz"), [f1 + ".gz", f2 + ".gz"], tmp)
    assert gz == want


def _digest(result):
    h = hashlib.sha256(result[0].encode())
    for o in result[1]:
        h.update(b"|" if o is None else hashlib.sha256(o).digest())
    return h.hexdigest()


def test_committed_digests_of_reference_outputs(exe, tmp_path):
    """the same cases with seed 0 against digests of the reference's outputs committed under tests/golden (generated by this test
    file where the harness exists: PBK_WRITE_SEQLIB_DIGESTS=1) -- runs on any box"""
    tmp = str(tmp_path)
    digests = {}
    for idx, case in enumerate(CASES):
        mode, _, fastq, mate, not_pair, T, _ = case
        files = _case_files(tmp, idx, case, 0)
        if os.environ.get("PBK_WRITE_SEQLIB_DIGESTS") and os.path.exists(REF):
            digests[str(idx)] = _digest(run(REF, mode, T, mate, fastq, not_pair, os.path.join(tmp, f"ref{idx}"), files, tmp))
        if "tagged" in mode:                                      # the ids setTagStringConverter gave the tags of these files
            tagfile = os.path.join(HERE, "golden", f"seqlib_tags_{idx}.tsv")
            produced = os.path.join(tmp, f"ref{idx}.tags")
            if os.environ.get("PBK_WRITE_SEQLIB_DIGESTS") and os.path.exists(produced):
                os.replace(produced, tagfile)
            with open(os.path.join(tmp, f"our{idx}.tags"), "w") as dst:
                dst.write(open(tagfile).read() if os.path.exists(tagfile) else "")
    if digests:
        json.dump(digests, open(DIGESTS, "w"), indent=1)
    want = json.load(open(DIGESTS))
    for idx, case in enumerate(CASES):
        mode, _, fastq, mate, not_pair, T, _ = case
        files = [os.path.join(tmp, os.path.basename(f)) for f in _case_files(tmp, idx, case, 0)]
        got = run(exe, mode, T, mate, fastq, not_pair, os.path.join(tmp, f"our{idx}"), files, tmp)
        assert _digest(got) == want[str(idx)], case
