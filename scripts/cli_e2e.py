#!/usr/bin/env python
"""Process-level end to end (T_e2e): FASTQ files on disk -> PREFIX_32merFrq.tsv + PREFIX_kmer_occ.bin.
The C++ program of this repo (pbk_assemble) next to the unmodified reference program on the same files and box.
Output: gpurun_out/cli_e2e.json"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O                      # noqa: E402  (checker + reference runner)
from platanus_b_b200 import build as pbuild        # noqa: E402
from platanus_b_b200 import synth                  # noqa: E402

scale = float(os.environ.get("CLI_SCALE", "1"))
k = int(os.environ.get("CLI_K", "32"))
run_ref = os.environ.get("CLI_REF", "1") == "1"
cli = pbuild.build_cli()
rs = synth.make_reads(synth.config("C1", scale=scale))
tmp = tempfile.mkdtemp(prefix="pbk_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
out = {"workload": f"C1 x {scale:g}: {rs.n_reads} reads x {rs.read_len} bp, k={k}", "cores": os.cpu_count()}
try:
    files = synth.write_fastq(rs, os.path.join(tmp, "r_1.fq"), os.path.join(tmp, "r_2.fq"))
    out["fastq_bytes"] = sum(os.path.getsize(f) for f in files)
    # --gpus 1,2: the same program on one GPU and on a pbk_group of two (PBK_NUM_GPUS; the library exchanges the keys itself)
    gpu_counts = [int(x) for x in sys.argv[sys.argv.index("--gpus") + 1].split(",")] if "--gpus" in sys.argv else [1]
    walls = []
    for n_gpus in gpu_counts:
        w = []
        for rep in range(3):
            t0 = time.time()
            p = subprocess.run([cli, "assemble", "-kmer_occ_only", "-k", str(k), "-t", str(min(16, os.cpu_count() or 2)), "-m", "16", "-tmp", tmp, "-o",
                                os.path.join(tmp, "gpu"), "-f", *files], capture_output=True, text=True,
                               env=dict(os.environ, PBK_NUM_GPUS=str(n_gpus), PBK_TIMING="1"))
            w.append(time.time() - t0)
            assert p.returncode == 0, p.stderr
        out[f"pbk_assemble_wall_s_{n_gpus}gpu"] = w
        walls = w if n_gpus == gpu_counts[0] else walls
    out["pbk_assemble_wall_s"] = walls
    out["pbk_assemble_stderr_tail"] = p.stderr.strip().splitlines()[-6:]
    ours = O.read_bin(os.path.join(tmp, "gpu_kmer_occ.bin"))
    out["kept_kmers"] = int(len(ours.counts))
    if run_ref:
        ref = O.run_reference(files, k, tmp, threads=os.cpu_count() or 1, mem_gb=16)
        assert ref.returncode == 0, ref.stderr
        out["reference_wall_s"] = ref.wall_s
        out["reference_cmd"] = f"platanus_b assemble -kmer_occ_only -k {k} -t {os.cpu_count()} -m 16"
        gk, gc = ours.sorted_dump()
        rk, rc = ref.table.sorted_dump()
        out["identical_sorted_dump"] = bool(np.array_equal(gk, rk) and np.array_equal(gc, rc))
        out["identical_tsv"] = open(os.path.join(tmp, f"gpu_{k}merFrq.tsv")).read() == ref.tsv
        out["speedup_wall"] = ref.wall_s / min(walls)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cli_e2e.json"), "w"), indent=1)
print(json.dumps(out))
