#!/usr/bin/env python
"""bench.py -- canonical k-mers counted per second at k=32 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1] [--scale s]

A "step" is one complete pass of the hot path over one batch of synthetic reads: fresh table ->
pack (ASCII -> 2-bit) -> rolling canonical k-mer extraction + table insert -> clamp + occurrence
histogram (+ hash-range exchange and histogram all-reduce when N > 1).

  value  whole-job k-mer instances/s with the reads already resident in HBM (pbk_push_reads_device)
  e2e    the same through the reference-facing C ABI with HOST buffers: H2D of the reads from pinned
         memory, D2H of the occurrence histogram, the coverage cutoff, and the sorted (key, count)
         entries >= cutoff back in host memory (pbk_push_reads + pbk_finalize + pbk_export) -- everything
         the reference's makeKmerReadDistributionMT .. sortedKeyFromKmerFile leave behind -- inside the
         timed region

N = 1 runs BASELINE config C1 (4.6 Mb genome, 2x150 bp, 100x, k=32).  N > 1 is weak scaling of the same
configuration: every rank counts its own C1-sized sample of the same genome, keys are owned by hash range;
by default (k <= 32) nothing is sent: every rank's Pass B reads the keys it owns straight out of its peers'
bucket stores over NVLink (--exchange pull; `keys` = NCCL all-to-all of the keys, `records` = all-to-all of
pre-aggregated (k-mer, count) records after counting).  `--workload C4` gives every rank a C1-sized slice of
the C4 metagenome mix instead, `--workload C4full` the whole of BASELINE config 4 split 1/N per rank.
After the timed steps one more step is CHECKED: instance sums, sum i*hist[i], and at N > 1 a single-GPU
recount of all ranks' reads on rank 0 (verify_result).

`--impl reference` times the unmodified reference (`oracle/_ref/platanus_b assemble -kmer_occ_only`,
OpenMP, all host cores) on a bounded sample of the same workload; the same run is embedded as
`cpu_baseline` in the default arm at N = 1.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 32
METRIC = "canonical k-mers counted/sec at k=32"
UNIT = "k-mers/s"
BYTES_PER_INSTANCE = {150: 65.26, 250: 65.14}          # SURVEY.md section 8d (k=32)
CPU_FULL_RUN_S = 20.0                                   # one reference run over the whole C1 workload on a 16-core host
REF_MEM_GB = 16                                         # the reference's default -m
CPU_BUDGET_S = 620.0                                    # the reference arm counts the WHOLE C1 read set for the driver's --steps 20 --warmup 5
                                                        # (25 runs x ~20 s on 16 cores): same config as our arm, not a flattering fraction


def algorithmic_bytes_per_instance(read_len: int, k: int) -> float:
    return read_len / (read_len - k + 1.0) + 64.0


def ncu_traffic_bytes(split_build=False):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step's Pass A + Pass B launches, from the committed
    `ncu --set full` capture of this workload (profiles/*_traffic.json); None when there is no capture of the form of
    Pass B that ran."""
    path = os.path.join(ROOT, "profiles", "r2k_traffic_split_build.json" if split_build else "r2a_traffic.json")
    try:
        return float(json.load(open(path))["dram_bytes_per_step"])
    except Exception:
        return None


def atomic_peak_gops():
    """Best measured rate of the operation Pass B consists of (profiles/r1b_atomics_sweep.json, mode "sweep 64 regions, atom64-ret")."""
    try:
        rows = json.load(open(os.path.join(ROOT, "profiles", "r1b_atomics_sweep.json")))["results"]
        return max(r["gops"] for r in rows if r["mode"].startswith("sweep 64 regions") and r["mode"].endswith("atom64-ret") and "stream" not in r["mode"])
    except Exception:
        return 130.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region
    of this bench is tens of milliseconds, so the sampler is started before the warm-up, polls every 20 ms, and
    the summary only uses the samples whose arrival time falls inside [t0, t1] (widened to the nearest samples
    under load if the window caught fewer than two)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self, wait_s: float = 3.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_end = time.time() + wait_s
            while not self.lines and time.time() < t_end:       # nvidia-smi takes a few hundred ms to come up
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float = 0.0, t1: float = float("inf")):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if len(inside) < 2 and rows:                            # nearest samples around the window (still under load)
            mid = 0.5 * (t0 + min(t1, time.time()))
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in inside:
            for name, on in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if on:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
                "reasons": sorted(reasons), "samples": len(inside)}


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------------

def reference_sample(workload: str, scale: float, n_runs: int):
    """The whole workload when `n_runs` reference runs of it fit the time budget (one run of C1 takes ~20 s on the
    GPU box's 16 host cores), otherwise the first 1/2, 1/4, ... of the read pairs.  (A small sample flatters us: the
    reference's fixed costs -- zero-filling and scanning its -m sized table -- do not shrink with the input.)"""
    from platanus_b_b200 import synth
    spec = synth.config(workload, scale=scale)
    div = 1
    while div < 64 and n_runs * CPU_FULL_RUN_S * scale / div > CPU_BUDGET_S:
        div *= 2
    rs = synth.make_reads(spec, max_pairs=max(1, spec.n_pairs // div) if div > 1 else None)
    what = f"the whole {workload} read set" if div == 1 else f"the first 1/{div} of the {workload} read pairs"
    return spec, rs, f"{what} ({rs.n_reads} reads x {rs.read_len} bp)"


def count_instances(rs, k: int) -> int:
    is_n = rs.reads == ord("N")
    n = rs.reads.shape[0]
    csum = np.concatenate([np.zeros((n, 1), np.int32), np.cumsum(is_n, axis=1, dtype=np.int32)], axis=1)
    return int(((csum[:, k:] - csum[:, :-k]) == 0).sum())


def run_reference_once(files, workdir, threads, mem_gb):
    from oracle import oracle as O
    r = O.run_reference(files, K, workdir, threads=threads, mem_gb=mem_gb, parse_bin=False)
    if r.returncode != 0:
        raise RuntimeError("reference failed: " + r.stderr[-400:])
    return r.wall_s


def cpu_baseline(workload: str, scale: float, steps: int, warmup: int):
    """The unmodified reference on the host cores of this box; a reported baseline, not the target."""
    from oracle import oracle as O
    from platanus_b_b200 import synth
    if not O.have_ref_binary():
        return None
    spec, rs, sample = reference_sample(workload, scale, max(1, steps) + warmup)
    n_inst = count_instances(rs, K)
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="pbk_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files = synth.write_fastq(rs, os.path.join(tmp, "r_1.fq"), os.path.join(tmp, "r_2.fq"))
        for _ in range(warmup):
            run_reference_once(files, tmp, cores, REF_MEM_GB)
        t = [run_reference_once(files, tmp, cores, REF_MEM_GB) for _ in range(max(1, steps))]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    wall = float(np.mean(t))
    return {"value": n_inst / wall, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": sample + f"; platanus_b assemble -kmer_occ_only -k {K} -t {cores} -m {REF_MEM_GB}, whole-process wall clock "
                               f"{wall:.2f} s incl. FASTQ parse and kmer_occ.bin write",
            "n_instances": n_inst, "ms_per_step": wall * 1e3}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base = cpu_baseline(args.workload, args.scale, args.steps, args.warmup)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/platanus_b has not been built"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": args.workload, "k": K, "scale": args.scale, "sample": base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def verify_result(kc, step, world, rank, n_inst_local, total_inst, hist_dev, d_bases, d_offs, n_reads, n_bases, recount=True):
    """After the timed steps: one more step whose result is CHECKED (a lost or duplicated exchange chunk must not print a
    throughput).  (1) every rank's own window count, all-reduced, equals the expected total; (2) sum of i * hist[i] over the
    all-reduced occurrence histogram equals it too (when nothing saturated); (3) N > 1: rank 0 gathers every rank's reads over
    NCCL and recounts them on ONE unsharded context -- the all-reduced histogram and the number of distinct k-mers must be identical."""
    import torch
    import torch.distributed as dist

    from platanus_b_b200 import KmerCounter
    step(True)
    hist = kc.occ_hist.astype(np.int64)
    inst, distinct = int(kc.n_instances), int(kc.n_distinct)
    if world > 1:
        hist = hist_dev.cpu().numpy()                               # step() left the all-reduced histogram there
        t = torch.tensor([inst, distinct], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        inst, distinct = int(t[0].item()), int(t[1].item())
    else:
        assert inst == n_inst_local, (inst, n_inst_local)
    assert inst == total_inst, ("instances counted over all ranks", inst, total_inst)
    weighted = int((hist * np.arange(hist.shape[0], dtype=np.int64)).sum())
    assert int(hist.sum()) == distinct, ("distinct k-mers vs histogram", int(hist.sum()), distinct)
    assert weighted == total_inst if hist[65534] == 0 else weighted <= total_inst, ("sum i*hist[i]", weighted, total_inst)
    out = {"instances": inst, "sum_i_hist_i": weighted, "distinct": distinct, "single_gpu_recount": None}
    if world > 1 and recount:
        sizes = torch.tensor([n_reads, n_bases], dtype=torch.int64, device="cuda")
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        if rank == 0:
            with KmerCounter(K, device=int(os.environ.get("LOCAL_RANK", "0"))) as one:
                one.push_reads_device(d_bases.data_ptr(), d_offs.data_ptr(), n_reads, n_bases)
                for src in range(1, world):
                    nr, nb = (int(x) for x in all_sizes[src].tolist())
                    rb = torch.empty(nb, dtype=torch.uint8, device="cuda")
                    ro = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
                    dist.recv(rb, src=src)
                    dist.recv(ro, src=src)
                    torch.cuda.synchronize()
                    one.push_reads_device(rb.data_ptr(), ro.data_ptr(), nr, nb)
                one.finalize()
                assert int(one.n_instances) == total_inst, (int(one.n_instances), total_inst)
                assert int(one.n_distinct) == distinct, ("distinct: sharded vs one GPU", distinct, int(one.n_distinct))
                assert np.array_equal(one.occ_hist.astype(np.int64), hist), "all-reduced histogram differs from the single-GPU recount"
            out["single_gpu_recount"] = "identical histogram and distinct count (all ranks' reads recounted unsharded on rank 0)"
        else:
            dist.send(d_bases, dst=0)
            dist.send(d_offs, dst=0)
        dist.barrier()
    return out


def main_ours(args):
    import torch
    import torch.distributed as dist

    from platanus_b_b200 import KmerCounter, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload -------------------------------------------------------------------------------
    import dataclasses
    if world == 1 or args.workload not in ("C4", "C4full"):
        # weak scaling of the metric's own configuration: every rank counts its own C1-sized sample of the SAME genome
        # (rank r draws its reads with seed + 7919 r), i.e. N GPUs = the C1 genome at N x 100x coverage, sharded by
        # hash range; N = 1 is plain C1.  Per-GPU work is identical for every N.
        spec = synth.config(args.workload, scale=args.scale)
        if world > 1:
            spec = dataclasses.replace(spec, read_seed=spec.read_seed + 7919 * rank)
        rs = synth.make_reads(spec)
        workload = f"{args.workload}: {spec.total_genome} bp genome, 2x{spec.read_len} bp PE, {spec.coverage:g}x, k={K}"
        if world > 1:
            workload += f"; one such read set per GPU (independent samples of the same genome, {world}x the coverage in total), keys owned by hash range"
    else:
        # --workload C4: BASELINE config 4, every rank a C1-sized slice of the 100 Mb metagenome mix (low coverage per
        # slice: almost every k-mer is new -- a different regime from the N = 1 number).  --workload C4full: the WHOLE of
        # config 4 (20 genomes, 100 Mb, 200x: ~66.7 M pairs, 15.9 G k-mer instances) split 1/N per rank -- strong scaling.
        c1 = synth.config("C1", scale=args.scale)
        spec = synth.config("C4", scale=args.scale)
        n_slices = world if args.workload == "C4full" else max(world, int(round(spec.n_pairs / c1.n_pairs)))
        rs = synth.make_reads(spec, pair_slice=(rank, n_slices))
        workload = (f"C4 metagenome mix (20 genomes, {spec.total_genome} bp, {spec.coverage:g}x), 2x{spec.read_len} bp PE: "
                    + (f"the whole read set, 1/{world} per GPU ({rs.n_reads} reads each)" if args.workload == "C4full" else
                       f"one C1-sized slice of {rs.n_reads} reads per GPU") + f", k={K}, keys owned by hash range")
    bases, offsets = rs.flat()
    n_reads, n_bases, L = rs.n_reads, int(bases.shape[0]), rs.read_len
    n_inst_local = count_instances(rs, K)

    h_bases = torch.from_numpy(bases.copy()).pin_memory()
    h_offs = torch.from_numpy(offsets.astype(np.int64)).pin_memory()
    d_bases = h_bases.cuda()
    d_offs = h_offs.cuda()
    torch.cuda.synchronize()

    # a step pushes the rank's reads in batches of at most 2^30 bases (one batch for C1-sized inputs; C4full: several)
    batch_ranges = sharding.chunk_read_ranges(offsets, max(1, -(-n_bases // (1 << 30))))
    if world > 1:
        batch_ranges += [(n_reads, n_reads)] * (sharding.max_windows_any_rank(len(batch_ranges), device="cuda") - len(batch_ranges))
    batches = []
    for r0, r1 in batch_ranges:
        o_b = (offsets[r0:r1 + 1] - offsets[r0]).astype(np.int64)
        h_o = torch.from_numpy(o_b.copy()).pin_memory()
        batches.append({"n_reads": r1 - r0, "b0": int(offsets[r0]), "n_bases": int(offsets[r1] - offsets[r0]), "h_offs": h_o, "d_offs": h_o.cuda()})
    kc = KmerCounter(K, device=local_rank, n_shards=world, shard_rank=rank, timing=True)
    W = kc.words
    exchange_mode = args.exchange if args.exchange != "auto" else ("pull" if W == 1 else "records")
    hist_dev = torch.zeros(65535, dtype=torch.int64, device="cuda")
    # --exchange keys (k <= 32): the k-mers travel BEFORE counting -- Pass A writes them straight into the all-to-all send
    # buffer, every rank runs Pass B over what it received (include/pbk.h, pbk_keyx_*).  Default is the record exchange,
    # the form measured in round 1.
    keyx = world > 1 and exchange_mode == "keys" and W == 1
    assert len(batches) == 1 or exchange_mode == "pull" or world == 1, "several batches per step: --exchange pull (or one GPU)"
    # --exchange pull (k <= 32): nobody sends keys.  Pass A fills the rank's own owner-major store, a barrier, and every rank's
    # Pass B reads the segments addressed to it out of its peers' HBM over NVLink while it inserts (pbk_keyx_pull_*).
    pull = world > 1 and exchange_mode == "pull" and W == 1
    exchange_note = None
    if pull:
        puller = sharding.KeyPull(kc, world, rank, max(max(bt["n_bases"] - bt["n_reads"] * (K - 1), 0) for bt in batches), device="cuda")
        if not puller.ok:                             # e.g. CUDA IPC not permitted between the processes of this box: all ranks agree
            exchange_note = f"--exchange pull unavailable ({puller.error or 'failed on another rank'}): fell back to records"
            pull, exchange_mode = False, "records"
            assert len(batches) == 1, exchange_note
    if keyx:
        # the batch in --keyx-chunks chunks (cut at read boundaries): the all-to-all of one chunk is in flight while the
        # next one is partitioned and the previous one inserted (sharding.pipelined_key_exchange)
        ranges = sharding.chunk_read_ranges(offsets, max(1, args.keyx_chunks))
        n_ch = sharding.max_windows_any_rank(len(ranges), device="cuda")          # same number of collectives on every rank
        ranges += [(n_reads, n_reads)] * (n_ch - len(ranges))
        ch = []
        for r0, r1 in ranges:
            o_c = (offsets[r0:r1 + 1] - offsets[r0]).astype(np.int64)
            h_o = torch.from_numpy(o_c.copy()).pin_memory()
            ch.append({"n_reads": r1 - r0, "b0": int(offsets[r0]), "n_bases": int(offsets[r1] - offsets[r0]), "h_offs": h_o, "d_offs": h_o.cuda()})
        kx = sharding.KeyExchange(kc, world, max(max(c["n_bases"] - c["n_reads"] * (K - 1), 0) for c in ch), device="cuda")

    record_bufs = {}

    def exchange():
        """hash-range all-to-all of pre-aggregated (k-mer, count) records (platanus_b_b200/sharding.py over NCCL)"""
        return sharding.exchange_staged_records(kc, world, "cuda", record_bufs, torch.cuda.current_stream().synchronize)

    def step_keyx(resident: bool):
        def partition(i, send_ptr, cur_ptr):
            c = ch[i]
            if resident:
                (kc.keyx_partition_device_async if args.keyx_async else kc.keyx_partition_device)(
                    d_bases.data_ptr() + c["b0"], c["d_offs"].data_ptr(), c["n_reads"], c["n_bases"], send_ptr, cur_ptr)
            else:
                kc.keyx_partition_ptr(h_bases.data_ptr() + c["b0"], c["h_offs"].data_ptr(), c["n_reads"], send_ptr, cur_ptr)

        ordered = args.keyx_async and resident         # chunks chained on the device: no host synchronisation inside the step
        return kx.step(len(ch), partition, torch.cuda.current_stream().synchronize,
                       caller_stream=(lambda: torch.cuda.current_stream().cuda_stream) if ordered else None)

    export_bufs = {}

    def export_kept(global_hist=None):
        """e2e only: the path's result as the reference's caller gets it -- coverage cutoff (assemble.cpp:318-321) from the
        (all-reduced) histogram, then this rank's entries >= cutoff, sorted on the GPU, copied into pinned host memory."""
        if global_hist is not None:
            occ = np.ascontiguousarray(global_hist.astype(np.uint64))
            nz = np.nonzero(occ[1:])[0]
            cutoff = int(kc._L.pbk_coverage_cutoff(occ.ctypes.data_as(C.c_void_p), int(nz[-1]) + 1 if len(nz) else 0, 0, 0))
        else:
            cutoff = kc.coverage_cutoff()
        n = kc.export_count(cutoff)
        if export_bufs.get("cap", 0) < n:
            cap = int(n * 1.25) + 1024
            export_bufs.update(cap=cap, keys=torch.empty((cap, W), dtype=torch.int64).pin_memory(), counts=torch.empty(cap, dtype=torch.int16).pin_memory())
        got = kc.export_into(cutoff, True, export_bufs["keys"].data_ptr(), export_bufs["counts"].data_ptr(), export_bufs["cap"])
        export_bufs.update(n=got, cutoff=cutoff)

    def step_pull(resident: bool):
        sent = 0
        for bt in batches:
            if resident:
                part = lambda: kc.keyx_pull_partition_device(d_bases.data_ptr() + bt["b0"], bt["d_offs"].data_ptr(), bt["n_reads"], bt["n_bases"], True)
                sent += puller.step(part, torch.cuda.current_stream().synchronize, caller_stream=lambda: torch.cuda.current_stream().cuda_stream)
            else:
                sent += puller.step(lambda: kc.keyx_pull_partition_ptr(h_bases.data_ptr() + bt["b0"], bt["h_offs"].data_ptr(), bt["n_reads"]),
                                    torch.cuda.current_stream().synchronize)
        return sent

    def step(resident: bool):
        kc.reset()
        if pull:
            sent = step_pull(resident)
        elif keyx:
            sent = step_keyx(resident)
        else:
            for bt in batches:
                if resident == "packed":                 # host buffers in the opt-in 2-bit form (pbk_pack_reads): 0.25 B/base over PCIe
                    kc.push_reads_packed_ptr(bt["h_words"].data_ptr(), bt["h_offs"].data_ptr(), bt["n_reads"], bt["h_npos"].data_ptr(), bt["n_n"])
                elif resident:
                    kc.push_reads_device(d_bases.data_ptr() + bt["b0"], bt["d_offs"].data_ptr(), bt["n_reads"], bt["n_bases"])
                else:
                    kc.push_reads_ptr(h_bases.data_ptr() + bt["b0"], bt["h_offs"].data_ptr(), bt["n_reads"])
            sent = exchange() if world > 1 else 0
        kc.finalize_light()                     # D2H of the occurrence histogram: the step's result
        if world > 1:
            hist_dev.copy_(torch.from_numpy(kc.occ_hist.astype(np.int64)), non_blocking=False)
            sharding.allreduce_histogram(hist_dev)
            torch.cuda.current_stream().synchronize()
        if not resident or resident == "packed":
            export_kept(hist_dev.cpu().numpy() if world > 1 else None)
        return sent

    def timed(resident: bool, steps: int, warmup: int):
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()                     # before the warm-up: it is already polling when the timed steps run
        for _ in range(warmup):
            step(resident)
        s0 = kc.stats()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        kc.timer_mark(0)
        t0 = time.perf_counter()
        w0 = time.time()
        sent = 0
        for _ in range(steps):
            sent += step(resident)
        kc.timer_mark(1)
        ms_dev = kc.timer_elapsed_ms(0, 1)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop(w0, time.time()) if rank == 0 else None
        ms = max(ms_dev, 0.0)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        s1 = kc.stats()
        delta = {k: s1[k] - s0[k] for k in s1 if isinstance(s1[k], (int, float))}
        return ms, wall_ms, delta, clocks, sent

    total_inst = n_inst_local
    if world > 1:
        t = torch.tensor([n_inst_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_inst = int(t.item())

    # throughput numbers without the per-launch CUDA events of PBK_F_TIMING (they cost 3-4 % of a step) ...
    kc.set_timing(False)
    ms_res, wall_res, d_plain, clocks, sent_res = timed(True, args.steps, args.warmup)
    ms_e2e, wall_e2e, d_e2e, clocks_e2e, _ = timed(False, args.steps, max(1, args.warmup // 2))
    e2e_packed = None
    if world == 1 and not args.no_packed:
        # the same end-to-end region with the host buffers in the opt-in packed form: what a parser that packs 2 bits per base as
        # it copies each record hands over (pbk_pack_reads does it here, once, outside the timed region -- like the parse itself)
        from platanus_b_b200 import capi
        for bt in batches:
            w, npos = capi.pack_reads(bases[bt["b0"]:bt["b0"] + bt["n_bases"]])
            bt["h_words"] = torch.from_numpy(w.view(np.int64)).pin_memory()
            bt["h_npos"] = torch.from_numpy(np.concatenate([npos, np.zeros(1, np.uint64)]).view(np.int64)).pin_memory()
            bt["n_n"] = len(npos)
        ms_p, _, d_p, _, _ = timed("packed", args.steps, max(1, args.warmup // 2))
        e2e_packed = {"value": total_inst * args.steps / (ms_p * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(d_p["h2d_bytes"] / args.steps),
                      "d2h_bytes_per_step": int(d_p["d2h_bytes"] / args.steps), "ms_per_step": ms_p / args.steps,
                      "input": "2-bit words + absolute N positions in pinned host memory (pbk_push_reads_packed, opt-in); same result region as e2e"}
    # ... then the same K device-resident steps once more WITH them, for the per-kernel durations (roofline, breakdown)
    kc.set_timing(True)
    ms_inst, _, d_res, _, _ = timed(True, args.steps, 1)

    # sanity: the counter saw exactly the windows we expect -- on every rank, and the exchanged result is the right one
    verified = verify_result(kc, step, world, rank, n_inst_local, total_inst, hist_dev, d_bases, d_offs, n_reads, n_bases,
                             recount=not args.no_verify_recount)

    value = total_inst * args.steps / (ms_res * 1e-3)
    e2e = total_inst * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the counting phase -----------------------------------------------------------
    # Every k-mer instance is handled once by Pass A (partition_kernel: extraction + canonical key + hash ->
    # hash-range bucket store) and once by Pass B (bucket_insert_compact_kernel: one 64-bit atomic on the
    # L2-resident table region), so the "launch" the algorithmic bytes are divided by is the pair: all Pass A
    # and Pass B launches of one step.  SURVEY.md 8d: L/(L-k+1) + 64 bytes per instance.
    peak, peak_src = measured_peaks()
    bpi = algorithmic_bytes_per_instance(L, K)
    # begin-to-end time of the counting phase: the sum of the launch durations when they are serial; when Pass B of one
    # sub-batch overlaps Pass A of the next (two streams), the measured elapsed time of the overlapped region
    ms_pair = d_res["ms_count_elapsed"] / args.steps
    direct = d_res["ms_count"] - d_res["ms_partition"] - d_res["ms_insert"]      # non-partitioned launches (small batches)
    achieved = bpi * n_inst_local / (ms_pair * 1e-3) / 1e9 if ms_pair > 0 else 0.0
    # which form of Pass B ran (k <= 32, one GPU): the L2-atomic one, or split + shared-memory build (pbk_stats.n_split_build)
    split_build = d_res.get("n_split_build", 0) > 0
    insert_name = "split_kernel + region_build_kernel" if split_build else "bucket_insert_compact_kernel"
    roofline = {"bound": "hbm", "kernel": ("partition_kernel<1, KEYX> + split_gather_kernel + region_build_kernel (Pass A into the rank's own owner-major store; Pass B in its "
                                          "second form: split_gather_kernel reads every peer's store over NVLink and scatters the keys by 64 KB sub-region of the "
                                          "table, region_build_kernel builds every sub-region in shared memory)" if (pull and split_build) else
                                          "partition_kernel<1, KEYX> + bucket_insert_gather_kernel (Pass A into the rank's own owner-major store + Pass B "
                                          "reading every peer's store over NVLink: each instance goes through both exactly once)" if pull else
                                          "partition_kernel<1, KEYX> + bucket_insert_gather_kernel (Pass A into the all-to-all send "
                                          "buffer + Pass B over the received keys: each instance goes through both exactly once)" if keyx else
                                          f"partition_kernel<{W}> + bucket_insert_wide_kernel<{W}> (Pass A + Pass B, multi-word keys: "
                                          "each instance goes through both exactly once)" if W > 1 else
                                          "partition_kernel<1> + split_kernel + region_build_kernel (Pass A, then Pass B in its second form: keys split by "
                                          "64 KB sub-region of the table, every sub-region built in shared memory; each instance goes through all three "
                                          "exactly once)" if split_build else
                                          "partition_kernel<1> + bucket_insert_compact_kernel<0> (Pass A + Pass B: "
                                          "each instance goes through both exactly once)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(split_build) if (W == 1 and world == 1 and args.workload == "C1" and args.scale == 1.0) else None,
                "traffic_source": "committed ncu --set full capture of this workload (profiles/), not measured in this run", "peak_source": peak_src, "algorithmic_bytes_per_instance": bpi,
                "instances_per_launch": n_inst_local, "ms_per_launch": ms_pair,
                "launches_per_step": {"partition_kernel": d_res["launches_partition"] / args.steps,
                                      insert_name: d_res["launches_insert"] / args.steps},
                "ms_per_step": {"partition_kernel": d_res["ms_partition"] / args.steps,
                                insert_name: d_res["ms_insert"] / args.steps,
                                "direct_count_kernel": direct / args.steps},
                "sub_batched_behind_h2d_copies": {"device_resident": bool(d_res["n_pipelined_batches"] > 0), "host_buffers": bool(d_e2e["n_pipelined_batches"] > 0)},
                "kernel_share_of_step": ms_pair * args.steps / max(ms_inst, 1e-9),
                "instrumented_pass_ms_per_step": ms_inst / args.steps,
                "frac_of_step": bpi * n_inst_local * args.steps / (ms_res * 1e-3) / 1e9 / peak,
                "atomic_bound_note": "random 64-bit atomics with return: 125 G/s on an L2-resident table, 22 G/s on a table >> L2 "
                                     "(profiles/r1a_atomics_microbench.json, profiles/r1b_atomics_sweep.json, profiles/r1c_warp_ops_microbench.jsonl); Pass B alone "
                                     "runs at instances / ms_per_step.bucket_insert_compact_kernel; its second form (split_kernel + region_build_kernel) "
                                     "sends no atomics to the L2 at all"}

    # north_star quotes the fraction of the "HBM/atomic roofline": next to the HBM figure above, Pass B (first form, one-word keys)
    # against the measured rate of what it is made of -- one 64-bit atomic with return per instance on an L2-resident table region
    if W == 1 and not split_build and d_res["ms_insert"] > 0:
        peak_atomics = atomic_peak_gops()
        ach = n_inst_local * args.steps / (d_res["ms_insert"] * 1e-3) / 1e9
        roofline["atomic"] = {"bound": "l2_atomic", "kernel": "bucket_insert_gather_kernel" if (pull or keyx) else "bucket_insert_compact_kernel<0>",
                              "achieved": ach, "peak": peak_atomics, "unit": "G atomics/s", "frac": ach / peak_atomics if peak_atomics else None,
                              "peak_source": "profiles/r1b_atomics_sweep.json (pbk_microbench_atomics: uniform-random 64-bit atomicAdd with return, 1 GiB table swept "
                                             "in 64 L2-resident regions; committed measurement on a B200 of this pool, not repeated in this run)"}

    line = None
    if rank == 0:
        base = cpu_baseline(args.workload, args.scale, 1, 0) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "C4full" else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "k": K, "instances_per_step": total_inst,
                       "reads_per_gpu": n_reads, "input_bytes_per_gpu": n_bases + (n_reads + 1) * 8,
                       "l2": "inputs (>= 460 MB per GPU) and table are larger than the 126 MB L2; no explicit flush",
                       "timed_region": "reset + push (pack, count) + finalize (clamp + histogram), device stopwatch "
                                       "on the library's compute stream, max over ranks; value and e2e are timed without "
                                       "per-launch events, kernel durations come from a third pass of the same K steps with them"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(d_e2e["h2d_bytes"] / args.steps),
                    "d2h_bytes_per_step": int(d_e2e["d2h_bytes"] / args.steps), "ms_per_step": ms_e2e / args.steps,
                    "result": "occurrence histogram + coverage cutoff + the sorted (key, count) entries >= cutoff in pinned host memory",
                    "coverage_cutoff": export_bufs.get("cutoff"), "entries_exported_per_gpu": export_bufs.get("n")},
            "e2e_packed2": e2e_packed,
            "gpu_launches": int(d_plain["launches_pack"] + d_plain["launches_count"] + d_plain["launches_other"]),
            "roofline": roofline,
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")} if clocks else None,
            "kernel_ms_per_step": {"pack": d_res["ms_pack"] / args.steps, "count": d_res["ms_count"] / args.steps,
                                   "other": d_res["ms_other"] / args.steps,
                                   "count_partition_pass": d_res["ms_partition"] / args.steps,
                                   "count_insert_pass": d_res["ms_insert"] / args.steps},
            "wall_ms_per_step": wall_res / args.steps,
            "verified": verified,
            "table": {"slots": int(kc.stats()["table_slots"]), "bytes": int(kc.stats()["table_bytes"]),
                      "distinct_local": int(kc.n_distinct), "grows_in_timed_region": int(d_res["n_grow"])},
        }
        if world > 1:
            line["exchange_bytes_sent_per_gpu_per_step"] = int(sent_res / args.steps)
            line["exchange"] = ("pull: Pass B reads the peers' bucket stores in place over NVLink (P2P loads, no collective moves the keys; "
                                "bytes = what this rank's peers read from it, pbk_keyx_pull_*)" if pull else
                                "keys before counting (8-byte hashes, equal splits, pbk_keyx_*)" if keyx else
                                "pre-aggregated (key, count) records after counting (pbk_shard_*)")
        if exchange_note:
            line["exchange_note"] = exchange_note
        if base is not None:
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if args.write_outputs:                      # the sharded count's files, as the single-GPU program writes them (untimed)
        step(True)
        if world > 1:
            sharding.write_outputs(kc, args.write_outputs, K, world, rank, 16 * 10 ** 9, device="cuda")
        else:
            kc.output_occurrence_distribution(f"{args.write_outputs}_{K}merFrq.tsv")
            kc.output_occurrence_table_binary(f"{args.write_outputs}_kmer_occ.bin", kc.coverage_cutoff(),
                                              int(kc._L.pbk_double_hash_size(16 * 10 ** 9, K)))
    kc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C1")
    ap.add_argument("--write-outputs", default="", metavar="PREFIX",
                    help="after the timed steps: write PREFIX_<k>merFrq.tsv and PREFIX_kmer_occ.bin of the (sharded) count, rank 0 (untimed)")
    ap.add_argument("--keyx-async", action="store_true", default=bool(os.environ.get("PBK_BENCH_KEYX_ASYNC")),
                    help="--exchange keys: order Pass A -> all-to-all -> Pass B by CUDA events (pbk_stream_signal/wait) instead of host syncs")
    ap.add_argument("--keyx-chunks", type=int, default=int(os.environ.get("PBK_BENCH_KEYX_CHUNKS", "4")),
                    help="--exchange keys: chunks per step (the all-to-all of one chunk overlaps the passes of its neighbours)")
    ap.add_argument("--exchange", default=os.environ.get("PBK_BENCH_EXCHANGE", "auto"), choices=["auto", "records", "keys", "pull"],
                    help="N > 1: what crosses NVLink -- pull: Pass B reads the peers' bucket stores in place (default for k <= 32: measured "
                         "0.96 weak-scaling efficiency at N = 2 against 0.67 for records); keys: all-to-all of the keys before counting; "
                         "records: (key, count) records after counting (the only form for k > 32)")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--k", type=int, default=32, help="k-mer length (BASELINE metric: 32; 75 = the multi-word target of north_star)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the extra e2e pass with the opt-in 2-bit host encoding")
    ap.add_argument("--no-verify-recount", action="store_true", help="N > 1: skip the single-GPU recount of all ranks' reads on rank 0 (the sum checks stay)")
    ap.add_argument("--ref-mem-gb", type=int, default=REF_MEM_GB, help="-m of the reference arm (its default is 16)")
    args = ap.parse_args()
    globals()["REF_MEM_GB"] = args.ref_mem_gb
    globals()["K"] = args.k
    globals()["METRIC"] = f"canonical k-mers counted/sec at k={args.k}"
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
