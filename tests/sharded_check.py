"""Shared body of the sharded-count checks: every exchange form (records after counting, keys before counting with host
synchronisation, keys with device-side ordering) around a real sharded KmerCounter, verified against the unsharded oracle --
per-rank tables, all-reduced histogram, instance count, and the output files written by rank 0.

Used by tests/test_sharded_gloo_cpu.py (gloo, host-emulated C ABI, CPU tensors) and by
`torchrun --nproc-per-node N tests/sharded_check.py` on real GPUs (NCCL, CUDA tensors): the round-2 gate for the key exchange."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_checks(rank: int, world: int, k: int, mode: str, n_chunks: int, out_dir: str, device=None, scale: float = 1 / 400, reps: int = 2):
    """mode: 'records' | 'keys' | 'keys_async' | 'pull' | 'pull_async'.  device None = host tensors (emulated ABI), 'cuda' = the real
    thing ('pull*' maps the peers' HBM through CUDA IPC and therefore only exists there)."""
    from oracle import oracle as O
    from platanus_b_b200 import KmerCounter, capi, sharding, synth
    L = capi.load_library()
    cuda = device is not None
    sync = torch.cuda.current_stream().synchronize if cuda else (lambda: None)
    rs = synth.make_reads(synth.config("C1", scale=scale))
    bases, offs = rs.flat()
    rd = O.Reads()
    rd.add_array(bases, offs)
    want = O.count(rd, k)
    n = len(offs) - 1
    lo, hi = n * rank // world, n * (rank + 1) // world
    b = np.ascontiguousarray(bases[int(offs[lo]):int(offs[hi])])
    o = (offs[lo:hi + 1] - offs[lo]).astype(np.uint64)
    with KmerCounter(k, device=(torch.cuda.current_device() if cuda else -1), n_shards=world, shard_rank=rank) as kc:
        puller = None
        for rep in range(reps):                                     # later passes: reset, tables and layout reused, queued inserts
            kc.reset()
            if mode.startswith("pull"):
                if puller is None:
                    puller = sharding.KeyPull(kc, world, rank, max(int(o[-1]) - (len(o) - 1) * (k - 1), 0), device=device)
                if mode == "pull" or rep == 0:
                    sent = puller.step(lambda: kc.keyx_pull_partition(b, o), sync)
                else:                                               # device-ordered: inputs resident, no host sync inside the step
                    t_b = torch.from_numpy(b).to(device)
                    t_o = torch.from_numpy(o.view(np.int64)).to(device)
                    sync()
                    sent = puller.step(lambda: kc.keyx_pull_partition_device(t_b.data_ptr(), t_o.data_ptr(), len(o) - 1, int(o[-1]), True), sync,
                                       caller_stream=lambda: torch.cuda.current_stream().cuda_stream)
                if rep >= 1:
                    assert kc.stats()["n_pipelined_batches"] >= 1
            elif mode.startswith("keys"):
                ranges = sharding.chunk_read_ranges(o, n_chunks)
                n_ch = sharding.max_windows_any_rank(len(ranges), device=device)
                ranges += [(len(o) - 1, len(o) - 1)] * (n_ch - len(ranges))
                chunks = [(np.ascontiguousarray(b[int(o[r0]):int(o[r1])]), (o[r0:r1 + 1] - o[r0]).astype(np.uint64)) for r0, r1 in ranges]
                kx = sharding.KeyExchange(kc, world, max(max(int(co[-1]) - (len(co) - 1) * (k - 1), 0) for _, co in chunks), device=device)
                if mode == "keys" or rep == 0:
                    sent = kx.step(len(chunks), lambda i, sp, cp: kc.keyx_partition(chunks[i][0], chunks[i][1], sp, cp), sync)
                else:                                               # device-ordered: inputs resident, no host sync inside the step
                    t_b = [torch.from_numpy(cb).to(device) for cb, _ in chunks]
                    t_o = [torch.from_numpy(co.view(np.int64)).to(device) for _, co in chunks]
                    sync()
                    stream = (lambda: torch.cuda.current_stream().cuda_stream) if cuda else (lambda: 0)
                    sent = kx.step(len(chunks), lambda i, sp, cp: kc.keyx_partition_device_async(
                        t_b[i].data_ptr(), t_o[i].data_ptr(), len(chunks[i][1]) - 1, int(chunks[i][1][-1]), sp, cp), sync,
                        caller_stream=stream)
                if rep >= 1:                                        # steady state: received chunks are inserted without a host round trip
                    assert kc.stats()["n_pipelined_batches"] >= 1
            else:
                kc.push_reads(b, o)
                sent = sharding.exchange_staged_records(kc, world, device, None, sync)
            assert sent > 0
            kc.finalize()
            hist = torch.from_numpy(kc.occ_hist.astype(np.int64)).to(device)
            sharding.allreduce_histogram(hist)
            assert np.array_equal(hist.cpu().numpy().astype(np.uint64), want.occ_hist), (mode, rep)
            keys, counts = kc.export(1, sorted=True)
            sel = np.array([L.pbk_shard_of_key(np.ascontiguousarray(r).ctypes.data_as(C.c_void_p), k, world) == rank
                            for r in want.keys], dtype=bool)
            assert np.array_equal(keys, want.keys[sel]) and np.array_equal(counts, want.counts[sel]), (mode, rep)
            inst = torch.tensor([kc.n_instances], dtype=torch.int64, device=device)
            dist.all_reduce(inst)
            assert int(inst.item()) == want.n_instances
        # the output files of the sharded count (rank 0): identical to what the unsharded oracle table gives
        cutoff = sharding.write_outputs(kc, os.path.join(out_dir, "sh"), k, world, rank, 10 ** 9, device=device)
        assert cutoff == O.coverage_cutoff(want.occ_hist, want.max_occ)
        if rank == 0:
            assert open(os.path.join(out_dir, f"sh_{k}merFrq.tsv")).read() == O.tsv_text(want.occ_hist, want.max_occ)
            t = O.read_bin(os.path.join(out_dir, "sh_kmer_occ.bin"))
            keep = want.counts >= cutoff
            bk, bc = t.sorted_dump()
            assert t.reachable and t.k == k and t.index_size == max(O.load_size(int(keep.sum())), O.double_hash_size(10 ** 9, k)) - 1
            assert np.array_equal(bk, want.keys[keep]) and np.array_equal(bc, want.counts[keep])


def main():
    """torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/sharded_check.py [scale]"""
    import tempfile
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1 / 20
    out = tempfile.mkdtemp()
    try:
        for k, mode, n_chunks in ((32, "records", 1), (75, "records", 1), (32, "keys", 3), (32, "keys_async", 4), (21, "keys_async", 2),
                                  (32, "pull", 1), (32, "pull_async", 1), (21, "pull_async", 1)):
            run_checks(rank, world, k, mode, n_chunks, out, device="cuda", scale=scale, reps=3)
            dist.barrier()
            if rank == 0:
                print(f"ok: world={world} k={k} exchange={mode} chunks={n_chunks}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
