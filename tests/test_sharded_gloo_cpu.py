"""The N > 1 path on CPU: two processes (gloo, world size 2) run the host side of the hash-range exchange
(platanus_b_b200/sharding.py, the same functions bench.py uses over NCCL) around the product kernels compiled
for the host (tests/cpu_emul: shard routing, remote staging, record packing, weighted inserts).  Every rank must
end up with exactly the keys it owns, with the global counts; the all-reduced histogram must be the oracle's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, k, fq, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes as C

    from emul_helper import emul_count, emul_insert_records
    from oracle import oracle as O
    from platanus_b_b200 import capi, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = capi.load_library()
        W = (k + 31) // 32
        rd = O.Reads()
        rd.add_file(fq)
        bases, offs = rd.arrays()
        n = len(offs) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        got = emul_count(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], k, n_shards=world, rank=rank)
        staged = got["remote"].astype(np.int64)              # grouped by destination, as pbk_shard_pack_device writes them
        owner = np.array([L.pbk_shard_of_key(np.ascontiguousarray(r[:W]).astype(np.uint64).ctypes.data_as(C.c_void_p), k, world)
                          for r in staged], dtype=np.int64) if len(staged) else np.zeros(0, np.int64)
        assert np.all(np.diff(owner) >= 0) and not np.any(owner == rank)
        send_counts = torch.from_numpy(np.bincount(owner, minlength=world).astype(np.int64))
        recv_counts = sharding.exchange_counts(send_counts)
        recv = sharding.exchange_records(torch.from_numpy(staged.reshape(-1, W + 1).copy()), send_counts.tolist(), recv_counts.tolist())
        mine = np.concatenate([got["keys"], got["counts"].astype(np.uint64)[:, None]], axis=1)
        keys, counts = emul_insert_records(np.concatenate([mine, recv.numpy().astype(np.uint64).reshape(-1, W + 1)]), k)
        # expected: the oracle's table restricted to the keys this rank owns
        want = O.count(rd, k)
        sel = np.array([L.pbk_shard_of_key(np.ascontiguousarray(r).ctypes.data_as(C.c_void_p), k, world) == rank
                        for r in want.keys], dtype=bool)
        assert np.array_equal(keys, want.keys[sel]) and np.array_equal(counts, want.counts[sel])
        hist = torch.from_numpy(np.bincount(counts.astype(np.int64), minlength=65535).astype(np.int64))
        sharding.allreduce_histogram(hist)
        assert np.array_equal(hist.numpy().astype(np.uint64), want.occ_hist)
        inst = torch.tensor([got["n_instances"]], dtype=torch.int64)
        dist.all_reduce(inst)
        assert int(inst.item()) == want.n_instances
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [32, 75])
def test_two_rank_exchange_over_gloo(oracle, k, tmp_path):
    import emul_helper
    emul_helper.lib()                                        # build the emulation once, before the workers race for it
    fq = os.path.join(HERE, "golden", "inputs", "small.fq")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, fq, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _keyx_worker(rank, world, port, k, fq, seg_cap, out_dir):
    """Second form of the exchange (pbk_keyx_*): the bucket store of Pass A is the all-to-all send buffer."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes as C

    from emul_helper import emul_keyx_insert, emul_keyx_partition
    from oracle import oracle as O
    from platanus_b_b200 import capi, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = capi.load_library()
        rd = O.Reads()
        rd.add_file(fq)
        bases, offs = rd.arrays()
        n = len(offs) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        b, o = bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo]
        windows = int(o[-1]) - (len(o) - 1) * (k - 1)
        assert sharding.max_windows_any_rank(windows) >= windows          # what pbk_keyx_plan is fed on every rank
        n_regions = 4
        send, cursors, n_inst, spilled = emul_keyx_partition(b, o, k, world, n_regions, seg_cap)
        t_send, t_cur = torch.from_numpy(send.view(np.int64)), torch.from_numpy(cursors.view(np.int64))
        recv, rcur = sharding.exchange_keys(t_send, t_cur, torch.empty_like(t_send), torch.empty_like(t_cur))
        extra = None
        if sharding.any_rank_staged(len(spilled)):                        # record route for keys whose segment was full
            owner = np.array([L.pbk_shard_of_key(np.array([x], np.uint64).ctypes.data_as(C.c_void_p), k, world)
                              for x in spilled[:, 0]], dtype=np.int64) if len(spilled) else np.zeros(0, np.int64)
            order = np.argsort(owner, kind="stable")
            staged = spilled[order].astype(np.int64)
            cnt = torch.from_numpy(np.bincount(owner, minlength=world).astype(np.int64))
            rc = sharding.exchange_counts(cnt)
            extra = sharding.exchange_records(torch.from_numpy(staged.reshape(-1, 2).copy()), cnt.tolist(), rc.tolist()).numpy().astype(np.uint64)
        else:
            assert seg_cap >= 4096
        keys, counts = emul_keyx_insert(recv.numpy().view(np.uint64), rcur.numpy().view(np.uint64), seg_cap, k, extra=extra)
        want = O.count(rd, k)
        sel = np.array([L.pbk_shard_of_key(np.ascontiguousarray(r).ctypes.data_as(C.c_void_p), k, world) == rank
                        for r in want.keys], dtype=bool)
        assert np.array_equal(keys, want.keys[sel]) and np.array_equal(counts, want.counts[sel])
        hist = torch.from_numpy(np.bincount(counts.astype(np.int64), minlength=65535).astype(np.int64))
        sharding.allreduce_histogram(hist)
        assert np.array_equal(hist.numpy().astype(np.uint64), want.occ_hist)
        inst = torch.tensor([n_inst], dtype=torch.int64)
        dist.all_reduce(inst)
        assert int(inst.item()) == want.n_instances
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _keyx_pipeline_worker(rank, world, port, k, fq, n_chunks, out_dir):
    """sharding.pipelined_key_exchange: the batch in chunks, the all-to-all of one chunk in flight (async_op) while the
    next chunk is partitioned and the previous one inserted; double-buffered send/receive buffers."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes as C

    from emul_helper import emul_insert_records, emul_keyx_insert, emul_keyx_partition
    from oracle import oracle as O
    from platanus_b_b200 import capi, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = capi.load_library()
        rd = O.Reads()
        rd.add_file(fq)
        bases, offs = rd.arrays()
        n = len(offs) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        b, o = bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo]
        ranges = sharding.chunk_read_ranges(o, n_chunks)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(o) - 1 and all(a[1] == c[0] for a, c in zip(ranges, ranges[1:]))
        n_ch = sharding.max_windows_any_rank(len(ranges))                 # every rank must issue the same collectives
        ranges += [(len(o) - 1, len(o) - 1)] * (n_ch - len(ranges))       # empty chunks: zero cursors travel
        n_regions, seg_cap = 4, 8192
        send = [torch.zeros((world, n_regions, seg_cap), dtype=torch.int64) for _ in range(2)]
        cur = [torch.zeros((world, n_regions), dtype=torch.int64) for _ in range(2)]
        recv = [torch.zeros_like(send[0]) for _ in range(2)]
        rcur = [torch.zeros_like(cur[0]) for _ in range(2)]
        inst, tables = [0], []

        def partition(i, send_buf, cur_buf):
            r0, r1 = ranges[i]
            s, c, n_inst, spilled = emul_keyx_partition(b[int(o[r0]):int(o[r1])], o[r0:r1 + 1] - o[r0], k, world, n_regions, seg_cap)
            assert len(spilled) == 0
            send_buf.copy_(torch.from_numpy(s.view(np.int64)))
            cur_buf.copy_(torch.from_numpy(c.view(np.int64)))
            inst[0] += n_inst

        def insert(recv_buf, rcur_buf):                                    # (the emulation has no persistent table: one table
            keys, counts = emul_keyx_insert(recv_buf.numpy().view(np.uint64).copy(), rcur_buf.numpy().view(np.uint64).copy(), seg_cap, k)
            tables.append(np.concatenate([keys, counts.astype(np.uint64)[:, None]], axis=1))     # per chunk, merged below)

        sharding.pipelined_key_exchange(len(ranges), partition, insert, send, cur, recv, rcur, lambda: None)
        keys, counts = emul_insert_records(np.concatenate(tables), k)
        want = O.count(rd, k)
        sel = np.array([L.pbk_shard_of_key(np.ascontiguousarray(r).ctypes.data_as(C.c_void_p), k, world) == rank
                        for r in want.keys], dtype=bool)
        assert np.array_equal(keys, want.keys[sel]) and np.array_equal(counts, want.counts[sel])
        t = torch.tensor([inst[0]], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t.item()) == want.n_instances
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [1, 3, 4])
def test_two_rank_pipelined_key_exchange_over_gloo(oracle, n_chunks, tmp_path):
    import emul_helper
    emul_helper.lib()
    fq = os.path.join(HERE, "golden", "inputs", "small.fq")
    world = 2
    mp.spawn(_keyx_pipeline_worker, args=(world, _free_port(), 32, fq, n_chunks, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


@pytest.mark.parametrize("k,seg_cap", [(32, 8192), (21, 64)])
def test_two_rank_key_exchange_over_gloo(oracle, k, seg_cap, tmp_path):
    import emul_helper
    emul_helper.lib()
    fq = os.path.join(HERE, "golden", "inputs", "small.fq")
    world = 2
    mp.spawn(_keyx_worker, args=(world, _free_port(), k, fq, seg_cap, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _abi_worker(rank, world, port, k, mode, n_chunks, out_dir):
    """Every exchange form as bench.py runs it (sharding.KeyExchange.step / sharding.exchange_staged_records) around a real
    sharded KmerCounter -- the C ABI compiled for the host (tests/cpu_emul/cuda_rt_shim.h), so a torch CPU tensor's
    data_ptr() is a valid "device" pointer -- over gloo.  The body is tests/sharded_check.py, which also runs under torchrun
    on real GPUs."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import emul_helper
    from platanus_b_b200 import build
    build.LIB = emul_helper.abi_lib_path()
    import sharded_check

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharded_check.run_checks(rank, world, k, mode, n_chunks, out_dir, device=None)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,mode,n_chunks", [(32, "keys", 3), (32, "keys_async", 4), (32, "records", 1), (75, "records", 1)])
def test_two_rank_exchanges_through_the_c_abi(oracle, k, mode, n_chunks, tmp_path):
    import emul_helper
    emul_helper.abi_lib_path()                               # build once, before the workers race for it
    world = 2
    mp.spawn(_abi_worker, args=(world, _free_port(), k, mode, n_chunks, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
