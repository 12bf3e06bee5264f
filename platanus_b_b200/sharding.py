"""Host side of the hash-range exchange between GPUs: plumbing over
torch.distributed -- NCCL over NVLink on the GPUs, gloo in the CPU tests.  The records themselves are
produced and consumed by libpbk (pbk_shard_pack_device / pbk_shard_insert_device)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def exchange_counts(send_counts: torch.Tensor, group=None) -> torch.Tensor:
    """send_counts[d] = records this rank holds for owner d  ->  recv_counts[s] = records rank s holds for us."""
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def exchange_records(send_buf: torch.Tensor, send_counts, recv_counts, recv_buf: torch.Tensor | None = None,
                     group=None) -> torch.Tensor:
    """All-to-all of (key words..., count) records.  `send_buf` is [n_send, W + 1] int64 grouped by destination in
    rank order (what pbk_shard_pack_device writes); returns the [n_recv, W + 1] records this rank owns."""
    send_counts = [int(x) for x in send_counts]
    recv_counts = [int(x) for x in recv_counts]
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    width = send_buf.shape[1]
    if recv_buf is None or recv_buf.shape[0] < n_recv:
        recv_buf = torch.empty((n_recv, width), dtype=send_buf.dtype, device=send_buf.device)
    out = recv_buf[:n_recv]
    dist.all_to_all_single(out, send_buf[:n_send], recv_counts, send_counts, group=group)
    return out


def allreduce_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-shard occurrence histograms: shards own disjoint key sets, so the sum is the global one."""
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


# ---- second form: the keys themselves are exchanged before counting (k <= 32, pbk_keyx_* in include/pbk.h) ----------

def max_windows_any_rank(n_windows: int, device=None, group=None) -> int:
    """Every rank must derive the same all-to-all layout: the segment size follows from the largest batch."""
    t = torch.tensor([int(n_windows)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def exchange_keys(send: torch.Tensor, cursors: torch.Tensor, recv: torch.Tensor, recv_cursors: torch.Tensor, group=None):
    """Equal-split all-to-all of the bucket store Pass A filled: `send` is [n_dest, n_regions, seg_cap] int64 (the
    hashes of this rank's k-mers, grouped by owner), `cursors` [n_dest, n_regions] the fill counts; afterwards
    `recv` / `recv_cursors` hold the same by source rank -- the input of pbk_keyx_insert_device."""
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    dist.all_to_all_single(recv_cursors.view(-1), cursors.view(-1), group=group)
    return recv, recv_cursors


def any_rank_staged(n_staged_here: int, device=None, group=None) -> bool:
    """True if some rank holds staged (key, count) records -- keys whose segment was full -- so that every rank joins
    the record exchange; False in the common case, and the record exchange is skipped by all."""
    t = torch.tensor([int(n_staged_here)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item()) > 0


def pipelined_key_exchange(n_chunks: int, partition, insert, send, cursors, recv, recv_cursors, wait_collectives, group=None,
                           after_partition=None, before_insert=None):
    """The key exchange over a batch cut into chunks, so that the all-to-all of chunk i runs while Pass A of chunk i + 1
    and Pass B of chunk i - 1 are computed.

      partition(i, send_buf, cursor_buf)   Pass A of chunk i into the buffers (pbk_keyx_partition*; returns when they
                                           are complete)
      insert(recv_buf, recv_cursor_buf)    Pass B over a received chunk (pbk_keyx_insert_device)
      send, cursors, recv, recv_cursors    two buffers each (double buffering), shaped as for exchange_keys
      wait_collectives()                   blocks the host until the collectives issued so far have finished (NCCL:
                                           synchronize the current stream after Work.wait(); gloo: nothing to do)

      after_partition(), before_insert()   optional, device-side ordering instead of host synchronisation: after_partition
                                           makes the collective's stream wait for Pass A (pbk_stream_signal, with a partition
                                           call that does not block: pbk_keyx_partition_device_async), before_insert makes the
                                           counter's stream wait for the collective (pbk_stream_wait) and replaces
                                           wait_collectives().  Everything is then ordered by the two streams alone.

    Buffer reuse: chunk i uses buffers i % 2.  Its send buffer is rewritten by partition(i + 2), which is called after
    chunk i's all-to-all has been waited for; its receive buffer is rewritten by the all-to-all of chunk i + 2, which
    is issued after insert() of chunk i has returned."""
    def finish(p):
        works, s = p
        for w in works:
            w.wait()
        if before_insert is not None:
            before_insert()
        else:
            wait_collectives()
        insert(recv[s], recv_cursors[s])

    pending = None
    for i in range(n_chunks):
        s = i & 1
        partition(i, send[s], cursors[s])
        if after_partition is not None:
            after_partition()
        works = [dist.all_to_all_single(recv[s].view(-1), send[s].view(-1), group=group, async_op=True),
                 dist.all_to_all_single(recv_cursors[s].view(-1), cursors[s].view(-1), group=group, async_op=True)]
        if pending is not None:
            finish(pending)
        pending = (works, s)
    if pending is not None:
        finish(pending)


def chunk_read_ranges(offsets, n_chunks: int):
    """[lo, hi) read ranges of roughly equal base counts (chunks are cut at read boundaries)."""
    import numpy as np
    offsets = np.asarray(offsets)
    n = len(offsets) - 1
    total = int(offsets[-1])
    cuts = [0]
    for c in range(1, n_chunks):
        cuts.append(int(np.searchsorted(offsets, total * c // n_chunks, side="left")))
    cuts.append(n)
    cuts = sorted(set(min(max(x, 0), n) for x in cuts))
    return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]


# ---- the two exchanges as whole steps around a KmerCounter (what bench.py runs per step; tests drive the same code) ----------

def exchange_staged_records(kc, world: int, device=None, bufs: dict | None = None, sync=None, group=None) -> int:
    """Record exchange after counting: all-to-all of the pre-aggregated (key, count) records a sharded KmerCounter staged
    (pbk_shard_send_counts / pack_device / insert_device).  Returns the bytes this rank sent."""
    import numpy as np
    bufs = bufs if bufs is not None else {}
    W = kc.words
    cnt = kc.shard_send_counts(world).astype(np.int64)
    rc = exchange_counts(torch.from_numpy(cnt).to(device) if device is not None else torch.from_numpy(cnt), group=group).cpu().numpy()
    n_send, n_recv = int(cnt.sum()), int(rc.sum())
    for name, need in (("send", n_send), ("recv", n_recv)):
        if bufs.get(name) is None or bufs[name].shape[0] < need + 1:
            bufs[name] = torch.empty((int(need * 1.2) + 1024, W + 1), dtype=torch.int64, device=device)
    kc.shard_pack_device(bufs["send"].data_ptr(), bufs["send"].shape[0])
    got = exchange_records(bufs["send"], cnt.tolist(), rc.tolist(), bufs["recv"], group=group)
    if sync is not None:
        sync()
    kc.shard_insert_device(got.data_ptr(), n_recv)
    return n_send * (W + 1) * 8


class KeyExchange:
    """Key exchange before counting (pbk_keyx_*) for one sharded KmerCounter: layout agreed between the ranks, double-buffered
    send / receive tensors, and the chunked, overlapped step."""

    def __init__(self, kc, world: int, max_windows_per_chunk: int, device=None, group=None):
        self.kc, self.world, self.device, self.group = kc, world, device, group
        self.lay = kc.keyx_plan(max_windows_any_rank(max_windows_per_chunk, device=device, group=group))
        shape = (world, int(self.lay.n_regions), int(self.lay.seg_cap))
        mk = lambda shp, zero: [(torch.zeros if zero else torch.empty)(shp, dtype=torch.int64, device=device) for _ in range(2)]
        self.send, self.recv = mk(shape, False), mk(shape, False)
        self.cur, self.rcur = mk(shape[:2], True), mk(shape[:2], True)
        self.record_bufs: dict = {}

    def step(self, n_chunks: int, partition_chunk, sync, caller_stream=None) -> int:
        """partition_chunk(i, d_send_ptr, d_cursors_ptr) runs pbk_keyx_partition* for chunk i.  Returns the bytes sent.
        caller_stream: a callable returning the raw cudaStream_t the collectives are issued from; when given, the step is
        ordered on the device (pbk_stream_signal / pbk_stream_wait) and partition_chunk should be the _async form."""
        hooks = {}
        if caller_stream is not None:
            hooks = {"after_partition": lambda: self.kc.stream_signal(caller_stream()),
                     "before_insert": lambda: self.kc.stream_wait(caller_stream())}
        pipelined_key_exchange(n_chunks, lambda i, s, c: partition_chunk(i, s.data_ptr(), c.data_ptr()),
                               lambda r, rc: self.kc.keyx_insert_device(r.data_ptr(), rc.data_ptr()),
                               self.send, self.cur, self.recv, self.rcur, sync, group=self.group, **hooks)
        sent = n_chunks * (self.world - 1) * int(self.lay.bytes_per_dest)
        if any_rank_staged(int(self.kc.shard_send_counts(self.world).sum()), device=self.device, group=self.group):
            sent += exchange_staged_records(self.kc, self.world, self.device, self.record_bufs, sync, self.group)
        return sent


class KeyPull:
    """Pull form of the key exchange (pbk_keyx_pull_*): no collective moves the keys.  Every rank's Pass A fills its own owner-major
    bucket store; after a barrier every rank's Pass B reads the segments addressed to it straight out of its peers' HBM over
    NVLink (the stores are mapped into each other's address space once, at set-up: CUDA IPC handles travel through one
    all_gather).  Per step the host side is: partition -> barrier (a one-element all_reduce on the caller's stream, ordered
    against the library's stream on the device) -> insert."""

    def __init__(self, kc, world: int, rank: int, max_windows: int, device=None, group=None):
        """Collective: every rank calls it.  The collectives inside are issued in the same order whether or not a local step
        fails, and `self.ok` is the AND over all ranks -- so a box that does not allow CUDA IPC between its processes makes
        every rank see ok == False (and fall back to another exchange form) instead of hanging some of them."""
        self.kc, self.world, self.rank, self.device, self.group = kc, world, rank, device, group
        self.error = ""
        max_w = max_windows_any_rank(max_windows, device=device, group=group)
        handle = bytes(kc.KEYX_HANDLE_BYTES)
        try:
            self.lay = kc.keyx_pull_setup(max_w)
            handle = kc.keyx_pull_handle()
        except Exception as e:                       # noqa: BLE001 -- reported through self.error, agreed on below
            self.error = str(e)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(device)
        handles = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        if not self.error:
            try:
                for r in range(world):
                    if r != rank:
                        kc.keyx_pull_connect_ipc(r, bytes(handles[r].cpu().numpy().tobytes()))
            except Exception as e:                   # noqa: BLE001
                self.error = str(e)
        flag = torch.tensor([0 if self.error else 1], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also the barrier: every store is mapped before anybody reads
        self.ok = bool(int(flag.item()))
        self._token = torch.zeros(2, dtype=torch.int64, device=device)
        self.record_bufs: dict = {}

    def step(self, partition, sync, caller_stream=None) -> int:
        """partition() runs pbk_keyx_pull_partition* for this rank's batch.  Returns the bytes this rank's peers read from it.
        The barrier between Pass A and Pass B is ONE all-reduce of two 8-byte words that every rank first sets to the number of
        records it has staged and of keys on its overflow list (keys that found their segment full): the sum also tells whether the record route has to run."""
        partition()
        if caller_stream is not None:                # device-ordered: the host does not wait inside the step
            self.kc.keyx_staged_count_device(self._token.data_ptr())
            self.kc.stream_signal(caller_stream())
            dist.all_reduce(self._token, group=self.group)
            self.kc.stream_wait(caller_stream())
            self.kc.keyx_pull_insert()               # queued behind the barrier; the host runs ahead
            staged_anywhere = int(self._token.sum().item()) > 0
        else:
            self._token.zero_()
            dist.all_reduce(self._token, group=self.group)
            sync()
            self.kc.keyx_pull_insert()
            staged_anywhere = any_rank_staged(int(self.kc.shard_send_counts(self.world).sum()), device=self.device, group=self.group)
        pulled = (self.world - 1) * int(self.lay.bytes_per_dest)
        if staged_anywhere:
            pulled += exchange_staged_records(self.kc, self.world, self.device, self.record_bufs, sync, self.group)
        return pulled


# ---- the outputs of a sharded count (collective 3) --------------------------------------------------

def gather_sorted_entries(kc, min_count: int, world: int, rank: int, dst: int = 0, group=None, device=None):
    """Every rank exports the entries it owns with count >= min_count (sorted by key on its GPU); rank `dst` receives them
    all and merges them into ONE ascending list -- what sortedKeyFromKmerFile / loadKmer see in the single-process program
    (counter.h:917-951, 600-640).  Shards are hash ranges, not key ranges, so a global merge is needed.
    Returns (keys [n, W] uint64, counts [n] uint16) on `dst`, (None, None) elsewhere."""
    import numpy as np
    keys, counts = kc.export(min_count, sorted=True)
    W = keys.shape[1] if keys.ndim == 2 else kc.words
    # `device`: where the collective's tensors live ("cuda" under NCCL, None = host under gloo)
    n_here = torch.tensor([len(counts)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_here, group=group)
    sizes = [int(s.item()) for s in sizes]
    # (entries >= cutoff are on their way to a file; the table itself never leaves the GPUs)
    payload = torch.from_numpy(np.concatenate([keys.reshape(-1).view(np.int64), counts.astype(np.int64)]) if len(counts) else np.zeros(0, np.int64)).to(device)
    if rank == dst:
        parts = [torch.zeros(s * (W + 1), dtype=torch.int64, device=device) for s in sizes]
        reqs = [dist.irecv(parts[src], src=src, group=group) for src in range(world) if src != dst and sizes[src]]
        parts[dst] = payload
        for r in reqs:
            r.wait()
        parts = [p.cpu() for p in parts]
        ks = [p.numpy()[:s * W].view(np.uint64).reshape(s, W) for p, s in zip(parts, sizes)]
        cs = [p.numpy()[s * W:].astype(np.uint16) for p, s in zip(parts, sizes)]
        allk, allc = np.concatenate(ks), np.concatenate(cs)
        order = np.lexsort(tuple(allk[:, w] for w in range(W)))       # reference order: top word first (binstr.h:460-466)
        return np.ascontiguousarray(allk[order]), np.ascontiguousarray(allc[order])
    if len(counts):
        dist.send(payload, dst=dst, group=group)
    return None, None


def write_outputs(kc, prefix: str, k: int, world: int, rank: int, memory_bytes: int, n_opt: int = 0, repeat: bool = False, group=None,
                  device=None):
    """PREFIX_<k>merFrq.tsv and PREFIX_kmer_occ.bin of a sharded count, written by rank 0 exactly as the single-GPU path
    writes them (pbk_write_frq_tsv, pbk_write_kmer_occ_bin): all-reduced occurrence histogram -> cutoff (assemble.cpp:318-321)
    -> merged entries >= cutoff.  Call after finalize() on every rank.  Returns the coverage cutoff."""
    import ctypes as C

    import numpy as np
    L = kc._L
    hist = torch.from_numpy(kc.occ_hist.astype(np.int64)).to(device)
    allreduce_histogram(hist, group=group)
    occ = np.ascontiguousarray(hist.cpu().numpy().astype(np.uint64))
    nz = np.nonzero(occ[1:])[0]
    max_occ = int(nz[-1]) + 1 if len(nz) else 0
    cutoff = int(L.pbk_coverage_cutoff(occ.ctypes.data_as(C.c_void_p), max_occ, int(n_opt), int(repeat)))
    keys, counts = gather_sorted_entries(kc, cutoff, world, rank, 0, group, device)
    if rank == 0:
        if L.pbk_write_frq_tsv(f"{prefix}_{k}merFrq.tsv".encode(), occ.ctypes.data_as(C.c_void_p), max_occ) != 0:
            raise OSError(f"cannot write {prefix}_{k}merFrq.tsv")
        dh = int(L.pbk_double_hash_size(int(memory_bytes), int(k)))
        if L.pbk_write_kmer_occ_bin(f"{prefix}_kmer_occ.bin".encode(), int(k), keys.ctypes.data_as(C.c_void_p),
                                    counts.ctypes.data_as(C.c_void_p), len(counts), dh, None) != 0:
            raise OSError(f"cannot write {prefix}_kmer_occ.bin")
    return cutoff
