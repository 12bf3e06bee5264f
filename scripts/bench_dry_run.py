#!/usr/bin/env python
"""Dry run of bench.py's control flow in the GPU-less container -- a check of the SCRIPT, not a measurement: the C ABI
compiled for the host (tests/cpu_emul), gloo instead of NCCL, torch.cuda stubbed out, the clock sampler replaced.  Every
number it prints is meaningless; what matters is that each flow (N = 1; N = 2 with the record exchange, the key exchange,
the device-ordered key exchange + --write-outputs) runs to its JSON line with every contract key.  (The default exchange at N > 1,
--exchange pull, maps the peers' device memory through CUDA IPC and cannot be dry-run between host processes: its logic is covered
in-process by tests/test_zz_keyx_gpu.py against the emulated ABI, and on real GPUs by tests/sharded_check.py.)

    python scripts/bench_dry_run.py [--scale 0.004] [--flows n1,records,keys,keys_async]"""
import argparse
import json
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOWS = {"n1": (1, []), "records": (2, ["--exchange", "records"]), "keys": (2, ["--exchange", "keys", "--keyx-chunks", "3"]),
         "keys_async": (2, ["--exchange", "keys", "--keyx-async"])}


def worker(rank, world, port, extra, scale, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      PBK_TEST_EMULATED_ABI="1")
    import emul_helper
    from platanus_b_b200 import build
    build.LIB = emul_helper.abi_lib_path()

    class _Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    src = open(os.path.join(ROOT, "bench.py")).read().replace('"cuda"', '"cpu"')
    src = src.replace('dist.init_process_group("nccl", device_id=torch.device("cpu", local_rank))', 'dist.init_process_group("gloo")')
    assert world == 1 or 'init_process_group("gloo")' in src
    src = src.replace("class ClockSampler", "class _RealClockSampler") + '''
class ClockSampler:
    def __init__(self, *a): pass
    def start(self): pass
    def stop(self, *a): return {"sm_mhz": 0, "sm_max_mhz": 0, "reasons": [], "samples": 0}
'''
    g = {"__name__": "bench_dry_run", "__file__": os.path.join(ROOT, "bench.py")}
    sys.argv = ["bench.py", "--gpus", str(world), "--steps", "2", "--warmup", "3", "--scale", str(scale), "--no-cpu-baseline"] + extra
    if rank == 0:
        sys.stdout = open(out_path, "w")
    exec(compile(src, os.path.join(ROOT, "bench.py"), "exec"), g)
    assert g["main"]() == 0
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.004)
    ap.add_argument("--flows", default="n1,records,keys,keys_async")
    ap.add_argument("--tmp", default="/tmp")
    args = ap.parse_args()
    for name in args.flows.split(","):
        world, extra = FLOWS[name]
        if name == "keys_async":
            extra = extra + ["--write-outputs", os.path.join(args.tmp, "bench_dry_out")]
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        out = os.path.join(args.tmp, f"bench_dry_{name}.json")
        mp.spawn(worker, args=(world, port, extra, args.scale, out), nprocs=world, join=True)
        line = json.loads(open(out).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert key in line, (name, key)
        assert line["n_gpus"] == world and line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
        print(f"flow {name}: ok (n_gpus={world}, exchange={line.get('exchange', '-')})", flush=True)


if __name__ == "__main__":
    main()
