#!/bin/bash
# single B200: A/B of the compacted multi-word Pass A, consumer timings (kernel-only), then the whole suite, smoke and the default bench at HEAD
mkdir -p gpurun_out
for k in 75 42; do for v in 1 0; do
  PBK_PART_COMPACT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-packed --k $k > gpurun_out/_c.json 2> gpurun_out/_c.err || tail -3 gpurun_out/_c.err
  python - "$k" "$v" <<'PY' | tee -a gpurun_out/r2h_compact_ab.jsonl
import json, sys
l = json.loads(open("gpurun_out/_c.json").read().strip().splitlines()[-1])
print(json.dumps({"k": int(sys.argv[1]), "compact": int(sys.argv[2]), "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "frac": round(l["roofline"]["frac"], 3), "verified": l["verified"]["instances"]}))
PY
done; done
timeout 600 python scripts/time_consumers.py > gpurun_out/r2h_consumers_k32.json 2> gpurun_out/r2h_consumers.err; echo "consumers rc=$?"; tail -2 gpurun_out/r2h_consumers.err; cat gpurun_out/r2h_consumers_k32.json
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2h_gpu_tests.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r2h_gpu_tests.log
tail -6 gpurun_out/r2h_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2h_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --k 75 > gpurun_out/r2h_bench_k75.json 2> gpurun_out/r2h_bench_k75.err; echo "bench k75 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2h_bench.json", "gpurun_out/r2h_bench_k75.json"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step")}, "e2e", l["e2e"]["value"], "packed", (l.get("e2e_packed2") or {}).get("value"), l["roofline"]["frac"], l["roofline"]["frac_of_step"])
    except Exception as e:
        print(f, "no line:", e)
PY
