"""bench.py's reference arm (`--impl reference`: the unmodified reference program on the host cores) runs without a GPU,
so its JSON-line contract can be checked here on a tiny workload.  Our own arm needs a B200 and must refuse to run
without one (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "platanus_b")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built")
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.002", "--steps", "1",
                        "--warmup", "0", "--ref-mem-gb", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "canonical k-mers counted/sec at k=32" and line["unit"] == "k-mers/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_committed_bench_line_has_every_contract_key():
    """profiles/r1c_bench.json is the line `python bench.py` printed on a B200 with the final code of the round: check
    the keys the driver and the judge read (base contract + roofline + cpu_baseline + clocks)."""
    line = json.load(open(os.path.join(ROOT, "profiles", "r1c_bench.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["metric"] == "canonical k-mers counted/sec at k=32" and line["unit"] == "k-mers/s" and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"] and "model" not in line["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 4e8 and line["e2e"]["value"] < line["value"]      # real host->device copies inside
    r = line["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    c = line["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] == "reference" and c["cores"] >= 1
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"]) and line["clocks"]["sm_mhz"] is not None
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert line["gpu_launches"] > 0 and line["warmup"] >= 3


def test_bench_flows_run_end_to_end_on_the_emulated_abi(tmp_path):
    """scripts/bench_dry_run.py: bench.py's own control flow -- N = 1, and N = 2 over gloo with the record exchange, the key
    exchange and the device-ordered key exchange + --write-outputs -- against the C ABI compiled for the host, torch.cuda
    stubbed out.  A check of the script (every flow reaches its JSON line with the contract keys), not a measurement."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # (the plain `keys` flow differs from `keys_async` only in who waits for the collective; run by hand: --flows keys)
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "bench_dry_run.py"), "--tmp", str(tmp_path), "--scale", "0.002",
                        "--flows", "n1,records,keys_async"], cwd=root, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    for flow in ("n1", "records", "keys_async"):
        assert f"flow {flow}: ok" in p.stdout
    assert os.path.exists(tmp_path / "bench_dry_out_32merFrq.tsv") and os.path.exists(tmp_path / "bench_dry_out_kmer_occ.bin")
