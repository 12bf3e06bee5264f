"""The host logic of the C ABI (platanus_b_b200/csrc/pbk_api.cu: batching, table sizing and growth, overflow handling,
sharding, key exchange, lookup, seeded counting) checked in the GPU-less build container: the product sources are
compiled for the host against a synchronous stand-in for the CUDA runtime (tests/cpu_emul/cuda_rt_shim.h, kernels run as
one sequential thread), and the `-m gpu` tests of the listed files run against that library in a child pytest -- including the
C++ program pbk_assemble (linked against the same library) against the unmodified reference program.

This is a logic check only -- never shipped, never timed, and no fallback: the product library has no path to it.  The real
parity tests are the same test functions run on a B200."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run_child(args, timeout, **extra_env):
    import emul_helper
    emul_helper.abi_lib_path()                       # built here, once: the child's workers only find them up to date
    emul_helper.abi_cli_path()
    env = dict(os.environ, PBK_TEST_EMULATED_ABI="1", **extra_env)
    workers = ["-n", str(min(6, os.cpu_count() or 1))] if (os.cpu_count() or 1) >= 4 else []      # pytest-xdist: the cases are independent
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--runxfail", "-p", "no:cacheprovider", "-x", *workers, *args],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_emulated_abi_builds_and_exports_the_abi():
    import ctypes as C

    import emul_helper
    from platanus_b_b200 import capi
    L = C.CDLL(emul_helper.abi_lib_path())
    for name in capi.SYMBOLS:
        assert hasattr(L, name), name


SELECTION = (
    ["tests/test_zz_keyx_gpu.py", "tests/test_zz_lookup_gpu.py", "tests/test_cli_gpu.py"]
    + [f"tests/test_gpu_parity.py::test_golden_cases_bit_exact[{c}]"
       for c in ("kat_k4", "smallfq_k32", "smallfq_k75", "smallfa_k200", "cov_k21_auto", "multi_k32_n2", "sat_k32")]
    + ["tests/test_gpu_parity.py::test_table_growth_from_a_tiny_hint", "tests/test_gpu_parity.py::test_error_behaviour",
       "tests/test_gpu_parity.py::test_hash_range_passes_add_up_to_the_whole_count", "tests/test_gpu_parity.py::test_config_layouts_and_pass_arguments"])
# (tests/test_gpu_parity.py::test_pipelined_pass_overlap_gives_the_identical_table also passes this way, but its batch of > 4 M
#  windows takes 80 s in the emulation: run by hand when the chained Pass A -> Pass B route of pbk_api.cu changes)


def test_gpu_tests_pass_against_the_emulated_abi():
    """key exchange, lookup / read filter / seeded counting / contig tables, pbk_assemble against the reference program, a
    subset of the golden cases, table growth, error behaviour -- one child pytest, xdist workers"""
    p = run_child(SELECTION, timeout=2000)
    tail = "\n".join(p.stdout.splitlines()[-25:])
    assert p.returncode == 0, tail + "\n" + p.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail


OWN_STORE_ROUTES = ["tests/test_gpu_parity.py::test_forced_partition_path_on_golden_cases", "tests/test_gpu_parity.py::test_direct_and_partitioned_paths_agree",
                    "tests/test_gpu_parity.py::test_packed_two_bit_host_input_equals_ascii_input"]
EXCHANGE_ROUTES = ["tests/test_zz_keyx_gpu.py", "tests/test_zz_group_gpu.py"]


@pytest.mark.parametrize("env,selection", [({"PBK_PASSB2": "0"}, OWN_STORE_ROUTES), ({"PBK_PASSB2_GATHER": "1"}, EXCHANGE_ROUTES)],
                         ids=["own_store_first_form", "exchange_second_form"])
def test_the_other_form_of_pass_b_against_the_emulated_abi(env, selection):
    """Pass B for one-word keys has two forms: one L2 atomic per instance (bucket_insert_compact / _gather_kernel), or the keys split by
    sub-region of the table and every sub-region built in shared memory (split_kernel + region_build_kernel).  The defaults -- second
    form for a context's own bucket store, first form for the key / pull exchange -- run in the selection above; here the
    partitioned routes go through the OTHER form of each."""
    p = run_child(selection, timeout=1500, **env)
    tail = "\n".join(p.stdout.splitlines()[-25:])
    assert p.returncode == 0, tail + "\n" + p.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail


def test_short_differential_fuzz_of_the_emulated_abi():
    """scripts/fuzz_emulated_abi.py for 20 s: random read sets and k, every counting route, seeded counting, lookup, both
    sharded exchanges -- the C ABI compiled for the host against the oracle."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_emulated_abi.py"), "--seconds", "20", "--seed", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert "cases ok" in p.stdout
