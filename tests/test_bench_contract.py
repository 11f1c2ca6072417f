"""bench.py's output contract: the JSON line of both arms carries every key the driver reads.

CPU part: the reference arm (the reference's own C programs from oracle/_ref on the host cores) and the clock sampler's
windowing.  GPU part: one short run of the B200 arm on a small grid through the same code path as the default run."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _bench_module():
    spec = importlib.util.spec_from_file_location("shll_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(args, env_extra):
    env = dict(os.environ, **env_extra)
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, pr.stderr[-2000:]
    lines = [l for l in pr.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, pr.stdout[-2000:]
    return json.loads(lines[0])


def test_clock_sampler_keeps_only_the_timed_region():
    b = _bench_module()
    c = b.ClockSampler()
    idle = ["120", "1965", "90.0", "Not Active", "Not Active", "Not Active", "Not Active"]
    busy = ["1905", "1965", "610.5", "Not Active", "Not Active", "Not Active", "Active"]
    c.rows = [(1.0, idle), (2.0, busy), (2.1, busy), (3.0, idle)]
    w = c.window(1.9, 2.2)
    assert w["samples"] == 2 and w["sm_mhz"] == 1905.0 and w["sm_max_mhz"] == 1965.0 and w["reasons"] == ["sw_power_cap"]
    near = c.window(2.4, 2.45)   # shorter than the sampling period: the nearest sample, flagged
    assert near["samples"] == 1 and near["sm_mhz"] == 1905.0 and "note" in near
    assert b.ClockSampler().window(0.0, 1.0)["samples"] == 0


def test_all_cores_baseline_uses_the_openmp_reference_for_the_second_order_scheme():
    from oracle import oracle as O
    if not O.ref_available("ref_omp_o2_128"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    b = _bench_module()
    r = b.cpu_all_cores_rate("2d_o2", budget_s=0.5, omp_n=128)   # the bench itself runs the 4096^2 build
    assert r["kind"] == "reference" and r["cores"] == b.host_cores() and r["value"] > 1e5 and "ref_omp_o2_128" in r["sample"]


@pytest.mark.parametrize("workload", ["1d_o1"])
def test_reference_arm_line(workload):
    from oracle import oracle as O
    if not O.ref_available("ref_1d_o1_65536"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    d = _run(["--impl", "reference", "--workload", workload, "--steps", "5", "--warmup", "3"], {"SHLL_BENCH_CPU_BUDGET": "0.02"})
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["value"] > 1e5
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/_ref/" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == workload


def test_reference_arm_other_ranks_do_nothing():
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT,
                        env=dict(os.environ, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert pr.returncode == 0 and pr.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line_small_grid():
    d = _run(["--workload", "2d_o1", "--nx", "512", "--ny", "512", "--steps", "40", "--warmup", "3"], {"SHLL_BENCH_CPU_BUDGET": "0.05"})
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "other_mode"} <= set(d)
    assert d["gpu_launches"] == 20 and d["steps"] == 40 and d["n_gpus"] == 1 and d["dtype"] == "f32"   # two steps per launch
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_launch"] == 2 * 32 * 512 * 512 and r["steps_per_launch"] == 2
    assert abs(d["value"] - 512 * 512 * 40 / (d["ms_per_step"] * 40e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["value"] > 0 and e["value"] < d["value"] and e["h2d_bytes_per_step"] == pytest.approx(4 * 4 * 512 * 512 / 40)
    assert d["clocks"]["samples"] >= 1 and d["clocks"]["sm_mhz"] > 100
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 1e5
