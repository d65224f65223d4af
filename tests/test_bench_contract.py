"""The bench.py JSON contract, checked on the lines recorded under profiles/ by the last GPU visit (CPU only):
every key the driver reads is present and consistent with its neighbours."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_line(name):
    path = os.path.join(ROOT, "profiles", name)
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1])


def test_default_line_has_every_contract_key():
    d = last_line("r01_bench_default_n125_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"].startswith("p121 EBE-PCG MDOF-iters/s") and d["unit"] == "MDOF*iterations/s"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert "BASELINE config C" in d["config"]["workload"] and d["config"]["neq"] == 23531000
    # value = neq * steps / time
    assert d["value"] == pytest.approx(d["config"]["neq"] / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-9)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.9 * d["value"] < e["value"] <= d["value"] * 1.001            # measured separately, host copies inside
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-9)
    assert r["algorithmic_bytes_per_launch"] == 1953125 * 3600 * 8 and r["launches_timed"] == d["steps"]
    assert r["traffic"] >= r["algorithmic_bytes_per_launch"]                # ncu dram bytes per launch
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] <= k["sm_max_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == 5 * d["steps"] + 1                          # five kernels per iteration + k_pcg_init
    assert d["time_to_solution"]["converged"] and d["time_to_solution"]["iters"] == 1561


def test_reference_arm_line():
    d = last_line("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"].startswith("p121 EBE-PCG") and d["unit"] == "MDOF*iterations/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0


def test_variant_lines_carry_their_own_roofline():
    d = last_line("r01_bench_default_n125_final.json")
    v = d["variants"]
    assert set(v) == {"stored_symmetric_packed", "matrix_free_geometric_factors", "matrix_free_rebuilt_from_coordinates"}
    assert v["stored_symmetric_packed"]["roofline"]["bound"] == "hbm"
    for name in ("matrix_free_geometric_factors", "matrix_free_rebuilt_from_coordinates"):
        r = v[name]["roofline"]
        assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1


def test_reference_arm_runs_on_cpu_and_under_torchrun():
    """`bench.py --impl reference` is CPU work: alone it prints ONE JSON line; launched like the GPU arm for N = 2
    (torch.distributed.run, rendezvous on 127.0.0.1) rank 0 alone prints it and the other rank exits 0 without work."""
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    common = ["--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-n", "8"]
    one = subprocess.run([sys.executable, bench] + common, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert one.returncode == 0, one.stderr[-2000:]
    lines = [l for l in one.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "8^3" in d["cpu_baseline"]["sample"]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29377", bench, "--gpus", "2"] + common,
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert two.returncode == 0, two.stderr[-2000:]
    lines = [l for l in two.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2 and json.loads(lines[0])["impl"] == "reference"
