"""The bench.py JSON contract, checked on the lines recorded under profiles/ by the last GPU visit (CPU only):
every key the driver reads is present and consistent with its neighbours."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LINE = "r02_bench_default_n125_final.json"      # the default run of the last GPU visit (scripts/gpu_r2_16.sh)


def last_line(name):
    path = os.path.join(ROOT, "profiles", name)
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1])


def test_default_line_has_every_contract_key():
    d = last_line(DEFAULT_LINE)
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
    assert "ONE pf_pcg_solve call" in e["note"]                             # says how the copies are amortised
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-9)
    assert r["algorithmic_bytes_per_launch"] == 1953125 * 3600 * 8 and r["launches_timed"] == d["steps"]
    assert r["traffic"] >= r["algorithmic_bytes_per_launch"]                # ncu dram bytes per launch
    assert abs(r["back_to_back"]["frac"] / r["frac"] - 1.0) < 0.05          # K launches between one pair of events agree
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] <= k["sm_max_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == 5 * d["steps"] + 1                          # five kernels per iteration + k_pcg_init
    t = d["time_to_solution"]
    assert t["converged"] and t["iters"] == 1561 and t["e2e_s"] >= t["solve_s"]  # the whole solve through pf_pcg_solve


def test_default_line_carries_parity_configs_weak_and_reference_kernel():
    """Round 2: the driver-run line itself holds the N-rank parity self-check, the named
    configurations B and D, the weak-scaling block and the reference's own CUDA kernel timed beside this library's."""
    d = last_line(DEFAULT_LINE)
    pc = d["parity_check"]
    assert pc["bit_equal"] is True and pc["nranks"] == d["n_gpus"] and len(pc["specs"]) == 2
    assert all(s["bit_equal"] and s["iters"] == s["oracle_iters"] for s in pc["specs"])
    cfg = d["configs"]
    assert set(cfg) == {"B_p123_100", "D_p121_hex8_200"}
    assert cfg["B_p123_100"]["config"]["neq"] == 1000000 and "config B" in cfg["B_p123_100"]["config"]["workload"]
    assert cfg["D_p121_hex8_200"]["config"]["nels"] == 8000000 and "config D" in cfg["D_p121_hex8_200"]["config"]["workload"]
    for c in cfg.values():
        r = c["roofline"]
        assert c["value"] > 0 and c["e2e"]["value"] > 0 and r["bound"] == "hbm" and r["traffic"] >= r["algorithmic_bytes_per_launch"]
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12) and c["time_to_solution"]["converged"]
    w = d["weak"]
    assert w["scaling"] == "weak" and w["config"]["nels"] == 1000000 * d["n_gpus"] and w["value"] > 0
    k = d["reference_gpu_kernel"]
    assert k["same_bits_as_reference_fmad_false"] is True and k["max_rel_diff_vs_reference_default_build"] < 1e-13
    for name in ("reference_cuda_helpers_default_build", "reference_cuda_helpers_fmad_false", "parafem_b200"):
        assert 0.5 < k[name]["frac_of_hbm_peak"] < 1.2 and k[name]["ms_per_call"] > 0
    assert k["parafem_b200"]["ms_per_call"] <= k["reference_cuda_helpers_default_build"]["ms_per_call"] * 1.02


def test_multi_gpu_lines_carry_the_parity_check():
    for name, n in (("r02_bench_g2.json", 2), ("r02_bench_g8.json", 8)):
        d = last_line(name)
        assert d["n_gpus"] == n and d["parity_check"]["nranks"] == n and d["parity_check"]["bit_equal"] is True
        assert d["parity_check"]["transport"] == "peer" and "peer" in d["partition"]
        assert d["configs"]["D_p121_hex8_200"]["n_gpus"] == n and d["weak"]["n_gpus"] == n
        assert d["roofline"]["traffic"] is not None and "scaled per element" in d["roofline"]["traffic_source"]
        assert d["gpu_launches"] == 5 * d["steps"] + 2                      # five kernels per iteration + init + first put


def test_reference_arm_line():
    d = last_line("r02_bench_reference_arm_cpu_n10.json")
    assert d["impl"] == "reference" and d["metric"].startswith("p121 EBE-PCG") and d["unit"] == "MDOF*iterations/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0
    assert d["host"]["cores"] == d["cpu_baseline"]["cores"] >= 1 and d["same_workload_as_gpu_arm"] is False
    # both arms describe the linear system with the same function: same keys as the GPU arm's config
    assert set(d["config"]) == set(last_line(DEFAULT_LINE)["config"])


def test_variant_lines_carry_their_own_roofline():
    d = last_line(DEFAULT_LINE)
    v = d["variants"]
    assert set(v) == {"stored_symmetric_packed", "matrix_free_geometric_factors", "matrix_free_rebuilt_from_coordinates"}
    assert v["stored_symmetric_packed"]["roofline"]["bound"] == "hbm"
    for name in ("matrix_free_geometric_factors", "matrix_free_rebuilt_from_coordinates"):
        r = v[name]["roofline"]
        assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1
        # the tensor-core kernel is judged against the (higher) measured tensor peak; the DFMA figure stays in the line
        assert r["kernel"].startswith("k_apply_mf4") and r["peak"] == r["peak_fp64_tensor"] > r["peak_dfma"] > 25
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
        assert r["frac_of_dfma_peak"] == pytest.approx(r["achieved"] / r["peak_dfma"], rel=1e-12)
        assert r["achieved"] == pytest.approx(r["algorithmic_flops_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e12, rel=1e-9)
        # the same kernel timed back to back, and its HBM side (factors / coordinates, indices, products, right-hand sides)
        assert abs(r["back_to_back"]["avg_launch_ms"] / r["avg_launch_ms"] - 1.0) < 0.1
        hs = r["hbm_side"]
        assert hs["bound"] == "hbm" and 0.3 < hs["frac"] < 1.0 and hs["frac"] == pytest.approx(hs["achieved"] / hs["peak"], rel=1e-12)
    hs = v["matrix_free_geometric_factors"]["roofline"]["hbm_side"]
    assert hs["bytes_per_element"] == 640 + 720 and abs(hs["traffic"] / hs["algorithmic_bytes_per_launch"] - 1.0) < 0.05
    assert v["matrix_free_rebuilt_from_coordinates"]["value"] >= 15000                     # config E as named
    assert v["matrix_free_geometric_factors"]["roofline"]["flops_per_element"] == 6696      # as executed, padding not counted
    assert v["matrix_free_geometric_factors"]["roofline"]["frac"] >= 0.65
    assert v["matrix_free_geometric_factors"]["value"] >= 19000


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
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))        # not the launcher's OMP_NUM_THREADS
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29377", bench, "--gpus", "2"] + common,
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert two.returncode == 0, two.stderr[-2000:]
    lines = [l for l in two.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2 and json.loads(lines[0])["impl"] == "reference"
