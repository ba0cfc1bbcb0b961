"""CPU checks of bench.py's accounting and of the committed bench lines (the contract of the driver's JSON line)."""
import importlib.util
import json
import os

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_match_the_survey_figure():
    """8849 B per env-step at c3 (N = 200, P = 6, belief on)."""
    b = _bench()
    assert b.algorithmic_bytes_per_env_step(200, 6, True) == 8849
    assert b.algorithmic_bytes_per_env_step(200, 6, False) == 8849 - 8 * 200
    wl = b.WORKLOADS
    assert wl["c3"]["N"] == 200 and wl["c3"]["P"] == 6 and wl["c3"]["B"] == 65536 and wl["c3"]["belief"]
    assert wl["c2"]["B"] == 1024 and wl["c4"]["N"] == 1000 and wl["c5"]["B"] * 8 == 1 << 20


def test_committed_bench_lines_carry_the_contract_keys():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for name in ("r01_bench_c3_final.json", "r01_bench_c3_n8.json", "r01_bench_c5_gnn_n1.json", "r01_bench_c4_n1.json"):
        line = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert need <= set(line), (name, need - set(line))
        assert line["metric"] == "batched_env_steps_per_sec" and line["higher_is_better"] is True and line["scaling"] == "weak"
        assert line["vs_baseline"] is None and "workload" in line["config"] and line["gpu_launches"] > 0
        r = line["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        e = line["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != line["value"]
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    final = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_c3_final.json")).read().strip().splitlines()[-1])
    assert final["n_gpus"] == 1 and {"value", "unit", "cores", "kind", "sample"} <= set(final["cpu_baseline"])
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_c3_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"


def test_round2_bench_lines_carry_the_contract_keys():
    """The round-2 lines add: median-of-5 passes, the reference-shaped e2e, the library record, cpu_baseline on every
    N = 1 workload line, and a traffic figure that is either null or stamped for the build it ran on."""
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "e2e_reference_shape", "gpu_launches", "roofline", "passes_ms", "library"}
    names = ["r02_bench_c3_n1.json", "r02_bench_c3_steps20.json", "r02_bench_c2_n1.json", "r02_bench_c4_n1.json",
             "r02_bench_c5_gnn_n1.json", "r02_bench_c5_mappo_n1.json", "r02_bench_c3_n2.json", "r02_bench_c3_n4.json",
             "r02_bench_c3_n8.json"]
    for name in names:
        line = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert need <= set(line), (name, need - set(line))
        assert line["metric"] == "batched_env_steps_per_sec" and line["higher_is_better"] is True and line["scaling"] == "weak"
        assert line["vs_baseline"] is None and "workload" in line["config"] and line["gpu_launches"] > 0
        assert len(line["passes_ms"]) == 5 and line["library"]["override"] is False
        r = line["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] is None or 0.5 < r["traffic"] / (r["algorithmic_bytes_per_env_step"] * line["config"]["envs_per_gpu"]) < 1.2
        for key in ("e2e", "e2e_reference_shape"):
            e = line[key]
            assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != line["value"]
        assert line["e2e_reference_shape"]["h2d_bytes_per_step"] >= line["e2e"]["h2d_bytes_per_step"]
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if line["n_gpus"] == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"]) and line["cpu_baseline"]["kind"] == "port"
    n1 = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_c3_steps20.json")).read().strip().splitlines()[-1])
    for n, name in ((2, "r02_bench_c3_n2.json"), (4, "r02_bench_c3_n4.json"), (8, "r02_bench_c3_n8.json")):
        line = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert line["n_gpus"] == n and 0.95 < line["value"] / (n * n1["value"]) < 1.05  # weak scaling of the device path
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_c3_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"
    assert ref["config"]["envs_per_gpu"] == n1["config"]["envs_per_gpu"]  # the CPU arm steps the same batch
