"""bench.py's reference arm (the CPU oracle timed on the host) runs without a GPU: check the JSON contract of its line, and that
the other ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line():
    out = _run()
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("particles/sec per SMC set") and d["unit"] == "particles/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["n_gpus"] == 1
    assert "C2" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_default_line_has_the_contract_keys():
    """The default `python bench.py` line as measured on the B200 (profiles/r02_bench_final_default.json): every key the driver reads."""
    path = os.path.join(ROOT, "profiles", "r02_bench_final_default.json")
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"].startswith("particles/sec per SMC set") and d["unit"] == "particles/s" and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert "C3" in d["config"]["workload"] and "model" not in d["config"] and d["warmup"] >= 3
    assert abs(d["value"] - 250000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]                  # particles of one set / device time
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] >= 250000 * (150 + 30) * 8 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] in ("GB/s", "TFLOP/s")
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"] and c["value"] > 0
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name in ("C2", "C5", "T1M"):                                                                 # the other shapes ride along as objects
        assert name in d["configs"] and d["configs"][name]["ms_per_step"] > 0
    s = d["sharded_weight_update"]
    assert s["max_rel_err_vs_oracle"] <= 1e-10
