"""The bench line contract (bench.py docstring): keys and internal consistency of the committed B200 line, and the
reference arm run here on the host CPU with a small batch."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LINES = sorted((ROOT / "profiles").glob("bench_r02[c-z]_roe3d_mixed.json"))


@pytest.mark.parametrize("path", LINES[-2:], ids=lambda p: p.name)
def test_committed_bench_line_is_self_consistent(path):
    d = json.loads(path.read_text())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["metric"].startswith("Jacobians/s") and d["unit"] == "Jacobians/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] == d["steps"]
    batch = d["config"]["batch_per_gpu"]
    assert "RoeFlux_3d" in d["config"]["workload"] and d["config"]["order"] == "mixed" and batch == 1 << 20
    assert d["config"]["rounding"] == "reference" and "larger than" in d["config"]["cache"]
    # value = Jacobians of the batch / device time per step; roofline = 240 algorithmic bytes per Jacobian over that time
    assert d["value"] == pytest.approx(batch / (d["ms_per_step"] * 1e-3), rel=1e-9)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["algorithmic_bytes_per_jacobian"] == 240
    assert r["achieved"] == pytest.approx(240 * batch / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12) and 0.4 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] < 1.05 * 240 * batch          # no re-reads
    e = d["e2e"]
    assert e["unit"] == "Jacobians/s" and e["h2d_bytes_per_step"] == 40 * batch and e["d2h_bytes_per_step"] == 200 * batch
    assert e["value"] < d["value"] and e["value"] <= e["floor"]["value"] * 1.001
    assert e["frac_of_floor"] == pytest.approx(e["value"] / e["floor"]["value"], rel=1e-9)
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == "Jacobians/s" and c["sample"]
    assert d["verify"]["bit_exact"] is True and d["verify"]["equal_to_reference_arithmetic"] is True
    modes = d["rounding_modes"]
    assert modes["reference"]["parity_full_batch"]["value_mismatches"] == 0
    assert modes["fma"]["parity_full_batch"]["samples_over_bound"] >= 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name, rec in d["other_configs"].items():
        assert "unavailable" not in rec, name


def test_reference_arm_line_on_the_host_cpu():
    """`bench.py --impl reference`: the reference's own sources (baseline/_ref on the NumPy stand-in, or the oracle's
    port when the install is absent) timed on the host cores; same metric / unit, e2e = value, no GPU launches."""
    done = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--batch", "8192"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert done.returncode == 0, done.stderr[-2000:]
    d = json.loads(done.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"].startswith("Jacobians/s") and d["unit"] == "Jacobians/s"
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "Jacobians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["order"] == "mixed" and d["config"]["batch_per_gpu"] == 8192


def test_committed_multi_gpu_lines_scale_and_verify_every_rank():
    """Weak scaling of the final code (profiles/bench_r02{f,g}_roe3d_mixed_{2,4,8}gpu.json): whole-job Jacobians/s within
    5 % of N x the one-GPU line, max-over-ranks device time, results of EVERY rank checked over the NCCL all-gather."""
    one = json.loads(LINES[-1].read_text())
    seen = []
    for path in sorted((ROOT / "profiles").glob("bench_r02[f-z]_roe3d_mixed_*gpu.json")):
        d = json.loads(path.read_text())
        n = d["n_gpus"]
        seen.append(n)
        assert d["scaling"] == "weak" and d["config"]["batch_per_gpu"] == 1 << 20
        assert 0.95 * n * one["value"] <= d["value"] <= 1.03 * n * one["value"], path.name
        assert d["verify"]["ranks"] == n and d["verify"]["rows"] == 1024 * n and d["verify"]["bit_exact"] is True
        assert "NCCL" in d["verify"]["collective"]
        assert d["e2e"]["value"] <= d["e2e"]["floor"]["value"] * 1.001 and d["e2e"]["frac_of_floor"] > 0.85
    assert {2, 4, 8} <= set(seen)
