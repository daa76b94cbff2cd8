"""The JSON line bench.py prints keeps its contract (CPU only).

Checks (1) the lines committed under profiles/ for this round -- N = 1, 2, 8 and the reference arm -- and (2) one live step of the
reference arm (`bench.py --impl reference`: the reference's own compiled kernels on the host cores; no GPU needed).  The B200 arm
itself needs a GPU and is exercised by the driver."""
import json
import os
import subprocess
import sys

import pytest

import _oracle as o

PROFILES = os.path.join(o.ROOT, "profiles")
BASE_KEYS = {"metric": str, "value": (int, float), "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": (int, float),
             "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict, "gpu_launches": int,
             "cpu_baseline": dict}


def _last_json_line(text):
    lines = [ln for ln in text.splitlines() if ln.startswith("{")]
    assert lines, "no JSON line"
    return json.loads(lines[-1])


def _check_common(d, n_gpus):
    for k, t in BASE_KEYS.items():
        assert k in d, k
        assert isinstance(d[k], t), (k, type(d[k]))
    assert "vs_baseline" in d and d["vs_baseline"] is None          # BASELINE.md holds no published number for this metric
    assert d["metric"] == "saliency_frames_per_sec_1080p" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["n_gpus"] == n_gpus
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")


@pytest.mark.parametrize("name,n_gpus", [("r02_bench_final.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n8.json", 8)])
def test_committed_b200_lines(name, n_gpus):
    d = _last_json_line(open(os.path.join(PROFILES, name)).read())
    if n_gpus > 1:
        d.setdefault("cpu_baseline", {"value": 0, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "rank 0 at N = 1 only"})
    _check_common(d, n_gpus)
    assert d["dtype"] == "f32" and d["gpu_launches"] > 0
    assert d["steps"] >= 1 and d["warmup"] >= 3
    # value = all ranks' pictures / the device time of the K steps
    frames = d["config"]["frames_per_step_per_gpu"] * n_gpus
    assert abs(d["value"] - frames / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    # e2e is a separate measurement with real copies
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"] and "f32_value" in d["e2e"] and "d2h_link_gbs" in d["e2e"]
    clk = d["clocks"]
    assert clk["sm_mhz"] >= 0.9 * clk["sm_max_mhz"]
    assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9
    if n_gpus == 1:
        assert d["parity_check"]["ok"] is True and d["parity_check"]["max_abs_err"] <= d["parity_check"]["tolerance"] == 1e-4
        assert len(d["sweep"]["cases"]) == 5
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0


def test_committed_reference_arm_line_has_the_same_config():
    ref = _last_json_line(open(os.path.join(PROFILES, "r02_bench_reference_arm.json")).read())
    own = _last_json_line(open(os.path.join(PROFILES, "r02_bench_final.json")).read())
    _check_common(ref, 1)
    assert ref["impl"] == "reference" and ref["gpu_launches"] == 0
    assert ref["config"] == own["config"]                          # both arms name the same workload, key by key
    assert ref["e2e"]["value"] == ref["value"] and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"]


@pytest.mark.skipif(not os.path.exists(os.path.join(o.ROOT, "oracle", "_ref", "libref_contrast.so")), reason="oracle/_ref is not built")
def test_reference_arm_runs_on_the_host_and_loads_no_product_library(tmp_path):
    env = dict(os.environ, LD_DEBUG="files", LD_DEBUG_OUTPUT=str(tmp_path / "ld"))
    r = subprocess.run([sys.executable, os.path.join(o.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=o.ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json_line(r.stdout)
    _check_common(d, 1)
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    loaded = "".join(open(os.path.join(tmp_path, f)).read() for f in os.listdir(tmp_path))
    assert "libcds_b200.so" not in loaded                          # the CUDA library is never mapped by this arm
    assert "libcds_synth.so" in loaded and "libref_contrast.so" in loaded
