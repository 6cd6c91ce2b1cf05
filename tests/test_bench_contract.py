"""bench.py against the driver's contract: ONE JSON line on stdout with the required keys, the reference arm on the host cores
(rank 0 only under torchrun), a loud failure of the native arm without a GPU, and -- on the B200 box -- the native line with
`roofline`, `clocks`, `e2e` and a non-zero `gpu_launches`."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run_bench(*args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-units", "8")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "audio-sec/sec (MFCC+spectral feats)" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("cfg4")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "8 x 2.0 s segments" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = run_bench("--steps", "1", "--warmup", "3", "--no-cpu")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_native_line_on_the_gpu():
    r = run_bench("--hours", "0.1", "--steps", "2", "--warmup", "3", "--no-cpu", "--e2e-steps", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f32" and d["scaling"] == "strong"
    assert d["gpu_launches"] > 0 and d["value"] > 1e5
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000 and 0 < rf["frac"] < 1
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    e = d["e2e"]
    n_seg = d["config"]["segments"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == int(0.1 * 3600 * 44100) * 2 and e["d2h_bytes_per_step"] == n_seg * 24 * 8
    assert e["matches_device_path"] is True and e["f32"]["h2d_bytes_per_step"] == int(0.1 * 3600 * 44100) * 4
    assert d["weak"]["value"] > 1e5 and d["roofline"]["aggregate_ms_per_step"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("wl,extra", [("cfg3", ["--units", "4000"]), ("cfg2", ["--units", "512", "--nfft", "512,2048"]), ("cfg5", ["--seconds", "20"])])
def test_secondary_workloads_on_the_gpu(wl, extra):
    """cfg2 / cfg3 / cfg5 under the same one-line contract (roofline + clocks + e2e)."""
    r = run_bench("--workload", wl, "--steps", "2", "--warmup", "3", "--no-cpu", "--e2e-steps", "1", *extra)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and d["config"]["workload"].startswith(wl)
    assert d["value"] > 0 and d["gpu_launches"] > 0 and 0 < d["roofline"]["frac"] < 1.2 and d["e2e"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    if wl == "cfg2":
        assert [s["n_fft"] for s in d["roofline"]["sweep"]] == [512, 2048]


@pytest.mark.parametrize("wl", ["cfg3", "cfg2", "cfg5"])
def test_reference_arm_secondary_workloads(wl):
    r = run_bench("--impl", "reference", "--workload", wl, "--steps", "1", "--warmup", "0", "--cpu-units", "8")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["workload"].startswith(wl) and d["cpu_baseline"]["kind"] == "port"
