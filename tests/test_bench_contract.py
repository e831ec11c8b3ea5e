"""The benchmark contract, as far as it can be checked without a GPU: the reference arm (`bench.py --impl
reference`, the oracle port timed on host cores) prints ONE JSON line with the keys the driver reads; ranks
other than 0 stay silent; the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BENCH = [sys.executable, str(ROOT / "bench.py")]


def _env(**extra):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(extra)
    return env


def test_reference_arm_prints_one_contract_line(built):
    r = subprocess.run(BENCH + ["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--ref-nx", "40",
                                "--ref-ny", "20"], capture_output=True, text=True, timeout=600, env=_env(), cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "Melem/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and d["scaling"] in ("weak", "strong") and d["value"] > 0 and d["ms_per_step"] > 0
    import argparse
    sys.path.insert(0, str(ROOT))
    import bench
    args = argparse.Namespace(workload="c4", nx=4000, ny=2000)
    assert d["config"] == bench.config_of(args, 1)                     # exactly the GPU arm's config object
    assert "4000x2000" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"]
    assert "strip 4000x" in cb["sample"] and "17203 iterations" in cb["sample"]             # a bounded sample of that workload
    det = cb["sample_detail"]
    est = det["scale_rows"] * (det["seconds_serial_phases"] + det["job_cg_iterations"] * det["seconds_per_cg_iteration"])
    assert abs(est - det["job_seconds_estimated"]) < 1e-6 * est
    assert abs(d["value"] - 16e6 / cb["job_seconds_estimated"] / 1e6) < 1e-9 * d["value"]
    assert d["ms_per_step"] < 60e3 and d["job_ms_estimated"] > d["ms_per_step"]              # the step is the sample, not the job
    assert cb["complete_small"]["cg_iters"] > 0 and "40x20" in cb["complete_small"]["workload"]
    assert cb["faithful_dense"]["cg_iters"] > 0 and cb["faithful_dense"]["seconds"] > 0
    ac = cb["all_cores"]                        # extra figure: the completely solved plate with the CG on every host core
    assert ac["threads"] >= 1 and ac["value"] > 0 and ac["unit"] == d["unit"] and "not the reference" in ac["note"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0                                                              # nothing of ours ran on a GPU


def test_reference_arm_is_silent_on_other_ranks(built):
    r = subprocess.run(BENCH + ["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=120, env=_env(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback(built):
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run(BENCH + ["--steps", "1", "--warmup", "3", "--nx", "16", "--ny", "8"], capture_output=True, text=True,
                       timeout=300, env=_env(), cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.lstrip().startswith("{") for ln in r.stdout.splitlines())              # no benchmark line was produced


def test_no_collective_after_the_ranks_part_ways():
    """bench.py's ranks != 0 leave (barrier, free, close) while rank 0 builds and prints the line: a collective
    helper called by rank 0 alone in that stretch pairs with the other ranks' final barrier and the job hangs at
    exit (it happened once, with a device-memory figure added to the line).  Static check of the source."""
    src = (Path(__file__).resolve().parent.parent / "bench.py").read_text()
    start = src.index("    if rank != 0:\n        barrier()")
    end = src.index("print(json.dumps(line)", start)
    stretch = src[start:end]
    after_return = stretch[stretch.index("return") :]
    for helper in ("max_over_ranks(", "sum_over_ranks(", "barrier()", "all_reduce(", "all_gather("):
        assert helper not in after_return, helper
