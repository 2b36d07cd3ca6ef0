"""bench.py contract on the CPU side: the reference arm (`--impl reference`) prints ONE JSON line with
the keys the driver reads, on a small sample; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-ncell", "24"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_json_line():
    lines = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "timesteps_per_s" and d["unit"] == "timesteps/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] * 1e-3 - 1.0) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None
    # same config as the GPU arm's default workload; the value is the bounded sample's, scaled in proportion to the cells
    import bench
    assert d["config"]["workload"] == bench.workload_name(4096, 4) and d["config"]["grid_nodes"] == [4097, 4097]
    assert d["config"]["sample_grid_nodes"] == [25, 25]
    f = cb["scale_factor"]
    assert abs(f - (4096 / 24) ** 2) < 1e-6 and abs(cb["sample_timesteps_per_s"] / f - d["value"]) < 1e-12 * d["value"] + 1e-18


def test_reference_arm_other_ranks_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_every_profiled_kernel_class_has_a_name():
    """bench.py names the kernel classes of csrc/common.cuh (PLB_K_*) in its breakdown; a class without a name
    once made rank 0 raise after the timed region while the other ranks waited (round 2)."""
    import re
    import bench
    src = open(os.path.join(ROOT, "pylamp_b200", "csrc", "common.cuh")).read()
    ids = [int(v) for v in re.findall(r"PLB_K_[A-Z0-9_]+ = (\d+),", src) if int(v) < 16]
    assert len(ids) >= 12
    for k in ids:
        assert k in bench.CLASS_NAMES, k
    assert bench.class_name(15).startswith("kernel class")
