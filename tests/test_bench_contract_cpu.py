"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints exactly one JSON line
with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    # seconds are the headline (lower is better); the value is the extrapolated build time of the fixed sample
    assert d["metric"] == "isdf_build_seconds" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["cpu_baseline"]["value"] == d["cpu_baseline"]["extrapolated_build_s"] == d["value"]
    for k in ("workload", "nk", "nao", "n0", "nip", "ng", "mesh", "kmesh", "c0", "fit"):   # same keys as the device arm
        assert k in d["config"], k


def test_reference_arm_value_does_not_depend_on_steps():
    """The sample is fixed: --steps / --warmup change how often it is repeated, not what is measured."""
    sys.path.insert(0, ROOT)
    import bench
    cell, kpts, w = bench.make_workload("tiny")
    tables = bench.ao_tables(cell, kpts, w["m0"])
    a = bench.cpu_sample(cell, kpts, w, tables)
    b = bench.cpu_sample(cell, kpts, w, tables)
    assert a["sample"].split(":")[0] == b["sample"].split(":")[0]          # identical sample definition
    assert a["nip"] == b["nip"] and abs(a["value"] - a["extrapolated_build_s"]) == 0.0


def test_flop_model_matches_oracle_model():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import isdf_oracle as O
    a = bench.flop_model(27, 26, 3375, 520, 50653)
    b = O.flop_model(27, 26, 3375, 520, 50653)
    assert a == b and 9.0e12 < a["total"] < 1.0e13      # cfg2 ~ 9.5e12


def test_sweep_mac_count_matches_the_launch_schedule():
    """roofline.achieved counts the MACs of the sweep launches: for a full-rank multiple of 64 that is r(r+64) per
    column (two triangular solves with 64-blocks), and a ragged tail only adds what it uses."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.sweep_macs_per_column(448) == 448 * (448 + 64)
    assert bench.sweep_macs_per_column(64) == 2 * 64 * 64
    r = 413
    dense = sum(min(64 * (i // 64 + 1), r) for i in range(r)) + sum(r - 64 * (i // 64) for i in range(r))
    assert bench.sweep_macs_per_column(r) == dense
    assert bench.sweep_macs_per_column(413) < bench.sweep_macs_per_column(448) * 0.87
