"""bench.py pieces that need no GPU: the reference arm's JSON line, the cached synthetic CPU input, and the per-leg DRAM
traffic table (profiles/traffic_rNN.json, newest first) being picked up for legs measured on the captured workload."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from quickdna_b200.synth import synth_np  # noqa: E402


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-reads", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_input_is_the_counter_based_stream_and_is_cached():
    a = bench._cpu_input(3, 5_000_000)
    assert a.dtype == np.uint8 and np.array_equal(a, synth_np(3, 0, 5_000_000))
    b = bench._cpu_input(3, 1_000_000)  # a prefix of the cached stream, not a new buffer
    assert np.shares_memory(a, b) and np.array_equal(b, a[:1_000_000])


def test_leg_traffic_is_filled_only_for_the_captured_workload():
    table, src = bench._traffic_table()  # the newest committed capture: what bench.py itself reads
    assert src == "profiles/" + bench.TRAFFIC_FILES[0]
    legs = table["legs"]
    name = "translate_self_frames"
    algo = legs[name]["algorithmic_bytes_per_launch"]
    extra = {name: {"roofline": {"traffic": None, "algorithmic_bytes_per_launch": algo}},
             "canonical_window_keys": {"packed_2bit": {"roofline": {"traffic": None, "algorithmic_bytes_per_launch":
                                                                    legs["canonical_window_keys/packed_2bit"]["algorithmic_bytes_per_launch"]}}},
             "all_frames_strict": {"roofline": {"traffic": None, "algorithmic_bytes_per_launch": 12345}}}
    bench.fill_leg_traffic(extra)
    assert extra[name]["roofline"]["traffic"] == legs[name]["dram_read"] + legs[name]["dram_write"]
    assert extra["canonical_window_keys"]["packed_2bit"]["roofline"]["traffic"] > 0
    assert extra["all_frames_strict"]["roofline"]["traffic"] is None  # another workload: no claim
    # traffic stays close to the algorithmic bytes for the streaming kernels (what the roofline argument rests on)
    for leg in ("translate_1_frame", "translate_self_frames", "all_frames_strict", "canonical_windows", "windows_all"):
        e = legs[leg]
        assert 0.95 < (e["dram_read"] + e["dram_write"]) / e["algorithmic_bytes_per_launch"] < 1.02, leg
    # round 2: the expansion path reads its windows once (1.26 x algorithmic with the two-pass form)
    e = legs["expansion_windows"]
    assert (e["dram_read"] + e["dram_write"]) / e["algorithmic_bytes_per_launch"] < 1.05


def test_stored_bench_line_keeps_the_contract_and_ends_with_the_roofline_table():
    """The committed N = 1 line (profiles/r02_bench_n1.json): every contract key, the whole-output checks all true, and the
    roofline -- with its compact per-leg table -- as the LAST key, so that a reader of the tail of the line sees it."""
    raw = open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1]
    line = json.loads(raw)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["n_gpus"] == 1 and line["gpu_launches"] == line["steps"]
    assert list(line)[-1] == "roofline" and raw.rfind('"roofline"') > len(raw) - 3000
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.85 < r["frac"] <= 1.0
    assert {"revcomp", "parse", "canonical_windows", "windows_all", "expansion_windows"} <= set(r["legs"])
    checks = line["config"]["checks"]
    assert checks and all(v is True for v in checks.values() if isinstance(v, bool))
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and not line["clocks"]["reasons"]
