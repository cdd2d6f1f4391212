"""bench.py keeps the driver's JSON contract: the reference arm on CPU, the B200 arm on a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the reference's own CPU implementation (oracle/_ref when built, else the
    oracle port), same metric/unit/config keys, no GPU work."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "foreman_8x8_pm12")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "foreman_8x8_pm12" and d["dtype"] == "u8"
    assert d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


@pytest.mark.parametrize("workload,kind", [("diamond_foreman_8x8_pm12", "port")])
def test_reference_arm_of_the_widened_rows(workload, kind):
    """Fast patterns have no reference implementation: their CPU arm is the definition's port."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", workload)
    assert d["impl"] == "reference" and d["config"]["workload"] == workload and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == kind and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "2", "--warmup", "3", "--pairs", "4", "--no-cpu-baseline", "--sustained-s", "0.3",
                  "--dropin-calls", "5")
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "1080p_frames_per_sec_full_search_pm32" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["config"]["workload"] == "1080p_16x16_pm32" and d["kernel"] == "tiled" and d["fallback_launches"] == 0
    assert d["parity_checked"] is True and d["parity"]["blocks"] == 8160
    assert d["e2e_dropin"]["ms_per_call"] > 0 and d["sustained"]["seconds"] > 0.2
    assert d["value"] > 1000 and d["gpu_launches"] >= 2
    r = d["roofline"]
    assert r["bound"] == "int_alu" and 0.3 < r["frac"] < 1.1 and r["peak"] > 30 and r["unit"] == "T lane-instr/s"
    e = d["e2e"]
    assert e["value"] > 500 and e["h2d_bytes_per_step"] == 2 * 4 * 1920 * 1080 and e["d2h_bytes_per_step"] > 0
    assert "clocks" in d and "reasons" in d["clocks"]


@pytest.mark.gpu
@pytest.mark.parametrize("workload,pairs", [("ssim_1080p_16x16_pm32", "4"), ("diamond_1080p_16x16_pm32", "8")])
def test_b200_arm_widened_rows(workload, pairs):
    d = run_bench("--steps", "2", "--warmup", "3", "--pairs", pairs, "--no-cpu-baseline", "--workload", workload,
                  "--sustained-s", "0.3")
    assert d["parity_checked"] is True
    assert BASE_KEYS <= set(d) and d["config"]["workload"] == workload
    assert d["value"] > 100 and d["e2e"]["value"] > 100 and d["gpu_launches"] >= 2
    assert d["config"]["cost"] in ("mse", "ssim") and d["config"]["search"] in ("full", "three_step", "diamond")
    if d["config"]["search"] != "full":
        assert d["candidate_evaluations_per_s"] > 1e8


def test_both_arms_build_the_same_config():
    """The driver compares the `config` objects of the two arms: both come from workload_config()."""
    sys.path.insert(0, ROOT)
    import bench
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "foreman_8x8_pm12", "--gpus", "1")
    assert d["config"] == bench.workload_config("foreman_8x8_pm12", 512, 1)
    assert d["config"]["l2"] and d["config"]["parallelism"] == "frame-pair sharding x1"


def test_ingest_helper_plan_is_conservative():
    """bench.plan_ingest_helpers: starved ranks borrow at most 80 % of a donor's spare host-link rate; the numbers are
    the ones measured on the pool's 8-GPU box (profiles/h2d_probe_r02.txt; GPUs 0-3 share one uplink)."""
    import bench
    bw = [24.3, 24.2, 24.3, 24.3, 37.6, 37.9, 38.1, 38.0]
    plan = bench.plan_ingest_helpers(bw, [29.5] * 8, 16)
    assert sorted(plan) == [0, 1, 2, 3]                          # the four GPUs behind the shared uplink
    assert sorted(h[0] for h in plan.values()) == [4, 5, 6, 7]   # one donor each
    assert all(h[1] == 3 for h in plan.values())                 # 3 of 16 pairs: 5.5 GB/s of 8.3 GB/s spare
    # a kernel that needs less than every link delivers: nobody detours
    assert bench.plan_ingest_helpers(bw, [20.0] * 8, 16) == {}
    # donors without spare capacity are not used
    assert bench.plan_ingest_helpers([24.0, 24.0, 30.0, 30.0], [29.5] * 4, 16) == {}
    # the same plan on every rank (pure function of the gathered numbers), and the forced variant for experiments
    assert bench.plan_ingest_helpers(bw, [29.5] * 8, 16, hp_force=2)[0][1] == 2
