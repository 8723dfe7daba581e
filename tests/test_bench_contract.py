"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the bench contract
(metric / unit / config of the product arm, impl = reference, cpu_baseline, e2e with zero copy bytes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-frames", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "orb_extract_match_stereo_frames_per_s" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["value"] > 0 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("kitti_stereo_frontend") and line["config"]["cpu_sample_frames_per_step"] == 4
    assert line["config"]["frames_per_step_per_gpu"] == 128   # the product arm's batch: same config keys on both arms
    cb = line["cpu_baseline"]
    # "reference" = oracle/_ref (the reference's own sources, prebuilt here), "port" = the C oracle when that library is absent
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libslamref.so")):
        assert cb["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["matches_per_s"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
