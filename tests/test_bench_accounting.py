"""bench.py's roofline accounting (no GPU): SURVEY.md section 8d's algorithmic bytes / flops per view."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_algorithmic_bytes_and_flops_match_the_survey():
    import bench
    a = bench.algorithmic_per_view(1152, 1600, 5, [48, 32, 8], "fp16")
    # CostRegNet 456.3 GFLOP, 4.46 GB of layer-by-layer activation traffic in 2-byte storage
    assert abs(a["conv"]["flops"] / 1e9 - 456.3) < 0.1
    assert abs(a["conv"]["bytes"] / 1e9 - 4.46) < 0.01
    # head: 2*V*4 read + V*4 + 3*h*w*4 written = 0.45 GB
    assert abs(a["head"]["bytes"] / 1e9 - 0.449) < 0.001
    # warp+aggregate with what the pipeline moves: fp16 features, fp32 hypotheses, 2-byte volume
    V = [48 * 288 * 400, 32 * 576 * 800, 8 * 1152 * 1600]
    want = sum(5 * c * h * w * 2 + v * 4 + c * v * 2 for (h, w, c), v in zip(((288, 400, 32), (576, 800, 16), (1152, 1600, 8)), V))
    assert a["warp_agg"]["bytes"] == want
    f32 = bench.algorithmic_per_view(1152, 1600, 5, [48, 32, 8], "fp32")
    assert abs(f32["warp_agg"]["bytes"] / 1e9 - 2.78) < 0.01          # the survey's fp32 figure
    assert f32["conv"]["bytes"] == 2 * a["conv"]["bytes"] and f32["repack"]["bytes"] == 0.0


def test_reference_arm_line_is_bounded_and_self_consistent():
    """--impl reference on a tiny shape: one JSON line, kind reference|port, timed steps within the budget."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--height", "64", "--width", "96",
                          "--nviews", "3", "--ndepths", "16,8,8", "--steps", "3", "--budget-s", "60"],
                         capture_output=True, text=True, timeout=300).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert 1 <= line["steps"] <= 3 and line["steps_requested"] == 3
    assert abs(line["ms_per_step"] * line["steps"] / 1e3 - line["timed_region_s"]) <= 0.5 * line["timed_region_s"] + 0.05
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
