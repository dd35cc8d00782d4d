"""CPU test: the reference arm of bench.py (the reference's own CPU implementation, no GPU involved) prints
exactly ONE JSON line on stdout -- library chatter such as the reference's "Completed N steps" goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "tiny_swe512_rk4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_ncu_figures_belong_to_the_committed_kernel_sources():
    """roofline.fp32_pipe and roofline.traffic come from profiles/instruction_mix.json and profiles/traffic.json; both
    carry the git blob hashes of the kernel sources they were captured from. A kernel edit without a new capture
    must be visible here, not only as `sources_match_capture: false` in the bench line."""
    sys.path.insert(0, ROOT)
    import bench
    now = bench.source_stamp()
    with open(os.path.join(ROOT, "profiles", "instruction_mix.json")) as f:
        mix = json.load(f)
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        traffic = json.load(f)
    assert mix["strict"]["stamp"] == now and mix["folded"]["stamp"] == now
    assert traffic["_stamp"] == now
