"""bench.py contract pieces that need no GPU: the reference arm (`--impl reference`) on a small mesh."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_line_small_mesh():
    lines = _run(["--impl", "reference", "--size", "48", "--steps", "5", "--warmup", "3"])
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["metric"] == "DoF-updates/s per timestep" and d["higher_is_better"] is True
    # the arm caps its counts and says so: what ran vs what was asked for
    assert d["steps"] == 2 and d["warmup"] == 1 and d["requested_steps"] == 5 and d["requested_warmup"] == 3
    assert d["config"]["dofs"] == 49 * 49 and d["config"]["cells"] == 2 * 48 * 48
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and abs(d["value"] - d["config"]["dofs"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_stay_silent():
    lines = _run(["--impl", "reference", "--gpus", "2", "--size", "32", "--steps", "1", "--warmup", "0"],
                 {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []
