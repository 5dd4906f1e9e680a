"""The driver's contract for bench.py that can be checked without a GPU: the reference arm prints one JSON line with
the agreed keys (and does no work on ranks other than 0), and the B200 arm refuses to run without a CUDA device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, text=True,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-rays", "64"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["metric"] == "rays/sec (render fwd)" and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["data"] == "synthetic"


def test_reference_arm_other_ranks_do_nothing():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-rays", "64"], env={"RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
