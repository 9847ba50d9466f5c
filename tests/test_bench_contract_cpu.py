"""bench.py's output contract, checked without a GPU through the `--impl reference` arm (the CPU restatement on a tiny
graph): stdout carries exactly ONE line and it is the JSON object the driver parses; nothing a C library writes to file
descriptor 1 can get in front of it; the arm survives the driver's own launch line (torchrun exports its rendezvous
environment to every rank -- the arm's gloo workers must not pick it up)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _check_line(stdout: str, n_gpus: int):
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["unit"] == "epochs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["steps"] >= 1 and d["steps"] <= d["steps_requested"]                 # epochs actually timed
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    return d


def test_reference_arm_prints_exactly_one_json_line(built):
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--shape", "tiny", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    _check_line(p.stdout, 1)


@pytest.mark.parametrize("world", [2])
def test_reference_arm_under_the_drivers_launch_line(built, world):
    """`python -m torch.distributed.run ... bench.py --impl reference --gpus N`: rank 0 alone works (as N gloo processes
    of its own), the other ranks leave with 0; one line on stdout."""
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29741", BENCH, "--impl", "reference", "--gpus", str(world), "--shape", "tiny",
           "--steps", "1", "--warmup", "0", "--watchdog", "240"]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=400)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _check_line(p.stdout, world)
    assert f"{world} gloo processes" in d["cpu_baseline"]["sample"]


def test_stdout_belongs_to_the_json_line():
    """After `protect_stdout()` a write to descriptor 1 from anywhere (here: os.write, standing in for NCCL's banner)
    lands on stderr; `emit` still reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "bench.protect_stdout()\n"
            "os.write(1, b'NCCL version 0.0.0+noise\\n')\n"
            "print('python-level print')\n"
            "sys.stdout.flush()\n"
            "bench.emit({'ok': True})\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == '{"ok": true}\n', p.stdout
    assert "NCCL version 0.0.0+noise" in p.stderr and "python-level print" in p.stderr
