"""bench.py's CPU-side contract (no GPU): the reference arm prints one JSON line with the keys the driver reads, runs the
reference's own classes when the notebook is available (kind "reference"), and non-zero ranks of a multi-rank launch exit
without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--batch", "8",
                        *args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line():
    out = _run({})
    line = json.loads(out.splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    from oracle import load_reference
    assert line["cpu_baseline"]["kind"] == ("reference" if load_reference.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1


def test_reference_arm_other_ranks_do_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == ""
