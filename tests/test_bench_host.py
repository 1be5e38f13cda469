"""Host-side logic of bench.py that can be checked without a GPU: the watchdog around the extra multi-GPU workloads.
(`bench.py --gpus N` first completes the headline line, then runs cfg 5 row-sharded and cfg 4 batch-split in the same
process group; an extra that stalls must not cost the headline.)"""
import json
import os
import subprocess
import sys
import time

from conftest import ROOT

SCRIPT = r"""
import json, sys, time
sys.path.insert(0, {root!r})
import bench
line = {{"metric": "m", "value": 1.0}}
mode, rank = sys.argv[1], int(sys.argv[2])
def ok(): return {{"MP/s": 5.0}}
def boom(): raise RuntimeError("exchange failed")
def hang(): time.sleep(60)
jobs = {{"fine": [("cfg5", ok), ("cfg4", ok)], "raises": [("cfg5", boom), ("cfg4", ok)], "stalls": [("cfg5", ok), ("cfg4", hang)]}}[mode]
extras = bench.run_extras_under_watchdog(jobs, 1.5, rank, line if rank == 0 else None)
if rank == 0:
    line["extra_workloads"] = extras
    print(json.dumps(line), flush=True)
"""


def run(mode, rank):
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT), mode, str(rank)], capture_output=True, text=True, timeout=120)
    return r, time.perf_counter() - t0


def json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_extras_finish_normally():
    r, _ = run("fine", 0)
    lines = json_lines(r.stdout)
    assert r.returncode == 0 and len(lines) == 1
    assert lines[0]["value"] == 1.0 and lines[0]["extra_workloads"] == {"cfg5": {"MP/s": 5.0}, "cfg4": {"MP/s": 5.0}}


def test_an_extra_that_raises_is_reported_not_fatal():
    r, _ = run("raises", 0)
    lines = json_lines(r.stdout)
    assert r.returncode == 0 and len(lines) == 1
    ex = lines[0]["extra_workloads"]
    assert "exchange failed" in ex["cfg5"]["error"] and ex["cfg4"] == {"MP/s": 5.0}


def test_an_extra_that_stalls_cannot_cost_the_headline():
    """rank 0: exactly ONE line, with the finished extra, the reason, and the intact headline, well before the stalled job
    would have returned; the other ranks leave silently with exit code 0."""
    r, dt = run("stalls", 0)
    lines = json_lines(r.stdout)
    assert r.returncode == 0 and len(lines) == 1 and dt < 40
    ex = lines[0]["extra_workloads"]
    assert lines[0]["value"] == 1.0 and ex["cfg5"] == {"MP/s": 5.0} and "did not finish" in ex["error"] and "cfg4" not in ex
    r1, dt1 = run("stalls", 1)
    assert r1.returncode == 0 and json_lines(r1.stdout) == [] and dt1 < 40
