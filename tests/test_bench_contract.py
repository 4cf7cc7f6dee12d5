"""bench.py host logic that runs without a GPU: the reference arm's JSON line (driver contract) and the CPU leg."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    calls = []

    def fake_run(affs, workers=None):
        calls.append(tuple(affs.shape[1:]))
        v = float(np.prod(affs.shape[1:]))
        return 5.0e6, {"total": v / 5.0e6, "workers": workers or 1, "blocks": 8}, None

    monkeypatch.setattr(bench, "cpu_oracle_run", fake_run)
    monkeypatch.setattr(bench, "sample_affs", lambda shape, seed, in_process=True: np.zeros((3,) + tuple(shape), np.uint8))
    monkeypatch.setenv("RANK", "0")
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="reference", quick=True, no_cpu=False, no_e2e=False, config=2)
    bench.run_reference(args)
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "voxels/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["dtype"] == "u8"
    assert line["value"] == 5.0e6 and line["ms_per_step"] > 0
    # the same workload string as the GPU arm prints for N = 1 (the driver's same_config check)
    assert line["config"]["workload"] == bench.workload_string(2, (50, 500, 500), 1, (50, 500, 500), bench.BLOCK, bench.CONTEXT)
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert calls.count((50, 500, 500)) == 2 and line["steps_run"] == 2   # whole volume per timed step, after one small warm-up
    # the other ranks of a torchrun launch exit without work and without a line
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args)
    assert capsys.readouterr().out == ""


def test_reference_arm_never_maps_the_cuda_library():
    """--impl reference must time the CPU implementation in a process that has not loaded libbsnative.so: the input
    comes from a child process (or the numpy generator)"""
    import subprocess
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--quick','--steps','1','--warmup','0'];"
            "import runpy; runpy.run_path('bench.py', run_name='__main__');"
            "print('MAPPED' if 'libbsnative' in open('/proc/self/maps').read() else 'CLEAN')")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "CLEAN" in out.stdout and "MAPPED" not in out.stdout
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["value"] > 0


def test_cpu_leg_runs_one_block_of_the_workload():
    from bootstrapper_b200.synth import synth_affs
    rate, tm, ref = bench.cpu_oracle_run(synth_affs(bench.BLOCK, seed=0), workers=1)
    assert rate > 0 and tm["blocks"] == 1 and tm["total"] > 0 and ref["fragments"].shape == bench.BLOCK


def test_algorithmic_bytes_per_voxel():
    # 3 (u8 affinities) + 8 (fragments) + 8 per threshold
    assert bench.BYTES_PER_VOXEL == 3 + 8 + 8 * len(bench.THRESHOLDS) == 35
