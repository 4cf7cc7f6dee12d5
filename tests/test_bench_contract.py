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

    def fake_rate(sample_shape, workers=None, seed=0):
        calls.append(tuple(sample_shape))
        return 5.0e6, {"total": float(np.prod(sample_shape)) / 5.0e6, "workers": workers or 1, "blocks": 8}

    monkeypatch.setattr(bench, "cpu_oracle_rate", fake_rate)
    monkeypatch.setenv("RANK", "0")
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="reference", quick=True, no_cpu=False, no_e2e=False, config=2)
    bench.run_reference(args)
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "voxels/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["dtype"] == "u8"
    assert line["value"] == 5.0e6 and line["ms_per_step"] > 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert calls.count((50, 500, 500)) == 2          # one bounded sample per timed step, after one small warm-up
    # the other ranks of a torchrun launch exit without work and without a line
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args)
    assert capsys.readouterr().out == ""


def test_cpu_leg_runs_one_block_of_the_workload():
    rate, tm = bench.cpu_oracle_rate(bench.BLOCK, workers=1)
    assert rate > 0 and tm["blocks"] == 1 and tm["total"] > 0


def test_algorithmic_bytes_per_voxel():
    # 3 (u8 affinities) + 8 (fragments) + 8 per threshold
    assert bench.BYTES_PER_VOXEL == 3 + 8 + 8 * len(bench.THRESHOLDS) == 35
