"""CPU: the JSON contract of bench.py's reference arm and its helpers (no GPU, the oracle step itself is stubbed out)."""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(bench, "cpu_port_step_rate", lambda batch, steps, warmup, threads: (batch * 2.0, 0.5))
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=2))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["unit"] == "samples/s" and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and 2 <= int(line["config"]["sample"].split()[3]) <= 32     # bounded sample of the 256-sample step
    # the other ranks of a torchrun launch exit without work or output
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=2))
    assert capsys.readouterr().out == ""


def test_peaks_come_from_the_driver_file_or_the_stated_fallback():
    import bench
    hbm, burst, sustained, src = bench.peaks()
    assert src in ("measured", "fallback") and 3000 < hbm < 9000 and 800 < sustained <= burst < 2600


def test_parity_summary_arithmetic():
    import parity_steps
    gold = np.tile(np.array([[1.0, 2.0, 0.5, 0.4, 0.6, 0.5]]), (20, 1))
    ours = gold * 1.01
    s = parity_steps.summarize(ours, gold)
    assert s["steps"] == 20
    for k in parity_steps.NAMES:
        assert abs(s[k]["max_rel_all_steps"] - 0.01) < 1e-9 and abs(s[k]["rel_of_mean_last10"] - 0.01) < 1e-9
