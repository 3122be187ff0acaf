"""CPU: the JSON contract of bench.py's reference arm and its helpers (no GPU, the oracle step itself is stubbed out)."""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))


def test_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(bench, "cpu_port_step_rate", lambda batch, steps, warmup, threads, variant="image", wtgdl=0.0, fine=128: (batch * 2.0, 0.5, threads))
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=2))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["unit"] == "samples/s" and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same workload string as our arm (the driver compares them), the full 256-sample batch per step when it fits the time budget
    assert line["config"]["workload"] == bench.WORKLOAD and line["config"]["batch_per_step"] == 256
    # a slow host: each step becomes a bounded sample of the batch, and the line says so
    monkeypatch.setattr(bench, "cpu_port_step_rate", lambda batch, steps, warmup, threads, variant="image", wtgdl=0.0, fine=128: (0.5, batch / 0.5, threads))
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=1))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert 2 <= line["config"]["batch_per_step"] < 256 and "bounded sample" in line["config"]["sample"] and line["config"]["workload"] == bench.WORKLOAD
    # the video workload keeps its own workload string
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=1, workload="video", wtgdl=0.5, batch=None))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["config"]["workload"] == bench.WORKLOAD_VIDEO % (0.5, 64)
    # the other ranks of a torchrun launch exit without work or output
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(types.SimpleNamespace(steps=5, warmup=1, gpus=2))
    assert capsys.readouterr().out == ""


def test_clock_sampler_window_and_class_rooflines():
    import bench
    cs = bench.ClockSampler(0)
    cs.samples = [(10.0, 1900.0, 1965.0, 300.0, []), (10.5, 1800.0, 1965.0, 900.0, ["sw_power_cap"]), (11.0, 1700.0, 1965.0, 950.0, []), (12.5, 1000.0, 1965.0, 100.0, ["hw_slowdown"])]
    cs.t_begin, cs.t_end = 10.2, 11.5
    out = cs.finish()
    assert out["samples"] == 2 and out["sm_mhz"] == 1750.0 and out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965.0
    prof = {"names": ["bn_fin_apply", "bn_bwd_fused", "adam", "conv_fwd"], "ms": np.array([0.02, 0.05, 0.4, 0.03]),
            "bytes": np.array([64e6, 160e6, 2.2e9, 0.0])}
    cr = bench.class_rooflines(prof, 6471.1, "measured")
    assert cr["classes"]["bn_fwd"]["achieved"] == 3200.0 and cr["classes"]["bn_bwd"]["launches"] == 1 and "act_bwd" not in cr["classes"]
    assert cr["best_large_layer_launch"]["kernel"] == "bn_fin_apply" and abs(cr["classes"]["adam"]["frac"] - 5500.0 / 6471.1) < 1e-3


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
