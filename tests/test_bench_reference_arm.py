"""The CPU arm of bench.py (`--impl reference`): the shape of its JSON line, with the oracle's pair loop stubbed out (no GPU, no
20-second CPU sample).  A step of that arm IS the bounded sample: ms_per_step must be the sample's measured wall time, not the
time the full configuration would take at the measured rate (that goes into ms_per_full_step_extrapolated)."""
import argparse
import json

import bench


def test_reference_line_reports_the_sampled_step(monkeypatch, capsys):
    calls = []

    def fake_rate(target_s=12.0):
        calls.append(target_s)
        fake_rate.last_seconds = 2.5
        return 4.0e9, 7, "123 codons x 10000 seqs of config 2 (stub)"

    monkeypatch.setattr(bench, "cpu_pair_rate", fake_rate)
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(argparse.Namespace(steps=2, warmup=1, gpus=1))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert len(calls) == 3                                    # warm-up + two timed steps
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["value"] == 4.0e9 and line["ms_per_step"] == 2500.0
    assert abs(line["ms_per_full_step_extrapolated"] - 1e3 * bench.nominal_pairs(bench.N_SEQS, 9677) / 4.0e9) < 1e-6
    assert line["cpu_baseline"] == {"value": 4.0e9, "unit": bench.UNIT, "cores": 7, "kind": "port", "sample": "123 codons x 10000 seqs of config 2 (stub)"}
    assert line["e2e"] == {"value": 4.0e9, "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"] == bench.workload_config(1)


def test_reference_arm_is_silent_on_other_ranks(monkeypatch, capsys):
    monkeypatch.setattr(bench, "cpu_pair_rate", lambda target_s=12.0: (_ for _ in ()).throw(AssertionError("rank 1 must not work")))
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(argparse.Namespace(steps=1, warmup=0, gpus=2))
    assert capsys.readouterr().out == ""
