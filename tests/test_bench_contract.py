"""CPU tests of bench.py's contract pieces that do not need a GPU: the reference-arm JSON line (task statement ④), rank gating
under torchrun, the measured-peaks loader and the clock-sampler parser."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _args(**kw):
    base = dict(gpus=1, steps=2, warmup=0, steps_ref=2, warmup_ref=0, workload="e2e", impl="reference")
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_reference_arm_line(monkeypatch, capsys):
    import bench
    calls = []
    monkeypatch.setattr(bench, "cpu_baseline_for", lambda workload, cores: (calls.append((workload, cores)) or (2.5, "stub sample")))
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference_arm(_args())
    line = json.loads(capsys.readouterr().out.strip())
    assert len(calls) == 2 and calls[0][0] == "e2e" and calls[0][1] == (os.cpu_count() or 1)
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] == 2.5 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"] == {"value": 2.5, "unit": bench.UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": "stub sample"}
    assert line["e2e"] == {"value": 2.5, "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == bench.WORKLOADS["e2e"] and line["steps"] == 2 and line["n_gpus"] == 1


def test_reference_arm_only_rank_zero_works(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(bench, "cpu_baseline_for", lambda workload, cores: (_ for _ in ()).throw(AssertionError("rank 1 must not run the CPU arm")))
    monkeypatch.setenv("RANK", "1")
    bench.run_reference_arm(_args(gpus=2))
    assert capsys.readouterr().out == ""


def test_workloads_cover_baseline_configs_and_peaks_load():
    import bench
    assert set(bench.WORKLOADS) == {"e2e", "forward-b1", "cascade", "cascade-sweep", "train"}
    with open(os.path.join(ROOT, "BASELINE.json")) as fp:
        baseline = json.load(fp)
    assert "frames/sec" in baseline["metric"] and bench.UNIT == "frames/s"
    peaks = bench.load_peaks()
    assert peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops_sustained"] <= peaks["bf16_tflops"] and peaks["source"] in ("measured", "fallback")
    # algorithmic bytes of the cascade (SURVEY.md §8 d): 786 432 B of int16 words in, 4 194 304 B of complex64 cube out per frame-sensor
    assert bench.FS_IN_BYTES == 192 * 4 * 256 * 2 * 2 and bench.FS_OUT_BYTES == 16 * 64 * 64 * 8 * 8


def test_clock_sampler_parses_nvidia_smi_rows(tmp_path):
    import bench
    s = bench.ClockSampler(0)
    s.path = str(tmp_path / "clocks.csv")
    with open(s.path, "w") as fp:
        fp.write("0, 345, 1965, 140.2, 0x0, Not Active, Not Active, Not Active, Not Active\n")      # idle sample: ignored for the median
        fp.write("0, 1950, 1965, 990.1, 0x4, Not Active, Not Active, Not Active, Active\n")
        fp.write("0, 1935, 1965, 985.0, 0x4, Not Active, Not Active, Not Active, Active\n")
        fp.write("garbage line\n")

    class _Done(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s.proc = _Done()
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0 and out["sm_mhz"] == 1942.5 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 3


def test_clock_sampler_reports_only_samples_after_mark(tmp_path):
    """The sampler is started (and waited for) before the timed region; mark() right before the region drops the idle-GPU samples taken
    until then, so that a 0.2 s region is described by the samples that fell inside it (a run whose first sample arrived late used to
    report none at all)."""
    import bench
    s = bench.ClockSampler(0)
    s.path = str(tmp_path / "clocks.csv")
    with open(s.path, "w") as fp:
        fp.write("0, 345, 1965, 140.2, 0x0, Not Active, Not Active, Not Active, Not Active\n")
        fp.write("0, 420, 1965, 150.0, 0x0, Not Active, Not Active, Not Active, Not Active\n")
    s.mark()
    assert s.skip == 2
    with open(s.path, "a") as fp:
        fp.write("0, 1800, 1965, 990.1, 0x4, Not Active, Not Active, Not Active, Active\n")
        fp.write("0, 1790, 1965, 985.0, 0x4, Not Active, Not Active, Not Active, Active\n")
        fp.write("0, 1780, 1965, 985.0, 0x4, Not Active, Not Active, Not Active, Active\n")

    class _Done(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s.proc = _Done()
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1790.0 and out["reasons"] == ["sw_power_cap"]
