"""CPU: host-side pieces of bench.py that decide what the JSON line says -- the clock sampler's in-region window, the
run-time parse of the committed ncu summary (`roofline.traffic`), the measured-peak lookup and the workload description."""
import importlib.util
import os

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


class _FakeProc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass

    def poll(self):
        return 0


def _line(mhz, power_cap="Not Active", thermal="Not Active"):
    return "%d, 1965, 900.0, Not Active, %s, Not Active, %s" % (mhz, thermal, power_cap)


def test_clock_sampler_counts_only_the_samples_of_the_timed_region():
    s = bench.ClockSampler(0)
    s.proc = _FakeProc()
    # start-up / warm-up samples (idle clocks, a thermal flag that must NOT leak into the report), then the timed region
    s.lines = [_line(345, thermal="Active"), _line(1965)]
    first = s.mark()
    s.lines += [_line(1700, "Active"), _line(1650, "Active"), _line(1600, "Active")]
    last = s.mark()
    s.lines += [_line(1580, "Active"), _line(300)]  # the sample in flight when the region ended, then idle again
    r = s.stop(first, last)
    assert r["samples"] == 4 and r["sm_mhz"] == 1625.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_in_region_samples_falls_back_to_the_latest_one():
    s = bench.ClockSampler(0)
    s.proc = _FakeProc()
    s.lines = [_line(1800, "Active")]
    r = s.stop(s.mark(), s.mark())  # a region shorter than the 200 ms sampling period
    assert r["samples"] == 1 and r["sm_mhz"] == 1800.0
    s2 = bench.ClockSampler(0)  # nvidia-smi missing
    assert s2.stop()["sm_mhz"] is None


def test_roofline_traffic_is_parsed_from_the_newest_committed_ncu_summary():
    traffic, src = bench.ncu_traffic()
    assert src is not None and src.startswith("profiles/r02_ncu_gemm_full_summary.txt")
    assert 5e7 < traffic < 1e9  # bytes per launch: operands + outputs of one trunk GEMM


def test_workload_description_names_the_baseline_config():
    cfg = bench.make_config(16, 2, 384, 512)
    assert "cfg-2" in cfg["workload"] and "512x384" in cfg["workload"]
    assert "model" not in cfg
