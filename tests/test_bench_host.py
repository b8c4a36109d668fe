"""Host-side pieces of bench.py that need no GPU: the clock sampler's summary and the argument defaults the driver
relies on."""
import importlib
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_clock_sampler_summary_and_reasons():
    bench = _bench()
    s = bench.ClockSampler(0)
    s.thread = threading.Thread(target=lambda: None)
    s.thread.start()
    s.source = "fake"
    # (sm MHz, max MHz, power W, nvmlClocksEventReason mask): sw_power_cap = 0x4, hw_slowdown = 0x8
    s.rows = [(1965.0, 1965.0, 600.0, 0x0), (1950.0, 1965.0, 990.0, 0x4), (1965.0, 1965.0, float("nan"), 0x1)]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["power_w_max"] == 990.0
    assert out["reasons"] == ["sw_power_cap"]          # 0x1 (gpu idle) is not a throttle reason the contract names
    s2 = bench.ClockSampler(0)
    s2.thread = threading.Thread(target=lambda: None)
    s2.thread.start()
    s2.rows = [(900.0, 1965.0, 300.0, 0x8 | 0x40 | 0x20)]
    assert s2.stop()["reasons"] == ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"]


def test_clock_sampler_start_stop_is_safe_with_or_without_a_gpu():
    bench = _bench()
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    if out["samples"] == 0:
        assert out["sm_mhz"] is None
    else:
        assert out["sm_mhz"] > 0 and out["sm_max_mhz"] >= out["sm_mhz"]


def test_bench_defaults(monkeypatch):
    bench = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 5 and a.workload == "c3" and a.impl == "b200"
