"""bench.py pieces that run without a GPU: the clock sampler's pacing and shutdown (with a stand-in NVML)."""
import sys
import time
import types

import bench


def _fake_nvml(query_seconds):
    nv = types.ModuleType("pynvml")
    nv.NVML_CLOCK_SM = 1
    nv.nvmlInit = lambda: None
    nv.nvmlDeviceGetHandleByIndex = lambda i: i
    nv.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965

    def clock(h, k):
        time.sleep(query_seconds)
        return 1950

    nv.nvmlDeviceGetClockInfo = clock
    nv.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4           # sw_power_cap
    return nv


def test_clock_sampler_paces_itself_and_stops(monkeypatch):
    for query, lo, hi in ((0.0002, 0.02, 0.021), (0.004, 0.09, 0.25)):
        monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(query))
        with bench.ClockSampler(0) as clk:
            time.sleep(0.6)
        t0 = time.perf_counter()
        s = clk.summary()
        assert not clk._thread.is_alive() and time.perf_counter() - t0 < 0.5
        assert s["sm_mhz"] == 1950.0 and s["sm_max_mhz"] == 1965 and s["reasons"] == ["sw_power_cap"]
        assert lo <= clk.period <= hi, clk.period
        assert 2 <= s["samples"] <= 0.6 / lo + 3
