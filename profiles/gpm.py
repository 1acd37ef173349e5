"""Steady-state hardware counters without a profiler: NVML GPM (GPU performance monitoring) samples
bracketing a region of the real, multi-lane run - DRAM bandwidth utilisation, SM activity /
occupancy, PCIe and NVLink byte rates.  ncu's per-kernel DRAM bytes come from serialised, cold-cache
replays; this is the traffic of the pipeline as it actually runs (8 lanes, warm L2).

    with GpmWindow(index) as w: ...run...
    w.metrics  ->  {"dram_bw_util_pct": ..., "sm_util_pct": ..., ...} or {"error": "..."}
"""
import time


class GpmWindow:
    IDS = {"graphics_util_pct": 1, "sm_util_pct": 2, "sm_occupancy_pct": 3, "integer_util_pct": 4,
           "dram_bw_util_pct": 10, "fp64_util_pct": 11, "fp32_util_pct": 12,
           "pcie_tx_MBps": 20, "pcie_rx_MBps": 21, "nvlink_rx_MBps": 60, "nvlink_tx_MBps": 61}

    def __init__(self, index: int):
        self.index, self.metrics, self._s1, self._h, self._nv = index, {}, None, None, None

    def __enter__(self):
        try:
            import pynvml as nv
            self._nv = nv
            nv.nvmlInit()
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self._s1 = nv.nvmlGpmSampleAlloc()
            nv.nvmlGpmSampleGet(self._h, self._s1)
            self._t0 = time.perf_counter()
        except Exception as e:                                   # GPM unsupported / not permitted
            self.metrics = {"error": f"{type(e).__name__}: {e}"}
            self._s1 = None
        return self

    def __exit__(self, *exc):
        if self._s1 is None:
            return False
        nv = self._nv
        try:
            s2 = nv.nvmlGpmSampleAlloc()
            nv.nvmlGpmSampleGet(self._h, s2)
            mg = nv.c_nvmlGpmMetricsGet_t()
            mg.version = nv.NVML_GPM_METRICS_GET_VERSION
            mg.numMetrics = len(self.IDS)
            mg.sample1, mg.sample2 = self._s1, s2
            for i, mid in enumerate(self.IDS.values()):
                mg.metrics[i].metricId = mid
            nv.nvmlGpmMetricsGet(mg)
            out = {"window_s": round(time.perf_counter() - self._t0, 4)}
            for i, name in enumerate(self.IDS):
                m = mg.metrics[i]
                out[name] = round(float(m.value), 3) if m.nvmlReturn == 0 else None
            self.metrics = out
            nv.nvmlGpmSampleFree(s2)
            nv.nvmlGpmSampleFree(self._s1)
        except Exception as e:
            self.metrics = {"error": f"{type(e).__name__}: {e}"}
        return False
