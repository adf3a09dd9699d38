"""Which NVML query stalls kernel submission?  Times each query while a stream of small kernels runs."""
import time

import pynvml as nv
import torch

nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(1 << 20, device="cuda")
for _ in range(100):
    x.mul_(1.0001)
torch.cuda.synchronize()
qs = {"clock_sm": lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
      "reasons": lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h),
      "power": lambda: nv.nvmlDeviceGetPowerUsage(h),
      "max_clock": lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)}
for name, q in qs.items():
    ts = []
    for rep in range(6):
        for _ in range(2000):   # keep the launch queue busy
            x.mul_(1.0001)
        t = time.perf_counter()
        q()
        ts.append((time.perf_counter() - t) * 1e3)
        torch.cuda.synchronize()
    print(f"{name:10s} ms per call: " + " ".join(f"{v:.2f}" for v in ts))
