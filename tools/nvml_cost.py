"""How long do the NVML queries bench.py's clock sampler makes take while the GPU is busy, and do they slow the enqueueing thread?"""
import os, sys, time, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda")
def busy(n):
    t0 = time.perf_counter()
    for _ in range(n):
        y = x @ x
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3, (time.perf_counter() - t0) / n * 1e3
print("no sampler: enqueue ms/iter, total ms/iter", busy(50))
res = {}
stop = False
def loop(which, period):
    while not stop:
        for name, fn in which:
            t0 = time.perf_counter(); fn(); res.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        time.sleep(period)
Q = dict(sm=lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx=lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
         reasons=lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h), power=lambda: nv.nvmlDeviceGetPowerUsage(h))
for names in (["sm"], ["reasons"], ["sm", "reasons"], ["sm", "mx", "reasons"]):
    res.clear(); stop = False
    th = threading.Thread(target=loop, args=([(n, Q[n]) for n in names], 0.05), daemon=True); th.start()
    r = busy(50); stop = True; th.join()
    print(names, "enqueue/total ms per iter", tuple(round(v, 3) for v in r), {k: (len(v), round(sum(v) / len(v), 3), round(max(v), 3)) for k, v in res.items()})
