"""Diagnose step time: wall (CUDA events) vs host enqueue time vs per-kernel sum, with the event profiler on / off."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianavatar_b200 import _lib
from gaussianavatar_b200.trainer import Stage1Trainer
from gaussianavatar_b200.workload import Stage1Workload

wl = Stage1Workload(3, 2, device="cuda:0")
wl.make_ground_truth()
tr = Stage1Trainer(wl.model)
def run(n, start):
    t0 = time.perf_counter()
    for i in range(n):
        tr.step(wl.device_batch(wl.frame_ids(start + i)), 5000 + i, epoch=1)
    return (time.perf_counter() - t0) / n * 1e3
run(3, 0); torch.cuda.synchronize()
for prof in (False, True, False):
    _lib.profile(prof); _lib.profile_report()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); host = run(10, 10); e1.record(); torch.cuda.synchronize()
    rep = _lib.profile_report()
    ksum = sum(ms for _, ms in rep.values()) / 10
    print(f"profile={prof}: wall {e0.elapsed_time(e1)/10:.2f} ms/step, host enqueue {host:.2f} ms/step, kernel sum {ksum:.2f} ms/step", flush=True)
_lib.profile(False)
# phase split
def timed(fn, n=10):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
m = wl.model
S = 512
print("decoder fwd only", timed(lambda: m.net.forward_packed(m.geo_feature.detach(), S, 2)), "ms")
dec = m.net.forward_packed(m.geo_feature, S, 2)
g = torch.randn_like(dec)
print("decoder fwd+bwd", timed(lambda: m.net.forward_packed(m.geo_feature, S, 2).backward(g)), "ms")
