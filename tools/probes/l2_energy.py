"""Joules per GB delivered to shared memory: from L2 (slices that stay resident) and from HBM (slices read once), by
plain bulk copies and by cluster multicast (one read, two CTAs).  See l2_energy.cu.   python tools/probes/l2_energy.py"""
import ctypes
import json
import sys
import time
from pathlib import Path

import pynvml
import torch

here = Path(__file__).resolve().parent
lib = ctypes.CDLL(str(here / "l2_energy.so"))
lib.l2_stream.restype = ctypes.c_longlong
lib.l2_stream.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream


def run(name, buf, slice_bytes, passes, mc, ctas, seconds=2.5):
    def fn():
        r = lib.l2_stream(buf.data_ptr(), slice_bytes, passes, mc, ctas, stream)
        assert r > 0
        return r
    for _ in range(3):
        delivered = fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fn()
    e.record()
    torch.cuda.synchronize()
    per = s.elapsed_time(e) / 5
    iters = max(5, int(seconds * 1e3 / per))
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    s.record()
    clocks = []
    for i in range(iters):
        fn()
        if i % 20 == 0:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    e.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    ms = s.elapsed_time(e) / iters
    joules = (e1 - e0) / 1e3 / iters
    out = {"ms": round(ms, 4), "delivered_gbs": round(delivered / ms / 1e6, 1), "watts": round((e1 - e0) / 1e3 / (t1 - t0), 1),
           "joules_per_delivered_gb": round(joules / (delivered / 1e9), 4), "sm_mhz": sorted(clocks)[len(clocks) // 2]}
    print(name, json.dumps(out), flush=True)
    return out


def idle(seconds=2.0):
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    time.sleep(seconds)
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    w = (e1 - e0) / 1e3 / (time.perf_counter() - t0)
    print("idle watts", round(w, 1), flush=True)
    return w


if __name__ == "__main__":
    torch.cuda.set_device(0)
    big = torch.empty(8 << 30, dtype=torch.uint8, device=dev)
    big.zero_()
    idle()
    ctas = 148
    l2_slice = 256 << 10   # 148 x 256 KB = 37 MB: L2-resident after the first pass
    run("l2_plain", big, l2_slice, 64, 0, ctas)
    run("l2_multicast", big, l2_slice, 64, 1, ctas)      # 74 clusters: half the slices, each delivered twice
    hbm_slice = (8 << 30) // 148 // 16384 * 16384
    run("hbm_plain", big, hbm_slice, 1, 0, ctas)
    run("hbm_multicast", big, hbm_slice, 1, 1, ctas)
    idle()
