"""Host<->device copy rates of this box for the sizes the host-buffer entry points move (pinned memory, CUDA events) --
a development probe.  Usage: python profiles/pcie_probe.py"""
import torch

torch.cuda.set_device(0)
for mb in (1.5, 4.3, 8.6, 13.8, 34.4, 128.0):
    n = int(mb * 1e6) // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%s %6.1f MB: %.3f ms  %.1f GB/s" % (name, mb, ms, n * 4 / ms / 1e6), flush=True)
