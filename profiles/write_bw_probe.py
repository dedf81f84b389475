"""How fast can a B200 absorb a pure WRITE stream of the size one target-assignment step produces?  (torch fill_/copy_ on rotating
buffers larger than L2, one CUDA graph of 8 launches, CUDA events.)  The encode kernel writes 34.4 MB per launch and reads ~9 MB."""
import torch
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)


def time_graph(fn, per, n=100):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * per) * 1e3


for mb in (34.4, 68.8, 137.6):
    nbytes = int(mb * 1e6) // 16 * 16
    bufs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(10)]
    srcs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(10)]
    us_fill = time_graph(lambda: [b.view(torch.float32).fill_(1.0) for b in bufs], len(bufs))
    us_copy = time_graph(lambda: [b.copy_(s) for b, s in zip(bufs, srcs)], len(bufs))
    print("%6.1f MB  fill %6.2f us = %6.0f GB/s written   copy %6.2f us = %6.0f GB/s read+written" % (
        mb, us_fill, nbytes / us_fill / 1e3, us_copy, 2 * nbytes / us_copy / 1e3), flush=True)
