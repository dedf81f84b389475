"""Fits the per-batch cost of target assignment, t = a*B + b*sum(G) + c*sum(ceil(G/64)) + d, on one GPU (CUDA graphs, CUDA
events) -- the weights sharding.local_targets balances with.  Development probe: python profiles/cost_model_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes  # noqa: E402

from jabd_b200 import _lib, _tensor, anchors, batched, config, synth  # noqa: E402
from jabd_b200._tensor import ptr  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pri = anchors.Anchors(config.cfg_mnet, image_size=(640, 640)).get_anchors()
P = int(pri.shape[0])
L = _lib.lib()
rows, ts = [], []
rng = np.random.default_rng(0)
cases = []
for B in (16, 24, 32, 40, 48):
    for g in (5, 30, 64, 65, 100, 128, 129, 200, 300):
        cases.append([g] * B)
for _ in range(20):
    B = int(rng.integers(16, 49))
    cases.append([int(1 + 299 * rng.random() ** 2) for _ in range(B)])
for counts in cases:
    tg = [synth.make_gt(2, 100 + i, (640, 640), count=c) for i, c in enumerate(counts)]
    gt, offs, _ = batched.pack_targets(tg, dev)
    nb, sumG = len(tg), int(gt.shape[0])
    ws = _tensor.workspace(L.jabd_assign_workspace_bytes(nb, P, sumG), dev)
    loc = torch.empty((nb, P, 4), dtype=torch.float32, device=dev)
    conf = torch.empty((nb, P), dtype=torch.int64, device=dev)
    landm = torch.empty((nb, P, 10), dtype=torch.float32, device=dev)

    def run():
        _lib.call("jabd_assign", ptr(pri), P, ptr(gt), ptr(offs), nb, sumG, 0.35, 0.1, 0.2, 0, 1, 0, ptr(loc), ptr(conf), ptr(landm),
                  None, None, None, None, ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            run()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 80 * 1e3
    rows.append([len(counts), sum(counts), sum((c + 63) // 64 for c in counts), 1.0])
    ts.append(us)
A, y = np.array(rows, float), np.array(ts)
coef, res, *_ = np.linalg.lstsq(A, y, rcond=None)
pred = A @ coef
print("us = %.4f*B + %.5f*sumG + %.4f*segments + %.3f   (max rel err %.1f%%, rms %.2f us)" %
      (coef[0], coef[1], coef[2], coef[3], 100 * np.max(np.abs(pred - y) / y), np.sqrt(np.mean((pred - y) ** 2))))
print("in GT-equivalents: image %.1f, segment %.1f" % (coef[0] / coef[1], coef[2] / coef[1]))
A2 = A[:, [0, 1, 3]]
c2, *_ = np.linalg.lstsq(A2, y, rcond=None)
p2 = A2 @ c2
print("without the segment term: us = %.4f*B + %.5f*sumG + %.3f (max rel err %.1f%%); image = %.1f GT" %
      (c2[0], c2[1], c2[2], 100 * np.max(np.abs(p2 - y) / y), c2[0] / c2[1]))
