"""cfg4 (dense tiny-face stress: 2048x2048, 172,032 priors, 1,500 GT per image) timings on one GPU: target assignment and
detection -- a development probe for DESIGN.md, not a bench line.  python profiles/cfg4_probe.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, _tensor, anchors, batched, config, synth  # noqa: E402
from jabd_b200._tensor import ptr  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
VAR = (0.1, 0.2)
size = (2048, 2048)
pri = anchors.Anchors(config.cfg_mnet, image_size=size).get_anchors()
P = int(pri.shape[0])
L = _lib.lib()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (1, 8):
    tg = [synth.make_gt(4, i, size) for i in range(B)]
    gt, offs, _ = batched.pack_targets(tg, dev)
    sumG = int(gt.shape[0])
    ws = _tensor.workspace(L.jabd_assign_workspace_bytes(B, P, sumG), dev)
    loc = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    conf = torch.empty((B, P), dtype=torch.int64, device=dev)
    landm = torch.empty((B, P, 10), dtype=torch.float32, device=dev)

    def run(flags=0):
        _lib.call("jabd_assign", ptr(pri), P, ptr(gt), ptr(offs), B, sumG, 0.35, 0.1, 0.2, 0, 1, flags, ptr(loc), ptr(conf), ptr(landm),
                  None, None, None, None, ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    ms = timed(run, 20)
    pairs = float(P) * sumG
    print("assign cfg4 B=%d: %.3f ms -> %.0f images/s, dense-equivalent %.1f TFLOP/s (14 ops x P x sumG = %.2f GFLOP), outputs %.1f MB"
          % (B, ms, B / ms * 1e3, 14 * pairs / ms / 1e9, 14 * pairs / 1e9, B * P * 72 / 1e6))
    if B == 1:
        print("   dense (no culling): %.3f ms" % timed(lambda: run(1), 5))

for B in (1, 8):
    locs, confs, lms = [], [], []
    for i in range(B):
        l, c, m = synth.make_preds_clustered(4, i, pri, synth.make_gt(4, i, size), VAR, device=dev)
        locs.append(l); confs.append(c); lms.append(m)
    loc_d, conf_d, lm_d = (torch.stack(x).contiguous().to(dev) for x in (locs, confs, lms))
    ms = timed(lambda: batched.detect(loc_d, conf_d, lm_d, pri, VAR), 10)
    d = batched.detect(loc_d, conf_d, lm_d, pri, VAR)
    ms_u = timed(lambda: batched.detect(loc_d, conf_d, lm_d, pri, VAR, conf_thres=0.5, strict=False, pre_nms_topk=0, nms_thres=0.3, keep_topk=0), 5)
    d2 = batched.detect(loc_d, conf_d, lm_d, pri, VAR, conf_thres=0.5, strict=False, pre_nms_topk=0, nms_thres=0.3, keep_topk=0)
    print("detect cfg4 B=%d: top-5000 / keep 750: %.3f ms (%.0f images/s, kept %.0f);  reference's live setting (>=0.5, IoU 0.3, uncapped): "
          "%.3f ms (kept %.0f)" % (B, ms, B / ms * 1e3, d[1].float().mean().item(), ms_u, d2[1].float().mean().item()))
