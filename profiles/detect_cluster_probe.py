"""Fused detect (jabd_detect) timed at every CTAs-per-image width of the thread-block-cluster kernel (1, 2, 4, 8 and the
automatic choice) on the 640^2 x 32 and cfg3 (1024^2 x 16) clustered synthetic predictions, plus single-image calls --
a development probe for DESIGN.md, not a bench line.  Usage: python profiles/detect_cluster_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, anchors, batched, config, synth  # noqa: E402

VAR = (0.1, 0.2)
torch.cuda.set_device(0)
for (size, B, gen) in ((640, 32, "B"), (1024, 16, "B"), (1024, 16, "A"), (640, 1, "B"), (1024, 1, "B"), (640, 64, "B"), (2048, 8, "B")):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(size, size)).get_anchors()
    P = pri.shape[0]
    ls, cs, ms = [], [], []
    for i in range(B):
        gt = synth.make_gt(3, i, (size, size), count=60)
        l, c, m = synth.make_preds_clustered(3, i, pri, gt, VAR, device="cuda") if gen == "B" else synth.make_preds_random(3, i, P)
        ls.append(l.cuda()); cs.append(c.cuda()); ms.append(m.cuda())
    loc, conf, landm = torch.stack(ls).contiguous(), torch.stack(cs).contiguous(), torch.stack(ms).contiguous()
    ref, line = None, []
    for width in (1, 2, 3, 4, 5, 6, 7, 8, 0):
        for _ in range(3):
            out = batched.detect(loc, conf, landm, pri, VAR, cluster=width)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = batched.detect(loc, conf, landm, pri, VAR, cluster=width)
        e1.record()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 20
        if ref is None:
            ref = out
        assert all(torch.equal(a, b) for a, b in zip(out, ref)), width
        line.append("C=%s %.3f ms (%.0f img/s)" % (width or "auto", ms_, B / ms_ * 1e3))
    print("%dx%d B=%d gen %s kept %.0f: %s" % (size, size, B, gen, ref[1].float().mean().item(), "; ".join(line)), flush=True)

# host-buffer pipeline depth (HostDetect.submit / wait): 640^2 x 32 and cfg3
for (size, B) in (((640, 32), (1024, 16)) if "--host" in sys.argv else ()):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(size, size)).get_anchors()
    P = pri.shape[0]
    ls, cs, ms = [], [], []
    for i in range(B):
        gt = synth.make_gt(3, i, (size, size), count=60)
        l, c, m = synth.make_preds_clustered(3, i, pri, gt, VAR, device="cuda")
        ls.append(l.cpu()); cs.append(c.cpu()); ms.append(m.cpu())
    lp, cp, mp = torch.stack(ls).contiguous().pin_memory(), torch.stack(cs).contiguous().pin_memory(), torch.stack(ms).contiguous().pin_memory()
    for width, depth in ((1, 1), (1, 2), (1, 3), (0, 1), (0, 2), (0, 3), (1, 2), (0, 2)):
        hd = batched.HostDetect(pri, B, depth=depth)
        for _ in range(3):
            hd(lp, cp, mp, cluster=width)
        n, pend = 40, []
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for k in range(n):
            pend.append(hd.submit(lp, cp, mp, cluster=width))
            if len(pend) >= depth:
                hd.wait(pend.pop(0))
        while pend:
            hd.wait(pend.pop(0))
        dt = (time.perf_counter() - t0) / n
        print("host pipeline %dx%d B=%d width %s depth %d: %.3f ms per batch (%.0f img/s), h2d %d B" % (size, size, B, width or "auto", depth, dt * 1e3, B / dt, hd.last_h2d), flush=True)
