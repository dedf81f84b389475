"""Development probe: fused detect of consecutive, independent batches on S streams inside one CUDA graph (fork/join) against
back to back on one stream -- does the next batch fill the SMs that the current batch's early-finishing images free?
640^2 x 32 and cfg3 (1024^2 x 16), clustered synthetic predictions, 4 rotating input/output sets."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import anchors, batched, config, synth  # noqa: E402

VAR = (0.1, 0.2)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
SETS = 8
for size, B in ((640, 32), (1024, 16)):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(size, size)).get_anchors()
    sets = []
    for s in range(SETS):
        ls, cs, ms = [], [], []
        for i in range(B):
            gt = synth.make_gt(3, s * B + i, (size, size), count=60)
            l, c, m = synth.make_preds_clustered(3, s * B + i, pri, gt, VAR, device="cuda")
            ls.append(l.cuda()); cs.append(c.cuda()); ms.append(m.cuda())
        loc, conf, landm = torch.stack(ls).contiguous(), torch.stack(cs).contiguous(), torch.stack(ms).contiguous()
        out = batched.detect(loc, conf, landm, pri, VAR)
        sets.append((loc, conf, landm, out))
    torch.cuda.synchronize()
    ref = [tuple(t.clone() for t in o) for (_, _, _, o) in sets]

    def run(st):
        loc, conf, landm, out = st
        batched.detect(loc, conf, landm, pri, VAR, out=out)

    line = []
    for ns in (1, 4, 8):
        for width in ((0,) if ns == 1 else (1, 2)):
            side = [torch.cuda.Stream(dev) for _ in range(ns - 1)]

            def runw(st):
                loc, conf, landm, out = st
                batched.detect(loc, conf, landm, pri, VAR, out=out, cluster=width)
            g = torch.cuda.CUDAGraph()
            for st in sets:
                runw(st)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                cur = torch.cuda.current_stream(dev)
                for x in side:
                    x.wait_stream(cur)
                for i, st in enumerate(sets):
                    if i % ns == 0:
                        runw(st)
                    else:
                        with torch.cuda.stream(side[i % ns - 1]):
                            runw(st)
                for x in side:
                    cur.wait_stream(x)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1) / (20 * SETS)
            ok = all(torch.equal(a, b) for (_, _, _, o), r in zip(sets, ref) for a, b in zip(o, r))
            line.append("S=%d C=%s %.3f ms (%.0f img/s)%s" % (ns, width or "auto", ms_, B / ms_ * 1e3, "" if ok else " MISMATCH"))
    for width in (0, 1, 2):          # jabd_detect_batches without lanes: one launch over the images of all batches
        plan = batched.DetectBatches(pri, [(l, c, m) for (l, c, m, _) in sets], VAR, lanes_n=0, cluster=width)
        outs = plan()
        torch.cuda.synchronize()
        ok = all(torch.equal(a, b) for o, r in zip(outs, ref) for a, b in zip(o, r))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / (20 * SETS)
        line.append("merged C=%s %.3f ms (%.0f img/s)%s" % (width or "auto", ms_, B / ms_ * 1e3, "" if ok else " MISMATCH"))
    print("%dx%d B=%d: %s" % (size, size, B, "; ".join(line)), flush=True)
