"""Times jabd_nms (pre-decoded boxes, one segment per image) with the division-free pair decision and with JABD_NMS_EXACT_DIV on the cfg3 / 640^2
clustered synthetic predictions -- a development probe, not a bench number.  Usage: python profiles/nms_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, _tensor, anchors, config, synth, utils_bbox  # noqa: E402
from jabd_b200._tensor import ptr  # noqa: E402

VAR = (0.1, 0.2)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
for (size, B, gen) in ((640, 32, "B"), (640, 1, "B"), (1024, 16, "B"), (1024, 16, "A")):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(size, size)).get_anchors()
    P = pri.shape[0]
    boxes, scores = [], []
    for i in range(B):
        gt = synth.make_gt(3, i, (size, size), count=60)
        if gen == "B":
            l, c, m = synth.make_preds_clustered(3, i, pri, gt, VAR, device="cuda")
        else:
            l, c, m = synth.make_preds_random(3, i, P)
        boxes.append(utils_bbox.decode(l.cuda(), pri, VAR))
        scores.append(c.cuda()[:, 1].contiguous())
    bx, sc = torch.stack(boxes).contiguous(), torch.stack(scores).contiguous()
    keep_cap = 750
    keep = torch.empty((B, keep_cap), dtype=torch.int32, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = _tensor.workspace(L.jabd_nms_workspace_bytes(B, P, keep_cap), dev)
    res = {}
    for mode, width in ((0, 1), (0, 2), (0, 0), (256, 0)):      # CTAs per image: 1, automatic

        def run():
            _lib.call("jabd_nms", ptr(bx), P * 4, 4, ptr(sc), P, 1, B, P, 0.02, 2, 5000, 0.4, mode | (width << 12), keep_cap, ptr(keep), ptr(cnt),
                      ptr(ws), ws.numel(), _tensor.stream_of(dev))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        res[(mode, width)] = (e0.elapsed_time(e1) / 20, keep.clone(), cnt.clone())
        if hasattr(L, "jabd_debug_detect_profile"):   # development build (make EXTRA=-DJABD_DET_PROFILE)
            import ctypes
            prof = (ctypes.c_longlong * 16)()
            L.jabd_debug_detect_profile(prof, 1)
            run()
            L.jabd_debug_detect_profile(prof, 1)
            ch = max(prof[15], 1)
            if os.environ.get("JABD_PROBE_WINDOW", "1") == "1":
                print("   mode %d width %s CTA0 cycles: select %d (histogram %d, compaction %d, sort %d + run exchange %d + rank merge %d), decode %d; "
                      "%d windows (mean alive %d), per window (thread 0): kept-list tests %d, wait %d, exchange+compaction %d, triangle %d, wait %d, "
                      "resolve %d" % (mode, width or "auto", prof[0], prof[5], prof[6], prof[2], prof[3], prof[4], prof[1], prof[15], prof[14] // ch,
                                      prof[8] // ch, prof[9] // ch, prof[10] // ch, prof[11] // ch, prof[12] // ch, prof[13] // ch))
                continue
            print("   mode %d width %s CTA0 cycles: select %d (histogram passes %d, compaction %d, sort %d + run exchange %d + rank merge %d), decode %d; %d chunks, per chunk: "
                  "query %d, wait for slowest warp %d, cluster exchange %d, resolve %d, wait for next triangle %d; warp 1: next triangle %d, look-ahead %d" %
                  (mode, width or "auto", prof[0], prof[5], prof[6], prof[2], prof[3], prof[4], prof[1], prof[15], prof[8] // ch, prof[9] // ch,
                   prof[10] // ch, prof[11] // ch, prof[12] // ch, prof[13] // ch, prof[14] // ch))
    a, b, c = res[(0, 1)], res[(0, 0)], res[(256, 0)]
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(b[1], c[1]) and torch.equal(b[2], c[2])
    print("%dx%d B=%d gen %s: one CTA per image %.3f ms, automatic cluster width %.3f ms, exact-div %.3f ms per batch; mean kept %.1f"
          % (size, size, B, gen, a[0], b[0], c[0], b[2].float().mean().item()))
