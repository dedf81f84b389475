"""Small driver for ncu captures: a few launches of each hot kernel on BASELINE shapes (no timing claims here;
bench.py times with CUDA events, this file only feeds the profiler).  Usage: python profiles/prof_run.py [assign|detect|all]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import anchors, batched, config, synth, utils_bbox  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
VAR = (0.1, 0.2)
torch.cuda.set_device(0)
if what in ("assign", "all"):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(640, 640)).get_anchors()
    tg = [t.cuda() for t in synth.make_gt_batch(2, 32, (640, 640))]
    for _ in range(4):
        out = batched.assign_targets(pri, tg)
    torch.cuda.synchronize()
    print("assign ok", int((out[1] != 0).sum()))
if what in ("detect", "all"):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(1024, 1024)).get_anchors()
    P = pri.shape[0]
    locs, confs, lms = [], [], []
    for i in range(16):
        gt = synth.make_gt(3, i, (1024, 1024), count=60)
        l, c, m = synth.make_preds_clustered(3, i, pri, gt, VAR, device="cuda")
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, lm = torch.stack(locs).cuda(), torch.stack(confs).cuda(), torch.stack(lms).cuda()
    for _ in range(3):
        d = batched.detect(loc, conf, lm, pri, VAR)
        b = utils_bbox.decode(loc, pri, VAR)
    torch.cuda.synchronize()
    print("detect ok", int(d[1].sum()))
