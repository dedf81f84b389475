"""Small driver for ncu captures: warm-up launches of every hot kernel on BASELINE shapes, then ONE profiled pass between
cudaProfilerStart/Stop (ncu --profile-from-start off).  No timing claims here; bench.py times with CUDA events, this file
only feeds the profiler.  Usage: python profiles/prof_run.py [assign|loss|detect|eval|all]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import anchors, batched, config, synth, utils_bbox, utils_map  # noqa: E402
from jabd_b200 import retinaface_training_DIOU as diou  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
VAR = (0.1, 0.2)
torch.cuda.set_device(0)
steps = []
if what in ("assign", "loss", "all"):
    pri = anchors.Anchors(config.cfg_mnet, image_size=(640, 640)).get_anchors()
    P = pri.shape[0]
    tg = [t.cuda() for t in synth.make_gt_batch(2, 32, (640, 640))]
    steps.append(("assign", lambda: batched.assign_targets(pri, tg)))
    # the shape the lanes of jabd_assign_batches run (192-GT work items), serialised by the profiler
    steps.append(("assign_lanes_shape", lambda: batched.assign_targets(pri, tg, tune=(192, 192, 100))))
if what in ("loss", "all"):
    raw = [synth.make_logits(2, i, P) for i in range(32)]
    preds = tuple(torch.stack([r[k] for r in raw]).cuda().requires_grad_(True) for k in range(3))
    tgt = batched.assign_targets(pri, tg)
    tgt_raw = batched.assign_targets(pri, tg, encode=False)

    def loss_step():
        for p in preds:
            p.grad = None
        l, c, m = batched.multibox_loss(preds, *tgt)
        (l + c + m).backward()
        l, c, m = batched.multibox_loss(preds, *tgt_raw, loc_loss="Diou", priors=pri)
        (l + c + m).backward()
    steps.append(("loss", loss_step))
if what in ("detect", "all"):
    pri3 = anchors.Anchors(config.cfg_mnet, image_size=(1024, 1024)).get_anchors()
    locs, confs, lms = [], [], []
    for i in range(16):
        gt = synth.make_gt(3, i, (1024, 1024), count=60)
        l, c, m = synth.make_preds_clustered(3, i, pri3, gt, VAR, device="cuda")
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, lm = torch.stack(locs).cuda(), torch.stack(confs).cuda(), torch.stack(lms).cuda()
    locb = torch.randn((64, pri3.shape[0], 4), device="cuda") * 0.5

    def detect_step():
        d = batched.detect(loc, conf, lm, pri3, VAR)
        utils_bbox.decode(locb, pri3, VAR)
        return d
    steps.append(("detect", detect_step))
    # eight batches of 16 in one launch (jabd_detect_batches without lanes)
    many = [(loc, conf, lm)]
    for j in range(1, 8):
        locs, confs, lms = [], [], []
        for i in range(16 * j, 16 * j + 16):
            gt = synth.make_gt(3, i, (1024, 1024), count=60)
            l, c, m = synth.make_preds_clustered(3, i, pri3, gt, VAR, device="cuda")
            locs.append(l); confs.append(c); lms.append(m)
        many.append((torch.stack(locs).cuda(), torch.stack(confs).cuda(), torch.stack(lms).cuda()))
    plan = batched.DetectBatches(pri3, many, VAR)
    steps.append(("detect_batches", plan))
if what in ("eval", "all"):
    imgs = [synth.make_eval_image(6, i) for i in range(256)]
    ev = ([im[2] for im in imgs], [im[0] for im in imgs], [im[1][2] for im in imgs])
    steps.append(("eval", lambda: utils_map.pr_counters(*ev, 0.4, 1000)))

for name, fn in steps:
    for _ in range(3):
        fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in steps:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled pass done:", ", ".join(n for n, _ in steps))
