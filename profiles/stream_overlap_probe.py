"""Development probe: does issuing consecutive, independent target-assignment batches on S streams (inside one CUDA graph,
fork/join) raise the throughput over issuing them back to back on one stream?  A persistent match launch ends on a tail
during which most SMs idle; another batch's prep/encode can fill it.  cfg2 batches, the bench's rotating buffer sets.

    python profiles/stream_overlap_probe.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, _tensor, anchors, batched, config, synth  # noqa: E402
from jabd_b200._tensor import ptr  # noqa: E402

VAR, THR, IMAGE, BATCH, SETS = (0.1, 0.2), 0.35, (640, 640), 32, 8
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L = _lib.lib()


def st():
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


pri = anchors.Anchors(config.cfg_mnet, image_size=IMAGE).get_anchors()
P = int(pri.shape[0])
pool = synth.make_gt_batch(2, 256, IMAGE)
sets = []
for s in range(SETS):
    tg = pool[s * BATCH:(s + 1) * BATCH]
    gt, offs, _ = batched.pack_targets(tg, dev)
    nb, sumG = len(tg), int(gt.shape[0])
    sets.append(dict(gt=gt, offs=offs, sumG=sumG, B=nb,
                     ws=_tensor.workspace(L.jabd_assign_workspace_bytes(nb, P, sumG), dev),
                     loc=torch.empty((nb, P, 4), dtype=torch.float32, device=dev),
                     conf=torch.empty((nb, P), dtype=torch.int64, device=dev),
                     landm=torch.empty((nb, P, 10), dtype=torch.float32, device=dev)))


def assign(s):
    _lib.call("jabd_assign", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1, 0,
              ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), st())


for s in sets:
    assign(s)
torch.cuda.synchronize()
ref = [(s["loc"].clone(), s["conf"].clone(), s["landm"].clone()) for s in sets]


def graph(nstreams, rounds=1):
    side = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream(dev)
        for x in side:
            x.wait_stream(cur)
        for i, s in enumerate(sets * rounds):      # set s always lands on lane s % nstreams: its reuse is stream-ordered
            k = i % nstreams
            if k == 0:
                assign(s)
            else:
                with torch.cuda.stream(side[k - 1]):
                    assign(s)
        for x in side:
            cur.wait_stream(x)
    return g


CASES = ((1, 1), (4, 4)) if "--quick" in sys.argv else ((1, 1), (2, 1), (3, 1), (4, 1), (8, 1), (4, 2), (4, 4), (4, 8), (2, 4), (8, 4))
print("library:", os.path.basename(_lib.SO_PATH), flush=True)
for ns, rounds in CASES:
    if SETS % ns and rounds > 1:
        continue
    g = graph(ns, rounds)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    reps = max(250 // rounds, 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * SETS * rounds)
    ok = all(torch.equal(a, s["loc"]) and torch.equal(b, s["conf"]) and torch.equal(c, s["landm"])
             for (a, b, c), s in zip(ref, sets))
    print("streams %d, %d steps per graph: %.2f us per step  (%.0f img/s)  outputs equal: %s"
          % (ns, SETS * rounds, us, BATCH / us * 1e6, ok), flush=True)
