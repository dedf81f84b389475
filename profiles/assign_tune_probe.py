"""Development probe: the matching kernel's work-list shape (JABD_ASSIGN_TUNE: GT per item in the coarse-tile regime, in the
fine-tile regime, share of coarse tiles) against (a) one step alone on one stream, (b) the matching kernel alone, (c) steps
overlapping on 4 lanes.  cfg2 batches, rotating buffer sets, CUDA graphs; outputs compared with the default shape's.

    python profiles/assign_tune_probe.py [segA,segB,pct ...]
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


def assign(s, flags):
    _lib.call("jabd_assign", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1, flags,
              ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), st())


def match(s, flags):
    _lib.call("jabd_assign_match", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], flags, ptr(s["ws"]), s["ws"].numel(), st())


def graph(fn, nstreams=1, rounds=1):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    side = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream(dev)
        for x in side:
            x.wait_stream(cur)
        for i, s in enumerate(sets * rounds):
            if i % nstreams == 0:
                fn(s)
            else:
                with torch.cuda.stream(side[i % nstreams - 1]):
                    fn(s)
        for x in side:
            cur.wait_stream(x)
    return g


def us_of(g, steps, reps):
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * steps)


for s in sets:
    assign(s, 0)
torch.cuda.synchronize()
ref = [(s["loc"].clone(), s["conf"].clone(), s["landm"].clone()) for s in sets]
us_prep = us_of(graph(lambda s: match(s, 2)), SETS, 150)
NL = [int(a[8:]) for a in sys.argv[1:] if a.startswith("--lanes=")] or [4]
cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:] if not a.startswith("--")] or [(64, 64, 100), (128, 128, 100), (192, 192, 100), (128, 64, 50), (128, 32, 75)]
print("prep alone %.2f us" % us_prep, flush=True)
for a, b, pct in cases:
    fl = (a << 8) | (b << 16) | (pct << 24)
    for s in sets:
        for t in (s["loc"], s["conf"], s["landm"]):
            t.zero_()
    serial = us_of(graph(lambda s: assign(s, fl)), SETS, 150)
    ok = all(torch.equal(x, s["loc"]) and torch.equal(y, s["conf"]) and torch.equal(z, s["landm"]) for (x, y, z), s in zip(ref, sets))
    m = us_of(graph(lambda s: match(s, fl)), SETS, 150) - us_prep
    lanes = ["%d lanes %6.2f us" % (nl, us_of(graph(lambda s: assign(s, fl), nl, 4 if SETS % nl == 0 else 1), SETS * (4 if SETS % nl == 0 else 1), 40)) for nl in NL]
    print("seg_a %3d seg_b %3d coarse %3d%%: step alone %6.2f us  match alone %6.2f us  %s  equal %s" % (a, b, pct, serial, m, "  ".join(lanes), ok), flush=True)

# one call over all 256 images (what merging the batches of a call into one launch trio would give)
if "--big" in sys.argv:
    gt, offs, _ = batched.pack_targets(pool, dev)
    nb, sumG = len(pool), int(gt.shape[0])
    big = [dict(gt=gt, offs=offs, sumG=sumG, B=nb, ws=_tensor.workspace(L.jabd_assign_workspace_bytes(nb, P, sumG), dev),
                loc=torch.empty((nb, P, 4), dtype=torch.float32, device=dev), conf=torch.empty((nb, P), dtype=torch.int64, device=dev),
                landm=torch.empty((nb, P, 10), dtype=torch.float32, device=dev)) for _ in range(2)]
    for a, b, pct in cases:
        fl = (a << 8) | (b << 16) | (pct << 24)
        for s in big:
            assign(s, fl)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for s in big:
                assign(s, fl)
        t = us_of(g, 2 * 8, 40)
        ok = torch.equal(big[0]["conf"][:32], ref[0][1]) and torch.equal(big[0]["loc"][32:64], ref[1][0])
        print("one call of 256 images, seg %d/%d/%d: %.2f us per 32 images  equal %s" % (a, b, pct, t, ok), flush=True)

# a SHORT burst: one graph of 20 steps on 4 lanes between two device synchronisations (what `bench.py --steps 20` times)
if "--burst" in sys.argv:
    import time
    for a, b, pct in cases:
        fl = (a << 8) | (b << 16) | (pct << 24)
        side = [torch.cuda.Stream(dev) for _ in range(3)]
        for s in sets:
            assign(s, fl)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream(dev)
            for x in side:
                x.wait_stream(cur)
            for i, s in enumerate((sets * 3)[:20]):
                if i % 4 == 0:
                    assign(s, fl)
                else:
                    with torch.cuda.stream(side[i % 4 - 1]):
                        assign(s, fl)
            for x in side:
                cur.wait_stream(x)
        ts = []
        for rep in range(12):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 20)
        ts.sort()
        print("burst of 20 steps, seg %d/%d/%d: median %.2f us per step (min %.2f)" % (a, b, pct, ts[len(ts) // 2], ts[0]), flush=True)
