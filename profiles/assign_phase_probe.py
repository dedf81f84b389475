"""Development probe for the target-assignment kernels: per-kernel and per-step times of the cfg2 batch (the bench's
rotating buffer sets, CUDA graphs, CUDA events) plus a parity check against the C oracle, for one build of the library.

    JABD_B200_LIB=/path/to/variant.so python profiles/assign_phase_probe.py [tag] [--dense] [--cfg4]

Several variants of the library (make EXTRA=-D... OUT=...) can be compared in one GPU call by running this once per build."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, _tensor, anchors, batched, config, synth  # noqa: E402
from jabd_b200._tensor import ptr  # noqa: E402
from oracle import oracle as orc  # noqa: E402

VAR, THR, IMAGE, BATCH, SETS = (0.1, 0.2), 0.35, (640, 640), 32, 8
tag = next((a for a in sys.argv[1:] if not a.startswith("--")), os.path.basename(_lib.SO_PATH))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L = _lib.lib()
WARM = int(os.environ.get("JABD_PROBE_FLAGS", "0"))


def st():
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def make_sets(pri, batches):
    P = int(pri.shape[0])
    out = []
    for tg in batches:
        gt, offs, _ = batched.pack_targets(tg, dev)
        nb, sumG = len(tg), int(gt.shape[0])
        out.append(dict(host=tg, gt=gt, offs=offs, sumG=sumG, B=nb, P=P, pri=pri,
                        ws=torch.zeros(max(int(L.jabd_assign_workspace_bytes(nb, P, sumG)), 256), dtype=torch.uint8, device=dev),
                        loc=torch.empty((nb, P, 4), dtype=torch.float32, device=dev),
                        conf=torch.empty((nb, P), dtype=torch.int64, device=dev),
                        landm=torch.empty((nb, P, 10), dtype=torch.float32, device=dev)))
    return out


def assign(s, flags=0):
    _lib.call("jabd_assign", ptr(s["pri"]), s["P"], ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1, flags | WARM,
              ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), st())


def match(s, flags=0):
    _lib.call("jabd_assign_match", ptr(s["pri"]), s["P"], ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], flags | WARM, ptr(s["ws"]),
              s["ws"].numel(), st())


def encode(s):
    _lib.call("jabd_assign_encode", ptr(s["pri"]), s["P"], ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1,
              ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), st())


def graph_of(fn, sets):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s in sets:
            fn(s)
    return g


def time_graph(g, n, per):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * per) * 1e3


def check(s):
    ref = orc.match_batch(THR, [t.numpy() for t in s["host"]], s["pri"].cpu().numpy(), list(VAR))
    assign(s)
    torch.cuda.synchronize()
    ok = np.array_equal(s["conf"].cpu().numpy(), ref["conf_t"]) and np.array_equal(s["landm"].cpu().numpy(), ref["landm_t"])
    lt = s["loc"].cpu().numpy()
    ok = ok and np.array_equal(lt[..., :2], ref["loc_t"][..., :2]) and np.allclose(lt, ref["loc_t"], rtol=1e-5, atol=1e-6)
    return ok


pri = anchors.Anchors(config.cfg_mnet, image_size=IMAGE).get_anchors()
sets = make_sets(pri, [synth.make_gt_batch(2, BATCH, IMAGE, first_image=s * BATCH) for s in range(SETS)])
ok = all(check(s) for s in sets[:2])
ok2 = check(sets[0])          # a second call on the same workspace (warm-workspace protocols)
n = 150
g_step = graph_of(assign, sets)
us_step = time_graph(g_step, n, SETS)
us_split = us_step
have_phases = True
try:
    g_prep = graph_of(lambda s: match(s, 2), sets)
    g_pm = graph_of(match, sets)
    for s in sets:
        assign(s)             # leave the workspaces in their post-call state before the encode-only graph
    g_enc = graph_of(encode, sets)
    us_prep, us_pm, us_enc = time_graph(g_prep, n, SETS), time_graph(g_pm, n, SETS), time_graph(g_enc, n, SETS)
except Exception as e:        # a variant may not support the split entry points
    have_phases = False
    us_prep = us_pm = us_enc = float("nan")
line = "%-28s parity %s/%s  step %6.2f us (split %6.2f)  prep %5.2f  match %6.2f  encode %6.2f" % (tag, "OK" if ok else "FAIL", "OK" if ok2 else "FAIL",
                                                                              us_step, us_split, us_prep, us_pm - us_prep, us_enc)
if "--dense" in sys.argv:
    g_d = graph_of(lambda s: match(s, 1), sets)
    line += "  dense-match %6.2f" % (time_graph(g_d, 40, SETS) - us_prep)
if "--cfg4" in sys.argv:
    pri4 = anchors.cached_priors(config.cfg_mnet, (2048, 2048), dev)
    s4 = make_sets(pri4, [synth.make_gt_batch(4, 1, (2048, 2048), first_image=i) for i in range(2)])
    g4 = graph_of(assign, s4)
    line += "  cfg4 1 image %6.2f us" % time_graph(g4, 50, 2)
    s1 = make_sets(pri, [[synth.make_gt(1, 0, IMAGE)]] * 4)
    line += "  cfg1 %5.2f us" % time_graph(graph_of(assign, s1), 200, 4)
print(line, flush=True)
