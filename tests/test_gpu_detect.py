"""GPU parity of the inference side (decode, threshold, top-k, NMS, fused detect) against the CPU oracle and
the golden vectors recorded from the reference.

bit-exact: landmark decode, top-k index lists, NMS keep lists (order included) on identical box inputs,
scores.  rtol 1e-5 / atol 1e-6: decoded boxes (exp half; see conftest.RTOL/ATOL).
The fused pipeline is checked twice: (i) keep lists bit-exact against the oracle fed with the GPU-decoded
boxes (isolates NMS decisions from the <= 1 ulp exp difference), and (ii) against the
golden outputs of the real reference.
"""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL, load_golden

pytestmark = pytest.mark.gpu

VAR = [0.1, 0.2]


@pytest.fixture(scope="module")
def mods():
    from jabd_b200 import _ops, anchors, batched, box_utils, config, synth, utils_bbox
    from oracle import oracle as orc
    return dict(ops=_ops, anchors=anchors, batched=batched, box_utils=box_utils, cfgs=config, synth=synth, ub=utils_bbox,
                orc=orc)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_decode_golden_and_oracle(mods):
    g = load_golden("decode.npz")
    ub, orc = mods["ub"], mods["orc"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(160, 160)).get_anchors()
    b = ub.decode(cuda(g["loc"]), pri, VAR)
    np.testing.assert_allclose(b.cpu().numpy(), g["boxes"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(ub.decode_landm(cuda(g["landm"]), pri, VAR).cpu().numpy(), g["landms"])
    # how close is the fp64-exp decode to the reference's bits?  (informational, printed with -s)
    print("decode boxes bit-equal to torch CPU: %.4f" % (b.cpu().numpy() == g["boxes"]).mean())
    pri640 = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    loc, conf, landm = mods["synth"].make_preds_random(1, 0, pri640.shape[0])
    np.testing.assert_allclose(ub.decode(loc.cuda(), pri640, VAR).cpu().numpy(), g["boxes640"], rtol=RTOL, atol=ATOL)
    lm = ub.decode_landm(landm.cuda(), pri640, VAR).cpu().numpy()
    assert np.array_equal(lm[:1024], g["landms640_head"])
    assert np.array_equal(lm, orc.decode_landm(landm.numpy(), pri640.cpu().numpy(), VAR))
    # batched [B,P,*] form equals per-image calls; numpy in -> numpy out
    locb = torch.stack([loc, loc * 0.5, -loc]).cuda()
    outb = ub.decode(locb, pri640, VAR)
    for i in range(3):
        assert torch.equal(outb[i], ub.decode(locb[i], pri640, VAR))
    lmb = torch.stack([landm, landm * 2]).cuda()
    outl = ub.decode_landm(lmb, pri640, VAR)
    assert torch.equal(outl[1], ub.decode_landm(lmb[1], pri640, VAR))
    assert isinstance(ub.decode(loc.numpy(), pri640.cpu().numpy(), VAR), np.ndarray)
    assert np.array_equal(mods["box_utils"].decode(loc.cuda(), pri640, VAR).cpu().numpy(), outb[0].cpu().numpy())


def test_topk_segmented(mods):
    ops = mods["ops"]
    rng = np.random.default_rng(3)
    # heavy ties (quantised scores), several segments, K below / above the survivor count
    s = (rng.integers(0, 50, size=(5, 20000)) / 50.0).astype(np.float32)
    s[3] = 0.25                       # one segment where every score ties
    s[4, ::3] = -s[4, ::3]            # negative scores order correctly
    s[2] = (0.5 + 0.06 * rng.random(20000)).astype(np.float32)   # all in one 11-bit bin: the refinement passes run
    for k, thr in ((1, None), (100, None), (5000, 0.5), (7000, None), (6144, None), (20000, None), (20000, 0.9)):
        idx, cnt = ops.topk(cuda(s), k, conf_thres=thr, strict=True)
        idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
        for r in range(s.shape[0]):
            sel = np.nonzero(s[r] > np.float32(thr))[0] if thr is not None else np.arange(s.shape[1])
            order = sel[np.argsort(-s[r][sel].astype(np.float64), kind="stable")][:k]
            assert cnt[r] == len(order), (k, thr, r)
            assert np.array_equal(idx[r, :cnt[r]], order), (k, thr, r)
            assert (idx[r, cnt[r]:] == -1).all()
    # the all-tied segment cannot be cut by the fine first histogram: the exact three-pass select + ordered compaction ran
    _, _, st = ops.topk(cuda(s), 5000, return_stats=True)
    st = st.cpu().numpy()
    assert st[3, 1] >= 1 and (st[:, 0] >= 1).all() and st[3, 2] == 5000
    idx, cnt = ops.topk(cuda(s[0]), 10, conf_thres=0.98, strict=False)   # >= keeps the 0.98 bucket
    ref = np.nonzero(s[0] >= np.float32(0.98))[0][:10]
    assert np.array_equal(idx.cpu().numpy()[:int(cnt)], ref)
    idx, cnt = ops.topk(cuda(s[0]), 10, conf_thres=2.0)                   # nothing passes
    assert int(cnt) == 0 and (idx.cpu().numpy() == -1).all()


def test_nms_torchvision_golden(mods):
    """Keep lists recorded from torchvision.ops.nms (CPU) through the reference's call site."""
    g = load_golden("nms.npz")
    ops = mods["ops"]
    for name in g["names"]:
        name = str(name)
        b, s = g[name + "_boxes"], g[name + "_scores"]
        n = b.shape[0]
        for thr in (0.3, 0.4, 0.5):
            ref = g["%s_keep_%d" % (name, int(thr * 100))]
            keep, cnt = ops.nms_indices(cuda(b), 4, cuda(s), 1, n, 0.0, ops.THRESH_NONE, 0, thr, ops.NMS_TV, n, torch.device("cuda", 0))
            c = int(cnt.item())
            assert c == len(ref), (name, thr)
            assert np.array_equal(keep[:c].cpu().numpy(), ref), (name, thr)
            assert (keep[c:].cpu().numpy() == -1).all()


def test_nms_ssd_golden(mods):
    g = load_golden("nms.npz")
    bu, ub = mods["box_utils"], mods["ub"]
    for name in ("rand500", "dense2000"):
        b, s = g[name + "_boxes"], g[name + "_scores"]
        for (ov, tk) in ((0.5, 200), (0.3, 50), (0.45, 5000)):
            keep, count = bu.nms(cuda(b), cuda(s), ov, tk)
            assert count == int(g["%s_ssd_%d_%d_count" % (name, int(ov * 100), tk)])
            assert keep.dtype == torch.int64 and keep.shape[0] == b.shape[0]
            assert np.array_equal(keep.cpu().numpy(), g["%s_ssd_%d_%d_keep" % (name, int(ov * 100), tk)])
            keep2, count2 = ub.nms_r(cuda(b), cuda(s), ov, tk)
            assert count2 == count and torch.equal(keep, keep2)
    e = bu.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda())
    assert isinstance(e, torch.Tensor) and e.numel() == 0      # bare keep on empty input (:397-398)


def test_nms_large_uncapped_vs_oracle(mods):
    """More candidates than one selection round (6144) and more keeps than the shared-memory kept cache (1536):
    multi-round streaming top-k feeding NMS, kept boxes spilling to the workspace."""
    ops, orc = mods["ops"], mods["orc"]
    rng = np.random.default_rng(17)
    n = 20000
    c = rng.random((n, 2), dtype=np.float32)
    wh = 0.004 + 0.02 * rng.random((n, 2), dtype=np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(n, dtype=np.float32)
    s[::7] = s[3]                                                # ties across rounds
    for thr in (0.3, 0.5):
        ref = orc.nms_tv(b, s, thr)
        assert len(ref) > 1536
        keep, cnt = ops.nms_indices(cuda(b), 4, cuda(s), 1, n, 0.0, ops.THRESH_NONE, 0, thr, ops.NMS_TV, n, torch.device("cuda", 0))
        c_ = int(cnt.item())
        assert c_ == len(ref) and np.array_equal(keep[:c_].cpu().numpy(), ref)
        # capped keep list == prefix of the uncapped one; pre-NMS top-k == NMS of the k best
        keep, cnt = ops.nms_indices(cuda(b), 4, cuda(s), 1, n, 0.0, ops.THRESH_NONE, 0, thr, ops.NMS_TV, 100, torch.device("cuda", 0))
        assert int(cnt.item()) == 100 and np.array_equal(keep.cpu().numpy(), ref[:100])
        order = orc.argsort_desc(s)[:9000]
        ref_k = order[orc.nms_tv(b[order], s[order], thr)]
        keep, cnt = ops.nms_indices(cuda(b), 4, cuda(s), 1, n, 0.0, ops.THRESH_NONE, 9000, thr, ops.NMS_TV, n, torch.device("cuda", 0))
        c_ = int(cnt.item())
        assert c_ == len(ref_k) and np.array_equal(keep[:c_].cpu().numpy(), ref_k)


def _pipeline_inputs(mods, g, tag, gen):
    size = (160, 160) if tag == "s160" else (640, 640)
    img = 1 if tag == "s160" else 2
    pri = mods["orc"].priors(mods["cfgs"].cfg_mnet, size)
    key = "%s_%s_" % (tag, gen)
    gt = torch.from_numpy(g[key + "gt"])
    if gen == "A":
        loc, conf, landm = mods["synth"].make_preds_random(3, img, pri.shape[0])
    else:
        loc, conf, landm = mods["synth"].make_preds_clustered(3, img, torch.from_numpy(pri), gt, VAR)
    return pri, loc.numpy(), conf.numpy(), landm.numpy()


@pytest.mark.parametrize("tag", ["s160", "s640"])
@pytest.mark.parametrize("gen", ["A", "B"])
def test_pipeline_golden(mods, tag, gen):
    """Drop-in non_max_suppression (>= 0.5 / 0.3, uncapped) and the cfg3 pipeline (> 0.02, top-k, 0.4, keep)
    against outputs of the real reference."""
    g = load_golden("pipeline.npz")
    ub, orc = mods["ub"], mods["orc"]
    pri, loc, conf, landm = _pipeline_inputs(mods, g, tag, gen)
    key = "%s_%s_" % (tag, gen)
    pri_c = cuda(pri)
    boxes = ub.decode(cuda(loc), pri_c, VAR)
    lms = ub.decode_landm(cuda(landm), pri_c, VAR)
    det = torch.cat([boxes, cuda(conf)[:, 1:2], lms], -1)               # R/predict.py:180
    for ct, nt in ((0.5, 0.3), (0.05, 0.3)):
        ref = g[key + "nms_%d_%d" % (int(ct * 100), int(nt * 100))]
        out = ub.non_max_suppression(det, ct, nt)
        out = np.zeros((0, 15), np.float32) if isinstance(out, list) else out
        # exact vs the oracle on the same (GPU-decoded) boxes
        exp = orc.non_max_suppression(det.cpu().numpy(), ct, nt)
        exp = np.zeros((0, 15), np.float32) if isinstance(exp, list) else exp
        assert np.array_equal(out, exp)
        # and vs the real reference
        assert out.shape == ref.shape and out.dtype == np.float32
        assert np.array_equal(out[:, 4], ref[:, 4])
        np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    assert ub.non_max_suppression(det, 1.5, 0.3) == []
    for (ct, topk, nt, keepk) in ((0.02, 5000, 0.4, 750), (0.02, 200, 0.4, 50)):
        k2 = key + "pipe_%d_%d_" % (topk, keepk)
        dets, counts, kidx = mods["batched"].detect(cuda(loc)[None], cuda(conf)[None], cuda(landm)[None], pri_c, VAR,
                                                    conf_thres=ct, strict=True, pre_nms_topk=topk, nms_thres=nt, keep_topk=keepk)
        c = int(counts[0].item())
        e_d, e_i = orc.detect(loc, conf, landm, pri, VAR, ct, True, topk, nt, keepk, boxes_override=boxes.cpu().numpy())
        assert c == len(e_i) and np.array_equal(kidx[0, :c].cpu().numpy(), e_i)
        assert np.array_equal(dets[0, :c].cpu().numpy(), e_d)
        assert (dets[0, c:] == 0).all() and (kidx[0, c:] == -1).all()
        assert np.array_equal(kidx[0, :c].cpu().numpy(), g[k2 + "idx"])
        np.testing.assert_allclose(dets[0, :c].cpu().numpy(), g[k2 + "dets"], rtol=RTOL, atol=ATOL)


def _ulp_jitter(rng, a, k=2):
    """a moved by up to +-k ulp per element."""
    out = a.copy()
    for _ in range(k):
        d = rng.integers(-1, 2, size=a.shape).astype(np.float32)
        out = np.where(d == 0, out, np.nextafter(out, np.where(d > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32)))
    return out.astype(np.float32)


def _margin_check(orc, loc, conf, landm, pn, kidx, counts, dets, params, trials=3):
    """Returns (margin-stable images, images whose GPU keep list differs from the oracle run on its own decode); asserts the
    rows of agreeing images within the coordinate tolerance."""
    ct, strict, topk, nt, keepk = params
    rng = np.random.default_rng(99)
    stable, flips = [], []
    for i in range(loc.shape[0]):
        li, ci, mi = loc[i].numpy(), conf[i].numpy(), landm[i].numpy()
        own = orc.decode(li, pn, VAR)
        p_d, p_i = orc.detect(li, ci, mi, pn, VAR, ct, strict, topk, nt, keepk)
        ok = True
        for _ in range(trials):
            _, q_i = orc.detect(li, ci, mi, pn, VAR, ct, strict, topk, nt, keepk, boxes_override=_ulp_jitter(rng, own))
            ok = ok and np.array_equal(q_i, p_i)
        if ok:
            stable.append(i)
        c = int(counts[i].item())
        if c == len(p_i) and np.array_equal(kidx[i, :c].cpu().numpy(), p_i):
            np.testing.assert_allclose(dets[i, :c].cpu().numpy(), p_d, rtol=RTOL, atol=ATOL)
        else:
            flips.append(i)
    return stable, flips


def test_detect_batches_lanes(mods):
    """jabd_detect_batches: several independent batches on side streams (0, 1, 3, 4 lanes; eagerly, from a side stream and
    replayed from a CUDA graph; automatic and pinned cluster widths) give, batch by batch, the rows of one jabd_detect call
    each; the first batch is also checked against the oracle."""
    import ctypes
    from jabd_b200 import _lib
    orc, synth, bt = mods["orc"], mods["synth"], mods["batched"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    P = pri.shape[0]
    sizes = (3, 1, 5, 2, 4)
    batches, first = [], 0
    for n in sizes:
        ls, cs, ms = [], [], []
        for i in range(first, first + n):
            gt = synth.make_gt(2, i, (640, 640), count=30 + 10 * i)
            l, c, m = synth.make_preds_clustered(2, i, pri.cpu(), gt, VAR) if i % 3 else synth.make_preds_random(2, i, P)
            ls.append(l); cs.append(c); ms.append(m)
        batches.append((torch.stack(ls).cuda(), torch.stack(cs).cuda(), torch.stack(ms).cuda() if n != 2 else None))
        first += n
    want = [bt.detect(l, c, m, pri, VAR) for (l, c, m) in batches]
    loc0, conf0, lm0 = batches[0]
    boxes = mods["ub"].decode(loc0, pri, VAR).cpu().numpy()
    for i in range(sizes[0]):
        e_d, e_i = orc.detect(loc0[i].cpu().numpy(), conf0[i].cpu().numpy(), lm0[i].cpu().numpy(), pri.cpu().numpy(), VAR, 0.02, True,
                              5000, 0.4, 750, boxes_override=boxes[i])
        c = int(want[0][1][i])
        assert c == len(e_i) and np.array_equal(want[0][2][i, :c].cpu().numpy(), e_i) and np.array_equal(want[0][0][i, :c].cpu().numpy(), e_d)

    def same(outs):
        torch.cuda.synchronize()
        return all(torch.equal(a, b) for o, w in zip(outs, want) for a, b in zip(o, w))

    def scrub(outs):
        for o in outs:
            for t in o:
                t.fill_(-3)

    for n_lanes, width in ((0, 0), (0, 1), (0, 3), (1, 0), (3, 0), (4, 0), (4, 1), (4, 3), (8, 2)):   # 0 lanes: all batches in one launch
        plan = bt.DetectBatches(pri, batches, VAR, lanes_n=n_lanes, cluster=width)
        assert same(plan()), (n_lanes, width)
        scrub(plan.outputs)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            outs = plan()
            tot = [o[1].sum() for o in outs]
        s.synchronize()
        assert [int(x) for x in tot] == [int(w[1].sum()) for w in want]
        assert same(outs)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan()
        scrub(plan.outputs)
        g.replay()
        assert same(plan.outputs), ("graph", n_lanes, width)
    assert same(bt.detect_batches(pri, batches[:2], variances=VAR, lanes_n=2))
    # refused before anything is enqueued: a lane equal to the calling stream, a bad cluster width, two lanes on one workspace
    plan = bt.DetectBatches(pri, batches[:2], VAR, lanes_n=1)
    L = _lib.lib()
    cur = torch.cuda.current_stream().cuda_stream
    lane = (ctypes.c_void_p * 1)(cur)
    args = (pri.data_ptr(), P, ctypes.cast(plan.arr, ctypes.c_void_p), 2, 0.1, 0.2, 0.02, 2, 5000, 0.4, 750)
    assert L.jabd_detect_batches(*args, 0, ctypes.cast(lane, ctypes.c_void_p), 1, ctypes.c_void_p(cur)) == -1
    assert "calling stream" in _lib.last_error()
    assert L.jabd_detect_batches(*args, 9, None, 0, ctypes.c_void_p(cur)) == -1 and "CTAs per image" in _lib.last_error()
    plan.arr[1].workspace = plan.arr[0].workspace
    two = (ctypes.c_void_p * 2)(*[x.cuda_stream for x in bt.lanes(pri.device, 2)])
    assert L.jabd_detect_batches(*args, 0, ctypes.cast(two, ctypes.c_void_p), 2, ctypes.c_void_p(cur)) == -1
    assert "share a workspace" in _lib.last_error()
    assert L.jabd_detect_batches(*args, 0, None, 0, ctypes.c_void_p(cur)) == -1 and "share a workspace" in _lib.last_error()
    # more batches than one launch's table holds (16): 19 batches of one or two images in two launches
    many = [(loc0[i % 3:i % 3 + 1 + i % 2], conf0[i % 3:i % 3 + 1 + i % 2], lm0[i % 3:i % 3 + 1 + i % 2]) for i in range(19)]
    got = bt.detect_batches(pri, many, variances=VAR, lanes_n=0)
    torch.cuda.synchronize()
    for i, o in enumerate(got):
        for t, w in zip(o, want[0]):
            assert torch.equal(t, w[i % 3:i % 3 + 1 + i % 2]), i


@pytest.mark.parametrize("gen", ["A", "B"])
def test_detect_cfg3_batch_vs_oracle(mods, gen):
    """BASELINE configs[2] at full size: 1024x1024 (43,008 priors), > 0.02, top-5000, IoU 0.4, keep 750."""
    orc, synth = mods["orc"], mods["synth"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(1024, 1024)).get_anchors()
    P = pri.shape[0]
    assert P == 43008
    B = 4
    locs, confs, lms = [], [], []
    for i in range(B):
        if gen == "A":
            l, c, m = synth.make_preds_random(3, i, P)
        else:
            gt = synth.make_gt(3, i, (1024, 1024), count=120)
            l, c, m = synth.make_preds_clustered(3, i, pri.cpu(), gt, VAR)
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, landm = torch.stack(locs), torch.stack(confs), torch.stack(lms)
    dets, counts, kidx = mods["batched"].detect(loc.cuda(), conf.cuda(), landm.cuda(), pri, VAR)
    boxes = mods["ub"].decode(loc.cuda(), pri, VAR).cpu().numpy()
    pn = pri.cpu().numpy()
    for i in range(B):
        e_d, e_i = orc.detect(loc[i].numpy(), conf[i].numpy(), landm[i].numpy(), pn, VAR, 0.02, True, 5000, 0.4, 750,
                              boxes_override=boxes[i])
        c = int(counts[i].item())
        assert c == len(e_i), (gen, i)
        assert np.array_equal(kidx[i, :c].cpu().numpy(), e_i), (gen, i)
        assert np.array_equal(dets[i, :c].cpu().numpy(), e_d), (gen, i)
    # (ii) the un-overridden fused pipeline against the oracle's OWN decode (margin check): a <= 2 ulp expf
    # difference may flip an IoU > thr decision, so the keep lists must agree on every image whose oracle keep list is stable
    # under +-2 ulp perturbations of the oracle's boxes; flips on the remaining images are counted and reported, not hidden.
    stable, flips = _margin_check(orc, loc, conf, landm, pn, kidx, counts, dets, (0.02, True, 5000, 0.4, 750))
    print("cfg3 gen %s: %d of %d images margin-stable, %d keep lists differ from the oracle's own decode (%d of them on "
          "margin-stable images)" % (gen, len(stable), B, len(flips), len(set(flips) & set(stable))))
    assert len(stable) >= 1 and not (set(flips) & set(stable))
    # shard invariance and the host-buffer entry
    d1, c1, k1 = mods["batched"].detect(loc[2:3].cuda(), conf[2:3].cuda(), landm[2:3].cuda(), pri, VAR)
    assert torch.equal(d1[0], dets[2]) and torch.equal(k1[0], kidx[2]) and int(c1[0]) == int(counts[2])
    h = mods["batched"].HostDetect(pri, B)
    hd, hc, hk = h(loc.contiguous(), conf.contiguous(), landm.contiguous())                  # pageable landmarks: copied
    assert torch.equal(hd, dets.cpu()) and torch.equal(hc, counts.cpu()) and torch.equal(hk, kidx.cpu())
    full_h2d = h.last_h2d
    hd, hc, hk = h(loc.pin_memory(), conf.pin_memory(), landm.contiguous().pin_memory())     # pinned: kept rows read in place
    assert torch.equal(hd, dets.cpu()) and torch.equal(hc, counts.cpu()) and torch.equal(hk, kidx.cpu())
    assert h.last_h2d < 0.45 * full_h2d
    # no landmarks, >= threshold, uncapped top-k and keep
    d2, c2, k2 = mods["batched"].detect(loc[:1].cuda(), conf[:1].cuda(), None, pri, VAR, conf_thres=0.5, strict=False,
                                        pre_nms_topk=0, nms_thres=0.3, keep_topk=0)
    e_d, e_i = orc.detect(loc[0].numpy(), conf[0].numpy(), landm[0].numpy(), pn, VAR, 0.5, False, 0, 0.3, 0, boxes_override=boxes[0])
    c = int(c2[0])
    assert c == len(e_i) and np.array_equal(k2[0, :c].cpu().numpy(), e_i)
    assert np.array_equal(d2[0, :c, :5].cpu().numpy(), e_d[:, :5]) and (d2[0, :c, 5:] == 0).all()


def test_correct_boxes_matches_reference_numpy(mods):
    """letterbox undo + pixel scaling on kept rows == the reference's numpy code
    (retinaface_correct_boxes R/utils/utils_bbox.py:9-24 via the host-side drop-in restatement, then
    R/predict.py:194-195), bit for bit, for ragged counts and non-square images."""
    from jabd_b200 import batched, utils_bbox
    g = torch.Generator().manual_seed(11)
    B, K = 5, 37
    dets = torch.rand((B, K, 15), generator=g, dtype=torch.float32)
    counts = torch.tensor([37, 0, 5, 20, 1], dtype=torch.int32)
    input_shape = (640, 640)
    shapes = [(480, 640), (1080, 1920), (333, 517), (640, 640), (1200, 800)]
    post = batched.letterbox_params(input_shape, shapes)
    out = batched.correct_boxes(dets.clone().cuda(), counts.cuda(), post).cpu().numpy()
    for b in range(B):
        c = int(counts[b])
        ref = dets[b, :c].numpy().copy()
        if c:
            ref = utils_bbox.retinaface_correct_boxes(ref, np.array(input_shape), np.array(shapes[b]))
            h, w = shapes[b]
            ref[:, :4] = ref[:, :4] * [w, h, w, h]
            ref[:, 5:] = ref[:, 5:] * ([w, h] * 5)
        assert np.array_equal(out[b, :c], ref)
        assert np.array_equal(out[b, c:], dets[b, c:].numpy())       # rows past the count are untouched
    only_px = batched.correct_boxes(dets.clone().cuda(), None, post, letterbox=False).cpu().numpy()
    ref = dets.numpy().copy()
    for b in range(B):
        h, w = shapes[b]
        ref[b, :, :4] = ref[b, :, :4] * [w, h, w, h]
        ref[b, :, 5:] = ref[b, :, 5:] * ([w, h] * 5)
    assert np.array_equal(only_px, ref)


def _adversarial_boxes(rng, n, nonfinite):
    """Mixed-size boxes over (and beyond) the unit square with the inputs the kept index must route around:
    huge boxes, zero-area and negative-side boxes, exact duplicates, far out-of-range and (optionally) non-finite
    coordinates."""
    c = rng.random((n, 2), dtype=np.float32) * 1.2 - 0.1
    scale = np.exp(rng.uniform(np.log(0.003), np.log(0.6), size=(n, 1))).astype(np.float32)
    wh = scale * (0.6 + 0.8 * rng.random((n, 2), dtype=np.float32))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    idx = rng.permutation(n)
    b[idx[:20], 2:] = b[idx[:20], :2]                              # zero area
    b[idx[20:30], 2] = b[idx[20:30], 0] - 0.01                      # negative width
    b[idx[30:60]] = b[idx[60:90]]                                   # exact duplicates
    b[idx[90:95]] *= 1000.0                                         # far out of the grid's coordinate bound
    b[idx[101:140]] = np.array([0.0, 0.0, 1.0, 1.0], np.float32) + 0.01 * rng.standard_normal((39, 4)).astype(np.float32)
    # pairs whose IoU is exactly 0.5 / 0.25 / within a few ulp of 0.3: the sliver where the quotient is evaluated
    b[idx[140:150]] = [0.25, 0.25, 0.5, 0.5]
    b[idx[150:160]] = [0.25, 0.25, 0.5, 0.375]
    b[idx[160:170]] = [0.25, 0.25, 0.3125, 0.5]
    b[idx[170:180]] = np.float32([0.6, 0.6, 0.7, 0.7])
    b[idx[180:190]] = np.float32([0.6, 0.6, 0.7, 0.63])
    b[idx[180:190], 3] += (np.arange(10, dtype=np.float32) - 5) * np.float32(6e-8)
    if nonfinite:
        b[idx[95:97], 0] = np.nan
        b[idx[97:99], 3] = np.inf
        b[idx[99:101]] = [-np.inf, 0.1, 0.5, 0.5]
    s = rng.random(n, dtype=np.float32)
    s[::5] = s[2]
    return b, s


@pytest.mark.parametrize("n", [700, 5000, 9000])
def test_nms_division_free_decision_equals_exact_and_oracle(mods, n):
    """suppresses() (detect.cu) decides most pairs by comparing inter with thr*union instead of dividing; that never
    changes a decision: default == JABD_NMS_EXACT_DIV == oracle for torchvision and SSD semantics, thresholds 0 / 0.3 /
    0.5 and a negative one, on mixed-size boxes with zero-area, negative-side, duplicate, far-away boxes and pairs whose
    IoU sits exactly on / within ulps of the threshold.  With non-finite coordinates (which the reference never
    produces) the two paths must still agree with each other."""
    ops, orc = mods["ops"], mods["orc"]
    dev = torch.device("cuda", 0)
    DENSE = 256   # JABD_NMS_EXACT_DIV
    for nonfinite in (False, True):
        b, s = _adversarial_boxes(np.random.default_rng(100 + n), n, nonfinite)
        cb, cs = cuda(b), cuda(s)
        for thr in (0.0, 0.3, 0.5, -0.25):
            out = []
            for mode in (ops.NMS_TV, ops.NMS_TV | DENSE):
                keep, cnt = ops.nms_indices(cb, 4, cs, 1, n, 0.0, ops.THRESH_NONE, 0, thr, mode, n, dev)
                out.append(keep[:int(cnt.item())].cpu().numpy())
            assert np.array_equal(out[0], out[1]), (thr, nonfinite)
            if not nonfinite:
                assert np.array_equal(out[0], orc.nms_tv(b, s, thr)), thr
        for (ov, tk) in ((0.5, 200), (0.3, n), (0.0, n)):
            out = []
            for mode in (ops.NMS_SSD, ops.NMS_SSD | DENSE):
                keep, cnt = ops.nms_indices(cb, 4, cs, 1, n, 0.0, ops.THRESH_NONE, tk, ov, mode, min(n, tk), dev)
                out.append(keep[:int(cnt.item())].cpu().numpy())
            assert np.array_equal(out[0], out[1]), (ov, tk, nonfinite)
            if not nonfinite:
                rk, rc = orc.nms_ssd(b, s, ov, tk)
                assert np.array_equal(out[0], rk[:rc]), (ov, tk)


def test_detect_cluster_widths_agree(mods):
    """One image on a thread-block cluster of 1..8 CTAs (split decode + kept-list slices, masks exchanged through
    distributed shared memory): keep lists, counts and rows are identical for every width and equal the oracle's; covers
    several images per launch, a ragged last chunk, an empty image and the landmark-less output stage."""
    orc, synth = mods["orc"], mods["synth"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    P = pri.shape[0]
    B = 5
    locs, confs, lms = [], [], []
    for i in range(B):
        gt = synth.make_gt(2, i, (640, 640), count=40 + 30 * i)
        l, c, m = synth.make_preds_clustered(2, i, pri.cpu(), gt, VAR) if i % 2 == 0 else synth.make_preds_random(2, i, P)
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, landm = torch.stack(locs).cuda(), torch.stack(confs).cuda(), torch.stack(lms).cuda()
    conf[3, :, 1] = 0.0                                          # nothing above the threshold in image 3
    conf[3, :, 0] = 1.0
    with pytest.raises(ValueError):
        mods["batched"].detect(loc, conf, landm, pri, VAR, cluster=9)
    ref = None
    for width in (1, 2, 3, 4, 5, 6, 7, 8, 0):
        out = [mods["batched"].detect(loc, conf, landm, pri, VAR, cluster=width),                    # cfg3 parameters
               mods["batched"].detect(loc, conf, None, pri, VAR, conf_thres=0.3, strict=False, pre_nms_topk=1234,
                                      nms_thres=0.3, keep_topk=97, cluster=width)]
        torch.cuda.synchronize()
        if ref is None:
            ref = out
            boxes = mods["ub"].decode(loc, pri, VAR).cpu().numpy()
            pn = pri.cpu().numpy()
            for i in range(B):
                e_d, e_i = orc.detect(loc[i].cpu().numpy(), conf[i].cpu().numpy(), landm[i].cpu().numpy(), pn, VAR, 0.02, True,
                                      5000, 0.4, 750, boxes_override=boxes[i])
                c = int(out[0][1][i])
                assert c == len(e_i) and np.array_equal(out[0][2][i, :c].cpu().numpy(), e_i)
                assert np.array_equal(out[0][0][i, :c].cpu().numpy(), e_d)
            assert int(out[0][1][3]) == 0
        else:
            for (d, c, k), (rd, rc_, rk) in zip(out, ref):
                assert torch.equal(c, rc_) and torch.equal(k, rk) and torch.equal(d, rd), width


@pytest.mark.parametrize("width", [2, 5, 8])
def test_nms_cluster_multi_round_and_workspace_spill(mods, width):
    """jabd_nms on a cluster: more candidates than one selection round (several decode exchanges) and more keeps than the
    shared-memory kept cache (slices read back from the workspace copy each CTA writes itself); SSD-legacy order too."""
    ops, orc = mods["ops"], mods["orc"]
    rng = np.random.default_rng(23)
    n = 15000
    c = rng.random((n, 2), dtype=np.float32)
    wh = 0.004 + 0.02 * rng.random((n, 2), dtype=np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(n, dtype=np.float32)
    s[::5] = s[1]
    ref = orc.nms_tv(b, s, 0.4)
    assert len(ref) > 1536
    keep, cnt, st = ops.nms_indices(cuda(b), 4, cuda(s), 1, n, 0.0, ops.THRESH_NONE, 0, 0.4, ops.NMS_TV, n, torch.device("cuda", 0),
                                    cluster=width, return_stats=True)
    assert int(st[0]) == 3 and int(st[2]) == n                   # three selection rounds consumed every candidate
    c_ = int(cnt.item())
    assert c_ == len(ref) and np.array_equal(keep[:c_].cpu().numpy(), ref)
    ref_s, cnt_s = orc.nms_ssd(b[:3000], s[:3000], 0.45, 200)
    keep, count = mods["box_utils"].nms(cuda(b[:3000]), cuda(s[:3000]), 0.45, 200)
    assert count == cnt_s and np.array_equal(keep.cpu().numpy(), ref_s)


# ---------------------------------------------------------------------- exact three-pass select on a cluster (tie blocks)
@pytest.mark.parametrize("width", [1, 2, 3, 4, 5, 6, 7, 8, 0])
def test_nms_tie_block_exact_select_on_cluster(mods, width):
    """More than 8192 equal scores around the selection cut: the fine first histogram cannot isolate a small cut bin, so the
    exact three-pass radix select + ordered compaction run (first histogram shared by the cluster, passes 2-3 and the
    compaction replicated) -- asserted through the call's statistics -- for torchvision order (ties: lower index first) and
    SSD order (higher index first), on every cluster width.  Keep lists equal the oracle's."""
    ops, orc = mods["ops"], mods["orc"]
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(41)
    n = 20000
    c = rng.random((n, 2), dtype=np.float32)
    wh = np.exp(rng.uniform(np.log(0.004), np.log(0.05), (n, 2))).astype(np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(n, dtype=np.float32)
    s[rng.permutation(n)[:9000]] = np.float32(0.5)              # ~5500 scores above the block: the cut at 6144 falls inside it
    cb, cs = cuda(b), cuda(s)
    ref = orc.nms_tv(b, s, 0.3)
    for cap in (n, 750):
        keep, cnt, st = ops.nms_indices(cb, 4, cs, 1, n, 0.0, ops.THRESH_NONE, 0, 0.3, ops.NMS_TV, cap, dev, cluster=width,
                                        return_stats=True)
        c_ = int(cnt.item())
        assert c_ == min(len(ref), cap) and np.array_equal(keep[:c_].cpu().numpy(), ref[:cap]), (width, cap)
        if cap == n:
            assert int(st[1]) >= 2 and int(st[0]) == 4 and int(st[2]) == n, st.tolist()   # rounds 1 and 2 cut inside the block
    # pre-NMS top-k that ends inside the tie block
    order = orc.argsort_desc(s)[:7000]
    ref_k = order[orc.nms_tv(b[order], s[order], 0.3)]
    keep, cnt, st = ops.nms_indices(cb, 4, cs, 1, n, 0.0, ops.THRESH_NONE, 7000, 0.3, ops.NMS_TV, n, dev, cluster=width,
                                    return_stats=True)
    c_ = int(cnt.item())
    assert c_ == len(ref_k) and np.array_equal(keep[:c_].cpu().numpy(), ref_k) and int(st[1]) >= 1, width
    # SSD-legacy order: ties are taken from the high-index end
    rk, rc = orc.nms_ssd(b, s, 0.45, 7000)
    keep, cnt, st = ops.nms_indices(cb, 4, cs, 1, n, 0.0, ops.THRESH_NONE, 7000, 0.45, ops.NMS_SSD, 7000, dev, cluster=width,
                                    return_stats=True)
    assert int(cnt.item()) == rc and np.array_equal(keep[:rc].cpu().numpy(), rk[:rc]) and int(st[1]) >= 1, width


@pytest.mark.parametrize("width", [1, 2, 3, 4, 6, 0])
def test_detect_saturated_scores_on_cluster(mods, width):
    """An untrained network that scores every prior the same (and one that saturates half of them at 1.0): the fused kernel
    reaches the exact select at its default launch shape; candidates are then the lowest prior indices, like a stable sort."""
    orc, synth = mods["orc"], mods["synth"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    P = pri.shape[0]
    locs, confs, lms = [], [], []
    for i in range(3):
        l, c, m = synth.make_preds_random(4, i, P)
        if i == 0:
            c[:, 1] = 0.5
        elif i == 1:
            sat = torch.rand(P, generator=torch.Generator().manual_seed(5)) < 0.6
            c[sat, 1] = 1.0
        c[:, 0] = 1.0 - c[:, 1]
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, landm = torch.stack(locs), torch.stack(confs), torch.stack(lms)
    dets, counts, kidx, st = mods["batched"].detect(loc.cuda(), conf.cuda(), landm.cuda(), pri, VAR, cluster=width, return_stats=True)
    st = st.cpu().numpy()
    assert st[0, 1] >= 1 and st[1, 1] >= 1 and st[2, 1] == 0, st.tolist()
    boxes = mods["ub"].decode(loc.cuda(), pri, VAR).cpu().numpy()
    pn = pri.cpu().numpy()
    for i in range(3):
        e_d, e_i = orc.detect(loc[i].numpy(), conf[i].numpy(), landm[i].numpy(), pn, VAR, 0.02, True, 5000, 0.4, 750,
                              boxes_override=boxes[i])
        c = int(counts[i].item())
        assert c == len(e_i) and np.array_equal(kidx[i, :c].cpu().numpy(), e_i), (width, i)
        assert np.array_equal(dets[i, :c].cpu().numpy(), e_d), (width, i)


# ------------------------------------------------------------------------------------------------ cfg4, inference side
def test_detect_cfg4_vs_oracle(mods):
    """BASELINE configs[3] on the inference side: 2048x2048 (172,032 priors), 1,500 faces per image, clustered predictions.
    cfg3 parameters and the reference's live ones (>= 0.5, IoU 0.3, nothing capped: R/predict.py:40,181 ->
    R/utils/utils_bbox.py:260-279 -- ~50 k candidates, several selection rounds, > 1536 keeps spilling to the workspace);
    keep lists and rows bit-exact against the oracle on the GPU-decoded boxes."""
    orc, synth = mods["orc"], mods["synth"]
    size = (2048, 2048)
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=size).get_anchors()
    P = pri.shape[0]
    assert P == 172032
    B = 2
    locs, confs, lms = [], [], []
    for i in range(B):
        gt = synth.make_gt(4, i, size)
        l, c, m = synth.make_preds_clustered(4, i, pri, gt, VAR, device="cuda")
        locs.append(l); confs.append(c); lms.append(m)
    loc, conf, landm = torch.stack(locs), torch.stack(confs), torch.stack(lms)
    boxes = mods["ub"].decode(loc.cuda(), pri, VAR).cpu().numpy()
    pn = pri.cpu().numpy()
    for params in ((0.02, True, 5000, 0.4, 750), (0.5, False, 0, 0.3, 0)):
        ct, strict, topk, nt, keepk = params
        dets, counts, kidx, st = mods["batched"].detect(loc.cuda(), conf.cuda(), landm.cuda(), pri, VAR, conf_thres=ct, strict=strict,
                                                        pre_nms_topk=topk, nms_thres=nt, keep_topk=keepk, return_stats=True)
        st = st.cpu().numpy()
        for i in range(B):
            e_d, e_i = orc.detect(loc[i].numpy(), conf[i].numpy(), landm[i].numpy(), pn, VAR, ct, strict, topk, nt, keepk,
                                  boxes_override=boxes[i])
            c = int(counts[i].item())
            assert c == len(e_i), (params, i, c, len(e_i))
            assert np.array_equal(kidx[i, :c].cpu().numpy(), e_i), (params, i)
            assert np.array_equal(dets[i, :c].cpu().numpy(), e_d), (params, i)
            assert (kidx[i, c:] == -1).all()
        if topk == 0:
            assert (st[:, 0] >= 4).all() and (st[:, 2] > 20000).all(), st.tolist()     # multi-round, tens of thousands of candidates
            assert int(counts.min()) > 750
    # one image alone (cluster of 4 or 8 is co-resident) equals its row in the batch
    d1, c1, k1 = mods["batched"].detect(loc[1:2].cuda(), conf[1:2].cuda(), landm[1:2].cuda(), pri, VAR)
    d2, c2, k2 = mods["batched"].detect(loc.cuda(), conf.cuda(), landm.cuda(), pri, VAR, cluster=1)
    assert torch.equal(d1[0], d2[1]) and torch.equal(k1[0], k2[1]) and int(c1[0]) == int(c2[1])


def test_dlpack_inputs(mods):
    """north_star: the box utilities take numpy or framework tensors via DLPack.  An object that exports ONLY
    __dlpack__ / __dlpack_device__ (no torch or numpy type) goes through the same kernels: CUDA-resident in place, host
    memory through an explicit staging copy."""
    from jabd_b200 import retinaface_training as rt
    ub = mods["ub"]

    class Capsule(object):
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, **kw):
            return self._t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(160, 160)).get_anchors()
    loc, conf, landm = mods["synth"].make_preds_random(7, 0, pri.shape[0])
    ref = ub.decode(loc.cuda(), pri, VAR)
    for wrap in (Capsule(loc.cuda()), Capsule(loc.clone())):                 # device-resident and host-resident exporters
        out = ub.decode(wrap, Capsule(pri), VAR)
        assert torch.equal(torch.as_tensor(out).cuda(), ref)
    gt = mods["synth"].make_gt(7, 1, (160, 160), count=9)
    a = rt.jaccard(Capsule(gt[:, :4].contiguous().cuda()), Capsule(rt.point_form(pri)))
    assert torch.equal(torch.as_tensor(a).cuda(), rt.jaccard(gt[:, :4].cuda(), rt.point_form(pri)))
    t1 = mods["batched"].assign_targets(Capsule(pri), [Capsule(gt.cuda())])
    t2 = mods["batched"].assign_targets(pri, [gt.cuda()])
    assert all(torch.equal(x, y) for x, y in zip(t1, t2))
    d1 = mods["batched"].detect(Capsule(loc.cuda()[None]), Capsule(conf.cuda()[None]), Capsule(landm.cuda()[None]), Capsule(pri), VAR)
    d2 = mods["batched"].detect(loc.cuda()[None], conf.cuda()[None], landm.cuda()[None], pri, VAR)
    assert all(torch.equal(x, y) for x, y in zip(d1, d2))
