#!/usr/bin/env python
"""Randomised GPU-vs-oracle comparison over many seeds.  tests/test_gpu_fuzz.py runs a reduced slice of every job under
``pytest -m gpu``; the full sweep is run by hand on a GPU box:

    python tests/fuzz_gpu.py [n_trials [assign,nms/topk,tieblock,detect,loss,eval]]

Target assignment (ragged batches, duplicate / touching / tiny / out-of-image GT, both label and encode modes, culled and
dense), top-k with ties, NMS (torchvision / SSD / DIoU semantics, capped and uncapped), the fused detect pipeline, the MultiBox
loss selection mask and the WIDER AP counters -- every comparison at the bar of the parity tests (indices, labels, keep lists
and counters bit-exact; coordinates rtol 1e-5 / atol 1e-6)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _ops, anchors, batched, config, synth, utils_bbox, utils_map  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
from oracle import wider_eval as ow  # noqa: E402

RTOL, ATOL = 1e-5, 1e-6
VAR = [0.1, 0.2]
dev = torch.device("cuda", 0)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def rand_gt(rng, size, g):
    """[g,15] rows with the oddities the matching code must survive."""
    t = synth.make_gt(2, int(rng.integers(0, 10 ** 6)), size, count=g).numpy().copy()
    k = rng.integers(0, 6)
    if g >= 4 and k == 1:
        t[1] = t[0]                                       # duplicate GT (tie in the row argmax -> last j wins the force-match)
    if g >= 4 and k == 2:
        t[2, :4] = [t[0, 2], t[0, 1], t[0, 2] + 0.05, t[0, 3]]   # touching boxes (zero intersection)
    if g >= 4 and k == 3:
        t[3, 2:4] = t[3, :2] + np.float32(1e-4)           # tiny face
    if g >= 4 and k == 4:
        t[0, :4] = [-0.2, -0.1, 0.1, 0.15]                # partly outside the image
    if g >= 4 and k == 5:
        t[1, :4] = t[0, :4] + np.float32(1e-7)            # near-duplicate
    return t


def fuzz_assign(rng, trial):
    size = [(96, 128), (160, 160), (320, 256), (640, 640)][trial % 4]
    cfg = [config.cfg_mnet, config.cfg_re50][trial % 2]
    pri = anchors.cached_priors(cfg, size, dev)
    pn = pri.cpu().numpy()
    B = int(rng.integers(1, 6))
    tg = [rand_gt(rng, size, int(rng.integers(1, 40) if trial % 3 else rng.integers(1, 301))) for _ in range(B)]
    thr = float(rng.choice([0.35, 0.2, 0.5]))
    lm, em = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    ref = orc.match_batch(thr, tg, pn, VAR, lm, em)
    # a random work-list shape (JABD_ASSIGN_TUNE) on every second trial: results must not depend on it
    tune = None if trial % 2 else (int(rng.integers(16, 193)), int(rng.integers(16, 193)), int(rng.integers(0, 101)))
    for dense in (False, True):
        loc_t, conf_t, landm_t, ex = batched.assign_targets(pri, [cuda(t) for t in tg], threshold=thr, variances=VAR, label_mode=lm,
                                                            encode=bool(em), return_match=True, dense=dense, tune=tune)
        assert np.array_equal(conf_t.cpu().numpy(), ref["conf_t"]), ("conf_t", trial, dense)
        assert np.array_equal(ex["best_truth_idx"].cpu().numpy(), ref["best_truth_idx"]), ("bti", trial, dense)
        assert np.array_equal(ex["best_truth_overlap"].cpu().numpy(), ref["best_truth_overlap"]), ("bto", trial, dense)
        assert np.array_equal(ex["best_prior_idx"].cpu().numpy(), ref["best_prior_idx"]), ("bpi", trial, dense)
        assert np.array_equal(landm_t.cpu().numpy(), ref["landm_t"]), ("landm_t", trial, dense)
        lt = loc_t.cpu().numpy()
        fin = np.isfinite(ref["loc_t"])
        assert np.array_equal(np.isfinite(lt), fin), ("loc finite", trial, dense)
        np.testing.assert_allclose(lt[fin], ref["loc_t"][fin], rtol=RTOL, atol=ATOL)
    if trial % 3 == 0 and B > 1:
        # the same images as single-image batches on lanes (jabd_assign_batches): the bytes of the one call above
        outs = batched.assign_batches(pri, [[cuda(t)] for t in tg], threshold=thr, variances=VAR, label_mode=lm, encode=bool(em),
                                      lanes_n=int(rng.integers(0, 4)))
        torch.cuda.synchronize()
        for i, (a, b_, c_) in enumerate(outs):
            assert torch.equal(a[0], loc_t[i]) and torch.equal(b_[0], conf_t[i]) and torch.equal(c_[0], landm_t[i]), ("lanes", trial, i)


def fuzz_nms(rng, trial):
    n = int(rng.integers(1, 9000))
    c = rng.random((n, 2), dtype=np.float32)
    wh = np.exp(rng.uniform(np.log(0.004), np.log(0.3), (n, 2))).astype(np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    if n > 50:
        b[rng.integers(0, n, 20)] = b[rng.integers(0, n, 20)]          # duplicates
    s = rng.random(n, dtype=np.float32)
    if trial % 2:
        s = np.round(s * 64) / np.float32(64)                          # heavy score ties
    thr = float(rng.choice([0.3, 0.4, 0.5, 0.0]))
    cb, cs = cuda(b), cuda(s)
    ref = orc.nms_tv(b, s, thr)
    cap = int(rng.choice([n, 750, 1]))
    keep, cnt = _ops.nms_indices(cb, 4, cs, 1, n, 0.0, _ops.THRESH_NONE, 0, thr, _ops.NMS_TV, cap, dev)
    c_ = int(cnt.item())
    assert c_ == min(len(ref), cap) and np.array_equal(keep[:c_].cpu().numpy(), ref[:cap]), ("tv", trial, n, thr, cap)
    tk = int(rng.choice([200, n]))
    rk, rc = orc.nms_ssd(b, s, thr, tk)
    keep, cnt = _ops.nms_indices(cb, 4, cs, 1, n, 0.0, _ops.THRESH_NONE, tk, thr, _ops.NMS_SSD, min(n, tk), dev)
    assert int(cnt.item()) == rc and np.array_equal(keep[:rc].cpu().numpy(), rk[:rc]), ("ssd", trial, n, thr, tk)
    dk, dc = orc.diounms(b, s, thr, tk, 1.0)
    k2, c2 = utils_bbox.diounms(cb, cs, thr, tk, 1.0)
    assert c2 == dc and np.array_equal(k2[:dc].cpu().numpy(), dk[:dc]), ("diou", trial, n, thr, tk)
    # top-k with ties
    k = int(rng.integers(1, n + 1))
    idx, cnt = _ops.topk(cs, k)
    order = np.argsort(-s.astype(np.float64), kind="stable")[:k]
    assert int(cnt) == len(order) and np.array_equal(idx.cpu().numpy()[:len(order)], order), ("topk", trial, n, k)


def fuzz_nms_tie_block(rng, trial):
    """More than 8192 equal scores around the selection cut: the fine first histogram cannot isolate a small cut bin, so the
    exact three-pass radix select + ordered compaction run -- on every cluster width (replicated there, after a shared
    first pass)."""
    n = int(rng.integers(12000, 30000))
    c = rng.random((n, 2), dtype=np.float32)
    wh = np.exp(rng.uniform(np.log(0.004), np.log(0.05), (n, 2))).astype(np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(n, dtype=np.float32)
    tie = rng.permutation(n)[: int(rng.integers(8500, min(n - 2000, 16000)))]
    s[tie] = np.float32(rng.uniform(0.2, 0.8))
    thr = float(rng.choice([0.3, 0.5]))
    ref = orc.nms_tv(b, s, thr)
    cb, cs = cuda(b), cuda(s)
    exact = 0
    for width in (1, 2, 3, 4, 5, 6, 7, 8, 0):
        for cap in (n, 750):
            keep, cnt, st = _ops.nms_indices(cb, 4, cs, 1, n, 0.0, _ops.THRESH_NONE, 0, thr, _ops.NMS_TV, cap, dev, cluster=width,
                                             return_stats=True)
            c_ = int(cnt.item())
            assert c_ == min(len(ref), cap) and np.array_equal(keep[:c_].cpu().numpy(), ref[:cap]), ("tie block", trial, n, width, cap)
            exact += int(st[1]) if cap == n else 0
    assert exact >= 5, ("tie block never reached the exact select", trial, n, exact)   # every width ran it at least once


def fuzz_detect(rng, trial):
    size = [(160, 160), (320, 320), (640, 640)][trial % 3]
    pri = anchors.cached_priors(config.cfg_mnet, size, dev)
    pn = pri.cpu().numpy()
    seed = int(rng.integers(0, 10 ** 6))
    gt = synth.make_gt(3, seed, size, count=int(rng.integers(1, 80)))
    l, c, m = (synth.make_preds_clustered(3, seed, pri.cpu(), gt, VAR) if trial % 2 else synth.make_preds_random(3, seed, pri.shape[0]))
    ct, strict, topk, nt, keep = [(0.02, True, 5000, 0.4, 750), (0.5, False, 0, 0.3, 0), (0.3, True, 300, 0.5, 50)][trial % 3]
    dets, counts, kidx = batched.detect(l.to(dev)[None], c.to(dev)[None], m.to(dev)[None], pri, VAR, conf_thres=ct, strict=strict,
                                        pre_nms_topk=topk, nms_thres=nt, keep_topk=keep)
    boxes = utils_bbox.decode(l.to(dev), pri, VAR).cpu().numpy()
    e_d, e_i = orc.detect(l.numpy(), c.numpy(), m.numpy(), pn, VAR, ct, strict, topk, nt, keep, boxes_override=boxes)
    n = int(counts[0])
    assert n == len(e_i) and np.array_equal(kidx[0, :n].cpu().numpy(), e_i), ("detect idx", trial)
    assert np.array_equal(dets[0, :n].cpu().numpy(), e_d), ("detect rows", trial)
    if trial % 3 == 0:
        # the image three times as separate batches of one call (jabd_detect_batches: one launch, or lanes): the same rows
        one = (l.to(dev)[None], c.to(dev)[None], m.to(dev)[None])
        outs = batched.detect_batches(pri, [one, one, one], variances=VAR, conf_thres=ct, strict=strict, pre_nms_topk=topk, nms_thres=nt,
                                      keep_topk=keep, lanes_n=int(rng.integers(0, 3)), cluster=int(rng.integers(0, 5)))
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o[1], counts) and torch.equal(o[2], kidx) and torch.equal(o[0], dets), ("detect batches", trial)


def fuzz_loss(rng, trial):
    size = [(96, 128), (160, 160), (320, 256)][trial % 3]
    pri = anchors.cached_priors(config.cfg_mnet, size, dev)
    P = pri.shape[0]
    B = int(rng.integers(1, 5))
    seed = int(rng.integers(0, 10 ** 6))
    tg = [synth.make_gt(6, seed + i, size, count=int(rng.integers(1, 30))) for i in range(B)]
    raw = [synth.make_logits(6, seed + i, P) for i in range(B)]
    preds_c = tuple(torch.stack([r[k] for r in raw]) for k in range(3))
    if trial % 4 == 0:
        preds_c[1][:, :, :] = torch.round(preds_c[1] * 2) / 2                 # heavy ties in the rank values
    kind = [None, "diou", None, "ciou"][trial % 4]
    cpu = tuple(t.clone().requires_grad_(True) for t in preds_c)
    l0, c0, m0, aux = tp.multibox_loss(cpu, pri.cpu(), tg, 0.35, VAR, 7, return_aux=True, loc_loss=kind)
    loc_t, conf_t, landm_t = batched.assign_targets(pri, [t.to(dev) for t in tg], threshold=0.35, variances=VAR, encode=kind is None)
    gpu = tuple(t.to(dev) for t in preds_c)
    l1, c1, m1, mask, norms = batched.multibox_loss(gpu, loc_t, conf_t, landm_t, 7, return_aux=True,
                                                    loc_loss={None: "smooth_l1", "diou": "Diou", "ciou": "Ciou"}[kind], priors=pri)
    mask = mask.cpu().numpy()
    assert np.array_equal((mask & 1).astype(bool), aux["pos"].numpy()), ("pos", trial)
    # mined negatives: the same number per image; the sets may differ only where rank values (two exp + one log in fp32, 1-2 ulp
    # apart between torch-CPU and CUDA) tie or nearly tie at the cut -- a couple of swaps per image, none without such ties
    g_neg = ((mask >> 2) & 1).astype(bool) & ~aux["pos"].numpy()
    r_neg = aux["neg"].numpy() & ~aux["pos"].numpy()
    assert np.array_equal(g_neg.sum(1), r_neg.sum(1)), ("neg count", trial)
    if trial % 4 != 0:
        assert int((g_neg != r_neg).sum()) <= 2 * B, ("neg set", trial, int((g_neg != r_neg).sum()))
    np.testing.assert_allclose([l1.item(), c1.item(), m1.item()], [l0.item(), c0.item(), m0.item()], rtol=2e-5, atol=1e-6)


def fuzz_eval(rng, trial):
    n_img = int(rng.integers(1, 30))
    imgs = [synth.make_eval_image(int(rng.integers(7, 1000)), int(rng.integers(0, 10 ** 6)), count=(None if rng.random() < 0.8 else 300))
            for _ in range(n_img)]
    preds = ow.norm_scores([im[2] for im in imgs])
    gts = [im[0] for im in imgs]
    keeps = [im[1][int(rng.integers(0, 3))] for im in imgs]
    thr, tn = float(rng.choice([0.4, 0.5, 0.3])), int(rng.choice([1000, 100, 7]))
    assert np.array_equal(utils_map.pr_counters(preds, gts, keeps, thr, tn), ow.pr_counters(preds, gts, keeps, thr, tn)), ("eval", trial)


JOBS = {"assign": (fuzz_assign, 1.0), "nms/topk": (fuzz_nms, 1.0), "tieblock": (fuzz_nms_tie_block, 1 / 6.0),
        "detect": (fuzz_detect, 0.5), "loss": (fuzz_loss, 0.5), "eval": (fuzz_eval, 1 / 3.0)}


def run_job(name, count, seed=20240611):
    fn, _ = JOBS[name]
    rng = np.random.default_rng(seed + sum(ord(c) for c in name))
    for t in range(count):
        fn(rng, t)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    for name, (fn, share) in JOBS.items():
        if only and name not in only:
            continue
        count = max(int(n * share), 1)
        run_job(name, count)
        print("%-9s %4d trials ok" % (name, count), flush=True)
    print("FUZZ OK")


if __name__ == "__main__":
    main()
