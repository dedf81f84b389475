#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Dev-container only: imports the unmodified reference from /root/reference
(read-only; PYTHONDONTWRITEBYTECODE is forced) on CPU tensors and records its
outputs for seeded inputs.  The fixtures (not the reference) travel to the GPU
box.  Re-run:  python tests/golden/make_golden.py [--check]

--check additionally fuzzes the C oracle and the torch port against the live
reference on extra random cases (nothing is written for those).

Recorded with every file: torch / torchvision / numpy versions (``meta``).
The reference itself has no tests or golden vectors; NMS
is torchvision.ops.nms (third party, unpinned by R/requirements.txt) -- the
version recorded here is the oracle of record.
"""
import contextlib
import hashlib
import io
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("JABD_REF", "/root/reference/JABD2080ti")
sys.path.insert(0, ROOT)

from jabd_b200 import synth  # noqa: E402
from jabd_b200 import config as our_cfg  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):   # utils/anchors.py prints at import
        import utils.anchors as r_anchors
    import utils.config as r_config
    import utils.box_utils as r_box_utils
    import utils.utils_bbox as r_utils_bbox
    import nets.retinaface_training as r_training
    import nets.retinaface_training_DIOU as r_diou
    sys.path.remove(REF)
    return r_anchors, r_config, r_box_utils, r_utils_bbox, r_training, r_diou


def sha(a):
    a = np.ascontiguousarray(a)
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8).copy()


def meta():
    return np.array(json.dumps({
        "torch": torch.__version__, "torchvision": torchvision.__version__, "numpy": np.__version__,
        "reference": REF, "generator": "tests/golden/make_golden.py v1",
    }))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, meta=meta(), **arrays)
    print("wrote %-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


VAR = [0.1, 0.2]
THR = 0.35


# ----------------------------------------------------------------------------- priors
PRIOR_SIZES = [(640, 640), (1024, 1024), (840, 840), (96, 128), (100, 75), (333, 517), (2048, 2048)]
PRIOR_FULL = [(96, 128), (100, 75)]


def gen_priors(R):
    r_anchors, r_config = R[0], R[1]
    out = {}
    names = sorted(our_cfg.ALL_CFGS)
    for name in names:
        rcfg = getattr(r_config, name)
        ocfg = our_cfg.ALL_CFGS[name]
        for k in ("min_sizes", "steps", "variance", "clip"):
            assert list(rcfg[k]) == list(ocfg[k]) if k != "clip" else rcfg[k] == ocfg[k], (name, k)
        for (h, w) in PRIOR_SIZES:
            if (h, w) == (2048, 2048) and name != "cfg_mnet":
                continue
            a = r_anchors.Anchors(rcfg, image_size=(h, w)).get_anchors().numpy()
            key = "%s_%dx%d" % (name, h, w)
            out[key + "_n"] = np.int64(a.shape[0])
            out[key + "_sha"] = sha(a)
            if (h, w) in PRIOR_FULL:
                out[key] = a
            if name == "cfg_mnet" and (h, w) == (100, 75):
                e = r_anchors.Anchors_eval(rcfg, image_size=(h, w)).get_anchors().numpy()
                assert np.array_equal(a, e)
    clip_cfg = dict(r_config.cfg_mnet)
    clip_cfg["clip"] = True
    out["clip_mnet_100x75"] = r_anchors.Anchors(clip_cfg, image_size=(100, 75)).get_anchors().numpy()
    save("priors.npz", **out)


# ----------------------------------------------------------------------------- match
def ref_match_full(R, thr, gt, priors, var):
    """Run the live 10-arg match and recover the intermediate indices with the
    reference's own jaccard/point_form (+ the override lines :127-130)."""
    tr = R[4]
    t = torch.from_numpy(gt)
    P = priors.shape[0]
    p = torch.from_numpy(priors)
    loc_t = torch.zeros(1, P, 4)
    landm_t = torch.zeros(1, P, 10)
    conf_t = torch.zeros(1, P, dtype=torch.long)
    tr.match(thr, t[:, :4], p, var, t[:, -1], t[:, 4:14], loc_t, conf_t, landm_t, 0)
    ov = tr.jaccard(t[:, :4], tr.point_form(p))
    bpo, bpi = ov.max(1)
    bto, bti = ov.max(0)
    bto = bto.clone()
    bti = bti.clone()
    bto.index_fill_(0, bpi, 2)
    for j in range(bpi.size(0)):
        bti[bpi[j]] = j
    return dict(loc_t=loc_t[0].numpy(), conf_t=conf_t[0].numpy(), landm_t=landm_t[0].numpy(),
                bti=bti.numpy(), bto=bto.numpy(), bpi=bpi.numpy(), bpo=bpo.numpy())


def small_match_cases(priors):
    """Hand-built edge cases on a 160x160 image (P=1050)."""
    cases = {}
    cases["rand7"] = synth.make_gt(2, 100, (160, 160), count=7).numpy()
    cases["single"] = synth.make_gt(2, 101, (160, 160), count=1).numpy()
    g = synth.make_gt(2, 102, (160, 160), count=5).numpy()
    dup = np.concatenate([g, g[1:3], g[1:2]], 0)          # duplicate GT rows: ties + shared best prior
    dup[6, 14] = -dup[6, 14]                               # different label on a duplicate
    cases["dups"] = dup
    far = synth.make_gt(2, 103, (160, 160), count=4).numpy()
    far[0, :4] = [2.0, 2.0, 2.1, 2.1]                      # no overlap with any prior: row max -> index 0
    far[2, :4] = [-0.5, -0.5, -0.4, -0.45]
    cases["far"] = far
    cases["crowd40"] = synth.make_gt(2, 104, (160, 160), count=40, side_px=(4, 24)).numpy()
    # a GT identical to a prior box (IoU exactly 1) and tiny faces smaller than any prior
    exact = synth.make_gt(2, 105, (160, 160), count=3, side_px=(2, 4)).numpy()
    pr = priors[517]
    exact[1, :4] = [pr[0] - pr[2] / 2, pr[1] - pr[3] / 2, pr[0] + pr[2] / 2, pr[1] + pr[3] / 2]
    cases["exact"] = exact
    cases["big"] = synth.make_gt(2, 106, (160, 160), count=6, side_px=(60, 150)).numpy()
    return cases


def gen_match(R):
    r_anchors, r_config, r_bu, _, tr, diou = R
    out = {}
    priors = r_anchors.Anchors(r_config.cfg_mnet, image_size=(160, 160)).get_anchors().numpy()
    out["priors_160"] = priors
    names = []
    for name, gt in small_match_cases(priors).items():
        r = ref_match_full(R, THR, gt, priors, VAR)
        out[name + "_gt"] = gt
        for k, v in r.items():
            out[name + "_" + k] = v
        names.append(name)
    # threshold exactly equal to an achieved overlap: `<` must stay strict in fp32
    gt = out["crowd40_gt"]
    bto = out["crowd40_bto"]
    cand = np.unique(bto[(bto > 0.2) & (bto < 0.6)])
    thr_edge = float(cand[len(cand) // 2])
    r = ref_match_full(R, thr_edge, gt, priors, VAR)
    out["edge_thr"] = np.float64(thr_edge)
    out["edge_gt"] = gt
    for k, v in r.items():
        out["edge_" + k] = v
    assert (r["conf_t"] != out["crowd40_conf_t"]).any() or True
    names.append("edge")
    out["names"] = np.array(names)
    # SSD 8-arg variants and match_iou on rand7
    gt = torch.from_numpy(out["rand7_gt"])
    p = torch.from_numpy(priors)
    P = p.shape[0]
    loc_t = torch.zeros(1, P, 4); conf_t = torch.zeros(1, P, dtype=torch.long)
    r_bu.match(THR, gt[:, :4], p, VAR, gt[:, -1], loc_t, conf_t, 0)
    out["ssd_match_loc_t"] = loc_t[0].numpy(); out["ssd_match_conf_t"] = conf_t[0].numpy()
    loc_t = torch.zeros(1, P, 4); conf_t = torch.zeros(1, P, dtype=torch.long)
    r_bu.match_ious(THR, gt[:, :4], p, VAR, gt[:, -1], loc_t, conf_t, 0)
    out["ssd_match_ious_loc_t"] = loc_t[0].numpy(); out["ssd_match_ious_conf_t"] = conf_t[0].numpy()
    loc_t = torch.zeros(1, P, 4); conf_t = torch.zeros(1, P, dtype=torch.long); landm_t = torch.zeros(1, P, 10)
    diou.match_iou(THR, gt[:, :4], p, VAR, gt[:, -1], gt[:, 4:14], loc_t, conf_t, landm_t, 0)
    out["diou_match_iou_loc_t"] = loc_t[0].numpy(); out["diou_match_iou_conf_t"] = conf_t[0].numpy()
    out["diou_match_iou_landm_t"] = landm_t[0].numpy()
    # jaccard / point_form / encode / encode_landm direct
    ov = tr.jaccard(gt[:, :4], tr.point_form(p))
    out["rand7_jaccard"] = ov.numpy()
    out["point_form_160"] = tr.point_form(p).numpy()
    m = gt[:, :4][torch.from_numpy(out["rand7_bti"])]
    out["rand7_encode"] = tr.encode(m, p, VAR).numpy()
    out["rand7_encode_landm"] = tr.encode_landm(gt[:, 4:14][torch.from_numpy(out["rand7_bti"])], p, VAR).numpy()
    save("match_small.npz", **out)

    # full-size cases: cfg1 (640^2, G=50) and a cfg2 image with a large G
    pri640 = r_anchors.Anchors(r_config.cfg_mnet, image_size=(640, 640)).get_anchors().numpy()
    big = {}
    gt1 = synth.make_gt(1, 0, (640, 640)).numpy()
    r = ref_match_full(R, THR, gt1, pri640, VAR)
    big["cfg1_gt"] = gt1
    for k, v in r.items():
        big["cfg1_" + k] = v
    gt2 = synth.make_gt(2, 6, (640, 640)).numpy()   # image 6 of cfg2: G = 211
    r = ref_match_full(R, THR, gt2, pri640, VAR)
    big["cfg2_gt"] = gt2
    for k, v in r.items():
        if k in ("landm_t",):
            big["cfg2_landm_t_sha"] = sha(v)
            big["cfg2_landm_t_head"] = v[:2048]
        else:
            big["cfg2_" + k] = v
    big["priors_640_sha"] = sha(pri640)
    save("match_640.npz", **big)


# ----------------------------------------------------------------------------- decode / nms
def gen_decode(R):
    r_anchors, r_config, r_bu, ub = R[0], R[1], R[2], R[3]
    out = {}
    priors = r_anchors.Anchors(r_config.cfg_mnet, image_size=(160, 160)).get_anchors()
    loc, conf, landm = synth.make_preds_random(3, 0, priors.shape[0])
    out["loc"] = loc.numpy(); out["landm"] = landm.numpy(); out["conf"] = conf.numpy()
    out["boxes"] = ub.decode(loc, priors, VAR).numpy()
    out["landms"] = ub.decode_landm(landm, priors, VAR).numpy()
    assert torch.equal(r_bu.decode(loc, priors, VAR), ub.decode(loc, priors, VAR))
    pri640 = r_anchors.Anchors(r_config.cfg_mnet, image_size=(640, 640)).get_anchors()
    loc, conf, landm = synth.make_preds_random(1, 0, pri640.shape[0])
    out["loc640_sha"] = sha(loc.numpy())
    out["boxes640"] = ub.decode(loc, pri640, VAR).numpy()
    out["landms640_sha"] = sha(ub.decode_landm(landm, pri640, VAR).numpy())
    out["landms640_head"] = ub.decode_landm(landm, pri640, VAR).numpy()[:1024]
    save("decode.npz", **out)


def nms_inputs():
    g = torch.Generator().manual_seed(77)
    cases = {}

    def rnd(n, spread, size):
        c = torch.rand((n, 2), generator=g) * spread
        wh = size * (0.5 + torch.rand((n, 2), generator=g))
        return torch.cat([c - wh / 2, c + wh / 2], 1)
    cases["rand500"] = (rnd(500, 1.0, 0.08), torch.rand(500, generator=g))
    cases["dense2000"] = (rnd(2000, 0.3, 0.05), torch.rand(2000, generator=g))
    b = rnd(600, 0.5, 0.1)
    cases["ties600"] = (b, torch.round(torch.rand(600, generator=g) * 20) / 20)     # many equal scores
    b = rnd(64, 0.5, 0.1)
    b[10] = b[3]; b[20] = b[3]                                                      # identical boxes
    b[30, 2:] = b[30, :2]; b[31] = b[30]                                            # zero-area duplicates (NaN IoU)
    cases["degenerate64"] = (b, torch.rand(64, generator=g))
    cases["one"] = (rnd(1, 1.0, 0.1), torch.rand(1, generator=g))
    cases["neg_scores"] = (rnd(100, 0.4, 0.1), torch.randn(100, generator=g))
    return cases


def gen_nms(R):
    from torchvision.ops import nms as tv_nms
    r_bu, ub = R[2], R[3]
    out = {}
    names = []
    for name, (b, s) in nms_inputs().items():
        out[name + "_boxes"] = b.numpy(); out[name + "_scores"] = s.numpy()
        for thr in (0.3, 0.4, 0.5):
            out["%s_keep_%d" % (name, int(thr * 100))] = tv_nms(b, s, thr).numpy()
        names.append(name)
    out["names"] = np.array(names)
    out["empty_keep"] = tv_nms(torch.zeros(0, 4), torch.zeros(0), 0.3).numpy()
    # SSD-legacy nms / nms_r (unique scores so the unstable sort has one answer)
    for name in ("rand500", "dense2000"):
        b = torch.from_numpy(out[name + "_boxes"]); s = torch.from_numpy(out[name + "_scores"])
        assert len(torch.unique(s)) == len(s)
        for (ov, tk) in ((0.5, 200), (0.3, 50), (0.45, 5000)):
            k1, c1 = r_bu.nms(b, s, ov, tk)
            k2, c2 = ub.nms_r(b, s, ov, tk)
            assert torch.equal(k1, k2) and c1 == c2
            out["%s_ssd_%d_%d_keep" % (name, int(ov * 100), tk)] = k1.numpy()
            out["%s_ssd_%d_%d_count" % (name, int(ov * 100), tk)] = np.int64(c1)
    save("nms.npz", **out)


def gen_pipeline(R):
    """Drop-in non_max_suppression and the composed cfg3 pipeline."""
    from torchvision.ops import nms as tv_nms
    r_anchors, r_config, ub = R[0], R[1], R[3]
    out = {}
    for tag, size, cfg_id, img in (("s160", (160, 160), 3, 1), ("s640", (640, 640), 3, 2)):
        priors = r_anchors.Anchors(r_config.cfg_mnet, image_size=size).get_anchors()
        gt = synth.make_gt(2, 40 + img, size, count=12 if tag == "s160" else 60)
        for gen in ("A", "B"):
            if gen == "A":
                loc, conf, landm = synth.make_preds_random(cfg_id, img, priors.shape[0])
            else:
                loc, conf, landm = synth.make_preds_clustered(cfg_id, img, priors, gt, VAR)
            key = "%s_%s_" % (tag, gen)
            if tag == "s160":
                out[key + "loc"] = loc.numpy(); out[key + "conf"] = conf.numpy(); out[key + "landm"] = landm.numpy()
            else:
                out[key + "in_sha"] = sha(np.concatenate([loc.numpy().ravel(), conf.numpy().ravel(), landm.numpy().ravel()]))
            out[key + "gt"] = gt.numpy()
            boxes = ub.decode(loc, priors, VAR)
            lms = ub.decode_landm(landm, priors, VAR)
            det = torch.cat([boxes, conf[:, 1:2], lms], -1)
            # live drop-in call (R/predict.py:181): conf 0.5, nms 0.3; also a low threshold
            for ct, nt in ((0.5, 0.3), (0.05, 0.3)):
                r = ub.non_max_suppression(det, ct, nt)
                r = np.zeros((0, 15), np.float32) if isinstance(r, list) else r
                out[key + "nms_%d_%d" % (int(ct * 100), int(nt * 100))] = r
            # composed pipeline: > thr, stable desc sort, top-k, tv nms, keep-k
            for (ct, topk, nt, keepk) in ((0.02, 5000, 0.4, 750), (0.02, 200, 0.4, 50)):
                s = conf[:, 1]
                idx = torch.nonzero(s > ct).squeeze(1)
                order = torch.sort(s[idx], stable=True, descending=True)[1][:topk]
                idx = idx[order]
                keep = tv_nms(boxes[idx], s[idx], nt)[:keepk]
                idx = idx[keep]
                k2 = key + "pipe_%d_%d_" % (topk, keepk)
                out[k2 + "idx"] = idx.numpy()
                out[k2 + "dets"] = torch.cat([boxes[idx], s[idx, None], lms[idx]], 1).numpy()
                out[k2 + "n_candidates"] = np.int64((s > ct).sum().item())
    save("pipeline.npz", **out)


# ----------------------------------------------------------------------------- fuzz check
def check(R):
    """Fuzz the C oracle and the torch port against the live reference."""
    from oracle import oracle as orc
    from oracle import torch_port as tp
    r_anchors, r_config, r_bu, ub, tr, diou = R
    from torchvision.ops import nms as tv_nms
    n_bad = 0
    for name in sorted(our_cfg.ALL_CFGS):
        for size in ((96, 128), (333, 517), (640, 640)):
            a = r_anchors.Anchors(getattr(r_config, name), image_size=size).get_anchors().numpy()
            b = orc.priors(our_cfg.ALL_CFGS[name], size)
            if not np.array_equal(a, b):
                print("PRIORS MISMATCH", name, size); n_bad += 1
    stats = {"loc_maxrel": 0.0, "dec_maxabs": 0.0}
    for trial in range(40):
        size = [(160, 160), (320, 256), (640, 640)][trial % 3]
        priors = r_anchors.Anchors(r_config.cfg_mnet, image_size=size).get_anchors()
        gt = synth.make_gt(2, 500 + trial, size, max_count=80)
        ref = ref_match_full(R, THR, gt.numpy(), priors.numpy(), VAR)
        o = orc.match(THR, gt[:, :4].numpy(), priors.numpy(), VAR, gt[:, -1].numpy(), gt[:, 4:14].numpy())
        for k_ref, k_o in (("conf_t", "conf_t"), ("bti", "best_truth_idx"), ("bto", "best_truth_overlap"),
                           ("bpi", "best_prior_idx"), ("landm_t", "landm_t")):
            if not np.array_equal(ref[k_ref], o[k_o]):
                print("MATCH MISMATCH (C oracle)", trial, k_ref); n_bad += 1
        if not np.array_equal(ref["loc_t"][:, :2], o["loc_t"][:, :2]):
            print("MATCH loc cxcy MISMATCH", trial); n_bad += 1
        rel = np.abs(ref["loc_t"] - o["loc_t"]) / np.maximum(np.abs(ref["loc_t"]), 1e-30)
        stats["loc_maxrel"] = max(stats["loc_maxrel"], float(rel[np.isfinite(rel)].max()))
        if not np.allclose(ref["loc_t"], o["loc_t"], rtol=1e-5, atol=1e-6):
            print("MATCH loc tol MISMATCH", trial); n_bad += 1
        # torch port
        P = priors.shape[0]
        loc_t = torch.zeros(1, P, 4); conf_t = torch.zeros(1, P, dtype=torch.long); landm_t = torch.zeros(1, P, 10)
        tp.assign_one(THR, gt[:, :4], priors, VAR, gt[:, -1], gt[:, 4:14], loc_t, conf_t, landm_t, 0)
        if not (np.array_equal(loc_t[0].numpy(), ref["loc_t"]) and np.array_equal(conf_t[0].numpy(), ref["conf_t"])
                and np.array_equal(landm_t[0].numpy(), ref["landm_t"])):
            print("MATCH MISMATCH (torch port)", trial); n_bad += 1
        # decode + nms
        loc, conf, landm = synth.make_preds_clustered(3, 500 + trial, priors, gt, VAR)
        rb = ub.decode(loc, priors, VAR).numpy()
        ob = orc.decode(loc.numpy(), priors.numpy(), VAR)
        stats["dec_maxabs"] = max(stats["dec_maxabs"], float(np.abs(rb - ob).max()))
        if not np.allclose(rb, ob, rtol=1e-5, atol=1e-6):
            print("DECODE tol MISMATCH", trial); n_bad += 1
        if not np.array_equal(ub.decode_landm(landm, priors, VAR).numpy(),
                              orc.decode_landm(landm.numpy(), priors.numpy(), VAR)):
            print("DECODE_LANDM MISMATCH", trial); n_bad += 1
        if not torch.equal(tp.decode_boxes(loc, priors, VAR), ub.decode(loc, priors, VAR)):
            print("DECODE MISMATCH (torch port)", trial); n_bad += 1
        s = conf[:, 1]
        sel = s > 0.02
        for thr in (0.3, 0.4):
            k_ref = tv_nms(torch.from_numpy(rb)[sel], s[sel], thr).numpy()
            k_o = orc.nms_tv(rb[sel.numpy()], s[sel].numpy(), thr)
            if not np.array_equal(k_ref, k_o):
                print("NMS MISMATCH", trial, thr); n_bad += 1
        det = torch.cat([torch.from_numpy(rb), conf[:, 1:2], ub.decode_landm(landm, priors, VAR)], -1)
        r1 = ub.non_max_suppression(det, 0.5, 0.3)
        r2 = orc.non_max_suppression(det.numpy(), 0.5, 0.3)
        r3 = tp.suppress(det, 0.5, 0.3)
        if len(r1) != len(r2) or (len(r1) and not (np.array_equal(r1, r2) and np.array_equal(r1, r3))):
            print("non_max_suppression MISMATCH", trial); n_bad += 1
    n_bad += check_next_rows(R)
    print("fuzz stats:", stats)
    print("CHECK", "FAILED (%d)" % n_bad if n_bad else "OK")
    return n_bad


def check_next_rows(R):
    """Fuzz of the SURVEY 8(f) oracles against the live reference: WIDER AP evaluation (utils_map), IoU-family overlaps
    and IouLoss gradients, DIoU-NMS."""
    from oracle import oracle as orc
    from oracle import torch_port as tp
    from oracle import wider_eval as ow
    _, _, r_bu, ub, _, diou = R
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils.utils_map as r_map
    sys.path.remove(REF)
    n_bad = 0
    for trial in range(60):
        gt, keeps, pred = synth.make_eval_image(9, trial, count=(None if trial % 5 else 200))
        if len(gt) == 0 or len(pred) == 0:
            continue
        keep = keeps[trial % 3]
        for thr in (0.4, 0.5):
            with np.errstate(divide="ignore", invalid="ignore"):
                rec, prop = r_map.image_eval(pred, gt, keep.astype(np.float64), thr)
            o_rec, o_prop = ow.image_eval(pred, gt, keep, thr)
            info = r_map.img_pr_info(200, pred, prop, rec)
            if not (np.array_equal(rec, o_rec) and np.array_equal(prop, o_prop) and
                    np.array_equal(info, ow.img_pr_info(200, pred, o_prop, o_rec))):
                print("WIDER EVAL MISMATCH", trial, thr); n_bad += 1
    rng = np.random.default_rng(4242)
    for trial in range(30):
        n = int(rng.integers(1, 400))
        a, b = family_boxes(n=max(n, 64), seed=1000 + trial)
        a, b = torch.from_numpy(a[:n]), torch.from_numpy(b[:n])
        for kind in ("iou", "giou", "diou", "ciou"):
            ref = getattr(r_bu, "bbox_overlaps_" + kind)(a, b).numpy()
            if not np.array_equal(ref, tp.overlaps_family(a, b, kind).numpy(), equal_nan=True):
                print("OVERLAPS MISMATCH", kind, trial); n_bad += 1
        bx = torch.from_numpy(np.ascontiguousarray(b.numpy()))
        sc = torch.from_numpy(rng.permutation(n).astype(np.float32) / n)          # unique scores: one sort order
        for (ov, tk, beta) in ((0.5, 200, 1.0), (0.3, n, 1.0)):
            k_ref, c_ref = ub.diounms(bx, sc, ov, tk, beta)
            k_o, c_o = orc.diounms(bx.numpy(), sc.numpy(), ov, tk, beta)
            if c_ref != c_o or not np.array_equal(k_ref.numpy()[:c_ref], k_o[:c_o]):
                print("DIOUNMS MISMATCH", trial, ov, tk); n_bad += 1
    return n_bad


# ----------------------------------------------------------------------------- MultiBox loss
LOSS_CASES = [("s160", (160, 160), 3, 12), ("s640", (640, 640), 2, None)]


def loss_inputs(r_anchors, cfg, size, batch, count, cfg_id=6):
    with contextlib.redirect_stdout(io.StringIO()):
        pri = r_anchors.Anchors(cfg, image_size=size).get_anchors()
    P = pri.shape[0]
    targets = [synth.make_gt(cfg_id, i, size, count=count) for i in range(batch)]
    preds = [synth.make_logits(cfg_id, i, P) for i in range(batch)]
    loc = torch.stack([p[0] for p in preds]).requires_grad_(True)
    conf = torch.stack([p[1] for p in preds]).requires_grad_(True)
    landm = torch.stack([p[2] for p in preds]).requires_grad_(True)
    return pri, targets, (loc, conf, landm)


def gen_loss(R):
    """MultiBoxLoss.forward + backward of the reference (cuda=False) on seeded logits and GT."""
    r_anchors, r_config, _, _, r_training, _ = R
    out = {}
    for tag, size, batch, count in LOSS_CASES:
        pri, targets, preds = loss_inputs(r_anchors, r_config.cfg_mnet, size, batch, count)
        crit = r_training.MultiBoxLoss(2, THR, 7, VAR, False)
        l, c, m = crit(preds, pri, targets)
        (1.0 * l + 2.0 * c + 0.5 * m).backward()
        out[tag + "_losses"] = np.array([l.item(), c.item(), m.item()], dtype=np.float32)
        g_loc, g_conf, g_landm = (p.grad.numpy() for p in preds)
        out[tag + "_sel"] = np.packbits((np.abs(g_conf).sum(2) != 0))       # priors that enter loss_c (pos | mined neg)
        out[tag + "_gsum"] = np.array([np.abs(g_loc).sum(), np.abs(g_conf).sum(), np.abs(g_landm).sum()], dtype=np.float64)
        if tag == "s160":
            out[tag + "_g_loc"], out[tag + "_g_conf"], out[tag + "_g_landm"] = g_loc, g_conf, g_landm
        else:
            out[tag + "_g_loc_sha"], out[tag + "_g_landm_sha"] = sha(g_loc), sha(g_landm)
            nz = np.flatnonzero(np.abs(g_conf).sum(2).reshape(-1))
            out[tag + "_g_conf_nz_idx"] = nz.astype(np.int32)
            out[tag + "_g_conf_nz"] = g_conf.reshape(-1, 2)[nz]
            nzl = np.flatnonzero(np.abs(g_loc).sum(2).reshape(-1))
            out[tag + "_g_loc_nz_idx"] = nzl.astype(np.int32)
            out[tag + "_g_loc_nz"] = g_loc.reshape(-1, 4)[nzl]
            out[tag + "_g_landm_nz"] = g_landm.reshape(-1, 10)[nzl]
    save("loss.npz", **out)


# ----------------------------------------------------------------------------- box post-processing
POST_SHAPES = [(480, 640), (1080, 1920), (333, 517), (640, 640), (1200, 800)]


def gen_post(R):
    """retinaface_correct_boxes (R/utils/utils_bbox.py:9-24) followed by the pixel scaling of R/predict.py:194-195."""
    r_utils_bbox = R[3]
    rng = np.random.default_rng(1234)
    out = {}
    for i, (h, w) in enumerate(POST_SHAPES):
        x = rng.random((23, 15)).astype(np.float32)
        y = r_utils_bbox.retinaface_correct_boxes(x.copy(), np.array([640, 640]), np.array([h, w]))
        scale = [w, h, w, h]
        scale_for_landmarks = [w, h] * 5
        z = y.copy()
        z[:, :4] = z[:, :4] * scale
        z[:, 5:] = z[:, 5:] * scale_for_landmarks
        out["in_%d" % i], out["letterbox_%d" % i], out["pixels_%d" % i] = x, y, z
    save("post.npz", **out)


# ----------------------------------------------------------------------------- WIDER AP evaluation
EVAL_IMAGES = 24


def gen_wider_eval(R):
    """image_eval / img_pr_info / norm_score / dataset_pr_info / voc_ap of R/utils/utils_map.py on a seeded synthetic
    evaluation set (synth.make_eval_image), composed exactly like evaluation() (:181-211)."""
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils.utils_map as r_map
    sys.path.remove(REF)
    imgs = [synth.make_eval_image(6, i) for i in range(EVAL_IMAGES)]
    pred = {"ev": {str(i): imgs[i][2].copy() for i in range(EVAL_IMAGES)}}
    r_map.norm_score(pred)
    out = {"n_images": np.array(EVAL_IMAGES)}
    thresh_num = 1000
    for i in range(EVAL_IMAGES):
        out["norm_%d" % i] = pred["ev"][str(i)]
    box = np.array([[0., 0., 10., 10.], [5., 5., 15., 20.], [0., 0., 0., 0.], [3., 3., 4., 9.]])
    out["overlaps_in"] = box
    with np.errstate(divide="ignore", invalid="ignore"):
        out["overlaps"] = r_map.bbox_overlaps(box, box[::-1].copy())
    for s, name in enumerate(("easy", "medium", "hard")):
        pr_curve = np.zeros((thresh_num, 2))
        count_face = 0
        for i in range(EVAL_IMAGES):
            gt, keeps, _ = imgs[i]
            p = pred["ev"][str(i)]
            keep_index = np.nonzero(keeps[s])[0] + 1           # 1-based like the .mat gt_list
            count_face += len(keep_index)
            if len(gt) == 0 or len(p) == 0:
                continue
            ignore = np.zeros(gt.shape[0])
            if len(keep_index) != 0:
                ignore[keep_index - 1] = 1
            with np.errstate(divide="ignore", invalid="ignore"):
                rec, prop = r_map.image_eval(p, gt, ignore, 0.4)
            info = r_map.img_pr_info(thresh_num, p, prop, rec)
            if s == 2:
                out["recall_%d" % i], out["proposal_%d" % i] = rec, prop
                if i < 4:
                    out["pr_info_%d" % i] = info
            pr_curve += info
        out["pr_curve_" + name] = pr_curve
        out["count_face_" + name] = np.array(count_face)
        with np.errstate(divide="ignore", invalid="ignore"):
            pc = r_map.dataset_pr_info(thresh_num, pr_curve, count_face)
        out["ap_" + name] = np.array(r_map.voc_ap(pc[:, 1], pc[:, 0]))
    save("wider_eval.npz", **out)


# ----------------------------------------------------------------------------- IoU family
def family_boxes(n=600, seed=77):
    """Aligned box pairs: jittered copies (high IoU), random pairs (mostly disjoint), nested, identical and
    edge-sharing pairs -- every branch of the clamps / min / max."""
    rng = np.random.default_rng(seed)
    c = rng.random((n, 2)).astype(np.float32)
    wh = (0.02 + 0.3 * rng.random((n, 2))).astype(np.float32)
    a = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    b = a + (0.05 * rng.standard_normal((n, 4)) * np.concatenate([wh, wh], 1)).astype(np.float32)
    b[: n // 4] = np.roll(a[: n // 4], 7, axis=0)                      # unrelated boxes
    b[n // 4: n // 4 + 20] = a[n // 4: n // 4 + 20]                      # identical
    k = n // 4 + 20
    b[k: k + 20, :2] = a[k: k + 20, :2] + 0.25 * wh[k: k + 20]          # nested
    b[k: k + 20, 2:] = a[k: k + 20, 2:] - 0.25 * wh[k: k + 20]
    b[k + 20: k + 40] = a[k + 20: k + 40] + np.float32([1, 0, 1, 0]) * (a[k + 20: k + 40, 2:3] - a[k + 20: k + 40, 0:1])  # touching
    b[:, 2:] = np.maximum(b[:, 2:], b[:, :2] + np.float32(1e-3))
    return a, b.astype(np.float32)


DIOU_LOSS_CASES = [("s160", (160, 160), 3, None), ("s320", (320, 320), 2, 40)]


def gen_iou_family(R):
    """bbox_overlaps_{iou,giou,diou,ciou} (R/utils/box_utils.py:5-158), IouLoss + autograd and the DIoU MultiBoxLoss
    (R/nets/retinaface_training_DIOU.py:491-665, cuda=False), diounms (R/utils/utils_bbox.py:182-258)."""
    r_anchors, r_config, r_box_utils, r_utils_bbox, _, r_diou = R
    out = {}
    a, b = family_boxes()
    out["a"], out["b"] = a, b
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    for kind in ("iou", "giou", "diou", "ciou"):
        out["ov_" + kind] = getattr(r_box_utils, "bbox_overlaps_" + kind)(ta, tb).numpy()
        assert np.array_equal(getattr(r_diou, "bbox_overlaps_" + kind)(ta, tb).numpy(), out["ov_" + kind], equal_nan=True), kind
    out["ov_diou_bcast"] = r_box_utils.bbox_overlaps_diou(ta[:1], tb).numpy()          # one row against many
    # IouLoss on decoded predictions, all loss types, with autograd
    pri = r_anchors.Anchors(r_config.cfg_mnet, image_size=(160, 160)).get_anchors()[:600].contiguous()
    g = torch.Generator().manual_seed(5)
    enc = r_diou.encode(tb.clamp(0.01, 0.99), pri, VAR) if hasattr(r_diou, "encode") else None
    loc = (enc + 0.3 * torch.randn(enc.shape, generator=g)).clamp(-4, 4)
    out["loss_loc"], out["loss_pri"] = loc.numpy(), pri.numpy()
    for lt in ("Iou", "Giou", "Diou", "Ciou"):
        for size_sum in (True, False):
            lp = loc.clone().requires_grad_(True)
            crit = r_diou.IouLoss(pred_mode='Center', size_sum=size_sum, variances=VAR, losstype=lt)
            val = crit(lp, tb, pri)
            val.backward()
            tag = "%s_%d" % (lt, int(size_sum))
            out["loss_" + tag], out["grad_" + tag] = np.array(val.item(), np.float32), lp.grad.numpy()
    lp = ta.clone().requires_grad_(True)
    val = r_diou.IouLoss(pred_mode='Corner', size_sum=True, variances=VAR, losstype='Giou')(lp, tb, pri)
    val.backward()
    out["loss_corner_giou"], out["grad_corner_giou"] = np.array(val.item(), np.float32), lp.grad.numpy()
    # the DIoU MultiBoxLoss
    for tag, size, batch, count in DIOU_LOSS_CASES:
        pri2, targets, preds = loss_inputs(r_anchors, r_config.cfg_mnet, size, batch, count)
        crit = r_diou.MultiBoxLoss(2, THR, 7, VAR, False)
        l, c, m = crit(preds, pri2, targets)
        (1.0 * l + 2.0 * c + 0.5 * m).backward()
        out["mbl_%s_losses" % tag] = np.array([l.item(), c.item(), m.item()], dtype=np.float32)
        g_loc, g_conf, g_landm = (p.grad.numpy() for p in preds)
        out["mbl_%s_sel" % tag] = np.packbits((np.abs(g_conf).sum(2) != 0))
        nzl = np.flatnonzero(np.abs(g_loc).sum(2).reshape(-1))
        out["mbl_%s_g_loc_nz_idx" % tag] = nzl.astype(np.int32)
        out["mbl_%s_g_loc_nz" % tag] = g_loc.reshape(-1, 4)[nzl]
    # diounms on the NMS fixtures' boxes
    nms = np.load(os.path.join(HERE, "nms.npz"))
    for name in ("rand500", "dense2000"):
        bx, sc = torch.from_numpy(nms[name + "_boxes"]), torch.from_numpy(nms[name + "_scores"])
        assert len(np.unique(nms[name + "_scores"])) == len(nms[name + "_scores"])   # the unstable sort has one answer
        for (ov, tk, beta) in ((0.5, 200, 1.0), (0.3, 5000, 1.0), (0.45, 5000, 0.6)):
            keep, count = r_utils_bbox.diounms(bx, sc, ov, tk, beta)
            out["dnms_%s_%d_%d_%d_keep" % (name, int(ov * 100), tk, int(beta * 10))] = keep.numpy()[:count]
    save("iou_family.npz", **out)


def main():
    R = import_reference()
    gen_priors(R)
    gen_match(R)
    gen_decode(R)
    gen_nms(R)
    gen_pipeline(R)
    gen_loss(R)
    gen_post(R)
    gen_wider_eval(R)
    gen_iou_family(R)
    if "--check" in sys.argv:
        sys.exit(1 if check(R) else 0)


if __name__ == "__main__":
    main()
