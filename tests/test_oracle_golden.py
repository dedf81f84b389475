"""Pins the CPU oracle (oracle/jabd_oracle.c) and the torch port against the
golden vectors recorded from the imported reference (tests/golden/make_golden.py).

bit-exact: priors, indices, labels, overlaps, encode cx/cy, landmark encode/decode,
NMS keep lists.  rtol 1e-5 / atol 1e-6: the log/exp halves (glibc logf/expf vs
torch's SLEEF differ by <= 1 ulp).
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL, load_golden
from jabd_b200 import config as cfgs
from jabd_b200 import synth
from oracle import oracle as orc
from oracle import torch_port as tp

VAR = [0.1, 0.2]
THR = 0.35


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


PRIOR_SIZES = [(640, 640), (1024, 1024), (840, 840), (96, 128), (100, 75), (333, 517)]


@pytest.mark.parametrize("name", sorted(cfgs.ALL_CFGS))
def test_priors_all_cfgs(name):
    g = load_golden("priors.npz")
    for (h, w) in PRIOR_SIZES:
        a = orc.priors(cfgs.ALL_CFGS[name], (h, w))
        key = "%s_%dx%d" % (name, h, w)
        assert a.shape[0] == int(g[key + "_n"]) == cfgs.num_priors(cfgs.ALL_CFGS[name], (h, w))
        assert np.array_equal(sha(a), g[key + "_sha"]), key
        if key in g.files:
            assert np.array_equal(a, g[key])


def test_priors_2048_and_clip():
    g = load_golden("priors.npz")
    a = orc.priors(cfgs.cfg_mnet, (2048, 2048))
    assert a.shape[0] == 172032 and np.array_equal(sha(a), g["cfg_mnet_2048x2048_sha"])
    c = dict(cfgs.cfg_mnet, clip=True)
    assert np.array_equal(orc.priors(c, (100, 75)), g["clip_mnet_100x75"])


def _check_match(o, g, prefix, exact_loc=False):
    assert np.array_equal(o["conf_t"], g[prefix + "conf_t"])
    assert np.array_equal(o["best_truth_idx"], g[prefix + "bti"])
    assert np.array_equal(o["best_truth_overlap"], g[prefix + "bto"])
    assert np.array_equal(o["best_prior_idx"], g[prefix + "bpi"])
    assert np.array_equal(o["best_prior_overlap"], g[prefix + "bpo"])
    assert np.array_equal(o["loc_t"][:, :2], g[prefix + "loc_t"][:, :2])
    np.testing.assert_allclose(o["loc_t"], g[prefix + "loc_t"], rtol=RTOL, atol=ATOL)
    if prefix + "landm_t" in g.files:
        assert np.array_equal(o["landm_t"], g[prefix + "landm_t"])


def test_match_small_cases():
    g = load_golden("match_small.npz")
    pri = g["priors_160"]
    for name in g["names"]:
        name = str(name)
        gt = g[name + "_gt"]
        thr = float(g["edge_thr"]) if name == "edge" else THR
        o = orc.match(thr, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14])
        _check_match(o, g, name + "_")
    # the threshold-edge case must differ from the default-threshold run on the same GT
    assert (g["edge_conf_t"] != g["crowd40_conf_t"]).any()


def test_match_small_torch_port():
    g = load_golden("match_small.npz")
    pri = torch.from_numpy(g["priors_160"])
    P = pri.shape[0]
    for name in g["names"]:
        name = str(name)
        gt = torch.from_numpy(g[name + "_gt"])
        thr = float(g["edge_thr"]) if name == "edge" else THR
        loc_t = torch.zeros(1, P, 4); conf_t = torch.zeros(1, P, dtype=torch.long); landm_t = torch.zeros(1, P, 10)
        tp.assign_one(thr, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14], loc_t, conf_t, landm_t, 0)
        assert np.array_equal(loc_t[0].numpy(), g[name + "_loc_t"])
        assert np.array_equal(conf_t[0].numpy(), g[name + "_conf_t"])
        assert np.array_equal(landm_t[0].numpy(), g[name + "_landm_t"])


def test_match_variants_and_pieces():
    g = load_golden("match_small.npz")
    pri = g["priors_160"]
    gt = g["rand7_gt"]
    o = orc.match(THR, gt[:, :4], pri, VAR, gt[:, -1], None, label_mode=1, encode_mode=1)   # box_utils.match
    assert np.array_equal(o["conf_t"], g["ssd_match_conf_t"])
    np.testing.assert_allclose(o["loc_t"], g["ssd_match_loc_t"], rtol=RTOL, atol=ATOL)
    o = orc.match(THR, gt[:, :4], pri, VAR, gt[:, -1], None, label_mode=1, encode_mode=0)   # match_ious
    assert np.array_equal(o["conf_t"], g["ssd_match_ious_conf_t"])
    assert np.array_equal(o["loc_t"], g["ssd_match_ious_loc_t"])
    o = orc.match(THR, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14], label_mode=0, encode_mode=0)  # match_iou
    assert np.array_equal(o["conf_t"], g["diou_match_iou_conf_t"])
    assert np.array_equal(o["loc_t"], g["diou_match_iou_loc_t"])
    assert np.array_equal(o["landm_t"], g["diou_match_iou_landm_t"])
    assert np.array_equal(orc.point_form(pri), g["point_form_160"])
    assert np.array_equal(orc.jaccard(gt[:, :4], orc.point_form(pri)), g["rand7_jaccard"])
    bti = g["rand7_bti"]
    e = orc.encode(gt[:, :4][bti], pri, VAR)
    assert np.array_equal(e[:, :2], g["rand7_encode"][:, :2])
    np.testing.assert_allclose(e, g["rand7_encode"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(orc.encode_landm(gt[:, 4:14][bti], pri, VAR), g["rand7_encode_landm"])


def test_match_640_cfg1_cfg2():
    g = load_golden("match_640.npz")
    pri = orc.priors(cfgs.cfg_mnet, (640, 640))
    assert np.array_equal(sha(pri), g["priors_640_sha"])
    gt1 = synth.make_gt(1, 0, (640, 640)).numpy()
    assert np.array_equal(gt1, g["cfg1_gt"]), "synthetic generator drifted from the recorded inputs"
    o = orc.match(THR, gt1[:, :4], pri, VAR, gt1[:, -1], gt1[:, 4:14])
    _check_match(o, g, "cfg1_")
    gt2 = synth.make_gt(2, 6, (640, 640)).numpy()
    assert np.array_equal(gt2, g["cfg2_gt"]) and gt2.shape[0] == 211
    o = orc.match(THR, gt2[:, :4], pri, VAR, gt2[:, -1], gt2[:, 4:14])
    _check_match(o, g, "cfg2_")
    assert np.array_equal(sha(o["landm_t"]), g["cfg2_landm_t_sha"])


def test_match_empty_gt_raises():
    pri = orc.priors(cfgs.cfg_mnet, (96, 128))
    with pytest.raises(ValueError):
        orc.match(THR, np.zeros((0, 4), np.float32), pri, VAR, np.zeros((0,), np.float32), np.zeros((0, 10), np.float32))


def test_decode():
    g = load_golden("decode.npz")
    pri = orc.priors(cfgs.cfg_mnet, (160, 160))
    np.testing.assert_allclose(orc.decode(g["loc"], pri, VAR), g["boxes"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(orc.decode_landm(g["landm"], pri, VAR), g["landms"])
    assert torch.equal(tp.decode_boxes(torch.from_numpy(g["loc"]), torch.from_numpy(pri), VAR), torch.from_numpy(g["boxes"]))
    assert torch.equal(tp.decode_points(torch.from_numpy(g["landm"]), torch.from_numpy(pri), VAR), torch.from_numpy(g["landms"]))
    pri640 = orc.priors(cfgs.cfg_mnet, (640, 640))
    loc, conf, landm = synth.make_preds_random(1, 0, pri640.shape[0])
    assert np.array_equal(sha(loc.numpy()), g["loc640_sha"]), "synthetic generator drifted"
    np.testing.assert_allclose(orc.decode(loc.numpy(), pri640, VAR), g["boxes640"], rtol=RTOL, atol=ATOL)
    lm = orc.decode_landm(landm.numpy(), pri640, VAR)
    assert np.array_equal(sha(lm), g["landms640_sha"]) and np.array_equal(lm[:1024], g["landms640_head"])


def test_nms_torchvision_semantics():
    g = load_golden("nms.npz")
    for name in g["names"]:
        name = str(name)
        b, s = g[name + "_boxes"], g[name + "_scores"]
        for thr in (0.3, 0.4, 0.5):
            k = orc.nms_tv(b, s, thr)
            assert np.array_equal(k, g["%s_keep_%d" % (name, int(thr * 100))]), (name, thr)
    assert orc.nms_tv(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.3).shape == (0,)
    assert g["empty_keep"].shape == (0,)


def test_nms_ssd_legacy():
    g = load_golden("nms.npz")
    for name in ("rand500", "dense2000"):
        b, s = g[name + "_boxes"], g[name + "_scores"]
        for (ov, tk) in ((0.5, 200), (0.3, 50), (0.45, 5000)):
            keep, count = orc.nms_ssd(b, s, ov, tk)
            assert count == int(g["%s_ssd_%d_%d_count" % (name, int(ov * 100), tk)])
            assert np.array_equal(keep, g["%s_ssd_%d_%d_keep" % (name, int(ov * 100), tk)])


def _pipeline_inputs(g, tag, gen):
    size = (160, 160) if tag == "s160" else (640, 640)
    img = 1 if tag == "s160" else 2
    pri = orc.priors(cfgs.cfg_mnet, size)
    key = "%s_%s_" % (tag, gen)
    gt = torch.from_numpy(g[key + "gt"])
    if gen == "A":
        loc, conf, landm = synth.make_preds_random(3, img, pri.shape[0])
    else:
        loc, conf, landm = synth.make_preds_clustered(3, img, torch.from_numpy(pri), gt, VAR)
    loc, conf, landm = loc.numpy(), conf.numpy(), landm.numpy()
    if tag == "s160":
        assert np.array_equal(loc, g[key + "loc"]) and np.array_equal(conf, g[key + "conf"])
    else:
        assert np.array_equal(sha(np.concatenate([loc.ravel(), conf.ravel(), landm.ravel()])), g[key + "in_sha"])
    return pri, loc, conf, landm


@pytest.mark.parametrize("tag", ["s160", "s640"])
@pytest.mark.parametrize("gen", ["A", "B"])
def test_pipeline_dropin_and_topk(tag, gen):
    g = load_golden("pipeline.npz")
    pri, loc, conf, landm = _pipeline_inputs(g, tag, gen)
    key = "%s_%s_" % (tag, gen)
    # boxes as the reference decoded them (torch exp) are not stored; NMS decisions are checked on the
    # oracle's own decode, which the fuzz check showed to be within 2.4e-7 abs of torch's.
    boxes = orc.decode(loc, pri, VAR)
    lms = orc.decode_landm(landm, pri, VAR)
    det = np.concatenate([boxes, conf[:, 1:2], lms], 1)
    for ct, nt in ((0.5, 0.3), (0.05, 0.3)):
        ref = g[key + "nms_%d_%d" % (int(ct * 100), int(nt * 100))]
        out = orc.non_max_suppression(det, ct, nt)
        out = np.zeros((0, 15), np.float32) if isinstance(out, list) else out
        assert out.shape == ref.shape
        assert np.array_equal(out[:, 4], ref[:, 4])          # same boxes kept, same order
        np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    for (ct, topk, nt, keepk) in ((0.02, 5000, 0.4, 750), (0.02, 200, 0.4, 50)):
        k2 = key + "pipe_%d_%d_" % (topk, keepk)
        dets, idx = orc.detect(loc, conf, landm, pri, VAR, ct, True, topk, nt, keepk)
        assert np.array_equal(idx, g[k2 + "idx"])
        np.testing.assert_allclose(dets, g[k2 + "dets"], rtol=RTOL, atol=ATOL)
        d2, i2 = tp.infer_one_topk(torch.from_numpy(loc), torch.from_numpy(conf), torch.from_numpy(landm),
                                   torch.from_numpy(pri), VAR, ct, topk, nt, keepk)
        assert np.array_equal(i2.numpy(), g[k2 + "idx"]) and np.array_equal(d2.numpy(), g[k2 + "dets"])


# ------------------------------------------------------------------ MultiBox loss
LOSS_CASES = [("s160", (160, 160), 3, 12), ("s640", (640, 640), 2, None)]


def _loss_inputs(size, batch, count, cfg_id=6):
    from jabd_b200 import synth
    pri = torch.from_numpy(orc.priors(cfgs.cfg_mnet, size))
    P = pri.shape[0]
    targets = [synth.make_gt(cfg_id, i, size, count=count) for i in range(batch)]
    preds = [synth.make_logits(cfg_id, i, P) for i in range(batch)]
    return pri, targets, tuple(torch.stack([p[k] for p in preds]).requires_grad_(True) for k in range(3))


@pytest.mark.parametrize("tag,size,batch,count", LOSS_CASES)
def test_multibox_loss_torch_port(tag, size, batch, count):
    """oracle/torch_port.multibox_loss == the reference's MultiBoxLoss (forward values, selection, gradients)."""
    g = load_golden("loss.npz")
    pri, targets, preds = _loss_inputs(size, batch, count)
    l, c, m = tp.multibox_loss(preds, pri, targets, 0.35, [0.1, 0.2], 7)
    (1.0 * l + 2.0 * c + 0.5 * m).backward()
    np.testing.assert_array_equal(np.array([l.item(), c.item(), m.item()], dtype=np.float32), g[tag + "_losses"])
    g_loc, g_conf, g_landm = (p.grad.numpy() for p in preds)
    np.testing.assert_array_equal(np.packbits(np.abs(g_conf).sum(2) != 0), g[tag + "_sel"])
    if tag == "s160":
        np.testing.assert_array_equal(g_loc, g[tag + "_g_loc"])
        np.testing.assert_array_equal(g_conf, g[tag + "_g_conf"])
        np.testing.assert_array_equal(g_landm, g[tag + "_g_landm"])
    else:
        np.testing.assert_array_equal(sha(g_loc), g[tag + "_g_loc_sha"])
        np.testing.assert_array_equal(g_conf.reshape(-1, 2)[g[tag + "_g_conf_nz_idx"]], g[tag + "_g_conf_nz"])


def test_correct_boxes_host_restatement():
    """the host-side drop-in of retinaface_correct_boxes (R/utils/utils_bbox.py:9-24) and the
    letterbox_params used by the CUDA kernel reproduce the reference's outputs bit for bit."""
    from jabd_b200 import batched, utils_bbox
    g = load_golden("post.npz")
    shapes = [(480, 640), (1080, 1920), (333, 517), (640, 640), (1200, 800)]
    post = batched.letterbox_params((640, 640), shapes)
    for i, (h, w) in enumerate(shapes):
        x = g["in_%d" % i]
        y = utils_bbox.retinaface_correct_boxes(x.copy(), np.array([640, 640]), np.array([h, w]))
        np.testing.assert_array_equal(y, g["letterbox_%d" % i])
        # the same arithmetic from the [B,6] parameter rows (what correct_boxes_kernel evaluates in fp64)
        q = post[i]
        z = x.copy()
        for c in range(15):
            if c == 4:
                continue
            xy = (c if c < 4 else c - 5) & 1
            z[:, c] = ((z[:, c].astype(np.float64) - q[xy]) * q[2 + xy]).astype(np.float32)
            z[:, c] = (z[:, c].astype(np.float64) * q[4 + xy]).astype(np.float32)
        np.testing.assert_array_equal(z, g["pixels_%d" % i])


def test_wider_eval_oracle_matches_reference():
    """oracle/wider_eval.py reproduces the reference's utils_map.py (image_eval, img_pr_info,
    norm_score, dataset_pr_info + voc_ap) on the seeded synthetic evaluation set, exactly (fp64, integer counters)."""
    from jabd_b200 import synth
    from oracle import wider_eval as ow
    g = load_golden("wider_eval.npz")
    n = int(g["n_images"])
    imgs = [synth.make_eval_image(6, i) for i in range(n)]
    normed = ow.norm_scores([im[2] for im in imgs])
    for i in range(n):
        np.testing.assert_array_equal(normed[i], g["norm_%d" % i])
    np.testing.assert_array_equal(ow.bbox_overlaps(g["overlaps_in"], g["overlaps_in"][::-1]), g["overlaps"])
    for i in range(n):
        gt, keeps, _ = imgs[i]
        if len(gt) == 0 or len(normed[i]) == 0:
            continue
        rec, prop = ow.image_eval(normed[i], gt, keeps[2], 0.4)
        np.testing.assert_array_equal(rec, g["recall_%d" % i])
        np.testing.assert_array_equal(prop, g["proposal_%d" % i])
        if i < 4:
            np.testing.assert_array_equal(ow.img_pr_info(1000, normed[i], prop, rec), g["pr_info_%d" % i])
    for s, name in enumerate(("easy", "medium", "hard")):
        pr = ow.pr_counters(normed, [im[0] for im in imgs], [im[1][s] for im in imgs], 0.4, 1000)
        np.testing.assert_array_equal(pr, g["pr_curve_" + name])
        assert ow.average_precision(pr, int(g["count_face_" + name])) == float(g["ap_" + name])


def test_iou_family_torch_port_and_c_oracle():
    """oracle/torch_port.overlaps_family / iou_loss / multibox_loss(loc_loss="diou") and
    oracle.diounms reproduce the reference (box_utils.bbox_overlaps_*, DIOU.IouLoss, DIOU.MultiBoxLoss, utils_bbox.diounms)."""
    g = load_golden("iou_family.npz")
    a, b = torch.from_numpy(g["a"]), torch.from_numpy(g["b"])
    for kind in ("iou", "giou", "diou", "ciou"):
        np.testing.assert_array_equal(tp.overlaps_family(a, b, kind).numpy(), g["ov_" + kind])
    np.testing.assert_array_equal(tp.overlaps_family(a[:1], b, "diou").numpy(), g["ov_diou_bcast"])
    loc, pri = torch.from_numpy(g["loss_loc"]), torch.from_numpy(g["loss_pri"])
    for lt in ("Iou", "Giou", "Diou", "Ciou"):
        for size_sum in (True, False):
            lp = loc.clone().requires_grad_(True)
            val = tp.iou_loss(lp, b, pri, [0.1, 0.2], lt.lower(), True, size_sum)
            val.backward()
            tag = "%s_%d" % (lt, int(size_sum))
            assert np.float32(val.item()) == g["loss_" + tag]
            np.testing.assert_array_equal(lp.grad.numpy(), g["grad_" + tag])
    for tag, size, batch, count in (("s160", (160, 160), 3, None), ("s320", (320, 320), 2, 40)):
        pri2, targets, preds = _loss_inputs(size, batch, count)
        l, c, m = tp.multibox_loss(preds, pri2, targets, 0.35, [0.1, 0.2], 7, loc_loss="diou")
        (1.0 * l + 2.0 * c + 0.5 * m).backward()
        np.testing.assert_array_equal(np.array([l.item(), c.item(), m.item()], dtype=np.float32), g["mbl_%s_losses" % tag])
        g_loc, g_conf = preds[0].grad.numpy(), preds[1].grad.numpy()
        np.testing.assert_array_equal(np.packbits(np.abs(g_conf).sum(2) != 0), g["mbl_%s_sel" % tag])
        np.testing.assert_array_equal(g_loc.reshape(-1, 4)[g["mbl_%s_g_loc_nz_idx" % tag]], g["mbl_%s_g_loc_nz" % tag])
    nms = load_golden("nms.npz")
    for name in ("rand500", "dense2000"):
        for (ov, tk, beta) in ((0.5, 200, 1.0), (0.3, 5000, 1.0), (0.45, 5000, 0.6)):
            keep, count = orc.diounms(nms[name + "_boxes"], nms[name + "_scores"], ov, tk, beta)
            np.testing.assert_array_equal(keep[:count], g["dnms_%s_%d_%d_%d_keep" % (name, int(ov * 100), tk, int(beta * 10))])
