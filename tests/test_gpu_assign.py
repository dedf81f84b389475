"""GPU parity of the training side (priors, IoU/argmax matching, encode) against the CPU oracle and the
golden vectors recorded from the reference.  Everything goes through the package -> ctypes -> C-ABI.

bit-exact: priors, conf_t, best_truth_idx, best_truth_overlap, best_prior_idx, best_prior_overlap, landm_t,
loc_t[:, :2]; rtol 1e-5 / atol 1e-6: loc_t[:, 2:] (log half; see conftest.RTOL/ATOL).
"""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL, load_golden

pytestmark = pytest.mark.gpu

VAR = [0.1, 0.2]
THR = 0.35


@pytest.fixture(scope="module")
def mods():
    from jabd_b200 import anchors, batched, box_utils, config, retinaface_training, synth
    from oracle import oracle as orc
    return dict(anchors=anchors, batched=batched, box_utils=box_utils, cfgs=config, rt=retinaface_training, synth=synth,
                orc=orc)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def check_assign(out, extra, ref, img=None):
    """out = (loc_t, conf_t, landm_t) CUDA; ref = oracle match_batch dict."""
    loc_t, conf_t, landm_t = [t.cpu().numpy() if t is not None else None for t in out]
    assert np.array_equal(conf_t, ref["conf_t"])
    assert np.array_equal(loc_t[..., :2], ref["loc_t"][..., :2])
    np.testing.assert_allclose(loc_t, ref["loc_t"], rtol=RTOL, atol=ATOL)
    if landm_t is not None and ref.get("landm_t") is not None:
        assert np.array_equal(landm_t, ref["landm_t"])
    if extra is not None:
        assert np.array_equal(extra["best_truth_idx"].cpu().numpy().astype(np.int64), ref["best_truth_idx"])
        assert np.array_equal(extra["best_truth_overlap"].cpu().numpy(), ref["best_truth_overlap"])
        assert np.array_equal(extra["best_prior_idx"].cpu().numpy().astype(np.int64), ref["best_prior_idx"])
        assert np.array_equal(extra["best_prior_overlap"].cpu().numpy(), ref["best_prior_overlap"])


# ------------------------------------------------------------------------------------------------ priors
SIZES = [(640, 640), (1024, 1024), (840, 840), (96, 128), (100, 75), (333, 517), (2048, 2048)]


def test_priors_all_cfgs_bit_exact(mods):
    for name, cfg in mods["cfgs"].ALL_CFGS.items():
        for size in SIZES:
            ref = mods["orc"].priors(cfg, size)
            got = mods["anchors"].Anchors(cfg, image_size=size).get_anchors()
            assert got.is_cuda and got.shape == (ref.shape[0], 4)
            assert np.array_equal(got.cpu().numpy(), ref), (name, size)
    c = dict(mods["cfgs"].cfg_mnet, clip=True)
    g = load_golden("priors.npz")
    assert np.array_equal(mods["anchors"].Anchors_eval(c, image_size=(100, 75)).get_anchors().cpu().numpy(), g["clip_mnet_100x75"])
    a = mods["anchors"].cached_priors(mods["cfgs"].cfg_mnet, (640, 640))
    assert mods["anchors"].cached_priors(mods["cfgs"].cfg_mnet, (640, 640)) is a and a.shape[0] == 16800


# ------------------------------------------------------------------------------- stand-alone operators
def test_box_operators(mods):
    orc, rt = mods["orc"], mods["rt"]
    g = load_golden("match_small.npz")
    pri = g["priors_160"]
    gt = g["rand7_gt"]
    pf = rt.point_form(cuda(pri))
    assert np.array_equal(pf.cpu().numpy(), g["point_form_160"])
    assert np.array_equal(rt.jaccard(cuda(gt[:, :4]), pf).cpu().numpy(), g["rand7_jaccard"])
    # numpy in -> numpy out (explicit staging), same values
    j_np = rt.jaccard(gt[:, :4], g["point_form_160"])
    assert isinstance(j_np, np.ndarray) and np.array_equal(j_np, g["rand7_jaccard"])
    inter = rt.intersect(cuda(gt[:, :4]), pf).cpu().numpy()
    a, b = gt[:, :4], g["point_form_160"]
    wh = np.clip(np.minimum(a[:, None, 2:], b[None, :, 2:]) - np.maximum(a[:, None, :2], b[None, :, :2]), 0, None)
    assert np.array_equal(inter, wh[..., 0] * wh[..., 1])
    bti = g["rand7_bti"]
    e = rt.encode(cuda(gt[:, :4][bti]), cuda(pri), VAR).cpu().numpy()
    assert np.array_equal(e[:, :2], g["rand7_encode"][:, :2])
    np.testing.assert_allclose(e, g["rand7_encode"], rtol=RTOL, atol=ATOL)
    el = rt.encode_landm(cuda(gt[:, 4:14][bti]), cuda(pri), VAR).cpu().numpy()
    assert np.array_equal(el, g["rand7_encode_landm"])
    # larger random jaccard against the oracle, odd sizes
    rng = np.random.default_rng(5)
    A = rng.random((37, 4), dtype=np.float32); A[:, 2:] += A[:, :2]
    Bx = rng.random((1001, 4), dtype=np.float32); Bx[:, 2:] = Bx[:, :2] + 0.2 * Bx[:, 2:]
    assert np.array_equal(rt.jaccard(cuda(A), cuda(Bx)).cpu().numpy(), orc.jaccard(A, Bx))


# ------------------------------------------------------------------------------------------ match goldens
def test_match_golden_small_cases(mods):
    """Edge cases recorded from the reference: ties, collisions, threshold edge, G=1, crowds."""
    g = load_golden("match_small.npz")
    pri = cuda(g["priors_160"])
    for name in g["names"]:
        name = str(name)
        gt = g[name + "_gt"]
        thr = float(g["edge_thr"]) if name == "edge" else THR
        for dense in (False, True):
            loc_t, conf_t, landm_t, ex = mods["batched"].assign_targets(pri, [cuda(gt)], threshold=thr, variances=VAR,
                                                                        return_match=True, dense=dense)
            assert np.array_equal(conf_t[0].cpu().numpy(), g[name + "_conf_t"]), (name, dense)
            assert np.array_equal(ex["best_truth_idx"][0].cpu().numpy(), g[name + "_bti"]), (name, dense)
            assert np.array_equal(ex["best_truth_overlap"][0].cpu().numpy(), g[name + "_bto"]), (name, dense)
            assert np.array_equal(ex["best_prior_idx"].cpu().numpy(), g[name + "_bpi"]), (name, dense)
            assert np.array_equal(ex["best_prior_overlap"].cpu().numpy(), g[name + "_bpo"]), (name, dense)
            assert np.array_equal(landm_t[0].cpu().numpy(), g[name + "_landm_t"]), (name, dense)
            assert np.array_equal(loc_t[0].cpu().numpy()[:, :2], g[name + "_loc_t"][:, :2]), (name, dense)
            np.testing.assert_allclose(loc_t[0].cpu().numpy(), g[name + "_loc_t"], rtol=RTOL, atol=ATOL)


def test_match_dropin_signatures(mods):
    """The reference's per-image signatures with in-place row writes (10-arg live, 8-arg SSD, match_iou)."""
    g = load_golden("match_small.npz")
    pri = cuda(g["priors_160"])
    P = pri.shape[0]
    gt = cuda(g["rand7_gt"])
    rt, bu = mods["rt"], mods["box_utils"]
    # live 10-arg match, CPU target tensors as MultiBoxLoss.forward allocates them (:197-199)
    loc_t, conf_t, landm_t = torch.zeros(2, P, 4), torch.zeros(2, P, dtype=torch.long), torch.zeros(2, P, 10)
    rt.match(THR, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14], loc_t, conf_t, landm_t, 1)
    assert np.array_equal(conf_t[1].numpy(), g["rand7_conf_t"]) and conf_t[0].abs().sum() == 0
    assert np.array_equal(landm_t[1].numpy(), g["rand7_landm_t"])
    np.testing.assert_allclose(loc_t[1].numpy(), g["rand7_loc_t"], rtol=RTOL, atol=ATOL)
    # CUDA target tensors
    loc_c, conf_c, landm_c = loc_t.cuda() * 0, conf_t.cuda() * 0, landm_t.cuda() * 0
    rt.match(THR, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14], loc_c, conf_c, landm_c, 0)
    assert np.array_equal(conf_c[0].cpu().numpy(), g["rand7_conf_t"])
    # match_iou (raw boxes)
    rt.match_iou(THR, gt[:, :4], pri, VAR, gt[:, -1], gt[:, 4:14], loc_c, conf_c, landm_c, 1)
    assert np.array_equal(loc_c[1].cpu().numpy(), g["diou_match_iou_loc_t"])
    assert np.array_equal(conf_c[1].cpu().numpy(), g["diou_match_iou_conf_t"])
    assert np.array_equal(landm_c[1].cpu().numpy(), g["diou_match_iou_landm_t"])
    # SSD 8-arg forms
    loc_s, conf_s = torch.zeros(1, P, 4).cuda(), torch.zeros(1, P, dtype=torch.long).cuda()
    bu.match(THR, gt[:, :4], pri, VAR, gt[:, -1], loc_s, conf_s, 0)
    assert np.array_equal(conf_s[0].cpu().numpy(), g["ssd_match_conf_t"])
    np.testing.assert_allclose(loc_s[0].cpu().numpy(), g["ssd_match_loc_t"], rtol=RTOL, atol=ATOL)
    bu.match_ious(THR, gt[:, :4], pri, VAR, gt[:, -1], loc_s, conf_s, 0)
    assert np.array_equal(conf_s[0].cpu().numpy(), g["ssd_match_ious_conf_t"])
    assert np.array_equal(loc_s[0].cpu().numpy(), g["ssd_match_ious_loc_t"])
    with pytest.raises((IndexError, ValueError)):
        rt.match(THR, gt[:0, :4], pri, VAR, gt[:0, -1], gt[:0, 4:14], loc_c, conf_c, landm_c, 0)


def test_match_640_golden(mods):
    g = load_golden("match_640.npz")
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    for cfg_id, img, key in ((1, 0, "cfg1_"), (2, 6, "cfg2_")):
        gt = mods["synth"].make_gt(cfg_id, img, (640, 640))
        assert np.array_equal(gt.numpy(), g[key + "gt"])
        loc_t, conf_t, landm_t, ex = mods["batched"].assign_targets(pri, [gt.cuda()], threshold=THR, variances=VAR, return_match=True)
        assert np.array_equal(conf_t[0].cpu().numpy(), g[key + "conf_t"])
        assert np.array_equal(ex["best_truth_idx"][0].cpu().numpy(), g[key + "bti"])
        assert np.array_equal(ex["best_truth_overlap"][0].cpu().numpy(), g[key + "bto"])
        assert np.array_equal(ex["best_prior_idx"].cpu().numpy(), g[key + "bpi"])
        assert np.array_equal(ex["best_prior_overlap"].cpu().numpy(), g[key + "bpo"])
        np.testing.assert_allclose(loc_t[0].cpu().numpy(), g[key + "loc_t"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(loc_t[0].cpu().numpy()[:, :2], g[key + "loc_t"][:, :2])


# ---------------------------------------------------------------------------------- batches vs the oracle
@pytest.mark.parametrize("dense", [False, True])
def test_assign_cfg2_batch_vs_oracle(mods, dense):
    """BASELINE configs[1]: batch 32 at 640x640, ragged GT (1..300 faces), full size."""
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    targets = mods["synth"].make_gt_batch(2, 32, (640, 640))
    ref = mods["orc"].match_batch(THR, [t.numpy() for t in targets], pri.cpu().numpy(), VAR)
    out = mods["batched"].assign_targets(pri, [t.cuda() for t in targets], threshold=THR, variances=VAR, return_match=True,
                                         dense=dense)
    check_assign(out[:3], out[3], ref)
    # packed (gt, offsets) form and numpy staging give the same bytes
    gt_packed, offs = mods["synth"].pack_gt(targets)
    out2 = mods["batched"].assign_targets(pri, (gt_packed.numpy(), offs.numpy()), threshold=THR, variances=VAR, dense=dense)
    for a, b in zip(out[:3], out2):
        assert torch.equal(a, b)


def test_assign_variants_vs_oracle(mods):
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet_4, image_size=(333, 517)).get_anchors()   # 4 levels, odd size
    targets = mods["synth"].make_gt_batch(2, 5, (333, 517), first_image=40)
    tn = [t.numpy() for t in targets]
    for label_mode, enc in ((1, 1), (1, 0), (0, 0)):
        ref = mods["orc"].match_batch(THR, tn, pri.cpu().numpy(), VAR, label_mode=label_mode, encode_mode=enc)
        out = mods["batched"].assign_targets(pri, [t.cuda() for t in targets], threshold=THR, variances=VAR,
                                             label_mode=label_mode, encode=bool(enc), return_match=True)
        check_assign(out[:3], out[3], ref)
        if not enc:
            assert np.array_equal(out[0].cpu().numpy(), ref["loc_t"])
    # no landmark output (8-arg forms)
    out = mods["batched"].assign_targets(pri, [t.cuda() for t in targets], with_landm=False)
    assert out[2] is None


def test_assign_edge_cases(mods):
    orc = mods["orc"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(96, 128)).get_anchors()
    pn = pri.cpu().numpy()
    rng = np.random.default_rng(11)

    def rows(boxes):
        G = boxes.shape[0]
        r = np.zeros((G, 15), np.float32)
        r[:, :4] = boxes
        r[:, 4:14] = rng.random((G, 10), dtype=np.float32)
        r[:, 14] = np.where(rng.random(G) < 0.5, 1.0, -1.0)
        return r

    pf = orc.point_form(pn)
    cases = {
        "single": rows(np.array([[0.2, 0.3, 0.45, 0.7]], np.float32)),
        # exact copies of prior boxes: IoU == 1 ties between the two sizes/positions, duplicate GT rows
        "on_priors": rows(np.concatenate([pf[[0, 1, 50, 51, 300]], pf[[50]], pf[[0]]])),
        # every GT outside the image: all-zero IoU rows and columns -> index-0 rule
        "outside": rows(np.array([[1.5, 1.5, 1.7, 1.8], [-0.9, -0.8, -0.5, -0.6], [2.0, 0.1, 2.2, 0.3]], np.float32)),
        # zero-area and one outside among normal ones
        "degenerate": rows(np.array([[0.3, 0.3, 0.3, 0.5], [0.1, 0.1, 0.4, 0.5], [0.6, 0.6, 0.6, 0.6], [1.2, 0.1, 1.4, 0.2]],
                                    np.float32)),
        # many GT sharing one best prior (collision -> the largest j wins)
        "collide": rows(np.tile(np.array([[0.40, 0.40, 0.52, 0.58]], np.float32), (9, 1))
                        + rng.random((9, 4), dtype=np.float32) * 1e-3),
        # inverted box (negative area): generic dense path
        "inverted": rows(np.array([[0.5, 0.2, 0.3, 0.6], [0.1, 0.1, 0.4, 0.5], [0.7, 0.8, 0.6, 0.7]], np.float32)),
        "tiny_crowd": rows(np.concatenate([(c := rng.random((200, 2), dtype=np.float32) * 0.9),
                                           c + 0.01 + 0.05 * rng.random((200, 2), dtype=np.float32)], 1)),
    }
    names = list(cases)
    targets = [cases[k] for k in names]
    ref = orc.match_batch(THR, targets, pn, VAR)
    for dense in (False, True):
        out = mods["batched"].assign_targets(pri, [cuda(t) for t in targets], threshold=THR, variances=VAR, return_match=True,
                                             dense=dense)
        loc_t, conf_t, landm_t, ex = out
        off = 0
        for i, k in enumerate(names):
            G = targets[i].shape[0]
            assert np.array_equal(conf_t[i].cpu().numpy(), ref["conf_t"][i]), (k, dense)
            assert np.array_equal(ex["best_truth_idx"][i].cpu().numpy(), ref["best_truth_idx"][i]), (k, dense)
            assert np.array_equal(ex["best_prior_idx"][off:off + G].cpu().numpy(), ref["best_prior_idx"][off:off + G]), (k, dense)
            np.testing.assert_array_equal(ex["best_truth_overlap"][i].cpu().numpy(), ref["best_truth_overlap"][i], err_msg=k)
            np.testing.assert_array_equal(ex["best_prior_overlap"][off:off + G].cpu().numpy(),
                                          ref["best_prior_overlap"][off:off + G], err_msg=k)
            assert np.array_equal(landm_t[i].cpu().numpy(), ref["landm_t"][i]), (k, dense)
            with np.errstate(invalid="ignore"):
                np.testing.assert_allclose(loc_t[i].cpu().numpy(), ref["loc_t"][i], rtol=RTOL, atol=ATOL, err_msg=k)
            off += G
    # empty image: the reference raises; allow_empty gives all-background targets for that image
    with pytest.raises(ValueError):
        mods["batched"].assign_targets(pri, [cuda(targets[0]), cuda(targets[0][:0])])
    out = mods["batched"].assign_targets(pri, [cuda(targets[0]), cuda(targets[0][:0]), cuda(targets[1])], allow_empty=True)
    assert out[1][1].abs().sum().item() == 0 and out[0][1].abs().sum().item() == 0 and out[2][1].abs().sum().item() == 0
    assert np.array_equal(out[1][0].cpu().numpy(), ref["conf_t"][0]) and np.array_equal(out[1][2].cpu().numpy(), ref["conf_t"][1])


def test_assign_encode_kernel_paths(mods):
    """The pipelined encode kernel (256 priors x 8 images per CTA) off its main road: an odd prior count (rows of odd images
    are only 8-byte aligned: scalar landmark stores), 11 images (a full group of 8 and a group of 3, an image without GT in
    each), priors and GT rows that fail the once-per-prior / once-per-row range test (generic divide), every kernel variant
    (landmarks on/off, encode on/off, extra outputs on/off)."""
    orc = mods["orc"]
    pri_full = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(160, 192)).get_anchors()
    pn = pri_full.cpu().numpy()[:1261].copy()                 # odd P, last tile partial
    pn[7] = [1e-13, 0.5, 0.1, 0.1]                            # |cx| < 2^-36: prior fails the range test
    pn[300] = [0.5, 0.5, 1e-11, 0.2]                          # var0*w below 2^-36 (area still within the match kernel's bounds)
    pri = torch.from_numpy(pn).cuda()
    rng = np.random.default_rng(5)
    targets = []
    for i in range(11):
        G = [3, 0, 40, 1, 70, 5, 2, 9, 0, 130, 4][i]
        c = rng.random((G, 2), dtype=np.float32) * 0.8
        wh = 0.02 + 0.2 * rng.random((G, 2), dtype=np.float32)
        r = np.zeros((G, 15), np.float32)
        r[:, :2], r[:, 2:4] = c, c + wh
        r[:, 4:14] = rng.random((G, 10), dtype=np.float32)
        r[:, 14] = np.where(rng.random(G) < 0.5, 1.0, -1.0)
        if G >= 40:
            r[1, 4] = 3e18                                    # a landmark beyond 2^59: the row fails the range test
            r[2, 2] = r[2, 0]                                 # zero width: log(0) = -inf like the reference
        targets.append(r)
    nonempty = [t for t in targets if len(t)]
    ref_all = orc.match_batch(THR, nonempty, pn, VAR)
    b = mods["batched"]
    for kw in (dict(), dict(return_match=True), dict(encode=False), dict(encode=False, return_match=True)):
        out = b.assign_targets(pri, [cuda(t) for t in targets], threshold=THR, variances=VAR, allow_empty=True, **kw)
        loc_t, conf_t, landm_t = out[0].cpu().numpy(), out[1].cpu().numpy(), out[2].cpu().numpy()
        ref = ref_all if kw.get("encode", True) else orc.match_batch(THR, nonempty, pn, VAR, encode_mode=0)
        j = 0
        for i, t in enumerate(targets):
            if len(t) == 0:
                assert not conf_t[i].any() and not loc_t[i].any() and not landm_t[i].any(), i
                continue
            assert np.array_equal(conf_t[i], ref["conf_t"][j]), (i, kw)
            with np.errstate(invalid="ignore"):
                np.testing.assert_array_equal(landm_t[i], ref["landm_t"][j], err_msg=str((i, kw)))
                np.testing.assert_array_equal(loc_t[i][:, :2], ref["loc_t"][j][:, :2], err_msg=str((i, kw)))
                np.testing.assert_allclose(loc_t[i], ref["loc_t"][j], rtol=RTOL, atol=ATOL, err_msg=str((i, kw)))
            if kw.get("return_match"):
                assert np.array_equal(out[3]["best_truth_idx"][i].cpu().numpy(), ref["best_truth_idx"][j]), (i, kw)
            j += 1


def test_assign_cfg4_dense_tiny_faces(mods):
    """BASELINE configs[3]: 2048x2048 (172,032 priors), 1,500 GT per image (3 GT chunks per CTA)."""
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(2048, 2048)).get_anchors()
    assert pri.shape[0] == 172032
    targets = mods["synth"].make_gt_batch(4, 2, (2048, 2048))
    assert targets[0].shape[0] == 1500
    ref = mods["orc"].match_batch(THR, [t.numpy() for t in targets], pri.cpu().numpy(), VAR)
    out = mods["batched"].assign_targets(pri, [t.cuda() for t in targets], threshold=THR, variances=VAR, return_match=True)
    check_assign(out[:3], out[3], ref)
    out_d = mods["batched"].assign_targets(pri, [t.cuda() for t in targets], threshold=THR, variances=VAR, return_match=True,
                                           dense=True)
    for a, b in zip(out[:3], out_d[:3]):
        assert torch.equal(a, b)
    for k in out[3]:
        assert torch.equal(out[3][k], out_d[3][k]), k


def test_assign_full_size_properties(mods):
    """cfg5-sized batch (256 images at 640x640): size-independent properties instead of a full oracle run."""
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    targets = [t.cuda() for t in mods["synth"].make_gt_batch(5, 256, (640, 640))]
    loc_t, conf_t, landm_t, ex = mods["batched"].assign_targets(pri, targets, return_match=True)
    # (1) culled == dense, bit for bit
    d = mods["batched"].assign_targets(pri, targets, return_match=True, dense=True)
    assert torch.equal(loc_t, d[0]) and torch.equal(conf_t, d[1]) and torch.equal(landm_t, d[2])
    # (2) shard invariance: an image's targets do not depend on the batch it is processed in
    for lo, hi in ((0, 1), (17, 49), (255, 256)):
        s = mods["batched"].assign_targets(pri, targets[lo:hi])
        assert torch.equal(s[0], loc_t[lo:hi]) and torch.equal(s[1], conf_t[lo:hi]) and torch.equal(s[2], landm_t[lo:hi])
    # (3) every GT owns its best prior after the force-match override, with overlap 2
    offs = np.cumsum([0] + [int(t.shape[0]) for t in targets])
    bpi = ex["best_prior_idx"].long()
    for b in (0, 100, 255):
        g = torch.arange(offs[b + 1] - offs[b], device="cuda")
        p = bpi[offs[b]:offs[b + 1]]
        # the largest j wins on collisions
        winners = torch.full((pri.shape[0],), -1, dtype=torch.long, device="cuda").scatter_reduce(0, p, g, "amax")
        assert torch.equal(ex["best_truth_idx"][b].long()[p], winners[p])
        assert (ex["best_truth_overlap"][b][p] == 2.0).all()
    # (4) labels are in {-1,0,1}; positives are exactly the priors with overlap >= threshold
    assert set(torch.unique(conf_t).tolist()) <= {-1, 0, 1}
    assert torch.equal(conf_t != 0, ex["best_truth_overlap"] >= np.float32(THR))
    # (5) idempotence
    again = mods["batched"].assign_targets(pri, targets)
    assert torch.equal(again[0], loc_t) and torch.equal(again[1], conf_t)


def test_assign_host_entry(mods):
    """jabd_assign_host: host buffers in and out (the e2e path of bench.py) equals the device path."""
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    targets = mods["synth"].make_gt_batch(2, 4, (640, 640))
    dev = mods["batched"].assign_targets(pri, [t.cuda() for t in targets])
    h = mods["batched"].HostAssign(pri, 4, 2000)
    loc_t, conf_t, landm_t = h(targets)
    assert not loc_t.is_cuda and torch.equal(loc_t, dev[0].cpu()) and torch.equal(conf_t, dev[1].cpu())
    assert torch.equal(landm_t, dev[2].cpu())
    assert h.last_d2h == 4 * 16800 * 64
    # JABD_ASSIGN_DEVICE_OUT: same call, targets written to device tensors, pipelined over two slots
    hd = mods["batched"].HostAssign(pri, 4, 2000, device_out=True)
    slots = [hd.submit(targets), hd.submit(targets[::-1])]
    a, b = hd.wait(slots[0]), hd.wait(slots[1])
    assert a[0].is_cuda and hd.last_d2h == 0
    assert torch.equal(a[0], dev[0]) and torch.equal(a[1], dev[1]) and torch.equal(a[2], dev[2])
    assert torch.equal(b[1], dev[1].flip(0)) and torch.equal(b[0], dev[0].flip(0))


def test_assign_batches_lanes(mods):
    """jabd_assign_batches: several independent batches on side streams (0, 1, 3, 4 lanes; eagerly and replayed from a CUDA
    graph captured through the caller's stream) give, batch by batch, the bytes of one jabd_assign call each; the first batch
    is also checked against the oracle."""
    import ctypes
    from jabd_b200 import _lib
    bt = mods["batched"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    sizes = (5, 3, 8, 1, 6)
    batches, first = [], 0
    for n in sizes:
        batches.append([t.cuda() for t in mods["synth"].make_gt_batch(2, n, (640, 640), first_image=first)])
        first += n
    want = [bt.assign_targets(pri, b, threshold=THR, variances=VAR) for b in batches]
    ref = mods["orc"].match_batch(THR, [t.cpu().numpy() for t in batches[0]], pri.cpu().numpy(), VAR)
    check_assign(want[0], None, ref)

    def same(outs):
        torch.cuda.synchronize()
        return all(torch.equal(a, b) for o, w in zip(outs, want) for a, b in zip(o, w))

    def scrub(outs):
        for o in outs:
            for t in o:
                t.fill_(7)

    for n_lanes in (0, 1, 3, 4, 9):
        plan = bt.AssignBatches(pri, batches, threshold=THR, variances=VAR, lanes_n=n_lanes)
        assert len(plan.lane_streams) == min(n_lanes, len(sizes))
        outs = plan()
        assert same(outs), n_lanes
        scrub(outs)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):                       # a non-default calling stream; later work on it sees the results
            outs = plan()
            sums = [o[1].sum() for o in outs]
        s.synchronize()
        assert [int(x) for x in sums] == [int(w[1].sum()) for w in want]
        assert same(outs)
        scrub(outs)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan()
        scrub(outs)
        g.replay()
        assert same(outs), ("graph", n_lanes)
    assert same(bt.assign_batches(pri, batches[:2], threshold=THR, variances=VAR, lanes_n=2))
    # refused before anything is enqueued: a lane equal to the calling stream, two batches on one workspace
    plan = bt.AssignBatches(pri, batches[:2], lanes_n=1)
    L = _lib.lib()
    cur = torch.cuda.current_stream().cuda_stream
    lane = (ctypes.c_void_p * 1)(cur)
    args = (pri.data_ptr(), plan.P, ctypes.cast(plan.arr, ctypes.c_void_p), 2, THR, 0.1, 0.2, 0, 1, 0)
    assert L.jabd_assign_batches(*args, ctypes.cast(lane, ctypes.c_void_p), 1, ctypes.c_void_p(cur)) == -1
    assert "calling stream" in _lib.last_error()
    plan.arr[1].workspace = plan.arr[0].workspace                     # fine on one stream (stream order), a race on two lanes
    two = (ctypes.c_void_p * 2)(*[x.cuda_stream for x in bt.lanes(pri.device, 2)])
    assert L.jabd_assign_batches(*args, ctypes.cast(two, ctypes.c_void_p), 2, ctypes.c_void_p(cur)) == -1
    assert "share a workspace" in _lib.last_error()


def test_assign_work_list_shapes(mods):
    """JABD_ASSIGN_TUNE: the matching kernel's work list cut into 16..192 GT per item, in one or two regimes over the tiles --
    every shape gives the bytes of the default one (checked against the oracle); bad shapes are refused."""
    bt = mods["batched"]
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(640, 640)).get_anchors()
    targets = mods["synth"].make_gt_batch(2, 9, (640, 640), first_image=64)
    targets[4] = torch.cat([targets[4]] * 3)[:500]                                # > 2 x 192 GT in one image, duplicates (ties)
    tg = [t.cuda() for t in targets]
    ref = mods["orc"].match_batch(THR, [t.numpy() for t in targets], pri.cpu().numpy(), VAR)
    want = bt.assign_targets(pri, tg, threshold=THR, variances=VAR, return_match=True)
    check_assign(want[:3], want[3], ref)
    for tune in ((192, 192, 100), (16, 16, 100), (128, 32, 70), (32, 192, 50), (64, 17, 0), (191, 64, 1), (100, 100, 99)):
        for dense in (False, True):
            got = bt.assign_targets(pri, tg, threshold=THR, variances=VAR, return_match=True, tune=tune, dense=dense)
            assert all(torch.equal(a, b) for a, b in zip(got[:3], want[:3])), tune
            assert all(torch.equal(got[3][k], want[3][k]) for k in want[3]), tune
    for bad in ((8, 64, 100), (64, 200, 100), (64, 64, 101)):
        with pytest.raises(ValueError):
            bt.assign_targets(pri, tg, tune=bad)
    # the lanes of jabd_assign_batches pick the large-item shape themselves: same bytes
    outs = bt.assign_batches(pri, [tg, tg[:3]], threshold=THR, variances=VAR, lanes_n=2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(outs[0], want[:3]))


def test_errors_are_loud(mods):
    from jabd_b200 import _lib
    pri = mods["anchors"].Anchors(mods["cfgs"].cfg_mnet, image_size=(96, 128)).get_anchors()
    with pytest.raises(ValueError):
        mods["batched"].assign_targets(pri[:, :3], [torch.zeros(1, 15).cuda()])
    with pytest.raises(ValueError):
        mods["batched"].assign_targets(pri, [torch.zeros(1, 14).cuda()])
    # workspace too small at the C-ABI
    L = _lib.lib()
    rc = L.jabd_assign_match(pri.data_ptr(), pri.shape[0], pri.data_ptr(), pri.data_ptr(), 1, 1, 0, pri.data_ptr(), 16, None)
    assert rc == -3 and "workspace" in _lib.last_error()


def test_shared_reciprocal_division_is_ieee(mods):
    """fdiv_shared() (one refined reciprocal per divisor) must equal the compiler's div.rn bit for bit:
    2^30 pseudo-random operand pairs incl. zeros, subnormals, infinities and NaNs in the numerator."""
    from jabd_b200 import _lib
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    bad = torch.zeros(4, dtype=torch.float32, device="cuda")
    for seed in (1, 0x9E3779B97F4A7C15):
        _lib.selftest_call("jabd_selftest_div", 1 << 30, seed, out.data_ptr(), bad.data_ptr(), None)
        torch.cuda.synchronize()
        assert int(out[0].item()) == 0, "mismatch (a, d, got, expected) = %s" % (bad.cpu().tolist(),)
