"""GPU parity of the IoU-family row against golden vectors recorded from the reference
(R/utils/box_utils.py:5-158, R/nets/retinaface_training_DIOU.py:491-665, R/utils/utils_bbox.py:182-258) and the oracle.

bit-exact: IoU / GIoU / DIoU values (only IEEE +,-,*,/,min,max), hard-negative selection, DIoU-NMS keep lists.
rtol 1e-5 / atol 1e-6: CIoU (atan), loss sums (fp32 summation order).  rtol 1e-4 / atol 1e-6: gradients (the reference's
come from autograd in fp32; same formulas, different association)."""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL, load_golden

pytestmark = pytest.mark.gpu

VAR = [0.1, 0.2]
G_RTOL, G_ATOL = 1e-4, 1e-6


@pytest.fixture(scope="module")
def env():
    from jabd_b200 import box_utils, retinaface_training_DIOU as diou, synth, utils_bbox, anchors, config
    from oracle import oracle as orc
    return dict(bu=box_utils, diou=diou, synth=synth, ub=utils_bbox, anchors=anchors, cfg=config, orc=orc, g=load_golden("iou_family.npz"))


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_bbox_overlaps_family_golden(env):
    bu, g = env["bu"], env["g"]
    a, b = cuda(g["a"]), cuda(g["b"])
    for kind in ("iou", "giou", "diou"):
        out = getattr(bu, "bbox_overlaps_" + kind)(a, b).cpu().numpy()
        np.testing.assert_array_equal(out, g["ov_" + kind])
        assert np.array_equal(getattr(env["diou"], "bbox_overlaps_" + kind)(a, b).cpu().numpy(), out)
    c = bu.bbox_overlaps_ciou(a, b).cpu().numpy()
    assert np.array_equal(np.isnan(c), np.isnan(g["ov_ciou"]))                  # identical boxes: alpha = 0/0
    ok = ~np.isnan(c)
    np.testing.assert_allclose(c[ok], g["ov_ciou"][ok], rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(bu.bbox_overlaps_diou(a[:1], b).cpu().numpy(), g["ov_diou_bcast"])
    assert tuple(bu.bbox_overlaps_iou(a[:0], b).shape) == (0, b.shape[0])       # rows * cols == 0 (:9-10)
    assert isinstance(bu.bbox_overlaps_giou(g["a"], g["b"]), np.ndarray)        # numpy in -> numpy out


def test_iou_loss_forward_backward_golden(env):
    diou, g = env["diou"], env["g"]
    b, pri = cuda(g["b"]), cuda(g["loss_pri"])
    for lt in ("Iou", "Giou", "Diou", "Ciou"):
        for size_sum in (True, False):
            tag = "%s_%d" % (lt, int(size_sum))
            lp = cuda(g["loss_loc"]).requires_grad_(True)
            val = diou.IouLoss(pred_mode='Center', size_sum=size_sum, variances=VAR, losstype=lt)(lp, b, pri)
            (3.0 * val).backward()
            np.testing.assert_allclose(val.item(), g["loss_" + tag], rtol=RTOL)
            np.testing.assert_allclose(lp.grad.cpu().numpy() / 3.0, g["grad_" + tag], rtol=G_RTOL, atol=G_ATOL * (1 if size_sum else 1e-2))
    lp = cuda(g["a"]).requires_grad_(True)
    val = diou.IouLoss(pred_mode='Corner', size_sum=True, variances=VAR, losstype='Giou')(lp, b, pri)
    val.backward()
    np.testing.assert_allclose(val.item(), g["loss_corner_giou"], rtol=RTOL)
    np.testing.assert_allclose(lp.grad.cpu().numpy(), g["grad_corner_giou"], rtol=G_RTOL, atol=1e-4)


@pytest.mark.parametrize("tag,size,batch,count", [("s160", (160, 160), 3, None), ("s320", (320, 320), 2, 40)])
def test_diou_multibox_loss_golden(env, tag, size, batch, count):
    """MultiBoxLoss of R/nets/retinaface_training_DIOU.py:527-665 (cuda=False run of the reference): losses, the set of
    priors entering loss_c, and the box gradients."""
    diou, synth, g = env["diou"], env["synth"], env["g"]
    pri = env["anchors"].Anchors(env["cfg"].cfg_mnet, image_size=size).get_anchors()
    P = pri.shape[0]
    targets = [synth.make_gt(6, i, size, count=count).cuda() for i in range(batch)]
    raw = [synth.make_logits(6, i, P) for i in range(batch)]
    preds = tuple(torch.stack([r[k] for r in raw]).cuda().requires_grad_(True) for k in range(3))
    crit = diou.MultiBoxLoss(2, 0.35, 7, VAR, True)
    l, c, m = crit(preds, pri, targets)
    (1.0 * l + 2.0 * c + 0.5 * m).backward()
    np.testing.assert_allclose(np.array([l.item(), c.item(), m.item()], np.float32), g["mbl_%s_losses" % tag], rtol=RTOL)
    g_loc, g_conf = preds[0].grad.cpu().numpy(), preds[1].grad.cpu().numpy()
    np.testing.assert_array_equal(np.packbits(np.abs(g_conf).sum(2) != 0), g["mbl_%s_sel" % tag])
    nz = g["mbl_%s_g_loc_nz_idx" % tag]
    np.testing.assert_allclose(g_loc.reshape(-1, 4)[nz], g["mbl_%s_g_loc_nz" % tag], rtol=G_RTOL, atol=G_ATOL)
    rest = np.ones(g_loc.shape[0] * g_loc.shape[1], bool)
    rest[nz] = False
    assert not g_loc.reshape(-1, 4)[rest].any()


def test_diounms_golden_and_oracle(env):
    ub, orc, g = env["ub"], env["orc"], env["g"]
    nms = load_golden("nms.npz")
    for name in ("rand500", "dense2000"):
        b, s = nms[name + "_boxes"], nms[name + "_scores"]
        for (ov, tk, beta) in ((0.5, 200, 1.0), (0.3, 5000, 1.0), (0.45, 5000, 0.6)):
            keep, count = ub.diounms(cuda(b), cuda(s), ov, tk, beta)
            ref = g["dnms_%s_%d_%d_%d_keep" % (name, int(ov * 100), tk, int(beta * 10))]
            assert keep.dtype == torch.int64 and keep.shape[0] == b.shape[0]
            assert count == len(ref) and np.array_equal(keep[:count].cpu().numpy(), ref), (name, ov, tk, beta)
            assert not keep[count:].any()
    rng = np.random.default_rng(9)                                              # larger, tied scores: against the oracle
    n = 7000
    c = rng.random((n, 2), dtype=np.float32)
    wh = 0.01 + 0.05 * rng.random((n, 2), dtype=np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(n, dtype=np.float32)
    s[::9] = s[4]
    rk, rc = orc.diounms(b, s, 0.4, n, 1.0)
    keep, count = ub.diounms(cuda(b), cuda(s), 0.4, n, 1.0)
    assert count == rc and np.array_equal(keep[:count].cpu().numpy(), rk[:rc])
    e = ub.diounms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda())
    assert isinstance(e, torch.Tensor) and e.numel() == 0
