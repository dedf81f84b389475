"""Torch-CPU port of the reference's eager code path -- TEST INFRASTRUCTURE ONLY.

This is the timed CPU baseline (``cpu_baseline.kind == "port"``) and the
``bench.py --impl reference`` arm: the reference tree is Python and cannot
travel to the GPU box, so its eager op sequence is restated here with the same
dense temporaries (the [G,P,2] corner tensors, the [G,P] IoU matrix, the two
``max`` passes, the per-GT Python loop with one tensor index per iteration),
the same per-image Python loop, and the same third-party NMS
(``torchvision.ops.nms``, R/utils/utils_bbox.py:275-279), so that its cost on
the host cores is the cost the reference pays with ``Cuda=False``.

Checked bit-for-bit against the imported reference by tests/golden/make_golden.py
(--check) in the dev container and against tests/golden/*.npz everywhere.
"""
import torch

try:  # third-party NMS used by the reference (not vendored there, unpinned)
    from torchvision.ops import nms as _tv_nms
except Exception:  # pragma: no cover
    _tv_nms = None


def corners(p):
    """(cx,cy,w,h) -> (x1,y1,x2,y2); R/nets/retinaface_training.py:8-10."""
    half = p[:, 2:] / 2
    return torch.cat((p[:, :2] - half, p[:, :2] + half), 1)


def overlap_matrix(a, b):
    """Dense IoU [A,B]; R/nets/retinaface_training.py:22-59 (materialises the
    [A,B,2] corner temporaries like the reference's expand+min/max)."""
    na, nb = a.size(0), b.size(0)
    hi = torch.min(a[:, None, 2:].expand(na, nb, 2), b[None, :, 2:].expand(na, nb, 2))
    lo = torch.max(a[:, None, :2].expand(na, nb, 2), b[None, :, :2].expand(na, nb, 2))
    d = torch.clamp(hi - lo, min=0)
    inter = d[:, :, 0] * d[:, :, 1]
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None].expand_as(inter)
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :].expand_as(inter)
    return inter / (area_a + area_b - inter)


def encode_boxes(m, p, var):
    """R/nets/retinaface_training.py:61-70."""
    c = (m[:, :2] + m[:, 2:]) / 2 - p[:, :2]
    c /= (var[0] * p[:, 2:])
    s = torch.log((m[:, 2:] - m[:, :2]) / p[:, 2:]) / var[1]
    return torch.cat([c, s], 1)


def encode_points(m, p, var):
    """R/nets/retinaface_training.py:72-84 (builds the [P,5,4] prior temporary)."""
    n = m.size(0)
    pts = m.reshape(n, 5, 2)
    pp = torch.cat([p[:, k][:, None].expand(n, 5)[:, :, None] for k in range(4)], dim=2)
    c = pts - pp[:, :, :2]
    c /= (var[0] * pp[:, :, 2:])
    return c.reshape(n, -1)


def assign_one(thr, truths, priors, var, labels, landms, loc_t, conf_t, landm_t, idx,
               label_mode=0, encode_mode=1):
    """One image of ``match``; R/nets/retinaface_training.py:93-162."""
    ov = overlap_matrix(truths, corners(priors))
    _, bp_idx = ov.max(1, keepdim=True)
    bp_idx.squeeze_(1)
    bt_ov, bt_idx = ov.max(0, keepdim=True)
    bt_idx.squeeze_(0)
    bt_ov.squeeze_(0)
    bt_ov.index_fill_(0, bp_idx, 2)
    for j in range(bp_idx.size(0)):           # sequential, last j wins (:129-130)
        bt_idx[bp_idx[j]] = j
    m = truths[bt_idx]
    c = labels[bt_idx]
    if label_mode:
        c = c + 1                             # R/utils/box_utils.py:315
    c[bt_ov < thr] = 0
    loc_t[idx] = encode_boxes(m, priors, var) if encode_mode else m
    conf_t[idx] = c
    if landm_t is not None:
        landm_t[idx] = encode_points(landms[bt_idx], priors, var)
    return bt_idx, bt_ov, bp_idx


def assign_batch(thr, targets, priors, var, encode=True):
    """The loop of MultiBoxLoss.forward, R/nets/retinaface_training.py:197-214 (``encode=False``: match_iou,
    R/nets/retinaface_training_DIOU.py:176-246)."""
    n, P = len(targets), priors.size(0)
    loc_t = torch.Tensor(n, P, 4)
    landm_t = torch.Tensor(n, P, 10)
    conf_t = torch.LongTensor(n, P)
    for i in range(n):
        t = targets[i]
        assign_one(thr, t[:, :4], priors, var, t[:, -1], t[:, 4:14], loc_t, conf_t, landm_t, i, 0, 1 if encode else 0)
    return loc_t, conf_t, landm_t


def global_lse(x):
    """log_sum_exp with the maximum of the WHOLE tensor; R/nets/retinaface_training.py:86-88."""
    m = x.data.max()
    return torch.log(torch.sum(torch.exp(x - m), 1, keepdim=True)) + m


def overlaps_family(b1, b2, kind):
    """bbox_overlaps_{iou,giou,diou,ciou}, R/utils/box_utils.py:5-158 (= R/nets/retinaface_training_DIOU.py:342-490):
    row i of b1 against row i of b2; differentiable like the reference."""
    import math
    if b1.shape[0] * b2.shape[0] == 0:
        return torch.zeros((b1.shape[0], b2.shape[0]))
    if b1.shape[0] > b2.shape[0]:
        b1, b2 = b2, b1
    wh1, wh2 = b1[:, 2:] - b1[:, :2], b2[:, 2:] - b2[:, :2]
    a1, a2 = wh1[:, 0] * wh1[:, 1], wh2[:, 0] * wh2[:, 1]
    isz = torch.clamp(torch.min(b1[:, 2:], b2[:, 2:]) - torch.max(b1[:, :2], b2[:, :2]), min=0)
    inter = isz[:, 0] * isz[:, 1]
    union = a1 + a2 - inter
    if kind == "iou":
        return torch.clamp(inter / union, min=0, max=1.0)
    osz = torch.clamp(torch.max(b1[:, 2:], b2[:, 2:]) - torch.min(b1[:, :2], b2[:, :2]), min=0)
    if kind == "giou":
        closure = osz[:, 0] * osz[:, 1]
        return torch.clamp(inter / union - (closure - union) / closure, min=-1.0, max=1.0)
    c1, c2 = (b1[:, 2:] + b1[:, :2]) / 2, (b2[:, 2:] + b2[:, :2]) / 2
    centre = (c2[:, 0] - c1[:, 0]) ** 2 + (c2[:, 1] - c1[:, 1]) ** 2
    diag = osz[:, 0] ** 2 + osz[:, 1] ** 2
    if kind == "diou":
        return torch.clamp(inter / union - centre / diag, min=-1.0, max=1.0)
    u, iou = centre / diag, inter / union
    v = (4 / (math.pi ** 2)) * torch.pow(torch.atan(wh2[:, 0] / wh2[:, 1]) - torch.atan(wh1[:, 0] / wh1[:, 1]), 2)
    with torch.no_grad():
        alpha = v / ((1 - iou) + v)
    return torch.clamp(iou - (u + alpha * v), min=-1.0, max=1.0)


def iou_loss(loc_p, loc_t, prior_data, var, kind, center=True, size_sum=True):
    """IouLoss.forward, R/nets/retinaface_training_DIOU.py:491-525."""
    boxes = decode_boxes(loc_p, prior_data, var) if center else loc_p
    loss = torch.sum(1.0 - overlaps_family(boxes, loc_t, kind))
    return loss if size_sum else loss / loc_p.shape[0]


def multibox_loss(predictions, priors, targets, thr=0.35, var=(0.1, 0.2), negpos_ratio=7, return_aux=False, loc_loss=None):
    """MultiBoxLoss.forward with cuda=False; R/nets/retinaface_training.py:183-303.  ``predictions`` =
    (loc_data [B,P,4], conf_data [B,P,2], landm_data [B,P,10]) CPU tensors (may require grad).  ``loc_loss`` "iou" / "giou" /
    "diou" / "ciou": the variant of R/nets/retinaface_training_DIOU.py:527-665 (match_iou targets, IouLoss box term)."""
    import torch.nn.functional as F
    loc_data, conf_data, landm_data = predictions
    n = loc_data.size(0)
    loc_t, conf_t, landm_t = assign_batch(thr, [t.data for t in targets], priors.data, list(var), encode=loc_loss is None)
    zero = torch.tensor(0)
    lm_pos = conf_t > zero                                           # :243
    lm_sel = lm_pos.unsqueeze(lm_pos.dim()).expand_as(landm_data)
    loss_landm = F.smooth_l1_loss(landm_data[lm_sel].view(-1, 10), landm_t[lm_sel].view(-1, 10), reduction='sum')
    pos = conf_t != zero                                             # :250
    box_sel = pos.unsqueeze(pos.dim()).expand_as(loc_data)
    if loc_loss is None:
        loss_l = F.smooth_l1_loss(loc_data[box_sel].view(-1, 4), loc_t[box_sel].view(-1, 4), reduction='sum')
    else:                                                            # DIOU.py:606-607
        pri_sel = priors.data.unsqueeze(0).expand_as(loc_data)[box_sel].view(-1, 4)
        loss_l = iou_loss(loc_data[box_sel].view(-1, 4), loc_t[box_sel].view(-1, 4), pri_sel, var, loc_loss)
    conf_t[pos] = 1                                                  # :259
    flat = conf_data.view(-1, 2)
    rank_val = global_lse(flat) - flat.gather(1, conf_t.view(-1, 1)) # :265
    rank_val[pos.view(-1, 1)] = 0
    rank_val = rank_val.view(n, -1)
    _, order = rank_val.sort(1, descending=True)                     # :270-271
    _, rank = order.sort(1)
    num_pos = pos.long().sum(1, keepdim=True)
    num_neg = torch.clamp(negpos_ratio * num_pos, max=pos.size(1) - 1)
    neg = rank < num_neg.expand_as(rank)
    both = (pos.unsqueeze(2).expand_as(conf_data) + neg.unsqueeze(2).expand_as(conf_data)).gt(0)
    loss_c = F.cross_entropy(conf_data[both].view(-1, 2), conf_t[(pos + neg).gt(0)], reduction='sum')
    N = max(num_pos.data.sum().float(), 1)
    loss_l = loss_l / N
    loss_c = loss_c / N
    N1 = max(lm_pos.long().sum(1, keepdim=True).data.sum().float(), 1)
    loss_landm = loss_landm / N1
    if return_aux:
        return loss_l, loss_c, loss_landm, dict(pos=pos, pos1=lm_pos, neg=neg, N=float(N), N1=float(N1), rank_val=rank_val.detach())
    return loss_l, loss_c, loss_landm


def decode_boxes(loc, p, var):
    """R/utils/utils_bbox.py:29-34."""
    b = torch.cat((p[:, :2] + loc[:, :2] * var[0] * p[:, 2:],
                   p[:, 2:] * torch.exp(loc[:, 2:] * var[1])), 1)
    b[:, :2] -= b[:, 2:] / 2
    b[:, 2:] += b[:, :2]
    return b


def decode_points(pre, p, var):
    """R/utils/utils_bbox.py:39-46."""
    return torch.cat([p[:, :2] + pre[:, 2 * k:2 * k + 2] * var[0] * p[:, 2:] for k in range(5)], dim=1)


def suppress(detection, conf_thres=0.5, nms_thres=0.3):
    """``non_max_suppression``; R/utils/utils_bbox.py:260-296."""
    detection = detection[detection[:, 4] >= conf_thres]
    if len(detection) <= 0:
        return []
    keep = _tv_nms(detection[:, :4], detection[:, 4], nms_thres)
    return detection[keep].cpu().numpy()


def infer_one(loc, conf, landm, priors, var, conf_thres=0.5, nms_thres=0.3):
    """Per-image post-processing of Retinaface.detect_image, R/predict.py:167-181."""
    boxes = decode_boxes(loc, priors, var)
    score = conf[:, 1:2]
    pts = decode_points(landm, priors, var)
    return suppress(torch.cat([boxes, score, pts], -1), conf_thres, nms_thres)


def infer_one_topk(loc, conf, landm, priors, var, conf_thres=0.02, pre_nms_topk=5000, nms_thres=0.4,
                   keep_topk=750):
    """cfg3's composed pipeline: decode -> score > thr -> stable
    descending sort[:topk] -> torchvision nms -> [:keep_topk].  Returns
    (dets [K,15], prior indices [K])."""
    boxes = decode_boxes(loc, priors, var)
    pts = decode_points(landm, priors, var)
    s = conf[:, 1]
    idx = torch.nonzero(s > conf_thres).squeeze(1)
    order = torch.sort(s[idx], stable=True, descending=True)[1]
    if pre_nms_topk > 0:
        order = order[:pre_nms_topk]
    idx = idx[order]
    keep = _tv_nms(boxes[idx], s[idx], nms_thres)
    if keep_topk > 0:
        keep = keep[:keep_topk]
    idx = idx[keep]
    return torch.cat([boxes[idx], s[idx, None], pts[idx]], 1), idx
