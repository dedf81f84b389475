"""Drop-in for the SSD-legacy helpers of ``utils/box_utils.py`` (R/utils/box_utils.py:160-448).

Nobody imports that file in the reference, but its signatures are part of the public surface
the north star names, so they are kept: 8-argument ``match`` / ``match_ious`` (labels + 1, no landmarks),
``encode`` / ``decode``, the pure-torch greedy ``nms`` (ascending sort, ``top_k``, zero-padded keep + count) and the
element-wise ``bbox_overlaps_{iou,giou,diou,ciou}`` (:5-158).
"""
import torch

from . import _ops, _tensor

__all__ = ["point_form", "intersect", "jaccard", "match", "match_ious", "encode", "decode", "nms", "bbox_overlaps_iou",
           "bbox_overlaps_giou", "bbox_overlaps_diou", "bbox_overlaps_ciou"]


def bbox_overlaps_iou(bboxes1, bboxes2):
    """R/utils/box_utils.py:96-119: IoU of row i with row i, clamped to [0, 1]."""
    return _ops.overlaps_family(bboxes1, bboxes2, _ops.IOU)


def bbox_overlaps_giou(bboxes1, bboxes2):
    """R/utils/box_utils.py:121-158: IoU - (closure - union) / closure, clamped to [-1, 1]."""
    return _ops.overlaps_family(bboxes1, bboxes2, _ops.GIOU)


def bbox_overlaps_diou(bboxes1, bboxes2):
    """R/utils/box_utils.py:5-46: IoU - centre distance^2 / enclosing diagonal^2, clamped to [-1, 1]."""
    return _ops.overlaps_family(bboxes1, bboxes2, _ops.DIOU)


def bbox_overlaps_ciou(bboxes1, bboxes2):
    """R/utils/box_utils.py:48-94: DIoU with the aspect-ratio term alpha * v."""
    return _ops.overlaps_family(bboxes1, bboxes2, _ops.CIOU)


def point_form(boxes):
    """R/utils/box_utils.py:160-170."""
    return _ops.point_form(boxes)


def intersect(box_a, box_b):
    """R/utils/box_utils.py:185-205."""
    return _ops.intersect(box_a, box_b)


def jaccard(box_a, box_b):
    """R/utils/box_utils.py:208-226."""
    return _ops.jaccard(box_a, box_b)


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """R/utils/box_utils.py:276-320: conf = labels[best_truth_idx] + 1, background 0, encoded loc."""
    _ops.match_one(threshold, truths, priors, variances, labels, None, loc_t, conf_t, None, idx, 1, 1)


def match_ious(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """R/utils/box_utils.py:229-273: as ``match`` but loc_t[idx] holds the raw matched x1y1x2y2."""
    _ops.match_one(threshold, truths, priors, variances, labels, None, loc_t, conf_t, None, idx, 1, 0)


def encode(matched, priors, variances):
    """R/utils/box_utils.py:323-344."""
    return _ops.encode(matched, priors, variances)


def decode(loc, priors, variances):
    """R/utils/box_utils.py:348-367."""
    return _ops.decode(loc, priors, variances)


def nms(boxes, scores, overlap=0.5, top_k=200):
    """R/utils/box_utils.py:384-448.  Returns ``(keep, count)``: ``keep`` is an int64 tensor of
    ``scores.size(0)`` entries, the first ``count`` hold the kept indices (highest score first), the rest 0.
    Like the reference it returns the bare ``keep`` tensor when ``boxes`` is empty (:397-398).
    Ties: the reference's ascending ``scores.sort(0)`` is unstable; this follows a stable sort (higher
    index first among equal scores)."""
    kind, dev = _tensor.kind_of(scores), _tensor.device_of(boxes, scores)
    b = _tensor.to_dev(boxes, dev)
    s = _tensor.to_dev(scores, dev).reshape(-1)
    n = int(s.shape[0])
    keep = torch.zeros((n,), dtype=torch.int64, device=dev)
    if b.numel() == 0:
        return _tensor.like(kind, keep)
    b = b.reshape(n, 4)
    cap = min(n, int(top_k))
    k32, cnt = _ops.nms_indices(b, 4, s, 1, n, 0.0, _ops.THRESH_NONE, int(top_k), float(overlap), _ops.NMS_SSD, cap, dev)
    count = int(cnt.item())
    keep[:count] = k32[:count].to(torch.int64)
    return _tensor.like(kind, keep), count
