"""CPU oracle of the WIDER-FACE AP evaluation -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A scalar restatement of R/utils/utils_map.py in plain Python / numpy float64: every function cites the lines it
follows.  Pinned against the imported reference by tests/golden/make_golden.py (wider_eval.npz) and
tests/test_oracle_golden.py.
"""
import numpy as np


def iou_f64(a, b):
    """intersect + bbox_overlaps for one pair of point-form boxes (R/utils/utils_map.py:7-27)."""
    w = max(min(a[2], b[2]) - max(a[0], b[0]), 0.0)
    h = max(min(a[3], b[3]) - max(a[1], b[1]), 0.0)
    inter = np.float64(w) * np.float64(h)
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(inter) / np.float64(union)


def bbox_overlaps(box_a, box_b):
    a, b = np.asarray(box_a, np.float64), np.asarray(box_b, np.float64)
    out = np.empty((a.shape[0], b.shape[0]))
    for i in range(a.shape[0]):
        for j in range(b.shape[0]):
            out[i, j] = iou_f64(a[i], b[j])
    return out


def image_eval(pred, gt, keep, iou_thresh):
    """R/utils/utils_map.py:100-132.  ``keep`` is the reference's ``ignore`` array (1 = counted).
    Returns int arrays (pred_recall, proposal_list)."""
    pred, gt = np.asarray(pred, np.float64), np.asarray(gt, np.float64)
    N, G = pred.shape[0], gt.shape[0]
    p = np.stack([pred[:, 0], pred[:, 1], pred[:, 2] + pred[:, 0], pred[:, 3] + pred[:, 1]], 1)   # :112-113
    g = np.stack([gt[:, 0], gt[:, 1], gt[:, 2] + gt[:, 0], gt[:, 3] + gt[:, 1]], 1)             # :114-115
    state = np.zeros(G, np.int64)          # recall_list
    proposal = np.ones(N, np.int64)
    recalled = np.zeros(N, np.int64)
    n_rec = 0
    for h in range(N):
        best, bi, is_nan = 0.0, 0, False
        for j in range(G):
            v = iou_f64(p[h], g[j])
            if j == 0:
                best, bi, is_nan = v, 0, v != v
            elif not is_nan and (v != v or v > best):   # np.max / np.argmax: NaN wins, else first maximum
                best, bi, is_nan = v, j, v != v
        if best >= iou_thresh:                           # :123
            if keep[bi] == 0:                            # :124-126
                state[bi] = -1
                proposal[h] = -1
            elif state[bi] == 0:                         # :127-128
                state[bi] = 1
                n_rec += 1
        recalled[h] = n_rec                              # :130-131
    return recalled, proposal


def img_pr_info(thresh_num, pred, proposal, recalled):
    """R/utils/utils_map.py:135-148."""
    out = np.zeros((thresh_num, 2))
    score = np.asarray(pred, np.float64)[:, 4]
    for t in range(thresh_num):
        thresh = 1 - (t + 1) / thresh_num
        last = -1
        for h in range(score.shape[0]):
            if score[h] >= thresh:
                last = h
        if last >= 0:
            out[t, 0] = int(np.count_nonzero(np.asarray(proposal)[:last + 1] == 1))
            out[t, 1] = recalled[last]
    return out


def norm_scores(pred_list):
    """R/utils/utils_map.py:75-98 on a list of [N,5] arrays; returns normalised copies."""
    mx, mn = 0, 1
    for v in pred_list:
        if len(v) == 0:
            continue
        mx, mn = max(np.max(v[:, -1]), mx), min(np.min(v[:, -1]), mn)
    diff = mx - mn
    out = []
    for v in pred_list:
        v = np.array(v, np.float64, copy=True)
        if len(v) != 0:
            v[:, -1] = (v[:, -1] - mn) / diff
        out.append(v)
    return out


def pr_counters(preds, gts, keeps, iou_thresh, thresh_num):
    """The pr_curve accumulation of evaluation() (R/utils/utils_map.py:185-203)."""
    pr = np.zeros((thresh_num, 2))
    for p, g, k in zip(preds, gts, keeps):
        if len(g) == 0 or len(p) == 0:                   # :196-197
            continue
        rec, prop = image_eval(p, g, k, iou_thresh)
        pr += img_pr_info(thresh_num, p, prop, rec)
    return pr


def average_precision(pr_curve, count_face):
    """dataset_pr_info + voc_ap (R/utils/utils_map.py:151-170)."""
    n = pr_curve.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.array([pr_curve[i, 1] / pr_curve[i, 0] for i in range(n)])
        rec = np.array([pr_curve[i, 1] / count_face for i in range(n)])
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1])
