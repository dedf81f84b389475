"""Drop-in for ``utils/utils_map.py`` of the reference (R/utils/utils_map.py): WIDER-FACE AP evaluation.

Same names and argument meaning (``bbox_overlaps``, ``image_eval``, ``img_pr_info``, ``norm_score``,
``dataset_pr_info``, ``voc_ap``, ``read_pred_file``, ``get_preds``, ``get_gt_boxes``, ``evaluation``); the per-image
work -- prediction x GT IoU, the sequential GT assignment and the 1000-threshold PR counters -- runs in
``libjabd_b200.so`` for all images of a call at once (``jabd_wider_eval``, fp64 like the reference's numpy code).
Additive entry points: ``evaluate_arrays`` (in-memory predictions / GT instead of the txt + .mat round trip) and
``dets_to_pred_rows`` (padded detections ``[B,K,15]`` + counts, e.g. from ``sharding.allgather_detections`` ->
the ``x y w h score`` rows of the txt wire format).  File parsing, ``dataset_pr_info`` and ``voc_ap`` (1000-element
arithmetic) are host code, as in the reference.
"""
import os

import numpy as np
import torch

from . import _lib, _tensor
from ._tensor import ptr

__all__ = ["bbox_overlaps", "image_eval", "img_pr_info", "norm_score", "dataset_pr_info", "voc_ap", "read_pred_file",
           "get_preds", "get_gt_boxes", "evaluation", "evaluate_arrays", "pr_counters", "dets_to_pred_rows"]

F64 = torch.float64


def _dev():
    _tensor.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _f64(x, dev, cols):
    t = _tensor.to_dev(np.asarray(x, dtype=np.float64).reshape(-1, cols) if not isinstance(x, torch.Tensor) else x, dev, F64)
    return t.reshape(-1, cols)


def bbox_overlaps(box_a, box_b):
    """R/utils/utils_map.py:16-27: IoU of point-form boxes ``[A,4]`` x ``[B,4]`` -> ``[A,B]`` float64."""
    kind, dev = _tensor.kind_of(box_a), _tensor.device_of(box_a, box_b)
    a, b = _f64(box_a, dev, 4), _f64(box_b, dev, 4)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=F64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_bbox_overlaps_f64", ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def _pack(rows_list, cols, dev):
    off = np.zeros(len(rows_list) + 1, dtype=np.int32)
    arrs = []
    for i, r in enumerate(rows_list):
        a = np.asarray(r.detach().cpu().numpy() if isinstance(r, torch.Tensor) else r, dtype=np.float64)
        a = a.reshape(-1, a.shape[-1])[:, :cols] if a.size else np.zeros((0, cols))
        arrs.append(a)
        off[i + 1] = off[i] + a.shape[0]
    flat = np.concatenate(arrs, 0) if arrs else np.zeros((0, cols))
    return torch.from_numpy(np.ascontiguousarray(flat)).to(dev), torch.from_numpy(off).to(dev), off


def pr_counters(preds, gts, keeps, iou_thresh=0.4, thresh_num=1000, return_image_eval=False):
    """Sum over images of ``img_pr_info(image_eval(...))`` -- the ``pr_curve`` that ``evaluation`` accumulates
    (R/utils/utils_map.py:185-203) -- for lists of per-image ``pred [N,5]`` (x y w h score, file order), ``gt [G,4]``
    (x y w h) and ``keep [G]`` (1 = listed in the subset's gt_list).  One launch for all images.
    Returns ``pr_curve [thresh_num,2]`` float64 numpy (and the packed per-prediction ``pred_recall`` / ``proposal_list``)."""
    dev = _dev()
    I = len(preds)
    if not (len(gts) == I and len(keeps) == I):
        raise ValueError("preds, gts and keeps must have one entry per image")
    p, poff, poff_h = _pack(preds, 5, dev)
    g, goff, goff_h = _pack(gts, 4, dev)
    k = np.concatenate([np.asarray(x, dtype=np.uint8).reshape(-1) for x in keeps]) if I else np.zeros(0, np.uint8)
    if k.shape[0] != goff_h[-1]:
        raise ValueError("keep flags must match the GT rows")
    kd = torch.from_numpy(np.ascontiguousarray(k)).to(dev)
    pr = torch.zeros((thresh_num, 2), dtype=F64, device=dev)
    sumN, sumG = int(poff_h[-1]), int(goff_h[-1])
    rec = torch.zeros((max(sumN, 1),), dtype=torch.int32, device=dev) if return_image_eval else None
    prop = torch.ones((max(sumN, 1),), dtype=torch.int32, device=dev) if return_image_eval else None
    L = _lib.lib()
    ws = _tensor.workspace(L.jabd_wider_eval_workspace_bytes(I, sumN, sumG, thresh_num), dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_wider_eval", ptr(p), ptr(poff), ptr(g), ptr(goff), ptr(kd), I, sumN, sumG, float(iou_thresh), int(thresh_num),
                  ptr(pr), ptr(rec), ptr(prop), ptr(ws), ws.numel(), _tensor.stream_of(dev))
    out = pr.cpu().numpy()
    if return_image_eval:
        return out, rec[:sumN].cpu().numpy(), prop[:sumN].cpu().numpy(), poff_h
    return out


def image_eval(pred, gt, ignore, iou_thresh):
    """R/utils/utils_map.py:100-132 for one image: ``pred [N,5]`` and ``gt [G,4]`` as x y w h (score), ``ignore [G]``
    (1 = counted).  Returns ``(pred_recall, proposal_list)`` as float64 arrays like the reference."""
    pred, gt = np.asarray(pred, dtype=np.float64), np.asarray(gt, dtype=np.float64)
    if pred.shape[0] == 0 or gt.shape[0] == 0:   # the reference never calls it with empty inputs (:196-197)
        return np.zeros(pred.shape[0]), np.ones(pred.shape[0])
    _, rec, prop, _ = pr_counters([pred], [gt], [np.asarray(ignore) != 0], iou_thresh, 1, return_image_eval=True)
    return rec.astype(np.float64), prop.astype(np.float64)


def img_pr_info(thresh_num, pred_info, proposal_list, pred_recall):
    """R/utils/utils_map.py:135-148: per-threshold (proposal count, recalled count) of one image."""
    dev = _dev()
    p = _f64(pred_info, dev, 5)
    n = int(p.shape[0])
    prop = _tensor.to_dev(np.asarray(proposal_list), dev, torch.int32)
    rec = _tensor.to_dev(np.asarray(pred_recall), dev, torch.int32)
    out = torch.empty((thresh_num, 2), dtype=F64, device=dev)
    ws = _tensor.workspace(2 * 256 + 4 * (thresh_num + n) + 512, dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_img_pr_info", ptr(p), n, ptr(prop), ptr(rec), int(thresh_num), ptr(out), ptr(ws), ws.numel(),
                  _tensor.stream_of(dev))
    return out.cpu().numpy()


def norm_score(pred):
    """R/utils/utils_map.py:75-98: in-place min-max normalisation of the scores of ``{event: {image: [N,5]}}``."""
    dev = _dev()
    arrs = [v for k in pred.values() for v in k.values() if len(v) != 0]
    if not arrs:
        return
    flat = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, 5) for a in arrs], 0))).to(dev)
    ws = _tensor.workspace(256, dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_norm_score", ptr(flat), flat.shape[0], ptr(ws), ws.numel(), _tensor.stream_of(dev))
    out = flat.cpu().numpy()
    o = 0
    for a in arrs:
        n = len(a)
        a[:, -1] = out[o:o + n, 4]
        o += n


def dataset_pr_info(thresh_num, pr_curve, count_face):
    """R/utils/utils_map.py:151-156 (host arithmetic on ``thresh_num`` rows)."""
    _pr_curve = np.zeros((thresh_num, 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        _pr_curve[:, 0] = pr_curve[:, 1] / pr_curve[:, 0]
        _pr_curve[:, 1] = pr_curve[:, 1] / count_face
    return _pr_curve


def voc_ap(rec, prec):
    """R/utils/utils_map.py:159-170 (host arithmetic): area under the monotone envelope of the PR curve."""
    mrec = np.concatenate(([0.], rec, [1.]))
    envelope = np.maximum.accumulate(np.concatenate(([0.], prec, [0.]))[::-1])[::-1]   # right-to-left running maximum
    step = np.flatnonzero(mrec[1:] != mrec[:-1])
    return np.sum((mrec[step + 1] - mrec[step]) * envelope[step + 1])


def read_pred_file(filepath):
    """R/utils/utils_map.py:45-58: first line = image path, second = count, then one ``x y w h score`` row per line
    (rows whose first token is empty are skipped).  Returns ``(image file name, [N,5] float64)``."""
    with open(filepath, 'r') as f:
        text = f.readlines()
    name = text[0].rstrip('\n\r').split('/')[-1]
    rows = []
    for raw in text[2:]:
        tok = raw.rstrip('\r\n').split(' ')
        if tok[0] != '':
            rows.append([float(t) for t in tok[:5]])
    return name, np.array(rows)


def get_preds(pred_dir):
    """R/utils/utils_map.py:60-73: ``{event: {image name without .jpg: [N,5]}}`` from a directory of event folders."""
    out = {}
    for event in os.listdir(pred_dir):
        per_image = {}
        folder = os.path.join(pred_dir, event)
        for txt in os.listdir(folder):
            name, rows = read_pred_file(os.path.join(folder, txt))
            per_image[name.rstrip('.jpg')] = rows
        out[event] = per_image
    return out


def get_gt_boxes(gt_dir):
    """R/utils/utils_map.py:29-43."""
    from scipy.io import loadmat
    gt_mat = loadmat(os.path.join(gt_dir, 'wider_face_val.mat'))
    hard_mat = loadmat(os.path.join(gt_dir, 'wider_hard_val.mat'))
    medium_mat = loadmat(os.path.join(gt_dir, 'wider_medium_val.mat'))
    easy_mat = loadmat(os.path.join(gt_dir, 'wider_easy_val.mat'))
    return (gt_mat['face_bbx_list'], gt_mat['event_list'], gt_mat['file_list'], hard_mat['gt_list'], medium_mat['gt_list'],
            easy_mat['gt_list'])


def evaluate_arrays(preds, gts, keeps_by_setting, iou_thresh=0.4, thresh_num=1000):
    """AP per setting from in-memory arrays: ``preds`` / ``gts`` lists over images (scores already normalised),
    ``keeps_by_setting`` a list (easy, medium, hard, ...) of per-image keep flags.  Same arithmetic as the loop of
    ``evaluation`` (R/utils/utils_map.py:181-211)."""
    aps = []
    for keeps in keeps_by_setting:
        count_face = int(sum(int(np.count_nonzero(k)) for k in keeps))
        pr_curve = pr_counters(preds, gts, keeps, iou_thresh, thresh_num)
        pr = dataset_pr_info(thresh_num, pr_curve, count_face)
        aps.append(voc_ap(pr[:, 1], pr[:, 0]))
    return aps


def evaluation(pred, gt_path, iou_thresh=0.4):
    """R/utils/utils_map.py:173-223: ``pred`` directory of per-event txt files, ``gt_path`` the WIDER .mat directory.
    Prints and returns the easy / medium / hard AP."""
    pred = get_preds(pred)
    norm_score(pred)
    facebox_list, event_list, file_list, hard_gt_list, medium_gt_list, easy_gt_list = get_gt_boxes(gt_path)
    thresh_num = 1000
    preds, gts, keeps = [], [], [[], [], []]
    for i in range(len(event_list)):
        event_name = str(event_list[i][0][0])
        img_list = file_list[i][0]
        pred_list = pred[event_name]
        gt_bbx_list = facebox_list[i][0]
        for j in range(len(img_list)):
            gt_boxes = gt_bbx_list[j][0].astype('float')
            preds.append(np.asarray(pred_list[str(img_list[j][0][0])], dtype=np.float64).reshape(-1, 5))
            gts.append(gt_boxes)
            for s, gl in enumerate((easy_gt_list, medium_gt_list, hard_gt_list)):
                keep_index = gl[i][0][j][0]
                flag = np.zeros(gt_boxes.shape[0], dtype=np.uint8)
                if len(keep_index) != 0:
                    flag[np.asarray(keep_index).reshape(-1) - 1] = 1
                keeps[s].append(flag)
    aps = evaluate_arrays(preds, gts, keeps, iou_thresh, thresh_num)
    print("==================== Results ====================")
    print("Easy   Val AP: {}".format(aps[0]))
    print("Medium Val AP: {}".format(aps[1]))
    print("Hard   Val AP: {}".format(aps[2]))
    print("=================================================")
    return aps


def dets_to_pred_rows(dets, counts):
    """Padded detections ``[B,K,15]`` (pixel x1 y1 x2 y2 score ...) + ``counts [B]`` -> list of ``[N,5]`` float64 rows
    ``x y w h score`` in score order, i.e. what the txt wire format (read_pred_file) carries, without the disk round trip."""
    d = dets.detach().cpu().numpy() if isinstance(dets, torch.Tensor) else np.asarray(dets)
    c = counts.detach().cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    out = []
    for b in range(d.shape[0]):
        r = d[b, :int(c[b]), :5].astype(np.float64)
        rows = np.empty((r.shape[0], 5))
        rows[:, 0], rows[:, 1] = r[:, 0], r[:, 1]
        rows[:, 2], rows[:, 3] = r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]
        rows[:, 4] = r[:, 4]
        out.append(rows)
    return out
