"""numpy-facing wrapper around ``jabd_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

All arrays in/out are C-contiguous numpy; fp32 unless stated.  Function names
follow the reference (R/ = JABD2080ti/); each docstring cites what it restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjabd_oracle.so")
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_d = ctypes.POINTER(ctypes.c_double)
c_i = ctypes.POINTER(ctypes.c_int)
c_l = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "jabd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libjabd_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_priors.restype = ctypes.c_int64
        L.orc_priors.argtypes = [c_i, c_d, c_i, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f]
        L.orc_point_form.restype = None
        L.orc_point_form.argtypes = [c_f, ctypes.c_int64, c_f]
        L.orc_jaccard.restype = None
        L.orc_jaccard.argtypes = [c_f, ctypes.c_int64, c_f, ctypes.c_int64, c_f]
        L.orc_encode.restype = None
        L.orc_encode.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_float, ctypes.c_float, c_f]
        L.orc_encode_landm.restype = None
        L.orc_encode_landm.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_float, c_f]
        L.orc_match.restype = ctypes.c_int
        L.orc_match.argtypes = [ctypes.c_float, c_f, c_f, ctypes.c_float, ctypes.c_float, c_f, c_f,
                                ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                c_f, c_l, c_f, c_l, c_f, c_l, c_f]
        L.orc_decode.restype = None
        L.orc_decode.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_float, ctypes.c_float, c_f]
        L.orc_decode_landm.restype = None
        L.orc_decode_landm.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_float, c_f]
        L.orc_argsort_desc.restype = None
        L.orc_argsort_desc.argtypes = [c_f, ctypes.c_int64, c_l]
        L.orc_nms_tv.restype = ctypes.c_int64
        L.orc_nms_tv.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_double, c_l]
        L.orc_nms_ssd.restype = ctypes.c_int64
        L.orc_nms_ssd.argtypes = [c_f, c_l, ctypes.c_int64, ctypes.c_float, ctypes.c_int64, c_l]
        L.orc_diounms.restype = ctypes.c_int64
        L.orc_diounms.argtypes = [c_f, c_l, ctypes.c_int64, ctypes.c_float, ctypes.c_int64, ctypes.c_float, c_l]
        L.orc_detect.restype = ctypes.c_int64
        L.orc_detect.argtypes = [c_f, c_f, c_f, c_f, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, ctypes.c_int, ctypes.c_int64, ctypes.c_double,
                                 ctypes.c_int64, c_f, c_f, c_l, ctypes.c_int64]
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _pf(a):
    return a.ctypes.data_as(c_f)


def _pl(a):
    return a.ctypes.data_as(c_l)


def priors(cfg, image_size):
    """``Anchors(cfg, image_size).get_anchors()`` (R/utils/anchors.py:9-42) -> [P,4] cx,cy,w,h."""
    steps = np.asarray(cfg["steps"], dtype=np.int32)
    ms = np.asarray([m for lvl in cfg["min_sizes"] for m in lvl], dtype=np.float64)
    off = np.zeros(len(steps) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(l) for l in cfg["min_sizes"]])
    H, W = int(image_size[0]), int(image_size[1])
    args = (steps.ctypes.data_as(c_i), ms.ctypes.data_as(c_d), off.ctypes.data_as(c_i), len(steps), H, W,
            1 if cfg.get("clip", False) else 0)
    n = lib().orc_priors(*args, None)
    out = np.empty((n, 4), dtype=np.float32)
    lib().orc_priors(*args, _pf(out))
    return out


def point_form(boxes):
    """R/nets/retinaface_training.py:8-10."""
    b = _f(boxes)
    out = np.empty_like(b)
    lib().orc_point_form(_pf(b), b.shape[0], _pf(out))
    return out


def jaccard(box_a, box_b):
    """R/nets/retinaface_training.py:41-59: dense IoU [A,B], both point-form."""
    a, b = _f(box_a), _f(box_b)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_jaccard(_pf(a), a.shape[0], _pf(b), b.shape[0], _pf(out))
    return out


def encode(matched, pri, variances):
    """R/nets/retinaface_training.py:61-70."""
    m, p = _f(matched), _f(pri)
    out = np.empty_like(m)
    lib().orc_encode(_pf(m), _pf(p), m.shape[0], variances[0], variances[1], _pf(out))
    return out


def encode_landm(matched, pri, variances):
    """R/nets/retinaface_training.py:72-84."""
    m, p = _f(matched), _f(pri)
    out = np.empty_like(m)
    lib().orc_encode_landm(_pf(m), _pf(p), m.shape[0], variances[0], _pf(out))
    return out


def match(threshold, truths, pri, variances, labels, landms=None, label_mode=0, encode_mode=1):
    """One image of ``match`` (R/nets/retinaface_training.py:93-162 and the
    SSD / match_iou variants, see ``orc_match``).  Returns a dict with loc_t,
    conf_t (int64), landm_t (or None), best_truth_idx, best_truth_overlap (after
    the force-match override), best_prior_idx, best_prior_overlap."""
    t, p, lab = _f(truths), _f(pri), _f(labels)
    G, P = t.shape[0], p.shape[0]
    lm = _f(landms) if landms is not None else None
    loc_t = np.empty((P, 4), np.float32)
    conf_t = np.empty((P,), np.int64)
    landm_t = np.empty((P, 10), np.float32) if lm is not None else None
    bti = np.empty((P,), np.int64)
    bto = np.empty((P,), np.float32)
    bpi = np.empty((max(G, 1),), np.int64)
    bpo = np.empty((max(G, 1),), np.float32)
    rc = lib().orc_match(threshold, _pf(t), _pf(p), variances[0], variances[1], _pf(lab),
                         _pf(lm) if lm is not None else None, G, P, label_mode, encode_mode,
                         _pf(loc_t), _pl(conf_t), _pf(landm_t) if landm_t is not None else None,
                         _pl(bti), _pf(bto), _pl(bpi), _pf(bpo))
    if rc != 0:
        raise ValueError("match: empty ground truth (the reference raises on max over an empty dim)")
    return dict(loc_t=loc_t, conf_t=conf_t, landm_t=landm_t, best_truth_idx=bti, best_truth_overlap=bto,
                best_prior_idx=bpi[:G], best_prior_overlap=bpo[:G])


def match_batch(threshold, targets, pri, variances, label_mode=0, encode_mode=1):
    """The per-image loop of ``MultiBoxLoss.forward`` (R/nets/retinaface_training.py:197-214)
    over a list of ``[G_i,15]`` targets.  Returns stacked loc_t/conf_t/landm_t plus per-image extras."""
    outs = []
    for t in targets:
        t = _f(t)
        outs.append(match(threshold, t[:, :4], pri, variances, t[:, -1], t[:, 4:14], label_mode, encode_mode))
    return dict(
        loc_t=np.stack([o["loc_t"] for o in outs]),
        conf_t=np.stack([o["conf_t"] for o in outs]),
        landm_t=np.stack([o["landm_t"] for o in outs]),
        best_truth_idx=np.stack([o["best_truth_idx"] for o in outs]),
        best_truth_overlap=np.stack([o["best_truth_overlap"] for o in outs]),
        best_prior_idx=np.concatenate([o["best_prior_idx"] for o in outs]),
        best_prior_overlap=np.concatenate([o["best_prior_overlap"] for o in outs]),
    )


def decode(loc, pri, variances):
    """R/utils/utils_bbox.py:29-34."""
    l, p = _f(loc), _f(pri)
    out = np.empty_like(l)
    lib().orc_decode(_pf(l), _pf(p), l.shape[0], variances[0], variances[1], _pf(out))
    return out


def decode_landm(pre, pri, variances):
    """R/utils/utils_bbox.py:39-46."""
    l, p = _f(pre), _f(pri)
    out = np.empty_like(l)
    lib().orc_decode_landm(_pf(l), _pf(p), l.shape[0], variances[0], _pf(out))
    return out


def argsort_desc(scores):
    s = _f(scores)
    out = np.empty((s.shape[0],), np.int64)
    lib().orc_argsort_desc(_pf(s), s.shape[0], _pl(out))
    return out


def nms_tv(boxes, scores, iou_threshold):
    """torchvision.ops.nms (CPU kernel, 0.26.0) as called at R/utils/utils_bbox.py:275-279."""
    b, s = _f(boxes).reshape(-1, 4), _f(scores)
    keep = np.empty((max(b.shape[0], 1),), np.int64)
    n = lib().orc_nms_tv(_pf(b), _pf(s), b.shape[0], float(iou_threshold), _pl(keep))
    return keep[:n].copy()


def nms_ssd(boxes, scores, overlap=0.5, top_k=200, order_asc=None):
    """SSD greedy NMS, R/utils/box_utils.py:384-448 (== nms_r R/utils/utils_bbox.py:116-180).
    ``order_asc``: ascending argsort to use (torch's unstable sort order is an
    input); default = stable ascending.  Returns (keep[n] zero-padded int64, count)."""
    b, s = _f(boxes).reshape(-1, 4), _f(scores)
    n = b.shape[0]
    if order_asc is None:
        order_asc = np.argsort(s, kind="stable")
    o = np.ascontiguousarray(order_asc, dtype=np.int64)
    keep = np.zeros((max(n, 1),), np.int64)
    c = lib().orc_nms_ssd(_pf(b), _pl(o), n, overlap, top_k, _pl(keep))
    return keep[:n], int(c)


def diounms(boxes, scores, overlap=0.5, top_k=200, beta1=1.0, order_asc=None):
    """DIoU-NMS, R/utils/utils_bbox.py:182-258.  Returns (keep[n] zero-padded int64, count); ties like ``nms_ssd``."""
    b, s = _f(boxes).reshape(-1, 4), _f(scores)
    n = b.shape[0]
    if order_asc is None:
        order_asc = np.argsort(s, kind="stable")
    o = np.ascontiguousarray(order_asc, dtype=np.int64)
    keep = np.zeros((max(n, 1),), np.int64)
    c = lib().orc_diounms(_pf(b), _pl(o), n, overlap, top_k, beta1, _pl(keep))
    return keep[:n], int(c)


def non_max_suppression(detection, conf_thres=0.5, nms_thres=0.3):
    """R/utils/utils_bbox.py:260-296: rows with score >= conf_thres, torchvision nms,
    kept rows in descending score order; ``[]`` if nothing passes."""
    d = _f(detection)
    thr = np.float32(conf_thres)
    d = d[d[:, 4] >= thr]
    if len(d) <= 0:
        return []
    keep = nms_tv(d[:, :4], d[:, 4], nms_thres)
    return d[keep]


def detect(loc, conf, landm, pri, variances, conf_thres, strict, pre_nms_topk, nms_thres, keep_topk,
           boxes_override=None):
    """Composed inference pipeline for one image (see ``orc_detect``).
    Returns (dets [K,15], keep_idx [K] prior indices)."""
    l, c, lm, p = _f(loc), _f(conf), _f(landm), _f(pri)
    P = p.shape[0]
    cap = P if keep_topk <= 0 else int(keep_topk)
    dets = np.zeros((max(cap, 1), 15), np.float32)
    kidx = np.zeros((max(cap, 1),), np.int64)
    bo = _f(boxes_override) if boxes_override is not None else None
    n = lib().orc_detect(_pf(l), _pf(c), _pf(lm), _pf(p), P, variances[0], variances[1], conf_thres,
                         1 if strict else 0, pre_nms_topk, float(nms_thres), keep_topk,
                         _pf(bo) if bo is not None else None, _pf(dets), _pl(kidx), cap)
    return dets[:n].copy(), kidx[:n].copy()
