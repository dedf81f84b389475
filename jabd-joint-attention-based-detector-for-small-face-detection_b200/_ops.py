"""Operator wrappers shared by the reference-named modules (``retinaface_training``, ``box_utils``,
``utils_bbox``).  Each one marshals its arguments, launches one entry point of ``libjabd_b200.so`` on the
current stream and returns the result in the caller's array kind."""
import torch

from . import _lib, _tensor
from ._tensor import ptr

THRESH_NONE, THRESH_GE, THRESH_GT = 0, 1, 2
NMS_TV, NMS_SSD = 0, 1
IOU, GIOU, DIOU, CIOU = 1, 2, 3, 4


def _rows(x, cols, name):
    if x.ndim != 2 or x.shape[1] != cols:
        raise ValueError("%s must be [n, %d]" % (name, cols))


def point_form(boxes):
    kind, dev = _tensor.kind_of(boxes), _tensor.device_of(boxes)
    b = _tensor.to_dev(boxes, dev)
    _rows(b, 4, "boxes")
    out = torch.empty_like(b)
    with torch.cuda.device(dev):
        _lib.call("jabd_point_form", ptr(b), b.shape[0], ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def _pairwise(entry, box_a, box_b):
    kind, dev = _tensor.kind_of(box_a), _tensor.device_of(box_a, box_b)
    a, b = _tensor.to_dev(box_a, dev), _tensor.to_dev(box_b, dev)
    _rows(a, 4, "box_a")
    _rows(b, 4, "box_b")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call(entry, ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def jaccard(box_a, box_b):
    return _pairwise("jabd_jaccard", box_a, box_b)


def intersect(box_a, box_b):
    return _pairwise("jabd_intersect", box_a, box_b)


def encode(matched, priors, variances):
    kind, dev = _tensor.kind_of(matched), _tensor.device_of(matched, priors)
    m, p = _tensor.to_dev(matched, dev), _tensor.to_dev(priors, dev)
    _rows(m, 4, "matched")
    if p.shape != m.shape:
        raise ValueError("matched and priors must both be [n, 4]")
    v0, v1 = _tensor.variances_of(variances)
    out = torch.empty_like(m)
    with torch.cuda.device(dev):
        _lib.call("jabd_encode", ptr(m), ptr(p), m.shape[0], v0, v1, ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def encode_landm(matched, priors, variances):
    kind, dev = _tensor.kind_of(matched), _tensor.device_of(matched, priors)
    m, p = _tensor.to_dev(matched, dev), _tensor.to_dev(priors, dev)
    _rows(m, 10, "matched")
    _rows(p, 4, "priors")
    if p.shape[0] != m.shape[0]:
        raise ValueError("matched [n,10] and priors [n,4] must have the same n")
    v0, _ = _tensor.variances_of(variances)
    out = torch.empty_like(m)
    with torch.cuda.device(dev):
        _lib.call("jabd_encode_landm", ptr(m), ptr(p), m.shape[0], v0, ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def _decode_common(entry, x, priors, variances, cols, both_var):
    kind, dev = _tensor.kind_of(x), _tensor.device_of(x, priors)
    t, p = _tensor.to_dev(x, dev), _tensor.to_dev(priors, dev)
    _rows(p, 4, "priors")
    P = int(p.shape[0])
    if t.ndim == 2:
        batch = 1
    elif t.ndim == 3:
        batch = int(t.shape[0])
    else:
        raise ValueError("expected [P,%d] or [B,P,%d]" % (cols, cols))
    if t.shape[-1] != cols or t.shape[-2] != P:
        raise ValueError("expected [..., %d, %d] to go with %d priors" % (P, cols, P))
    v0, v1 = _tensor.variances_of(variances)
    out = torch.empty_like(t)
    with torch.cuda.device(dev):
        if both_var:
            _lib.call(entry, ptr(t), ptr(p), P, batch, v0, v1, ptr(out), _tensor.stream_of(dev))
        else:
            _lib.call(entry, ptr(t), ptr(p), P, batch, v0, ptr(out), _tensor.stream_of(dev))
    return _tensor.like(kind, out)


def decode(loc, priors, variances):
    return _decode_common("jabd_decode", loc, priors, variances, 4, True)


def decode_landm(pre, priors, variances):
    return _decode_common("jabd_decode_landm", pre, priors, variances, 10, False)


def match_one(threshold, truths, priors, variances, labels, landms, loc_t, conf_t, landm_t, idx, label_mode, encode_mode):
    """One image of ``match`` with the reference's in-place row write ``loc_t[idx] = ...``."""
    from .batched import assign_targets
    dev = _tensor.device_of(truths, priors)
    tr = _tensor.to_dev(truths, dev)
    G = int(tr.shape[0])
    if tr.ndim != 2 or tr.shape[1] != 4:
        raise ValueError("truths must be [G, 4]")
    if G == 0:
        raise IndexError("max(): Expected reduction dim 1 to have non-zero size (no ground truth)")
    lab = _tensor.to_dev(labels, dev).reshape(G, 1)
    lm = _tensor.to_dev(landms, dev).reshape(G, 10) if landms is not None else torch.zeros((G, 10), dtype=torch.float32, device=dev)
    rows = torch.cat([tr, lm, lab], 1)
    loc, conf, landm = assign_targets(priors, [rows], threshold=threshold, variances=variances, label_mode=label_mode,
                                      encode=bool(encode_mode), with_landm=landm_t is not None)
    loc_t[idx] = loc[0].to(loc_t.device) if isinstance(loc_t, torch.Tensor) else loc[0].cpu().numpy()
    conf_t[idx] = conf[0].to(conf_t.device) if isinstance(conf_t, torch.Tensor) else conf[0].cpu().numpy()
    if landm_t is not None:
        landm_t[idx] = landm[0].to(landm_t.device) if isinstance(landm_t, torch.Tensor) else landm[0].cpu().numpy()


def nms_indices(boxes, box_stride, scores, score_stride, n, conf_thres, thresh_mode, pre_nms_topk, nms_thres, nms_mode,
                keep_cap, dev, cluster=0, return_stats=False):
    """Single-segment ``jabd_nms``; returns (keep_idx i32 [keep_cap] CUDA, count i32 [1] CUDA).  ``cluster``: CTAs per segment
    (``JABD_NMS_CLUSTER``, 0 = automatic); ``return_stats`` adds the selection statistics ``[4]`` i32 (see jabd_b200.h)."""
    L = _lib.lib()
    keep = torch.empty((max(keep_cap, 1),), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = _tensor.workspace(L.jabd_nms_workspace_bytes(1, n, keep_cap), dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_nms", ptr(boxes), 0, box_stride, ptr(scores), 0, score_stride, 1, n, float(conf_thres), thresh_mode,
                  int(pre_nms_topk), float(nms_thres), int(nms_mode) | (int(cluster) << 12), keep_cap, ptr(keep), ptr(count),
                  ptr(ws), ws.numel(), _tensor.stream_of(dev))
    if return_stats:
        off = int(L.jabd_nms_stats_offset(1, keep_cap))
        return keep, count, ws[off:off + 16].view(torch.int32)
    return keep, count


def topk(scores, k, conf_thres=None, strict=True, return_stats=False):
    """Segmented top-k: scores [S,N] or [N]; stable descending order, ties -> lower index.
    Returns (idx [S,k] i32 padded with -1, count [S] i32)."""
    kind, dev = _tensor.kind_of(scores), _tensor.device_of(scores)
    s = _tensor.to_dev(scores, dev)
    squeeze = s.ndim == 1
    if squeeze:
        s = s[None]
    S, N = int(s.shape[0]), int(s.shape[1])
    out = torch.empty((S, k), dtype=torch.int32, device=dev)
    cnt = torch.empty((S,), dtype=torch.int32, device=dev)
    mode = THRESH_NONE if conf_thres is None else (THRESH_GT if strict else THRESH_GE)
    ws = _tensor.workspace(_lib.lib().jabd_topk_workspace_bytes(S, N, int(k)), dev) if return_stats else None
    with torch.cuda.device(dev):
        _lib.call("jabd_topk", ptr(s), N, 1, S, N, float(conf_thres or 0.0), mode, int(k), ptr(out), ptr(cnt), ptr(ws),
                  ws.numel() if ws is not None else 0, _tensor.stream_of(dev))
    if squeeze:
        out, cnt = out[0], cnt[0]
    if return_stats:
        return _tensor.like(kind, out), _tensor.like(kind, cnt), ws[:16 * S].view(torch.int32).reshape(S, 4)
    return _tensor.like(kind, out), _tensor.like(kind, cnt)


def overlaps_family(bboxes1, bboxes2, kind):
    """``bbox_overlaps_{iou,giou,diou,ciou}`` (R/utils/box_utils.py:5-158): row i of ``bboxes1`` against row i of
    ``bboxes2`` (one row broadcasts, like the reference's torch.min / torch.max), result ``[N]`` fp32."""
    knd, dev = _tensor.kind_of(bboxes1), _tensor.device_of(bboxes1, bboxes2)
    a, b = _tensor.to_dev(bboxes1, dev).reshape(-1, 4), _tensor.to_dev(bboxes2, dev).reshape(-1, 4)
    rows, cols = int(a.shape[0]), int(b.shape[0])
    if rows * cols == 0:
        return _tensor.like(knd, torch.zeros((rows, cols), dtype=torch.float32, device=dev))   # :9-10
    if rows > cols:                                                                           # :12-15
        a, b = b, a
    if a.shape[0] != b.shape[0]:
        if a.shape[0] != 1:
            raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton dimension 0"
                               % (a.shape[0], b.shape[0]))
        a = a.expand(b.shape[0], 4).contiguous()
    out = torch.empty((a.shape[0],), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_bbox_overlaps_family", ptr(a), ptr(b), a.shape[0], int(kind), ptr(out), _tensor.stream_of(dev))
    return _tensor.like(knd, out)


def diounms_indices(boxes, scores, n, top_k, overlap, beta1, keep_cap, dev):
    """Single-segment ``jabd_diounms``; returns (keep_idx i32 [keep_cap] CUDA, count i32 [1] CUDA)."""
    L = _lib.lib()
    keep = torch.empty((max(keep_cap, 1),), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = _tensor.workspace(L.jabd_nms_workspace_bytes(1, n, keep_cap), dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_diounms", ptr(boxes), 0, 4, ptr(scores), 0, 1, 1, n, int(top_k), float(overlap), float(beta1), keep_cap,
                  ptr(keep), ptr(count), ptr(ws), ws.numel(), _tensor.stream_of(dev))
    return keep, count
