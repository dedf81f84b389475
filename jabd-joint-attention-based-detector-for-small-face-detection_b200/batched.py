"""Batched entry points of the hot path (additive to the reference's per-image API).

``assign_targets`` replaces the per-image Python loop of ``MultiBoxLoss.forward``
(R/nets/retinaface_training.py:197-227: CPU target tensors, one ``match`` per
image, three D2H copies per image, three H2D copies per batch) with three kernel
launches for the whole batch, outputs left on the GPU.

``detect`` replaces the per-image post-processing of ``Retinaface.detect_image``
(R/predict.py:167-181: decode, decode_landm, cat, threshold, torchvision NMS,
``.cpu()``) with one launch for the whole batch.
"""
import ctypes

import numpy as np
import torch

from . import _lib, _tensor
from ._tensor import ptr

__all__ = ["assign_targets", "assign_batches", "AssignBatches", "detect_batches", "DetectBatches", "detect", "pack_targets", "assign_targets_host", "detect_host", "multibox_loss", "correct_boxes", "letterbox_params"]

THRESH_NONE, THRESH_GE, THRESH_GT = 0, 1, 2
FLAG_DENSE = 1
FLAG_ASYNC = 4
FLAG_DEVICE_OUT = 16


def pack_targets(targets, device):
    """list of per-image ``[G_i, 15]`` arrays/tensors (R/utils/dataloader.py:177-186) or a
    ``(gt_packed [sumG,15], gt_offsets [B+1])`` pair -> (gt CUDA f32 [sumG,15], offsets CUDA i32 [B+1],
    host offsets list)."""
    if isinstance(targets, tuple) and len(targets) == 2 and not isinstance(targets[0], (list, tuple)) \
            and getattr(targets[1], "ndim", 1) == 1 and getattr(targets[0], "ndim", 0) == 2:
        gt = _tensor.to_dev(targets[0], device)
        offs_in = targets[1]
        offs_host = [int(v) for v in (offs_in.tolist() if hasattr(offs_in, "tolist") else list(offs_in))]
        offs = _tensor.to_dev(offs_in, device, torch.int32)
    else:
        # numpy arrays and torch tensors are used as they are; anything else that speaks DLPack is imported without a copy
        rows = [t if isinstance(t, (np.ndarray, torch.Tensor)) else (torch.from_dlpack(t) if hasattr(t, "__dlpack__") else
                                                                     torch.as_tensor(t)) for t in targets]
        offs_host = [0]
        for t in rows:
            if t.ndim != 2 or t.shape[1] != 15:
                raise ValueError("each target must be [G, 15] (x1 y1 x2 y2, 10 landmark coords, label)")
            offs_host.append(offs_host[-1] + int(t.shape[0]))
        if rows and all(isinstance(t, np.ndarray) for t in rows):
            gt = _tensor.to_dev(np.concatenate(rows, 0) if rows else np.zeros((0, 15), np.float32), device)
        elif rows:
            gt = torch.cat([_tensor.to_dev(t, device) for t in rows], 0).contiguous()
        else:
            gt = torch.zeros((0, 15), dtype=torch.float32, device=device)
        offs = torch.tensor(offs_host, dtype=torch.int32).to(device, non_blocking=True)
    if gt.ndim != 2 or gt.shape[1] != 15:
        raise ValueError("gt_packed must be [sumG, 15]")
    if offs_host[0] != 0 or offs_host[-1] != gt.shape[0] or any(b < a for a, b in zip(offs_host, offs_host[1:])):
        raise ValueError("gt_offsets must start at 0, be non-decreasing and end at sumG")
    return gt, offs, offs_host


def tune_flags(tune):
    """``JABD_ASSIGN_TUNE(seg_a, seg_b, coarse_pct)``: shape of the matching kernel's work list (a per-call option, results do
    not depend on it); ``None``: the library's choice."""
    if tune is None:
        return 0
    a, b, pct = (int(x) for x in tune)
    if not (0 < a < 256 and 0 < b < 256 and 0 <= pct <= 100):
        raise ValueError("tune must be (seg_a, seg_b, coarse_pct)")
    return (a << 8) | (b << 16) | (pct << 24)


def assign_targets(priors, targets, threshold=0.35, variances=(0.1, 0.2), label_mode=0, encode=True, dense=False,
                   with_landm=True, return_match=False, allow_empty=False, out=None, tune=None):
    """Target assignment for a batch.

    priors   [P,4] (cx,cy,w,h); targets: see ``pack_targets``.
    Returns ``(loc_t [B,P,4] f32, conf_t [B,P] i64, landm_t [B,P,10] f32 | None)`` on the GPU and, with
    ``return_match``, a dict with best_truth_idx/overlap ``[B,P]`` (after the force-match override) and
    best_prior_idx/overlap ``[sumG]``.  ``out=(loc_t, conf_t, landm_t)`` writes into existing CUDA tensors.
    An image without GT raises ``ValueError`` like the reference's ``max`` over an empty dim
    (R/nets/retinaface_training.py:111) unless ``allow_empty`` (then its targets are all zero).
    """
    dev = _tensor.device_of(priors, targets[0] if isinstance(targets, (list, tuple)) and len(targets) else None)
    pri = _tensor.to_dev(priors, dev)
    if pri.ndim != 2 or pri.shape[1] != 4:
        raise ValueError("priors must be [P, 4]")
    gt, offs, offs_host = pack_targets(targets, dev)
    B, P, sumG = len(offs_host) - 1, int(pri.shape[0]), int(gt.shape[0])
    if not allow_empty and any(b == a for a, b in zip(offs_host, offs_host[1:])):
        raise ValueError("assign_targets: an image has no ground truth (the reference raises on max over an empty dim)")
    v0, v1 = _tensor.variances_of(variances)
    if out is not None:
        loc_t, conf_t, landm_t = out
        for t, shp, dt in ((loc_t, (B, P, 4), torch.float32), (conf_t, (B, P), torch.int64)):
            if not (t.is_cuda and t.is_contiguous() and tuple(t.shape) == shp and t.dtype == dt):
                raise ValueError("out tensors must be contiguous CUDA tensors of shape [B,P,4] f32 / [B,P] i64 / [B,P,10] f32")
        if landm_t is not None and not (landm_t.is_cuda and landm_t.is_contiguous() and tuple(landm_t.shape) == (B, P, 10)
                                        and landm_t.dtype == torch.float32):
            raise ValueError("landm_t must be a contiguous CUDA [B,P,10] f32 tensor")
    else:
        loc_t = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
        conf_t = torch.empty((B, P), dtype=torch.int64, device=dev)
        landm_t = torch.empty((B, P, 10), dtype=torch.float32, device=dev) if with_landm else None
    extra = None
    bti = bto = bpi = bpo = None
    if return_match:
        bti = torch.empty((B, P), dtype=torch.int32, device=dev)
        bto = torch.empty((B, P), dtype=torch.float32, device=dev)
        bpi = torch.empty((sumG,), dtype=torch.int32, device=dev)
        bpo = torch.empty((sumG,), dtype=torch.float32, device=dev)
        extra = dict(best_truth_idx=bti, best_truth_overlap=bto, best_prior_idx=bpi, best_prior_overlap=bpo)
    L = _lib.lib()
    ws = _tensor.workspace(L.jabd_assign_workspace_bytes(B, P, sumG), dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_assign", ptr(pri), P, ptr(gt), ptr(offs), B, sumG, float(threshold), v0, v1, int(label_mode),
                  1 if encode else 0, (FLAG_DENSE if dense else 0) | tune_flags(tune), ptr(loc_t), ptr(conf_t), ptr(landm_t), ptr(bti), ptr(bto),
                  ptr(bpi), ptr(bpo), ptr(ws), ws.numel(), _tensor.stream_of(dev))
    if return_match:
        return loc_t, conf_t, landm_t, extra
    return loc_t, conf_t, landm_t


_LANES = {}


def lanes(device, n):
    """``n`` side streams of ``device`` for ``assign_batches`` (created once per device, reused by every call)."""
    have = _LANES.setdefault(torch.device(device), [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device))
    return have[:n]


class AssignBatches(object):
    """Target assignment of several independent batches per call (``jabd_assign_batches``): batch i runs on side stream
    ``i % lanes``, so that one batch's staging and encode kernels fill the SMs another batch's persistent matching kernel
    leaves idle on its ramp and tail; the caller's stream is ordered before and after all of them (no host
    synchronisation, capturable into a CUDA graph).  The reference has one ``MultiBoxLoss.forward`` call per batch
    (R/nets/retinaface_training.py:197-227) and nothing couples two batches -- a prefetching data loader or
    gradient-accumulation micro-batches give several at once.

    ``plan = AssignBatches(priors, [targets_0, targets_1, ...])`` packs the GT and owns outputs and workspaces;
    ``plan()`` enqueues everything and returns ``[(loc_t, conf_t, landm_t), ...]`` (the same tensors on every call).
    """

    def __init__(self, priors, batches, threshold=0.35, variances=(0.1, 0.2), label_mode=0, encode=True, dense=False,
                 with_landm=True, lanes_n=4, device=None, tune=None):
        first = batches[0] if len(batches) else None
        dev = device or _tensor.device_of(priors, first[0] if isinstance(first, (list, tuple)) and len(first) else None)
        self.dev = torch.device(dev)
        self.pri = _tensor.to_dev(priors, self.dev)
        if self.pri.ndim != 2 or self.pri.shape[1] != 4:
            raise ValueError("priors must be [P, 4]")
        self.P = int(self.pri.shape[0])
        self.opts = (float(threshold),) + _tensor.variances_of(variances) + (int(label_mode), 1 if encode else 0,
                                                                             (FLAG_DENSE if dense else 0) | tune_flags(tune))
        L = _lib.lib()
        self.items, self.outputs = [], []
        arr = (_lib.AssignBatch * max(len(batches), 1))()
        for i, tg in enumerate(batches):
            gt, offs, offs_host = pack_targets(tg, self.dev)
            B, sumG = len(offs_host) - 1, int(gt.shape[0])
            if any(b == a for a, b in zip(offs_host, offs_host[1:])):
                raise ValueError("AssignBatches: an image of batch %d has no ground truth" % i)
            ws = _tensor.workspace(L.jabd_assign_workspace_bytes(B, self.P, sumG), self.dev)
            loc_t = torch.empty((B, self.P, 4), dtype=torch.float32, device=self.dev)
            conf_t = torch.empty((B, self.P), dtype=torch.int64, device=self.dev)
            landm_t = torch.empty((B, self.P, 10), dtype=torch.float32, device=self.dev) if with_landm else None
            self.items.append((gt, offs, ws))
            self.outputs.append((loc_t, conf_t, landm_t))
            arr[i] = _lib.AssignBatch(gt.data_ptr() or None, offs.data_ptr(), B, sumG, loc_t.data_ptr(), conf_t.data_ptr(),
                                      landm_t.data_ptr() if with_landm else None, ws.data_ptr(), ws.numel())
        self.n = len(batches)
        self.arr = arr
        self.set_lanes(lanes_n)

    def set_lanes(self, lanes_n):
        self.lane_streams = lanes(self.dev, max(0, min(int(lanes_n), self.n)))
        self.lane_arr = (ctypes.c_void_p * max(len(self.lane_streams), 1))(*[s.cuda_stream for s in self.lane_streams])

    def __call__(self):
        t, v0, v1, lm, enc, fl = self.opts
        with torch.cuda.device(self.dev):
            _lib.call("jabd_assign_batches", ptr(self.pri), self.P, ctypes.cast(self.arr, ctypes.c_void_p), self.n, t, v0, v1, lm,
                      enc, fl, ctypes.cast(self.lane_arr, ctypes.c_void_p), len(self.lane_streams), _tensor.stream_of(self.dev))
        return self.outputs


def assign_batches(priors, batches, **kw):
    """One-shot form of ``AssignBatches``: ``[(loc_t, conf_t, landm_t), ...]`` for a list of batches."""
    return AssignBatches(priors, batches, **kw)()


def detect(loc, conf, landm, priors, variances=(0.1, 0.2), conf_thres=0.02, strict=True, pre_nms_topk=5000,
           nms_thres=0.4, keep_topk=750, cluster=0, return_stats=False, out=None):
    """Fused decode -> class-1 score threshold -> top-k -> NMS -> keep for a batch.

    loc [B,P,4], conf [B,P,2] (softmax probabilities, R/nets/retinaface_eca_nonlocal.py:355-359),
    landm [B,P,10] or None, priors [P,4].  ``strict``: score > conf_thres (cfg3) else >= (R/utils/utils_bbox.py:266).
    Returns ``(dets [B,keep_topk,15] zero padded, counts [B] i32, keep_idx [B,keep_topk] i32, -1 padded)`` on the GPU.
    ``cluster``: CTAs (SMs) per image, 0 = automatic (``JABD_DET_CLUSTER``; results do not depend on it).
    ``return_stats`` adds the call's selection statistics ``[B,4]`` i32 (rounds, exact three-pass rounds, candidates, chunks).
    ``out=(dets, counts, keep_idx)`` writes into existing contiguous CUDA tensors of those shapes (e.g. the send buffer of
    ``sharding.DetectionGather``).
    """
    dev = _tensor.device_of(loc, conf, priors)
    loc_d, conf_d, pri = _tensor.to_dev(loc, dev), _tensor.to_dev(conf, dev), _tensor.to_dev(priors, dev)
    if loc_d.ndim == 2:
        loc_d, conf_d = loc_d[None], conf_d[None]
        landm = landm[None] if landm is not None else None
    landm_d = _tensor.to_dev(landm, dev) if landm is not None else None
    B, P = int(loc_d.shape[0]), int(loc_d.shape[1])
    if tuple(loc_d.shape) != (B, P, 4) or tuple(conf_d.shape) != (B, P, 2) or tuple(pri.shape) != (P, 4) or \
            (landm_d is not None and tuple(landm_d.shape) != (B, P, 10)):
        raise ValueError("detect: expected loc [B,P,4], conf [B,P,2], landm [B,P,10], priors [P,4]")
    keep_cap = int(keep_topk) if keep_topk and keep_topk > 0 else P
    v0, v1 = _tensor.variances_of(variances)
    if out is not None:
        dets, counts, keep_idx = out
        for t, shp, dt in ((dets, (B, keep_cap, 15), torch.float32), (counts, (B,), torch.int32), (keep_idx, (B, keep_cap), torch.int32)):
            if not (t.is_cuda and t.is_contiguous() and tuple(t.shape) == shp and t.dtype == dt):
                raise ValueError("out must be contiguous CUDA tensors dets [B,keep,15] f32, counts [B] i32, keep_idx [B,keep] i32")
    else:
        dets = torch.empty((B, keep_cap, 15), dtype=torch.float32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        keep_idx = torch.empty((B, keep_cap), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = _tensor.workspace(L.jabd_detect_workspace_bytes(B, P, keep_cap), dev)
    with torch.cuda.device(dev):
        _lib.call("jabd_detect", ptr(loc_d), ptr(conf_d), ptr(landm_d), ptr(pri), B, P, v0, v1, float(conf_thres),
                  THRESH_GT if strict else THRESH_GE, int(pre_nms_topk) if pre_nms_topk else 0, float(nms_thres), keep_cap,
                  int(cluster), ptr(dets), ptr(counts), ptr(keep_idx), ptr(ws), ws.numel(), _tensor.stream_of(dev))
    if return_stats:
        off = int(L.jabd_nms_stats_offset(B, keep_cap))
        return dets, counts, keep_idx, ws[off:off + 16 * B].view(torch.int32).reshape(B, 4)
    return dets, counts, keep_idx


class DetectBatches(object):
    """Fused detect of several independent batches per call (``jabd_detect_batches``).  ``lanes_n=0`` (default): ONE launch
    whose grid covers the images of all batches -- the block scheduler hands the next image to whichever SM falls free;
    ``lanes_n>0``: batch i is its own launch on side stream ``i % lanes``.  Either way the cluster width is chosen for all
    the images in flight together (narrower than a lone call's: less redundant work per image, the other images keep the
    rest of the GPU busy).  No host synchronisation; capturable into a CUDA graph.  The reference post-processes one image
    per call (R/predict.py:167-181); a validation pass over a data set (R/evaluate_utils.py) has many images queued.

    ``plan = DetectBatches(priors, [(loc, conf, landm), ...])`` owns outputs and workspaces; ``plan()`` enqueues everything
    and returns ``[(dets, counts, keep_idx), ...]`` (the same tensors on every call).  All batches share P and the options.
    """

    def __init__(self, priors, batches, variances=(0.1, 0.2), conf_thres=0.02, strict=True, pre_nms_topk=5000, nms_thres=0.4,
                 keep_topk=750, cluster=0, lanes_n=0, device=None):
        first = batches[0][0] if len(batches) else None
        self.dev = torch.device(device) if device is not None else _tensor.device_of(priors, first)
        self.pri = _tensor.to_dev(priors, self.dev)
        if self.pri.ndim != 2 or self.pri.shape[1] != 4:
            raise ValueError("priors must be [P, 4]")
        self.P = P = int(self.pri.shape[0])
        self.keep_cap = keep_cap = int(keep_topk) if keep_topk and keep_topk > 0 else P
        self.opts = _tensor.variances_of(variances) + (float(conf_thres), THRESH_GT if strict else THRESH_GE,
                                                        int(pre_nms_topk) if pre_nms_topk else 0, float(nms_thres), keep_cap, int(cluster))
        L = _lib.lib()
        self.inputs, self.outputs, self.ws = [], [], []
        arr = (_lib.DetectBatch * max(len(batches), 1))()
        for i, (loc, conf, landm) in enumerate(batches):
            loc_d, conf_d = _tensor.to_dev(loc, self.dev), _tensor.to_dev(conf, self.dev)
            landm_d = _tensor.to_dev(landm, self.dev) if landm is not None else None
            B = int(loc_d.shape[0])
            if tuple(loc_d.shape) != (B, P, 4) or tuple(conf_d.shape) != (B, P, 2) or \
                    (landm_d is not None and tuple(landm_d.shape) != (B, P, 10)):
                raise ValueError("DetectBatches: batch %d: expected loc [B,P,4], conf [B,P,2], landm [B,P,10]" % i)
            dets = torch.empty((B, keep_cap, 15), dtype=torch.float32, device=self.dev)
            counts = torch.empty((B,), dtype=torch.int32, device=self.dev)
            keep_idx = torch.empty((B, keep_cap), dtype=torch.int32, device=self.dev)
            ws = _tensor.workspace(L.jabd_detect_workspace_bytes(B, P, keep_cap), self.dev)
            self.inputs.append((loc_d, conf_d, landm_d))
            self.outputs.append((dets, counts, keep_idx))
            self.ws.append(ws)
            arr[i] = _lib.DetectBatch(loc_d.data_ptr(), conf_d.data_ptr(), landm_d.data_ptr() if landm_d is not None else None, B,
                                      dets.data_ptr(), counts.data_ptr(), keep_idx.data_ptr(), ws.data_ptr(), ws.numel())
        self.n = len(batches)
        self.arr = arr
        self.lane_streams = lanes(self.dev, max(0, min(int(lanes_n), self.n)))
        self.lane_arr = (ctypes.c_void_p * max(len(self.lane_streams), 1))(*[s.cuda_stream for s in self.lane_streams])

    def __call__(self):
        v0, v1, thr, mode, topk, nms, keep_cap, cluster = self.opts
        with torch.cuda.device(self.dev):
            _lib.call("jabd_detect_batches", ptr(self.pri), self.P, ctypes.cast(self.arr, ctypes.c_void_p), self.n, v0, v1, thr, mode,
                      topk, nms, keep_cap, cluster, ctypes.cast(self.lane_arr, ctypes.c_void_p), len(self.lane_streams),
                      _tensor.stream_of(self.dev))
        return self.outputs


def detect_batches(priors, batches, **kw):
    """One-shot form of ``DetectBatches``: ``[(dets, counts, keep_idx), ...]`` for a list of ``(loc, conf, landm)`` batches."""
    return DetectBatches(priors, batches, **kw)()


def letterbox_params(input_shape, image_shapes):
    """Per-image ``[B,6]`` float64 rows {offset_x, offset_y, scale_x, scale_y, width, height} for ``correct_boxes``:
    the quantities ``retinaface_correct_boxes`` derives from ``input_shape = (H_in, W_in)`` and each image's
    ``(H, W)`` (R/utils/utils_bbox.py:9-13) plus the pixel scale of R/predict.py:130-137, computed with the same numpy
    float64 expressions."""
    inp = np.array(input_shape, dtype=np.float64)
    rows = []
    for shp in image_shapes:
        img = np.array(shp, dtype=np.float64)
        new_shape = img * np.min(inp / img)
        offset = (inp - new_shape) / 2. / inp
        scale = inp / new_shape
        rows.append([offset[1], offset[0], scale[1], scale[0], float(int(shp[1])), float(int(shp[0]))])
    return np.array(rows, dtype=np.float64).reshape(-1, 6)


def correct_boxes(dets, counts, post, letterbox=True, to_pixels=True):
    """In-place post-processing of detection rows ``dets [B,K,15]`` (CUDA): undo the letterbox
    (R/utils/utils_bbox.py:9-24) and / or scale to pixels (R/predict.py:194-195).  ``post``: ``letterbox_params``."""
    if not (dets.is_cuda and dets.is_contiguous() and dets.dtype == torch.float32 and dets.ndim == 3 and dets.shape[2] == 15):
        raise ValueError("dets must be a contiguous CUDA f32 tensor [B, K, 15]")
    dev = dets.device
    B, K = int(dets.shape[0]), int(dets.shape[1])
    p = torch.as_tensor(np.ascontiguousarray(post, dtype=np.float64)).to(dev) if not isinstance(post, torch.Tensor) else post
    if tuple(p.shape) != (B, 6) or p.dtype != torch.float64 or not p.is_cuda:
        raise ValueError("post must be [B, 6] float64 (see letterbox_params)")
    cnt = None
    if counts is not None:
        cnt = counts.to(dev, torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.call("jabd_correct_boxes", ptr(dets), ptr(cnt), ptr(p.contiguous()), B, K, 1 if letterbox else 0,
                  1 if to_pixels else 0, _tensor.stream_of(dev))
    return dets


class _MultiBoxLossFn(torch.autograd.Function):
    """Hard-negative mining + the three loss reductions of ``MultiBoxLoss.forward``
    (R/nets/retinaface_training.py:229-303) as one autograd node over the network outputs."""

    @staticmethod
    def forward(ctx, loc_data, conf_data, landm_data, loc_t, conf_t, landm_t, negpos_ratio, loc_loss=0, priors=None, var0=0.1,
                var1=0.2):
        dev = loc_data.device
        B, P = int(loc_data.shape[0]), int(loc_data.shape[1])
        ld, cd, md = (t.detach().contiguous().float() for t in (loc_data, conf_data, landm_data))
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        norms = torch.empty((2,), dtype=torch.float32, device=dev)
        mask = torch.empty((B, P), dtype=torch.uint8, device=dev)
        L = _lib.lib()
        ws = _tensor.workspace(L.jabd_multibox_loss_workspace_bytes(B), dev)
        with torch.cuda.device(dev):
            _lib.call("jabd_multibox_loss_forward_ex", ptr(ld), ptr(cd), ptr(md), ptr(loc_t), ptr(conf_t), ptr(landm_t), B, P,
                      int(negpos_ratio), int(loc_loss), ptr(priors), float(var0), float(var1), ptr(losses), ptr(norms), ptr(mask),
                      ptr(ws), ws.numel(), _tensor.stream_of(dev))
        ctx.save_for_backward(ld, cd, md, loc_t, landm_t, mask, norms)
        ctx.loc_loss, ctx.priors, ctx.var = int(loc_loss), priors, (float(var0), float(var1))
        ctx.mark_non_differentiable(mask, norms)
        return losses[0], losses[1], losses[2], mask, norms

    @staticmethod
    def backward(ctx, g_l, g_c, g_landm, _gm, _gn):
        ld, cd, md, loc_t, landm_t, mask, norms = ctx.saved_tensors
        dev = ld.device
        B, P = int(ld.shape[0]), int(ld.shape[1])
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        g = torch.stack([(x if x is not None else zero).to(dev, torch.float32).reshape(()) for x in (g_l, g_c, g_landm)]).contiguous()
        g_loc, g_conf, g_lm = torch.empty_like(ld), torch.empty_like(cd), torch.empty_like(md)
        with torch.cuda.device(dev):
            _lib.call("jabd_multibox_loss_backward_ex", ptr(ld), ptr(cd), ptr(md), ptr(loc_t), ptr(landm_t), ptr(mask), ptr(norms),
                      ptr(g), B, P, ctx.loc_loss, ptr(ctx.priors), ctx.var[0], ctx.var[1], ptr(g_loc), ptr(g_conf), ptr(g_lm),
                      _tensor.stream_of(dev))
        return g_loc, g_conf, g_lm, None, None, None, None, None, None, None, None


LOC_LOSS = {"smooth_l1": 0, "Iou": 1, "Giou": 2, "Diou": 3, "Ciou": 4}


def multibox_loss(predictions, loc_t, conf_t, landm_t, negpos_ratio=7, return_aux=False, loc_loss="smooth_l1", priors=None,
                  variances=(0.1, 0.2)):
    """``(loss_l, loss_c, loss_landm)`` of R/nets/retinaface_training.py:229-303 for CUDA predictions
    ``(loc_data [B,P,4], conf_data [B,P,2] logits, landm_data [B,P,10])`` and the targets of ``assign_targets``.
    Differentiable w.r.t. the predictions.  ``return_aux`` adds the selection mask ``[B,P]`` u8 (bit0 pos, bit1
    pos1, bit2 mined negative) and ``(N, N1)``.  ``loc_loss`` "Iou" / "Giou" / "Diou" / "Ciou" replaces the smooth-L1 box
    term by ``IouLoss`` on ``decode(loc_data, priors)`` against raw matched boxes in ``loc_t`` (``assign_targets(...,
    encode=False)``), i.e. the MultiBoxLoss of R/nets/retinaface_training_DIOU.py:527-665."""
    loc_data, conf_data, landm_data = predictions
    B, P = int(loc_data.shape[0]), int(loc_data.shape[1])
    if not loc_data.is_cuda:
        raise RuntimeError("multibox_loss needs CUDA predictions; there is no CPU path")
    if tuple(conf_data.shape) != (B, P, 2) or tuple(landm_data.shape) != (B, P, 10) or tuple(loc_data.shape) != (B, P, 4):
        raise ValueError("expected loc_data [B,P,4], conf_data [B,P,2] (num_classes == 2), landm_data [B,P,10]")
    for t, shp, dt in ((loc_t, (B, P, 4), torch.float32), (conf_t, (B, P), torch.int64), (landm_t, (B, P, 10), torch.float32)):
        if not (t.is_cuda and t.is_contiguous() and tuple(t.shape) == shp and t.dtype == dt):
            raise ValueError("targets must be the contiguous CUDA tensors returned by assign_targets")
    kind = LOC_LOSS[loc_loss] if isinstance(loc_loss, str) else int(loc_loss)
    pri = None
    if kind:
        if priors is None:
            raise ValueError("the IoU-family box loss decodes loc_data and needs the priors")
        pri = _tensor.to_dev(priors, loc_data.device)
        if tuple(pri.shape) != (P, 4):
            raise ValueError("priors must be [P, 4]")
    v0, v1 = _tensor.variances_of(variances)
    l, c, m, mask, norms = _MultiBoxLossFn.apply(loc_data, conf_data, landm_data, loc_t, conf_t, landm_t, int(negpos_ratio), kind,
                                                 pri, v0, v1)
    return (l, c, m, mask, norms) if return_aux else (l, c, m)


class HostAssign(object):
    """End-to-end host-buffer path (``jabd_assign_host``) for a fixed (B, P, max sumG) shape: pinned staging buffers and
    the device scratch are allocated once; each call copies GT in, assigns, copies the three target tensors out.

    ``h(targets)`` is synchronous.  ``submit`` / ``wait`` expose the same call as a ``depth``-slot pipeline (one
    stream, scratch area and pinned output set per slot, ``JABD_ASSIGN_ASYNC``): while slot k's 34 MB of targets
    drain over PCIe, slot k+1's GT upload and kernels already run.  The tensors returned by ``wait(slot)`` stay
    valid until that slot is submitted again.

    ``device_out=True``: the three target tensors are CUDA tensors and stay in HBM (``JABD_ASSIGN_DEVICE_OUT``) -- the
    training flow, where ``MultiBoxLoss`` consumes them on the device; only the GT rows cross the bus.  Work that
    consumes them must be ordered after ``wait(slot)`` or enqueued on ``slot_stream(slot)``.

    The three host outputs of a slot are views of ONE pinned block laid out like the device staging area
    (``jabd_assign_host_out_offsets``), which lets ``jabd_assign_host`` bring them back with a single D2H copy.
    The slot streams are ordered after the stream that produced ``priors`` (constructor) -- submit never reads priors
    before they are written."""

    def __init__(self, priors, B, max_sum_g, with_landm=True, device=None, depth=2, device_out=False):
        _tensor.require_cuda()
        self.dev = torch.device(device) if device is not None else _tensor.device_of(priors)
        self.pri = _tensor.to_dev(priors, self.dev)
        self.B, self.P, self.cap = int(B), int(self.pri.shape[0]), int(max_sum_g)
        self.with_landm = bool(with_landm)
        self.device_out = bool(device_out)
        L = _lib.lib()
        nbytes = L.jabd_assign_host_scratch_bytes(self.B, self.P, self.cap, 1 if with_landm else 0)

        offs = (ctypes.c_size_t * 4)()
        L.jabd_assign_host_out_offsets(self.B, self.P, 1 if with_landm else 0, offs)
        BP = self.B * self.P

        def outputs():
            if self.device_out:
                return (torch.empty((self.B, self.P, 4), dtype=torch.float32, device=self.dev),
                        torch.empty((self.B, self.P), dtype=torch.int64, device=self.dev),
                        torch.empty((self.B, self.P, 10), dtype=torch.float32, device=self.dev) if with_landm else None, None)
            block = torch.empty((int(offs[3]),), dtype=torch.uint8).pin_memory()
            loc = block[int(offs[0]):int(offs[0]) + BP * 16].view(torch.float32).reshape(self.B, self.P, 4)
            conf = block[int(offs[1]):int(offs[1]) + BP * 8].view(torch.int64).reshape(self.B, self.P)
            lm = block[int(offs[2]):int(offs[2]) + BP * 40].view(torch.float32).reshape(self.B, self.P, 10) if with_landm else None
            return loc, conf, lm, block
        self.slots = []
        cur = torch.cuda.current_stream(self.dev)
        for _ in range(max(int(depth), 1)):
            loc, conf, lm, block = outputs()
            st = torch.cuda.Stream(self.dev)
            st.wait_stream(cur)          # self.pri may still be in flight on the caller's stream (priors kernel / H2D copy)
            self.slots.append(dict(
                scratch=_tensor.workspace(nbytes, self.dev),
                gt=torch.empty((self.cap, 15), dtype=torch.float32).pin_memory(),
                off=torch.empty((self.B + 1,), dtype=torch.int32).pin_memory(),
                loc_t=loc, conf_t=conf, landm_t=lm, block=block, stream=st, done=torch.cuda.Event()))
        self.next_slot = 0
        self.last_h2d = self.last_d2h = 0

    def submit(self, targets, threshold=0.35, variances=(0.1, 0.2), label_mode=0, encode=True, dense=False):
        """Enqueue one batch on the next slot; returns the slot id for ``wait``."""
        if len(targets) != self.B:
            raise ValueError("expected %d images" % self.B)
        k = self.next_slot
        self.next_slot = (k + 1) % len(self.slots)
        sl = self.slots[k]
        sl["done"].synchronize()            # the slot's previous outputs may still be in flight
        offs = sl["off"]
        # pack the per-image arrays into the slot's pinned buffer: one C call (jabd_pack_gt_rows) instead of torch.cat + cumsum
        rows = (ctypes.c_void_p * self.B)()
        counts = (ctypes.c_int * self.B)()
        keep = []
        f32 = torch.float32
        for i, t in enumerate(targets):
            if not (type(t) is torch.Tensor and t.dtype is f32 and not t.is_cuda and t.is_contiguous() and t.ndim == 2
                    and t.shape[1] == 15):
                t = torch.as_tensor(t).detach().to("cpu", f32)
                if t.ndim != 2 or t.shape[1] != 15:      # jabd_pack_gt_rows copies 15 floats per row: anything else is an
                    raise ValueError("each target must be [G, 15] (x1 y1 x2 y2, 10 landmark coords, label)")   # out-of-bounds read
                t = t.contiguous()
                keep.append(t)
            rows[i] = t.data_ptr()
            counts[i] = t.shape[0]
        total = int(_lib.lib().jabd_pack_gt_rows(rows, counts, self.B, ptr(sl["gt"]), self.cap, ptr(offs)))
        if total < 0:
            _lib.check(int(total), "jabd_pack_gt_rows")
        v0, v1 = _tensor.variances_of(variances)
        with torch.cuda.device(self.dev):
            _lib.call("jabd_assign_host", ptr(self.pri), self.P, ptr(sl["gt"]), ptr(offs), self.B, float(threshold),
                      v0, v1, int(label_mode), 1 if encode else 0,
                      (FLAG_DENSE if dense else 0) | FLAG_ASYNC | (FLAG_DEVICE_OUT if self.device_out else 0), ptr(sl["loc_t"]),
                      ptr(sl["conf_t"]), ptr(sl["landm_t"]), ptr(sl["scratch"]), sl["scratch"].numel(),
                      ctypes.c_void_p(sl["stream"].cuda_stream))
            sl["done"].record(sl["stream"])
        self.last_h2d = total * 15 * 4 + (self.B + 1) * 4
        self.last_d2h = 0 if self.device_out else self.B * self.P * (16 + 8 + (40 if self.with_landm else 0))
        return k

    def slot_stream(self, slot):
        return self.slots[slot]["stream"]

    def wait(self, slot):
        sl = self.slots[slot]
        sl["done"].synchronize()
        return sl["loc_t"], sl["conf_t"], sl["landm_t"]

    def __call__(self, targets, **kw):
        return self.wait(self.submit(targets, **kw))


def assign_targets_host(priors, targets, **kw):
    """One-shot host-buffer call; for repeated use build a ``HostAssign`` once."""
    sum_g = sum(int(t.shape[0]) for t in targets)
    h = HostAssign(priors, len(targets), max(sum_g, 1), with_landm=kw.pop("with_landm", True))
    return h(targets, **kw)


class HostDetect(object):
    """End-to-end host-buffer path (``jabd_detect_host``) for a fixed (B, P, keep_cap) shape.  ``h(loc, conf, landm)`` is
    synchronous; ``submit`` / ``wait`` run the same call as a ``depth``-slot pipeline (``jabd_detect_host_async``, one
    stream + scratch + pinned output set per slot) so that the next batch's 44 MB upload overlaps this batch's NMS."""

    def __init__(self, priors, B, keep_topk=750, with_landm=True, device=None, depth=2):
        _tensor.require_cuda()
        self.dev = torch.device(device) if device is not None else _tensor.device_of(priors)
        self.pri = _tensor.to_dev(priors, self.dev)
        self.B, self.P = int(B), int(self.pri.shape[0])
        self.keep_cap = int(keep_topk) if keep_topk and keep_topk > 0 else self.P
        self.with_landm = with_landm
        L = _lib.lib()
        nbytes = L.jabd_detect_host_scratch_bytes(self.B, self.P, self.keep_cap, 1 if with_landm else 0)
        self.slots = []
        for _ in range(max(int(depth), 1)):
            self.slots.append(dict(scratch=_tensor.workspace(nbytes, self.dev),
                                   dets=torch.empty((self.B, self.keep_cap, 15), dtype=torch.float32).pin_memory(),
                                   counts=torch.empty((self.B,), dtype=torch.int32).pin_memory(),
                                   keep_idx=torch.empty((self.B, self.keep_cap), dtype=torch.int32).pin_memory(),
                                   stream=torch.cuda.Stream(self.dev), done=torch.cuda.Event()))
        cur = torch.cuda.current_stream(self.dev)
        for sl in self.slots:
            sl["stream"].wait_stream(cur)   # self.pri may still be in flight on the caller's stream
        self.next_slot = 0
        self.last_h2d = self.B * self.P * 4 * (4 + 2 + (10 if self.with_landm else 0))
        self.last_d2h = self.B * (self.keep_cap * (60 + 4) + 4)

    def submit(self, loc, conf, landm, variances=(0.1, 0.2), conf_thres=0.02, strict=True, pre_nms_topk=5000, nms_thres=0.4,
               cluster=0):
        """``loc``/``conf``/``landm``: contiguous CPU f32 tensors (pinned for the copies to overlap).  Returns the slot id.
        The call is asynchronous: the copies -- and, for pinned ``loc`` / ``landm``, the kernel itself, which reads candidate
        and kept rows in place from host memory -- run after ``submit`` returns, so the three input buffers must stay
        unmodified until ``wait(slot)``."""
        for t, shp in ((loc, (self.B, self.P, 4)), (conf, (self.B, self.P, 2))):
            if t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shp or t.dtype != torch.float32:
                raise ValueError("HostDetect expects contiguous CPU f32 tensors loc [B,P,4], conf [B,P,2], landm [B,P,10]")
        k = self.next_slot
        self.next_slot = (k + 1) % len(self.slots)
        sl = self.slots[k]
        sl["done"].synchronize()
        v0, v1 = _tensor.variances_of(variances)
        # pinned loc / landmarks are not uploaded: the kernel reads the <= pre_nms_topk candidate rows (16 B) and the
        # <= keep_cap kept rows (40 B) from host memory
        topk = int(pre_nms_topk) if pre_nms_topk else 0
        loc_bytes = self.B * topk * 16 if (topk > 0 and self.P >= 6 * topk and loc.is_pinned()) else self.B * self.P * 16
        lm_bytes = 0
        if self.with_landm and landm is not None:
            lm_bytes = self.B * self.keep_cap * 40 if landm.is_pinned() else self.B * self.P * 40
        self.last_h2d = self.B * self.P * 8 + loc_bytes + lm_bytes
        with torch.cuda.device(self.dev):
            _lib.call("jabd_detect_host_async", ptr(loc), ptr(conf), ptr(landm if self.with_landm else None), ptr(self.pri),
                      self.B, self.P, v0, v1, float(conf_thres), THRESH_GT if strict else THRESH_GE,
                      int(pre_nms_topk) if pre_nms_topk else 0, float(nms_thres), self.keep_cap, int(cluster), ptr(sl["dets"]),
                      ptr(sl["counts"]), ptr(sl["keep_idx"]), ptr(sl["scratch"]), sl["scratch"].numel(),
                      ctypes.c_void_p(sl["stream"].cuda_stream))
            sl["done"].record(sl["stream"])
        return k

    def wait(self, slot):
        sl = self.slots[slot]
        sl["done"].synchronize()
        return sl["dets"], sl["counts"], sl["keep_idx"]

    def __call__(self, loc, conf, landm, **kw):
        return self.wait(self.submit(loc, conf, landm, **kw))


def detect_host(loc, conf, landm, priors, **kw):
    h = HostDetect(priors, loc.shape[0], keep_topk=kw.pop("keep_topk", 750), with_landm=landm is not None)
    return h(loc, conf, landm, **kw)
