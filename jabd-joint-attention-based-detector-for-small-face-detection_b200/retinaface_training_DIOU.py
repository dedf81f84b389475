"""Drop-in for ``nets/retinaface_training_DIOU.py`` of the reference (R/nets/retinaface_training_DIOU.py): the DIoU
training variant -- ``match_iou`` (raw matched boxes as regression targets, :176-246), the element-wise
``bbox_overlaps_{iou,giou,diou,ciou}`` (:342-490, same functions as R/utils/box_utils.py:5-158), ``IouLoss`` (:491-525)
and the ``MultiBoxLoss`` whose box term is ``IouLoss(losstype="Diou")`` (:527-665).  Everything else of that file
(``point_form`` ... ``decode``) is shared with ``retinaface_training`` / ``utils_bbox``.
"""
import torch

from . import _lib, _ops, _tensor
from ._tensor import ptr
from .batched import LOC_LOSS, assign_targets, multibox_loss
from .box_utils import bbox_overlaps_ciou, bbox_overlaps_diou, bbox_overlaps_giou, bbox_overlaps_iou  # noqa: F401
from .retinaface_training import encode, encode_landm, intersect, jaccard, match, match_iou, point_form  # noqa: F401
from .utils_bbox import decode, decode_landm  # noqa: F401

__all__ = ["IouLoss", "MultiBoxLoss", "match_iou", "match", "decode", "bbox_overlaps_iou", "bbox_overlaps_giou",
           "bbox_overlaps_diou", "bbox_overlaps_ciou", "install"]


class _IouLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc_p, loc_t, priors, var0, var1, kind, size_sum):
        dev = loc_p.device
        lp = loc_p.detach().contiguous().float()
        n = int(lp.shape[0])
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("jabd_iou_loss_forward", ptr(lp), ptr(loc_t), ptr(priors), n, var0, var1, kind, 1 if size_sum else 0,
                      ptr(None), ptr(loss), _tensor.stream_of(dev))
        ctx.save_for_backward(lp, loc_t, priors if priors is not None else lp.new_empty(0))
        ctx.cfg = (var0, var1, kind, size_sum, priors is not None)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lp, loc_t, pri = ctx.saved_tensors
        var0, var1, kind, size_sum, has_pri = ctx.cfg
        dev = lp.device
        n = int(lp.shape[0])
        out = torch.empty_like(lp)
        gg = g.to(dev, torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            _lib.call("jabd_iou_loss_backward", ptr(lp), ptr(loc_t), ptr(pri if has_pri else None), n, var0, var1, kind,
                      1 if size_sum else 0, ptr(gg), ptr(out), _tensor.stream_of(dev))
        return out, None, None, None, None, None, None


class IouLoss(torch.nn.Module):
    """R/nets/retinaface_training_DIOU.py:491-525: ``sum(1 - overlap(decode(loc_p, prior_data), loc_t))`` (divided by the
    row count unless ``size_sum``); ``losstype`` 'Iou' | 'Giou' | 'Diou' | anything else = CIoU, like the reference."""

    def __init__(self, pred_mode='Center', size_sum=True, variances=None, losstype='Diou'):
        super(IouLoss, self).__init__()
        self.size_sum = size_sum
        self.pred_mode = pred_mode
        self.variances = variances
        self.loss = losstype

    def forward(self, loc_p, loc_t, prior_data):
        if not loc_p.is_cuda:
            raise RuntimeError("IouLoss needs CUDA tensors; there is no CPU path")
        dev = loc_p.device
        kind = LOC_LOSS.get(self.loss, _ops.CIOU)
        kind = kind if kind else _ops.CIOU
        lp = loc_p.reshape(-1, 4)
        n = int(lp.shape[0])
        lt = _tensor.to_dev(loc_t, dev).reshape(-1, 4)
        pri, v0, v1 = None, 0.0, 0.0
        if self.pred_mode == 'Center':
            pri = _tensor.to_dev(prior_data, dev).reshape(-1, 4)
            v0, v1 = _tensor.variances_of(self.variances)
        # the kernels index loc_t[i] and priors[i] for i < n: a row-count mismatch is the reference's broadcast error
        # (decode(loc_p, prior_data) :338 / torch.min(bboxes1, bboxes2) :354), not a device out-of-bounds read
        if int(lt.shape[0]) != n or (pri is not None and int(pri.shape[0]) != n):
            raise ValueError("IouLoss: loc_p has %d rows, loc_t %d%s -- they must agree" % (
                n, int(lt.shape[0]), "" if pri is None else ", prior_data %d" % int(pri.shape[0])))
        return _IouLossFn.apply(lp, lt, pri, v0, v1, kind, bool(self.size_sum))


class MultiBoxLoss(torch.nn.Module):
    """R/nets/retinaface_training_DIOU.py:527-665: same constructor and ``forward(predictions, priors, targets)`` as the
    reference; targets come from ``match_iou`` semantics (raw boxes), the box loss is DIoU on the decoded predictions."""

    def __init__(self, num_classes, overlap_thresh, neg_pos, variance, cuda=True):
        super(MultiBoxLoss, self).__init__()
        if int(num_classes) != 2:
            raise ValueError("MultiBoxLoss: num_classes must be 2 (face / background), like the reference's RetinaFace heads")
        self.num_classes = int(num_classes)
        self.threshold = overlap_thresh
        self.negpos_ratio = neg_pos
        self.variance = variance
        self.cuda = cuda
        self.gious = IouLoss(pred_mode='Center', size_sum=True, variances=self.variance, losstype="Diou")

    def forward(self, predictions, priors, targets):
        loc_data = predictions[0]
        pri = priors.data.to(loc_data.device)
        with torch.no_grad():
            loc_t, conf_t, landm_t = assign_targets(pri, [t.data for t in targets], threshold=self.threshold,
                                                    variances=self.variance, encode=False)
        return multibox_loss(predictions, loc_t, conf_t, landm_t, self.negpos_ratio, loc_loss=self.gious.loss, priors=pri,
                             variances=self.variance)


def install(module):
    """Point the reference module's globals at these implementations (no edit of ``nets/``)."""
    for name in ("point_form", "intersect", "jaccard", "encode", "encode_landm", "match", "match_iou", "decode",
                 "bbox_overlaps_iou", "bbox_overlaps_giou", "bbox_overlaps_diou", "bbox_overlaps_ciou", "IouLoss",
                 "MultiBoxLoss"):
        if hasattr(module, name):
            setattr(module, name, globals()[name])
    return module
