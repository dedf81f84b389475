"""Drop-in for the box-geometry functions of ``nets/retinaface_training.py`` (R/nets/retinaface_training.py:8-162)
and the ``match_iou`` variant of ``nets/retinaface_training_DIOU.py`` (:176-246).

Same names, argument order and in-place row-write behaviour; every function runs a hand-written sm_100a
kernel from ``libjabd_b200.so``.  ``MultiBoxLoss.forward`` resolves ``match`` through its module globals at
call time (R/nets/retinaface_training.py:214), so::

    import nets.retinaface_training as ref
    from jabd_b200 import retinaface_training as fast
    fast.install(ref)            # ref.match = fast.match (and friends); nets/ itself is untouched

``assign_batch`` is the batched replacement of the loop at :197-227.
"""
from . import _ops
import torch

from .batched import assign_targets, multibox_loss

__all__ = ["point_form", "intersect", "jaccard", "encode", "encode_landm", "match", "match_iou", "assign_batch", "install",
           "MultiBoxLoss"]


def point_form(boxes):
    """(cx,cy,w,h) -> (x1,y1,x2,y2); R/nets/retinaface_training.py:8-10."""
    return _ops.point_form(boxes)


def intersect(box_a, box_b):
    """Intersection areas [A,B]; R/nets/retinaface_training.py:22-39."""
    return _ops.intersect(box_a, box_b)


def jaccard(box_a, box_b):
    """Dense IoU [A,B] of point-form boxes; R/nets/retinaface_training.py:41-59."""
    return _ops.jaccard(box_a, box_b)


def encode(matched, priors, variances):
    """R/nets/retinaface_training.py:61-70."""
    return _ops.encode(matched, priors, variances)


def encode_landm(matched, priors, variances):
    """R/nets/retinaface_training.py:72-84."""
    return _ops.encode_landm(matched, priors, variances)


def match(threshold, truths, priors, variances, labels, landms, loc_t, conf_t, landm_t, idx):
    """R/nets/retinaface_training.py:93-162: writes rows ``idx`` of loc_t / conf_t / landm_t, returns None."""
    _ops.match_one(threshold, truths, priors, variances, labels, landms, loc_t, conf_t, landm_t, idx, 0, 1)


def match_iou(threshold, truths, priors, variances, labels, landms, loc_t, conf_t, landm_t, idx):
    """R/nets/retinaface_training_DIOU.py:176-246: like ``match`` but ``loc_t[idx]`` holds the raw matched boxes."""
    _ops.match_one(threshold, truths, priors, variances, labels, landms, loc_t, conf_t, landm_t, idx, 0, 0)


def assign_batch(threshold, targets, priors, variances):
    """The whole loop of ``MultiBoxLoss.forward`` (R/nets/retinaface_training.py:197-227) in three launches:
    ``targets`` is the list of ``[G_i,15]`` tensors the data loader yields; returns CUDA
    ``(loc_t [B,P,4], conf_t [B,P] int64, landm_t [B,P,10])``."""
    return assign_targets(priors, targets, threshold=threshold, variances=variances)


class MultiBoxLoss(torch.nn.Module):
    """Drop-in for ``MultiBoxLoss`` (R/nets/retinaface_training.py:165-303): same constructor, same
    ``forward(predictions, priors, targets) -> (loss_l, loss_c, loss_landm)``.  Target assignment is the three
    batched launches of ``assign_batch``; positives, hard-negative mining (:259-281) and the loss sums (:229-301)
    are ``batched.multibox_loss`` -- targets never leave the GPU and nothing is sorted."""

    def __init__(self, num_classes, overlap_thresh, neg_pos, variance, cuda=True):
        super(MultiBoxLoss, self).__init__()
        if int(num_classes) != 2:
            raise ValueError("MultiBoxLoss: num_classes must be 2 (face / background), like the reference's RetinaFace heads")
        self.num_classes = int(num_classes)
        self.threshold = overlap_thresh
        self.negpos_ratio = neg_pos
        self.variance = variance
        self.cuda = cuda  # kept for signature parity; the computation always runs on the predictions' CUDA device

    def forward(self, predictions, priors, targets):
        loc_data = predictions[0]
        with torch.no_grad():
            loc_t, conf_t, landm_t = assign_targets(priors.data.to(loc_data.device), [t.data for t in targets],
                                                    threshold=self.threshold, variances=self.variance)
        return multibox_loss(predictions, loc_t, conf_t, landm_t, self.negpos_ratio)


def install(module):
    """Point the reference module's globals at these implementations (no edit of ``nets/``)."""
    for name in ("point_form", "intersect", "jaccard", "encode", "encode_landm", "match"):
        setattr(module, name, globals()[name])
    if hasattr(module, "match_iou"):
        module.match_iou = match_iou
    if hasattr(module, "MultiBoxLoss"):
        module.MultiBoxLoss = MultiBoxLoss
    return module
