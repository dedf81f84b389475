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
from .batched import assign_targets

__all__ = ["point_form", "intersect", "jaccard", "encode", "encode_landm", "match", "match_iou", "assign_batch", "install"]


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


def install(module):
    """Point the reference module's globals at these implementations (no edit of ``nets/``)."""
    for name in ("point_form", "intersect", "jaccard", "encode", "encode_landm", "match"):
        setattr(module, name, globals()[name])
    if hasattr(module, "match_iou"):
        module.match_iou = match_iou
    return module
