"""Drop-in for ``utils/utils_bbox.py`` of the reference (R/utils/utils_bbox.py): ``decode``, ``decode_landm``,
``non_max_suppression`` (torchvision NMS semantics), ``nms_r`` (SSD-legacy greedy NMS) and
``retinaface_correct_boxes``.  ``predict.py`` imports these three names (R/predict.py:12-14); putting this
package's ``utils`` shim first on ``sys.path`` swaps them in.
"""
import numpy as np
import torch

from . import _ops, _tensor
from . import box_utils as _box_utils

__all__ = ["decode", "decode_landm", "non_max_suppression", "nms_r", "diounms", "retinaface_correct_boxes"]


def decode(loc, priors, variances):
    """R/utils/utils_bbox.py:29-34: [P,4] (or [B,P,4]) offsets -> x1y1x2y2."""
    return _ops.decode(loc, priors, variances)


def decode_landm(pre, priors, variances):
    """R/utils/utils_bbox.py:39-46."""
    return _ops.decode_landm(pre, priors, variances)


def non_max_suppression(detection, conf_thres=0.5, nms_thres=0.3):
    """R/utils/utils_bbox.py:260-296: rows with ``detection[:,4] >= conf_thres``, greedy NMS with
    torchvision semantics, kept rows in descending score order as a float32 numpy array ``[K, C]``;
    ``[]`` when nothing passes the threshold (:269-270).  One kernel launch + one D2H of the kept rows."""
    dev = _tensor.device_of(detection)
    det = _tensor.to_dev(detection, dev)
    if det.ndim != 2 or det.shape[1] < 5:
        raise ValueError("detection must be [N, >=5] (x1 y1 x2 y2 score ...)")
    n, c = int(det.shape[0]), int(det.shape[1])
    if n == 0:
        return []
    keep, cnt = _ops.nms_indices(det, c, det[:, 4], c, n, float(conf_thres), _ops.THRESH_GE, 0, float(nms_thres),
                                 _ops.NMS_TV, n, dev)
    count = int(cnt.item())
    if count <= 0:
        return []
    return det[keep[:count].long()].cpu().numpy()


def nms_r(boxes, scores, overlap=0.5, top_k=200):
    """R/utils/utils_bbox.py:116-180 (same function as ``box_utils.nms``)."""
    return _box_utils.nms(boxes, scores, overlap, top_k)


def diounms(boxes, scores, overlap=0.5, top_k=200, beta1=1.0):
    """R/utils/utils_bbox.py:182-258: greedy NMS on ``IoU - (d / c) ** beta1 <= overlap`` (DIoU-NMS).  Returns
    ``(keep, count)`` like the reference; the bare zero ``keep`` tensor when ``boxes`` is empty (:196-198)."""
    kind, dev = _tensor.kind_of(scores), _tensor.device_of(boxes, scores)
    b = _tensor.to_dev(boxes, dev)
    s = _tensor.to_dev(scores, dev).reshape(-1)
    n = int(s.shape[0])
    keep = torch.zeros((n,), dtype=torch.int64, device=dev)
    if b.numel() == 0:
        return _tensor.like(kind, keep)
    cap = min(n, int(top_k))
    k32, cnt = _ops.diounms_indices(b.reshape(n, 4), s, n, int(top_k), float(overlap), float(beta1), cap, dev)
    count = int(cnt.item())
    keep[:count] = k32[:count].to(torch.int64)
    return _tensor.like(kind, keep), count


def retinaface_correct_boxes(result, input_shape, image_shape):
    """R/utils/utils_bbox.py:9-24: undo the letterbox padding on the kept rows (host-side numpy on the
    [K,15] result; x columns use offset/scale index 1, y columns index 0)."""
    input_shape = np.array(input_shape, dtype=np.float64)
    image_shape = np.array(image_shape, dtype=np.float64)
    new_shape = image_shape * np.min(input_shape / image_shape)
    offset = (input_shape - new_shape) / 2. / input_shape
    scale = input_shape / new_shape
    off_xy, sc_xy = [offset[1], offset[0]], [scale[1], scale[0]]
    box_off, box_sc = off_xy * 2, sc_xy * 2
    lm_off, lm_sc = off_xy * 5, sc_xy * 5
    result[:, :4] = (result[:, :4] - np.array(box_off)) * np.array(box_sc)
    result[:, 5:] = (result[:, 5:] - np.array(lm_off)) * np.array(lm_sc)
    return result
