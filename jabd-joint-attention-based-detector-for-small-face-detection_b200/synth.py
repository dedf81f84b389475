"""Seeded synthetic inputs for the box-geometry hot path.

Everything here is *data synthesis*, not product code and not the oracle: the
generators draw from a CPU ``torch.Generator`` so that the same seed gives the
same bytes in the dev container, in the tests and on the GPU box.

GT layout is the reference data loader's ``[G, 15]`` row
(R/utils/dataloader.py:37-58): x1 y1 x2 y2, five landmark (x, y) pairs,
label (+1 = has landmarks, -1 = landmarks zeroed), all normalised to [0, 1].
"""
import math

import torch

__all__ = [
    "gt_count", "make_gt", "make_gt_batch", "pack_gt", "make_preds_random",
    "make_preds_clustered", "dense_iou_for_synthesis", "make_eval_image",
]


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def gt_count(cfg_id, seed_gen):
    """Face count for one image: cfg1 50, cfg2/5 skewed 1..300, cfg4 1500."""
    if cfg_id == 1:
        return 50
    if cfg_id == 4:
        return 1500
    u = torch.rand((), generator=seed_gen).item()
    return 1 + int(math.floor(299.0 * u * u))


def make_gt(cfg_id, image_idx, image_size, count=None, side_px=None, max_count=None):
    """One image's ``[G, 15]`` fp32 target array ("tiny-face" distribution).

    side ~ logU[lo, hi] px, aspect h/w ~ U[1, 1.4], centre uniform with the box
    fully inside the image, boxes with w or h <= 1 px rejected
    (R/utils/dataloader.py:143), five landmarks uniform inside the box, label +1
    w.p. 0.5 else -1 with landmarks zeroed (R/utils/dataloader.py:55-58,145).
    """
    H, W = int(image_size[0]), int(image_size[1])
    g = _gen(1000 * int(cfg_id) + int(image_idx))
    G = gt_count(cfg_id, g) if count is None else int(count)
    if max_count is not None:
        G = min(G, int(max_count))
    if side_px is None:
        side_px = (4.0, 128.0) if cfg_id == 4 else (4.0, 64.0)
    lo, hi = float(side_px[0]), float(side_px[1])
    chunks, have = [], 0
    while have < G:
        n = max(2 * (G - have), 8)
        u = torch.rand((n, 5), generator=g, dtype=torch.float64)
        w = torch.exp(math.log(lo) + u[:, 0] * (math.log(hi) - math.log(lo)))
        h = w * (1.0 + 0.4 * u[:, 1])
        w = torch.clamp(w, max=W - 1.0)
        h = torch.clamp(h, max=H - 1.0)
        x1 = u[:, 2] * (W - w)
        y1 = u[:, 3] * (H - h)
        lm = torch.rand((n, 10), generator=g, dtype=torch.float64)
        row = torch.zeros((n, 15), dtype=torch.float64)
        row[:, 0] = x1 / W
        row[:, 1] = y1 / H
        row[:, 2] = (x1 + w) / W
        row[:, 3] = (y1 + h) / H
        row[:, 4:14:2] = (x1[:, None] + lm[:, 0::2] * w[:, None]) / W
        row[:, 5:14:2] = (y1[:, None] + lm[:, 1::2] * h[:, None]) / H
        has_lm = u[:, 4] < 0.5
        row[:, 14] = torch.where(has_lm, 1.0, -1.0).to(torch.float64)
        row[~has_lm, 4:14] = 0.0
        row = row[(w > 1.0) & (h > 1.0)]
        chunks.append(row)
        have += row.shape[0]
    rows = torch.cat(chunks, 0)[:G]
    return rows.to(torch.float32).reshape(G, 15).contiguous()


def make_gt_batch(cfg_id, batch, image_size, first_image=0, **kw):
    """List of per-image ``[G_i, 15]`` targets, as ``detection_collate`` returns
    them (R/utils/dataloader.py:177-186)."""
    return [make_gt(cfg_id, first_image + i, image_size, **kw) for i in range(batch)]


def pack_gt(targets):
    """Ragged list -> (``gt_packed [sumG, 15]`` fp32, ``gt_offsets [B+1]`` int32)."""
    offs = [0]
    for t in targets:
        offs.append(offs[-1] + int(t.shape[0]))
    packed = torch.cat([t.reshape(-1, 15) for t in targets], 0).contiguous() if targets else torch.zeros(0, 15)
    return packed.to(torch.float32), torch.tensor(offs, dtype=torch.int32)


def make_preds_random(cfg_id, image_idx, num_priors):
    """Gen A: loc ~ N(0, 0.5^2), score = u^8, landm ~ N(0, 1), conf = [1-s, s]."""
    g = _gen(2000 * int(cfg_id) + int(image_idx))
    loc = torch.randn((num_priors, 4), generator=g) * 0.5
    s = torch.rand((num_priors,), generator=g) ** 8
    landm = torch.randn((num_priors, 10), generator=g)
    conf = torch.stack([1.0 - s, s], 1).contiguous()
    return loc.contiguous(), conf, landm.contiguous()


def make_logits(cfg_id, image_idx, num_priors):
    """Training-time network outputs for one image (inputs of MultiBoxLoss.forward, R/nets/retinaface_training.py:183):
    loc ~ N(0, 0.5^2) [P,4], class logits ~ N(0, 1.5^2) [P,2], landm ~ N(0, 1) [P,10].  Seed 3000*cfg_id + image_idx."""
    g = _gen(3000 * int(cfg_id) + int(image_idx))
    loc = torch.randn((num_priors, 4), generator=g) * 0.5
    conf = torch.randn((num_priors, 2), generator=g) * 1.5
    landm = torch.randn((num_priors, 10), generator=g)
    return loc.contiguous(), conf.contiguous(), landm.contiguous()


def dense_iou_for_synthesis(truths, priors):
    """Plain broadcast IoU used only to *shape* the clustered predictions."""
    pf = torch.cat([priors[:, :2] - priors[:, 2:] / 2, priors[:, :2] + priors[:, 2:] / 2], 1)
    lt = torch.maximum(truths[:, None, :2], pf[None, :, :2])
    rb = torch.minimum(truths[:, None, 2:4], pf[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    aa = ((truths[:, 2] - truths[:, 0]) * (truths[:, 3] - truths[:, 1]))[:, None]
    ab = ((pf[:, 2] - pf[:, 0]) * (pf[:, 3] - pf[:, 1]))[None, :]
    return inter / (aa + ab - inter)


def make_preds_clustered(cfg_id, image_idx, priors, gt, variances=(0.1, 0.2), device=None):
    """Gen B: predictions clustered around the GT faces (primary NMS workload).

    For priors whose best-GT IoU > 0.1: ``loc = encode(gt, prior) + N(0, 0.05^2)``
    and ``score ~ U[0.3, 1)``; the rest is Gen A with ``score = 0.05 * u^8``.
    """
    P = int(priors.shape[0])
    g = _gen(2000 * int(cfg_id) + int(image_idx))
    loc = torch.randn((P, 4), generator=g) * 0.5
    s = 0.05 * torch.rand((P,), generator=g) ** 8
    landm = torch.randn((P, 10), generator=g)
    noise = torch.randn((P, 4), generator=g) * 0.05
    s_hi = 0.3 + 0.7 * torch.rand((P,), generator=g)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    pr = priors.to(dev, torch.float32)
    tr = gt[:, :4].to(dev, torch.float32)
    best, idx = dense_iou_for_synthesis(tr, pr).max(0)
    m = tr[idx]
    v0, v1 = float(variances[0]), float(variances[1])
    cxcy = ((m[:, :2] + m[:, 2:]) / 2 - pr[:, :2]) / (v0 * pr[:, 2:])
    wh = torch.log((m[:, 2:] - m[:, :2]) / pr[:, 2:]) / v1
    enc = torch.cat([cxcy, wh], 1).cpu()
    near = (best > 0.1).cpu()
    loc = torch.where(near[:, None], enc + noise, loc)
    s = torch.where(near, s_hi, s)
    conf = torch.stack([1.0 - s, s], 1).contiguous()
    return loc.contiguous(), conf, landm.contiguous()


def make_eval_image(cfg_id, image_idx, image_size=(640, 640), count=None, max_dets=750):
    """One image of a WIDER-style evaluation set, in the units of R/utils/utils_map.py: ``gt [G,4]`` float64 pixels
    (x y w h), ``keeps`` = three uint8 flag arrays (easy: side >= 32 px, medium: >= 16 px, hard: all -- nested like
    the WIDER subsets) and ``pred [N,5]`` float64 (x y w h score) sorted by descending score: jittered copies of most
    GT (some twice, so a GT is hit by several predictions), plus random false positives; a few images come out with no
    predictions or no GT."""
    import numpy as np
    H, W = int(image_size[0]), int(image_size[1])
    rng = np.random.default_rng(7000 * int(cfg_id) + int(image_idx))
    kind = int(image_idx) % 11
    G = 0 if kind == 7 else (int(count) if count is not None else int(1 + rng.integers(0, 60)))
    side = np.exp(rng.uniform(np.log(6.0), np.log(120.0), G))
    w, h = np.round(side), np.round(side * rng.uniform(1.0, 1.4, G))
    x, y = np.floor(rng.uniform(0, W - w)), np.floor(rng.uniform(0, H - h))
    gt = np.stack([x, y, w, h], 1).astype(np.float64).reshape(-1, 4)
    hard = np.ones(G, np.uint8)
    keeps = [(np.minimum(w, h) >= 32).astype(np.uint8), (np.minimum(w, h) >= 16).astype(np.uint8), hard]
    rows = []
    for g in range(G):
        for _ in range(int(rng.integers(0, 3))):
            j = rng.normal(0.0, 0.08, 4) * np.array([w[g], h[g], w[g], h[g]])
            rows.append([x[g] + j[0], y[g] + j[1], max(w[g] + j[2], 1.0), max(h[g] + j[3], 1.0), rng.uniform(0.3, 1.0)])
    for _ in range(int(rng.integers(0, 40))):
        s2 = np.exp(rng.uniform(np.log(6.0), np.log(120.0)))
        rows.append([rng.uniform(0, W - s2), rng.uniform(0, H - s2), s2, s2 * rng.uniform(1.0, 1.4), rng.uniform(0.02, 0.6)])
    pred = np.array(rows, dtype=np.float64).reshape(-1, 5)
    if kind == 3:
        pred = pred[:0]
    pred[:, :4] = np.round(pred[:, :4], 1)            # what a "%.1f" txt writer would leave
    pred[:, 4] = np.round(pred[:, 4], 8)
    if pred.shape[0] > 2 and kind == 5:
        pred[1, 4] = pred[0, 4]                        # tied scores
    pred = pred[np.argsort(-pred[:, 4], kind="stable")][:max_dets]
    return gt, keeps, pred
