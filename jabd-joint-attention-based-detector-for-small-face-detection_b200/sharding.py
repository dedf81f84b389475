"""Image sharding across ranks.

Every image's target assignment and detection is a pure function of (priors, that image's GT or
predictions), so a batch shards by contiguous image ranges with **no collective on the hot path**; priors
are regenerated per device.  The only exchange is one fixed-shape all-gather of the padded detections
(``dets [B_local, keep, 15]`` + ``counts [B_local]``) for AP evaluation, replacing the reference's
pickle -> pad -> two-all-gather pattern (R/utils.py:49-92).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_range", "local_targets", "allgather_detections", "world_info"]


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, world, weights=None):
    """Boundaries ``[world+1]`` of contiguous shards.  Without ``weights``: equal counts (first shards take the
    remainder).  With ``weights`` (e.g. GT count per image): greedy prefix split so that each rank's weight is as
    close as possible to total/world -- matching cost is proportional to sum(G), not to the image count."""
    n_items, world = int(n_items), int(world)
    if world <= 0:
        raise ValueError("world must be positive")
    if weights is None:
        base, rem = divmod(n_items, world)
        b = [0]
        for r in range(world):
            b.append(b[-1] + base + (1 if r < rem else 0))
        return b
    w = [float(x) for x in weights]
    if len(w) != n_items:
        raise ValueError("weights must have one entry per item")
    total = sum(w)
    b, acc, i = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        # leave at least one item for every remaining rank when possible
        while i < n_items - (world - r) and acc + w[i] / 2.0 <= target:
            acc += w[i]
            i += 1
        b.append(i)
    b.append(n_items)
    return b


def shard_range(n_items, rank=None, world=None, weights=None):
    r, w = world_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    b = shard_bounds(n_items, world, weights)
    return b[rank], b[rank + 1]


# Cost model of one target-assignment call on a B200, fitted by profiles/cost_model_probe.py over batches of 16..48 images with
# 5..300 GT each:  t [us] = 0.241*B + 0.00373*sum(G) + 0.0517*sum(ceil(G/64)) + 16.5.  In GT-equivalents an image costs 65
# (prep + encode are per prior) and every started segment of 64 GT 14 (one work item per prior tile).
IMAGE_COST_GT = 65
SEGMENT_COST_GT = 14
SEGMENT_GT = 64


def local_targets(targets, rank=None, world=None, balance=True, image_cost=IMAGE_COST_GT, segment_cost=SEGMENT_COST_GT):
    """This rank's slice of a list of per-image ``[G_i,15]`` targets (+ its global image range).  ``balance``: contiguous
    shards of equal estimated cost ``sum(G_i + segment_cost * ceil(G_i / 64) + image_cost)`` instead of equal image counts."""
    weights = None
    if balance:
        weights = [int(t.shape[0]) + segment_cost * ((int(t.shape[0]) + SEGMENT_GT - 1) // SEGMENT_GT) + image_cost for t in targets]
    lo, hi = shard_range(len(targets), rank, world, weights)
    return targets[lo:hi], (lo, hi)


def allgather_detections(dets, counts, group=None):
    """All-gather fixed-shape padded detections.  ``dets [B_local, keep, 15]``, ``counts [B_local]`` must have
    the same shape on every rank (pad the last shard).  Returns ``(dets [world*B_local, keep, 15],
    counts [world*B_local])`` in rank order.  NCCL for CUDA tensors, gloo for CPU tensors."""
    if not (dist.is_available() and dist.is_initialized()):
        return dets, counts
    world = dist.get_world_size(group)
    if world == 1:
        return dets, counts
    dets = dets.contiguous()
    counts = counts.contiguous()
    if dets.dtype == torch.float32 and counts.dtype == torch.int32:
        # one collective instead of two (a 1.4 MB message is latency-bound on NVLink): the int32 counts travel bit-cast
        # as an extra float32 column of the flattened rows
        b = int(dets.shape[0])
        flat = torch.cat([dets.reshape(b, -1), counts.view(torch.float32).reshape(b, 1)], 1)
        gathered = torch.empty((world * b, flat.shape[1]), dtype=torch.float32, device=dets.device)
        if dets.is_cuda:
            dist.all_gather_into_tensor(gathered, flat, group=group)
        else:
            dist.all_gather(list(gathered.chunk(world, 0)), flat, group=group)
        return gathered[:, :-1].reshape((world * b,) + tuple(dets.shape[1:])), gathered[:, -1].contiguous().view(torch.int32)
    out_d = torch.empty((world * dets.shape[0],) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    out_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    if dets.is_cuda:
        dist.all_gather_into_tensor(out_d, dets, group=group)
        dist.all_gather_into_tensor(out_c, counts, group=group)
    else:
        dist.all_gather(list(out_d.chunk(world, 0)), dets, group=group)
        dist.all_gather(list(out_c.chunk(world, 0)), counts, group=group)
    return out_d, out_c
