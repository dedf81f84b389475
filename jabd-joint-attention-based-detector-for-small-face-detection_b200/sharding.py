"""Image sharding across ranks.

Every image's target assignment and detection is a pure function of (priors, that image's GT or
predictions), so a batch shards by images with **no collective on the hot path**; priors are regenerated per
device.  Shards are cut by estimated cost (ragged GT counts), as an LPT bin-packing over images -- nothing downstream
needs contiguous ranges, a permutation index restores the global order (``lpt_shards`` / ``gather_order``).  The only
exchange is one fixed-shape all-gather of the padded detections (``dets [B_local, keep, 15]`` + ``counts [B_local]``)
for AP evaluation, replacing the reference's pickle -> pad -> two-all-gather pattern (R/utils.py:49-92);
``DetectionGather`` issues it on a side stream so that it overlaps the next batch's detection.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_range", "lpt_shards", "gather_order", "image_costs", "local_targets", "allgather_detections",
           "DetectionGather", "PeerGather", "world_info"]


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


def image_costs(targets, image_cost=IMAGE_COST_GT, segment_cost=SEGMENT_COST_GT):
    """Estimated assignment cost of every image in GT-equivalents: ``G + segment_cost * ceil(G / 64) + image_cost``."""
    return [int(t.shape[0]) + segment_cost * ((int(t.shape[0]) + SEGMENT_GT - 1) // SEGMENT_GT) + image_cost for t in targets]


def lpt_shards(weights, world, max_items=None):
    """Longest-processing-time bin packing: items by descending weight, each to the currently lightest rank (ties: lowest
    rank; equal weights keep their index order, so every rank derives the same partition).  Returns ``world`` ascending index
    lists that cover ``range(len(weights))`` once.  A contiguous prefix split cannot do better than the granularity of the
    largest image at a boundary; LPT is within 4/3 of the optimum and, for hundreds of images, within a fraction of a percent
    of perfectly even.  ``max_items`` caps the items per rank (fixed-shape consumers such as the detection all-gather)."""
    n, world = len(weights), int(world)
    if world <= 0:
        raise ValueError("world must be positive")
    if max_items is not None and max_items * world < n:
        raise ValueError("max_items * world < number of items")
    order = sorted(range(n), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        open_ranks = [r for r in range(world) if max_items is None or len(shards[r]) < max_items]
        r = min(open_ranks, key=lambda q: (load[q], q))
        shards[r].append(i)
        load[r] += float(weights[i])
    return [sorted(sh) for sh in shards]


def gather_order(shards):
    """Index ``perm`` with ``global[perm[k]] = gathered[k]`` for results concatenated in rank order (what an all-gather returns):
    ``out = torch.empty_like(gathered); out[perm] = gathered`` restores the global image order."""
    return [i for sh in shards for i in sh]


def local_targets(targets, rank=None, world=None, balance=True, image_cost=IMAGE_COST_GT, segment_cost=SEGMENT_COST_GT,
                  contiguous=False):
    """This rank's share of a list of per-image ``[G_i,15]`` targets and the global indices of its images.  ``balance``: shards
    of equal estimated cost (``image_costs``) by LPT bin packing; ``contiguous=True`` keeps the round-1 behaviour, a contiguous
    prefix split (returns the ``(lo, hi)`` range instead of an index list); ``balance=False``: equal image counts."""
    r, w = world_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    weights = image_costs(targets, image_cost, segment_cost) if balance else None
    if contiguous or not balance:
        lo, hi = shard_range(len(targets), rank, world, weights)
        return targets[lo:hi], ((lo, hi) if contiguous else list(range(lo, hi)))
    mine = lpt_shards(weights, world)[rank]
    return [targets[i] for i in mine], mine


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


class DetectionGather(object):
    """The validation flow's exchange, off the critical path ("keep it off the critical path (separate stream)").

    ``depth`` slots, each one flat send buffer ``[B*keep*15 floats | B int32 counts]`` whose views ``dets(slot)
    [B,keep,15]`` / ``counts(slot) [B]`` the detect kernel writes directly (no packing kernel), and one receive buffer
    ``[world, L]``.  ``launch(slot)`` records an event on the producer stream and enqueues ONE ``all_gather_into_tensor``
    (NCCL) on a side stream behind it; the producer stream goes straight on to the next batch.  ``result(slot)`` makes the
    caller's stream wait for that slot's gather and returns views ``(dets [world,B,keep,15], counts [world,B])`` in rank
    order.  A slot may be refilled once its ``result`` has been consumed (or ``launch`` of the same slot waits for it)."""

    def __init__(self, b_local, keep, device, depth=2, group=None):
        self.B, self.keep, self.dev, self.group, self.depth = int(b_local), int(keep), torch.device(device), group, int(depth)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.transport = "nccl" if self.dev.type == "cuda" else "gloo"
        self.n_det = self.B * self.keep * 15
        self.L = self.n_det + self.B
        self.send = [torch.zeros((self.L,), dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.recv = [torch.empty((self.world, self.L), dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.side = torch.cuda.Stream(self.dev) if self.dev.type == "cuda" else None
        self.ready = [torch.cuda.Event() for _ in range(depth)] if self.side is not None else None
        self.done = [torch.cuda.Event() for _ in range(depth)] if self.side is not None else None
        self.work = [None] * depth

    def dets(self, slot):
        return self.send[slot][:self.n_det].view(self.B, self.keep, 15)

    def counts(self, slot):
        return self.send[slot][self.n_det:].view(torch.int32)

    def launch(self, slot):
        if self.world == 1:
            self.recv[slot][0].copy_(self.send[slot], non_blocking=True)
            return
        if self.side is None:                       # CPU tensors (gloo; tests): synchronous
            dist.all_gather(list(self.recv[slot].unbind(0)), self.send[slot], group=self.group)
            return
        cur = torch.cuda.current_stream(self.dev)
        self.ready[slot].record(cur)
        self.side.wait_event(self.ready[slot])
        with torch.cuda.stream(self.side):
            self.work[slot] = dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group, async_op=True)
            self.work[slot].wait()                  # orders the SIDE stream after the collective; the host does not block
            self.done[slot].record(self.side)

    def acquire(self, slot):
        """The caller's stream waits until the slot's send buffer may be refilled (its last gather has read it)."""
        if self.side is not None and self.world > 1 and self.work[slot] is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.done[slot])

    def result(self, slot):
        if self.side is not None and self.world > 1:
            torch.cuda.current_stream(self.dev).wait_event(self.done[slot])
        r = self.recv[slot]
        return (r[:, :self.n_det].view(self.world, self.B, self.keep, 15), r[:, self.n_det:].view(torch.int32))


class _RawCuda(object):
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather(object):
    """``DetectionGather``'s interface with the exchange done by this library's own kernel over peer memory
    (``jabd_p2p_allgather`` / ``jabd_p2p_wait``, csrc/p2p.cu) instead of NCCL: every rank stores its ``[B*keep*15 floats | B
    counts]`` block straight into slot ``rank`` of every rank's receive buffer -- one NVSwitch hop per peer instead of the
    ``world - 1`` latency-bound steps of a ring all-gather of a 1.4 MB message -- and raises a flag there.

    Set-up (once): one IPC-exportable allocation per rank holding ``depth`` receive buffers ``[world, L]`` plus the flag words,
    ``jabd_p2p_alloc``; the 64-byte handles are exchanged with ``all_gather_object`` and mapped with ``jabd_p2p_open``.
    If any rank cannot map its peers (no IPC in the container, no peer access) every rank falls back to the NCCL path of
    ``DetectionGather``; ``transport`` says which one runs ("p2p" or "nccl").

    ``launch(slot)``: on a side stream behind the producer stream -- acknowledge the slot's previous contents as read (flags),
    wait for every peer's acknowledgement, scatter.  ``result(slot)``: the caller's stream waits for every peer's flag of that
    slot's exchange (device side, no host sync) and gets views ``(dets [world,B,keep,15], counts [world,B])``; reads of them must
    be enqueued before the same slot is launched again.  ``status()`` is non-zero if a wait ever timed out (2 s)."""

    def __init__(self, b_local, keep, device, depth=3, group=None):
        import ctypes
        from . import _lib
        self._ct, self._lib = ctypes, _lib
        self.B, self.keep, self.dev, self.group, self.depth = int(b_local), int(keep), torch.device(device), group, int(depth)
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.n_det = self.B * self.keep * 15
        self.L = self.n_det + self.B
        self.Lpad = (self.L * 4 + 255) // 256 * 256                      # bytes per block in a receive buffer
        self.send = [torch.zeros((self.L,), dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.side = torch.cuda.Stream(self.dev)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.sent = [torch.cuda.Event() for _ in range(depth)]
        self.seq = [0] * depth
        self.fallback = None
        self.transport = "p2p"
        L = _lib.lib()
        if self.world > L_MAX_PEERS:
            raise ValueError("PeerGather supports at most %d ranks" % L_MAX_PEERS)
        # layout of the allocation: depth x [world x Lpad] | depth x 16 data flags | depth x 16 ack flags | status
        self.o_flags = depth * self.world * self.Lpad
        self.o_acks = self.o_flags + depth * 16 * 8
        self.o_status = self.o_acks + depth * 16 * 8
        nbytes = self.o_status + 256
        with torch.cuda.device(self.dev):
            own = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            _lib.call("jabd_p2p_alloc", ctypes.c_size_t(nbytes), ctypes.byref(own), handle)
            self.own = own.value
            self.bases = [None] * self.world
            self.bases[self.rank] = self.own
            ok = True
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
                for j in range(self.world):
                    if j == self.rank:
                        continue
                    p = ctypes.c_void_p()
                    h = (ctypes.c_ubyte * 64).from_buffer_copy(handles[j])
                    if L.jabd_p2p_open(h, ctypes.byref(p)) != 0:
                        ok = False
                        self.open_error = _lib.last_error()
                        break
                    self.bases[j] = p.value
                oks = [None] * self.world
                dist.all_gather_object(oks, ok, group=group)
                ok = all(oks)
            if not ok:
                self.close()
                self.transport = "nccl"
                self.fallback = DetectionGather(b_local, keep, device, depth=min(depth, 2) if depth > 1 else 1, group=group)
                self.send = self.fallback.send
                self.depth = len(self.send)
                return
        self.counters = [torch.zeros((16,), dtype=torch.int32, device=self.dev) for _ in range(depth)]
        self.view = torch.as_tensor(_RawCuda(self.own, nbytes), device=self.dev)
        vp = ctypes.c_void_p * self.world
        self.bufs = [vp(*[b + s * self.world * self.Lpad for b in self.bases]) for s in range(depth)]
        self.flags = [vp(*[b + self.o_flags + s * 128 for b in self.bases]) for s in range(depth)]
        self.acks = [vp(*[b + self.o_acks + s * 128 for b in self.bases]) for s in range(depth)]
        # argument objects built once: launch / result are on the per-batch path
        self._send_ptr = [ctypes.c_void_p(t.data_ptr()) for t in self.send]
        self._cnt_ptr = [ctypes.c_void_p(t.data_ptr()) for t in self.counters]
        self._own_acks = [ctypes.c_void_p(self.own + self.o_acks + s * 128) for s in range(depth)]
        self._own_flags = [ctypes.c_void_p(self.own + self.o_flags + s * 128) for s in range(depth)]
        self._status_ptr = ctypes.c_void_p(self.own + self.o_status)
        self._nbytes = ctypes.c_size_t(self.L * 4)
        self._dst_off = ctypes.c_size_t(self.rank * self.Lpad)

    def dets(self, slot):
        return self.send[slot][:self.n_det].view(self.B, self.keep, 15)

    def counts(self, slot):
        return self.send[slot][self.n_det:].view(torch.int32)

    def _st(self, stream):
        return self._ct.c_void_p(stream.cuda_stream)

    def launch(self, slot):
        if self.fallback is not None:
            return self.fallback.launch(slot)
        ct, call = self._ct, self._lib.call
        cur = torch.cuda.current_stream(self.dev)
        self.ready[slot].record(cur)       # the send buffer is written AND this stream's reads of the slot's old contents are done
        self.side.wait_event(self.ready[slot])
        e = self.seq[slot] + 1
        self.seq[slot] = e
        with torch.cuda.device(self.dev):
            # one kernel: "slot read" to every peer, wait for theirs (nobody overwrites a block that is still being read), scatter
            call("jabd_p2p_allgather", self._send_ptr[slot], self._nbytes, self.bufs[slot], self._dst_off, self.flags[slot],
                 self.acks[slot], self._own_acks[slot], self.world, self.rank, ct.c_uint64(e), ct.c_uint64(e - 1), self._cnt_ptr[slot],
                 2.0, self._status_ptr, self._st(self.side))
            self.sent[slot].record(self.side)

    def acquire(self, slot):
        """The caller's stream waits until the slot's send buffer may be refilled (its last scatter has read it)."""
        if self.fallback is not None:
            return self.fallback.acquire(slot)
        if self.seq[slot] > 0:
            torch.cuda.current_stream(self.dev).wait_event(self.sent[slot])

    def result(self, slot):
        if self.fallback is not None:
            return self.fallback.result(slot)
        ct = self._ct
        with torch.cuda.device(self.dev):
            self._lib.call("jabd_p2p_wait", self._own_flags[slot], self.world, ct.c_uint64(self.seq[slot]), 2.0, self._status_ptr,
                           self._st(torch.cuda.current_stream(self.dev)))
        o = slot * self.world * self.Lpad
        blk = self.view[o:o + self.world * self.Lpad].view(torch.float32).view(self.world, self.Lpad // 4)   # row stride Lpad bytes
        return blk[:, :self.n_det].view(self.world, self.B, self.keep, 15), blk[:, self.n_det:self.L].view(torch.int32)

    def status(self):
        if self.fallback is not None:
            return 0
        torch.cuda.synchronize(self.dev)
        return int(self.view[self.o_status:self.o_status + 4].view(torch.int32).item())

    def close(self):
        """Unmap the peers and free the own allocation (every rank, after a barrier: peers may still be storing)."""
        L = self._lib.lib()
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize(self.dev)
            for j, b in enumerate(getattr(self, "bases", [])):
                if b is not None and j != self.rank:
                    L.jabd_p2p_close(self._ct.c_void_p(b))
            if getattr(self, "own", None):
                self.view = None
                L.jabd_p2p_free(self._ct.c_void_p(self.own))
                self.own = None
            self.bases = []


L_MAX_PEERS = 16
