"""Development probe for sharding.PeerGather under torchrun: times the pieces of one exchange (CUDA events, max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 profiles/p2p_probe.py"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jabd_b200 import _lib, sharding  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pg = sharding.PeerGather(32, 750, dev, depth=3)
ct = ctypes


def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(fn, n=50):
    for _ in range(3):
        fn()
    bar()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    bar()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


st = lambda: ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
seq = [1000]


def scatter_only():
    seq[0] += 1
    _lib.call("jabd_p2p_allgather", ct.c_void_p(pg.send[0].data_ptr()), ct.c_size_t(pg.L * 4), pg.bufs[0], ct.c_size_t(pg.rank * pg.Lpad),
              pg.flags[0], None, None, pg.world, pg.rank, ct.c_uint64(seq[0]), ct.c_uint64(0), ct.c_void_p(pg.counters[0].data_ptr()), 2.0,
              None, st())


def scatter_wait():
    scatter_only()
    _lib.call("jabd_p2p_wait", ct.c_void_p(pg.own + pg.o_flags), pg.world, ct.c_uint64(seq[0]), 2.0, ct.c_void_p(pg.own + pg.o_status), st())


def signal_only():
    seq[0] += 1
    _lib.call("jabd_p2p_allgather", None, ct.c_size_t(0), pg.bufs[0], ct.c_size_t(0), pg.flags[0], None, None, pg.world, pg.rank,
              ct.c_uint64(seq[0]), ct.c_uint64(0), ct.c_void_p(pg.counters[0].data_ptr()), 2.0, None, st())


k = [0]


def full():
    s = k[0] % pg.depth
    k[0] += 1
    pg.acquire(s)
    pg.launch(s)
    pg.result(s)


rows = [("transport " + pg.transport, 0.0), ("scatter kernel alone", timed(scatter_only)), ("signal-only kernel", timed(signal_only)),
        ("scatter + wait, one stream", timed(scatter_wait)), ("PeerGather acquire/launch/result", timed(full))]
if world > 1:
    send = pg.send[0]
    recv = torch.empty((world, pg.L), dtype=torch.float32, device=dev)
    rows.append(("nccl all_gather_into_tensor", timed(lambda: dist.all_gather_into_tensor(recv, send))))
if rank == 0:
    for name, us in rows:
        print("%-40s %10.1f us" % (name, us), flush=True)
    print("status", pg.status())
bar()
pg.close()
if world > 1:
    dist.destroy_process_group()
