"""What the HOST side of this box gives N GPUs copying at the same time -- the question behind the end-to-end scaling of the
host-buffer entry points (jabd_assign_host moves 34.4 MB of targets per rank and step device -> pinned host memory).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/pcie_scaling_probe.py [--no-affinity]

Every rank copies the same 34.4 MB block (and, separately, 13.8 MB host -> device: the detect inputs) between its own GPU and its
own pinned buffer, all ranks starting together behind a barrier; CUDA events per rank, rank 0 prints per-rank and aggregate
GB/s for three kinds of host memory: cudaHostAlloc default, cudaHostAllocWriteCombined (D2H target only written by the
device, read later by the CPU), and torch's pin_memory().  No library code of this repo is involved: the numbers are the
ceiling any implementation of the host-buffer API meets on this box."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
cpus_before = sorted(os.sched_getaffinity(0))
if "--no-affinity" not in sys.argv and world > 1:
    per = len(cpus_before) // world
    if per >= 1:
        os.sched_setaffinity(0, cpus_before[local * per:(local + 1) * per])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
D2H, H2D = 2, 1


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
    assert rc == 0, "cudaHostAlloc failed: %d" % rc
    return p


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def gather(x):
    if world == 1:
        return [x]
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    out = torch.empty((world,), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t)
    return [float(v) for v in out.tolist()]


def timed(copy, reps=30):
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3):
        copy(st)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        copy(st)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / reps


out = {"world": world, "cpus_visible": len(cpus_before), "cpus_this_rank": len(os.sched_getaffinity(0)), "rows": []}
for name, nbytes, kind in (("d2h_targets_34.4MB", 34406400, D2H), ("h2d_detect_inputs_13.8MB", 13862400, H2D)):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for mem, flags in (("cudaHostAllocDefault", 0), ("cudaHostAllocWriteCombined", 4), ("torch.pin_memory", None)):
        if flags is None:
            keep = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hp = ctypes.c_void_p(keep.data_ptr())
        else:
            hp = host_alloc(nbytes, flags)
        dp = ctypes.c_void_p(d.data_ptr())
        if kind == D2H:
            ms = timed(lambda st: rt.cudaMemcpyAsync(hp, dp, nbytes, D2H, st))
        else:
            ms = timed(lambda st: rt.cudaMemcpyAsync(dp, hp, nbytes, H2D, st))
        per_rank = gather(nbytes / ms / 1e6)
        out["rows"].append({"copy": name, "host_memory": mem, "per_rank_GBps": [round(x, 1) for x in per_rank],
                            "aggregate_GBps": round(sum(per_rank), 1)})
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
