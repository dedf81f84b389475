#!/usr/bin/env python
"""Benchmark of the JABD box-geometry hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (torch port)

A *step* is one pass of target assignment (prior x GT IoU, both argmaxes, force-match, SSD encode) over one
batch of BASELINE.json configs[1]: 32 images per GPU at 640x640 (16,800 priors), 1..300 synthetic faces per image.
Weak scaling: N ranks process N*32 DISTINCT images per step, drawn from one pool of 256 images (= the cfg5 batch) at
every N, cut into shards of equal estimated cost (LPT bin packing over images) -- no collective on the path.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VAR = (0.1, 0.2)
THR = 0.35
IMAGE = (640, 640)
BATCH = 32            # images per GPU per step (cfg2; cfg5 = 256 over 8 GPUs)
LANES = 4             # side streams the steps of one graph are dealt over (jabd_assign_batches): independent batches overlap
DET_SETS = 8          # distinct batches per jabd_detect_batches call
ROUNDS = 4            # a lanes graph holds ROUNDS x SETS steps (the lanes drain at every graph boundary)
SETS = 8              # rotating buffer sets: 8 x ~45 MB of outputs+workspace > 126 MB L2
POOL = 256            # distinct images every N draws its global batches from (N = 8: one step = the whole pool = cfg5)
METRIC = "images/s for prior match+encode and decode+NMS @640^2 (16.8k priors), 1-8 GPU"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = tempfile.mktemp(prefix="jabd_clocks_", suffix=".csv")
        self.index = index
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        """windows: list of (t0, t1) wall-clock intervals under load."""
        import datetime
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[2]), float(f[3]), f[4], f[5:9]))
                except Exception:
                    continue
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        sel = [r for r in rows if any(a <= r[0] <= b for a, b in windows)]
        window = "timed+load-hold regions"
        if not sel:
            sel, window = rows, "whole run (no sample fell inside the timed regions)"
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "nvidia-smi unavailable"}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in sel for n, v in zip(names, r[4]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in sel), "sm_max_mhz": max(r[2] for r in sel), "reasons": reasons,
                "samples": len(sel), "window": window}


# ------------------------------------------------------------------------------------------ reference
def cpu_assign_rate(tp, torch, targets, pri, threads, budget_s, max_reps=200):
    """images/s of oracle/torch_port.assign_batch on `threads` host threads over repeated passes of `targets` (~budget_s)."""
    torch.set_num_threads(threads)
    tp.assign_batch(THR, targets[:1], pri, list(VAR))
    t0 = time.perf_counter()
    reps = 0
    while True:
        tp.assign_batch(THR, targets, pri, list(VAR))
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= max_reps:
            break
    return len(targets) * reps / (time.perf_counter() - t0), reps


def run_reference(args):
    """The reference's own CPU implementation of the path: the per-image loop of MultiBoxLoss.forward
    (R/nets/retinaface_training.py:197-214) restated op for op in torch (oracle/torch_port.py; the reference is
    Python and cannot travel to the GPU box), on all host threads.  Each step is a bounded sample of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from jabd_b200 import config, synth
    from oracle import oracle as orc
    from oracle import torch_port as tp
    cores = host_cores()
    torch.set_num_threads(cores)
    pri = torch.from_numpy(orc.priors(config.cfg_mnet, IMAGE))
    targets = synth.make_gt_batch(2, BATCH, IMAGE)
    t0 = time.perf_counter()
    tp.assign_batch(THR, targets[:2], pri, list(VAR))
    t_img = (time.perf_counter() - t0) / 2
    budget = 150.0
    per_step = int(max(1, min(BATCH, budget / max(args.steps + args.warmup, 1) / max(t_img, 1e-4))))
    sample = targets[:per_step]
    for _ in range(args.warmup):
        tp.assign_batch(THR, sample, pri, list(VAR))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tp.assign_batch(THR, sample, pri, list(VAR))
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    one_thread, reps1 = cpu_assign_rate(tp, torch, targets[:4], pri, 1, 4.0)
    desc = "first %d of the %d images of the cfg2 batch per step, torch %s CPU, %d threads" % (per_step, BATCH, torch.__version__, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 training target assignment (match+encode): 640x640, 16800 priors, 1..300 GT/image; "
                               "CPU sample of %d images per step" % per_step, "global_batch": per_step, "image": list(IMAGE),
                   "priors": int(pri.shape[0])},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": desc,
                         "one_thread": {"value": one_thread, "unit": "images/s", "cores": 1,
                                        "sample": "%d x the first 4 images, torch.set_num_threads(1)" % reps1},
                         "os_cpu_count": os.cpu_count(), "sched_getaffinity": cores},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------- main arm
def emit(line):
    """Print the one JSON line on the real stdout (see main(): fd 1 is pointed at stderr while the run lasts, so that
    banners printed by C libraries -- NCCL's version line, for one -- cannot land in front of it)."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    sys.stdout.flush()
    os.write(fd, data)


_REAL_STDOUT = None


def bind_cpus(local, local_world):
    """One disjoint slice of the host's vCPUs per rank (eight ranks submitting a graph launch every ~40 us and driving PCIe
    copies otherwise share -- and migrate across -- the same cores).  Returns (cpus of this rank, cpus visible before)."""
    if not hasattr(os, "sched_getaffinity"):
        return None, None
    cpus = sorted(os.sched_getaffinity(0))
    per = len(cpus) // max(local_world, 1)
    if local_world > 1 and per >= 1:
        mine = cpus[local * per:(local + 1) * per]
        try:
            os.sched_setaffinity(0, mine)
            return mine, cpus
        except Exception:
            pass
    return cpus, cpus


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the detect / dense / loss / cfg1 / cfg4 / cfg5 side measurements")
    ap.add_argument("--only-cfg5", action="store_true", help="of the side measurements keep the cfg5 validation flow only (probing)")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind each rank to its own slice of the host's vCPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    my_cpus, all_cpus = (None, None) if args.no_affinity else bind_cpus(local, local_world)   # before torch spawns its threads

    import ctypes
    import torch
    import torch.distributed as dist
    from jabd_b200 import _lib, _tensor, anchors, batched, config, sharding, synth
    from jabd_b200._tensor import ptr

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    L = _lib.lib()

    def cur_stream():
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = torch.empty((world,), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.tolist()]

    last_rank_ms = [None]

    def timed_loop(fn, n):
        """n calls of fn(k) between two events on the current stream, a barrier + device sync on both sides; returns the
        slowest rank's milliseconds (this rank's own in last_rank_ms) and the wall-clock window."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record()
        for k in range(n):
            fn(k)
        e1.record()
        barrier()
        last_rank_ms[0] = e0.elapsed_time(e1)
        return max_over_ranks(last_rank_ms[0]), (t0, time.time())

    # ---- workload: every N draws its global batches (N*32 images per step) from the same pool of 256 distinct images, so
    # the per-image GT statistics -- hence the work per image -- are identical at every N; set s of a rank is its LPT shard
    # (equal estimated cost, sharding.lpt_shards) of global batch s.  N = 1: the shard is the whole batch (the cfg2 batches
    # s*32 .. s*32+31); N = 8: every step is the whole pool = BASELINE configs[4], 256 distinct images over 8 ranks.
    pri = anchors.Anchors(config.cfg_mnet, image_size=IMAGE).get_anchors()
    P = int(pri.shape[0])
    pool = synth.make_gt_batch(2, POOL, IMAGE)
    costs = sharding.image_costs(pool)

    def global_batch(s):
        return [(s * world * BATCH + j) % POOL for j in range(world * BATCH)]

    def shard_of(s, r):
        g = global_batch(s)
        return [g[i] for i in sharding.lpt_shards([costs[i] for i in g], world)[r]]

    def make_set(images):
        tg = [pool[i] for i in images]
        gt, offs, _ = batched.pack_targets(tg, dev)
        nb, sumG = len(tg), int(gt.shape[0])
        return dict(images=images, host=tg, gt=gt, offs=offs, sumG=sumG, B=nb,
                    ws=_tensor.workspace(L.jabd_assign_workspace_bytes(nb, P, sumG), dev),
                    loc=torch.empty((nb, P, 4), dtype=torch.float32, device=dev),
                    conf=torch.empty((nb, P), dtype=torch.int64, device=dev),
                    landm=torch.empty((nb, P, 10), dtype=torch.float32, device=dev))

    sets = [make_set(shard_of(s, rank)) for s in range(SETS)]
    mean_g = sum(s["sumG"] for s in sets) / float(sum(s["B"] for s in sets))

    def assign(s, flags=0):
        _lib.call("jabd_assign", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1, flags,
                  ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), cur_stream())

    def phase_match(s, flags=0):
        _lib.call("jabd_assign_match", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], flags, ptr(s["ws"]),
                  s["ws"].numel(), cur_stream())

    def phase_encode(s):
        _lib.call("jabd_assign_encode", ptr(pri), P, ptr(s["gt"]), ptr(s["offs"]), s["B"], s["sumG"], THR, VAR[0], VAR[1], 0, 1,
                  ptr(s["loc"]), ptr(s["conf"]), ptr(s["landm"]), None, None, None, None, ptr(s["ws"]), s["ws"].numel(), cur_stream())

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def step_graphs(ss, lanes_n):
        """One graph per buffer set (a step = 3 kernel launches) and one graph holding a step of every set back to back: the
        timed loop replays the long one while >= SETS steps remain, so that the host's launch cadence (one cudaGraphLaunch
        per ~40 us step, from N processes on shared vCPUs) is not what is measured."""
        for s_ in ss:
            assign(s_)
        torch.cuda.synchronize(dev)
        singles = [capture(lambda s_=s_: assign(s_)) for s_ in ss]
        chunk = capture(lambda: assign_batches(ss * (ROUNDS if lanes_n else 1), lanes_n))   # set s always lands on lane s % LANES
        # the steps that do not fill a whole graph (K = 20 is what the driver runs) go through the same call on the same lanes:
        # one graph per remainder that will be asked for, captured here, outside every timed region
        tails = {}
        if lanes_n:
            for m in sorted(set(x % (SETS * ROUNDS) for x in (W, K)) - {0}):
                tails[m] = capture(lambda m=m: assign_batches((ss * ROUNDS)[:m], lanes_n))
        # first launch of an instantiated graph uploads it to the device: done here, not inside a timed region
        for g_ in singles + [chunk] + list(tails.values()):
            g_.replay()
        torch.cuda.synchronize(dev)
        return singles, chunk, tails

    def assign_batches(ss, n_lanes):
        """jabd_assign_batches: the steps of ss (independent batches, own outputs and workspaces) on n_lanes side streams,
        forked from and joined back into the current stream -- n_lanes = 0: back to back on the current stream."""
        arr = (_lib.AssignBatch * len(ss))(*[
            _lib.AssignBatch(s_["gt"].data_ptr(), s_["offs"].data_ptr(), s_["B"], s_["sumG"], s_["loc"].data_ptr(), s_["conf"].data_ptr(),
                             s_["landm"].data_ptr(), s_["ws"].data_ptr(), s_["ws"].numel()) for s_ in ss])
        ls = batched.lanes(dev, n_lanes)
        la = (ctypes.c_void_p * max(len(ls), 1))(*[x.cuda_stream for x in ls])
        _lib.call("jabd_assign_batches", ptr(pri), P, ctypes.cast(arr, ctypes.c_void_p), len(ss), THR, VAR[0], VAR[1], 0, 1, 0,
                  ctypes.cast(la, ctypes.c_void_p), len(ls), cur_stream())

    singles, chunk, tails = step_graphs(sets, LANES)
    _, serial_chunk, _ = step_graphs(sets, 0)

    def run_steps(n, singles=singles, chunk=chunk, tails=tails):
        k = 0
        while n - k >= SETS * ROUNDS:
            chunk.replay()
            k += SETS * ROUNDS
        if n - k in tails:
            tails[n - k].replay()
            k = n
        while k < n:
            singles[k % SETS].replay()
            k += 1

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []
    run_steps(W)
    ms, win = timed_loop(lambda k: run_steps(K) if k == 0 else None, 1)
    windows.append(win)
    rank_ms = per_rank(last_rank_ms[0])                     # every rank's own device time for the K steps
    rank_images = per_rank(float(sum(s_["B"] for s_ in sets)) / SETS)
    rank_gt = per_rank(float(sum(s_["sumG"] for s_ in sets)) / SETS)
    rank_cost = per_rank(float(sum(costs[i] for s_ in sets for i in s_["images"])) / SETS)
    value = world * BATCH * K / (ms / 1e3)
    # hold the same load for ~1.5 s so that the 50 ms clock sampler sees the GPU under this workload
    hold = max(int(1.5e3 / max(ms / K, 1e-3)), SETS)
    hold = (hold + SETS * ROUNDS - 1) // (SETS * ROUNDS) * (SETS * ROUNDS)      # whole graphs
    _, win = timed_loop(lambda k: run_steps(hold) if k == 0 else None, 1)
    windows.append(win)

    # ---- N > 1: (a) what a rank computed for its shard is what one GPU computes for the same images (per-image signature:
    # positives, label sum, bit-pattern sums of loc_t and landm_t), checked on rank 0 for every image of global batch 0;
    # (b) scaling control: the N = 1 batch replicated on every rank (identical work), to separate imbalance from the rest.
    def signature(loc, conf, landm):
        return torch.stack([(conf != 0).sum(1), conf.sum(1), loc.view(torch.int32).long().sum((1, 2)),
                            landm.view(torch.int32).long().sum((1, 2))], 1).cpu()

    shard_check, control = None, None
    if world > 1:
        assign(sets[0])
        mine = (sets[0]["images"], signature(sets[0]["loc"], sets[0]["conf"], sets[0]["landm"]).tolist())
        got = [None] * world
        dist.all_gather_object(got, mine)
        if rank == 0:
            g0 = global_batch(0)
            full = make_set(g0)
            assign(full)
            want = {img: row for img, row in zip(g0, signature(full["loc"], full["conf"], full["landm"]).tolist())}
            seen, bad = [], []
            for r, (imgs, rows) in enumerate(got):
                for img, row in zip(imgs, rows):
                    seen.append(img)
                    if want[img] != row:
                        bad.append((r, img))
            covered = sorted(seen) == sorted(g0)
            shard_check = {"images": len(g0), "equal": not bad and covered, "covered_once": covered, "mismatches": bad[:8],
                           "what": "per-image (positives, sum conf_t, bit sums of loc_t and landm_t) of every rank's LPT shard of global "
                                   "batch 0 == the same images assigned on rank 0's GPU alone",
                           "positives_total": int(sum(row[0] for row in want.values()))}
            del full
            assert shard_check["equal"], "sharded result differs from the single-GPU result: %s" % (shard_check,)
        ctrl_sets = [make_set(list(range(s * BATCH, (s + 1) * BATCH))) for s in range(SETS)]
        c_singles, c_chunk, c_tails = step_graphs(ctrl_sets, LANES)
        run_steps(W, c_singles, c_chunk, c_tails)
        ms_c, _ = timed_loop(lambda k: run_steps(K, c_singles, c_chunk, c_tails) if k == 0 else None, 1)
        control = {"value": world * BATCH * K / (ms_c / 1e3), "unit": "images/s", "ms_per_step": ms_c / K,
                   "per_rank_ms_per_step": [x / K for x in per_rank(last_rank_ms[0])],
                   "what": "identical work on every rank: the N = 1 batches (images s*32..s*32+31) once per rank -- round 1's weak-"
                           "scaling workload, kept as the control that separates shard imbalance from everything else"}
        del ctrl_sets, c_singles, c_chunk, c_tails

    # ---- per-kernel phases (CUDA events on the launching stream, same rotating buffers)
    # One CUDA graph per phase holding that phase for all SETS buffer sets back to back, so that the host's launch
    # rate never limits a 5-15 us kernel: time per launch = graph time / SETS.
    n_ph = min(max(K // SETS, 25), 250)

    def phase_graph(fn):
        for s_ in sets:
            fn(s_)
        torch.cuda.synchronize(dev)
        return capture(lambda: [fn(s_) for s_ in sets])

    g_prep = phase_graph(lambda s_: phase_match(s_, 2))        # JABD_ASSIGN_PREP_ONLY
    g_pm = phase_graph(lambda s_: phase_match(s_))             # prep + match
    g_enc = phase_graph(phase_encode)
    for g in (g_prep, g_pm, g_enc):
        g.replay()
    ms_prep, _ = timed_loop(lambda k: g_prep.replay(), n_ph)
    ms_match, _ = timed_loop(lambda k: g_pm.replay(), n_ph)
    ms_enc, _ = timed_loop(lambda k: g_enc.replay(), n_ph)
    us_prep, us_pm, us_enc = (x / (n_ph * SETS) * 1e3 for x in (ms_prep, ms_match, ms_enc))
    serial_chunk.replay()
    ms_serial, _ = timed_loop(lambda k: serial_chunk.replay(), n_ph)
    us_serial = ms_serial / (n_ph * SETS) * 1e3               # the same steps back to back on one stream (rounds 1-2's number)
    us_match = max(us_pm - us_prep, 1e-3)

    # the matching kernel as the lanes run it (work list of 192-GT items, launches of different batches overlapping):
    # (prep + match) - (prep alone), both dealt over the lanes like the steps; and the same shape alone on one stream
    TUNE_LANES = (192 << 8) | (192 << 16) | (100 << 24)          # JABD_ASSIGN_TUNE(192, 192, 100): what jabd_assign_batches picks

    def lanes_graph(fn):
        side = batched.lanes(dev, LANES)
        for s_ in sets:
            fn(s_)
        torch.cuda.synchronize(dev)

        def body():
            cur = torch.cuda.current_stream(dev)
            for x in side:
                x.wait_stream(cur)
            for i, s_ in enumerate(sets * ROUNDS):
                with torch.cuda.stream(side[i % LANES]):
                    fn(s_)
            for x in side:
                cur.wait_stream(x)
        return capture(body)

    n_pl = max(n_ph // ROUNDS, 10)
    g_prep_l = lanes_graph(lambda s_: phase_match(s_, 2 | TUNE_LANES))
    g_pm_l = lanes_graph(lambda s_: phase_match(s_, TUNE_LANES))
    g_pm_a = phase_graph(lambda s_: phase_match(s_, TUNE_LANES))
    for g in (g_prep_l, g_pm_l, g_pm_a):
        g.replay()
    ms_prep_l, _ = timed_loop(lambda k: g_prep_l.replay(), n_pl)
    ms_pm_l, _ = timed_loop(lambda k: g_pm_l.replay(), n_pl)
    ms_pm_a, _ = timed_loop(lambda k: g_pm_a.replay(), n_ph)
    us_match_lanes = max((ms_pm_l - ms_prep_l) / (n_pl * SETS * ROUNDS) * 1e3, 1e-3)
    us_match_lanes_shape_alone = max(ms_pm_a / (n_ph * SETS) * 1e3 - us_prep, 1e-3)
    hbm_peak, peak_src = peaks()
    sum_g = sum(s["sumG"] for s in sets) / SETS
    pairs = float(P) * sum_g                                   # prior x GT pairs per launch
    flops_launch = 14.0 * pairs                                # 14 fp32 ops per pair, no FMA
    mean_b = sum(s["B"] for s in sets) / float(SETS)          # == BATCH on one GPU; cost-balanced shards vary by a few images
    bytes_step = mean_b * (80.0 * P) + 60.0 * sum_g            # 80P + 60G per image

    # measured FP32 (non-FMA) peak: dependency-free FMUL/FADD chains on every SM, same clocks as the run
    # (libjabd_b200_selftest.so: a bench hook, not part of the product ABI)
    sms = ctypes.c_int(0)
    _lib.call("jabd_device_info", ctypes.byref(sms), None, None)
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    probe_ctas, probe_iters = sms.value * 16, 4096

    def probe(k=0):
        _lib.selftest_call("jabd_selftest_fp32_probe", probe_ctas, probe_iters, ptr(sink), cur_stream())
    for _ in range(3):
        probe()
    ms_probe, _ = timed_loop(probe, 20)
    fp32_peak = probe_ctas * 256 * 32.0 * probe_iters / (ms_probe / 20 * 1e-3) / 1e12      # Tops/s
    fp32_nominal = sms.value * 128 * 1.965e9 / 1e12

    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    match_tflops = flops_launch / (us_match * 1e-6) / 1e12
    enc_gbs = bytes_step / (us_enc * 1e-6) / 1e9
    t_fp32_us = flops_launch / (fp32_peak * 1e12) * 1e6
    t_hbm_us = bytes_step / (hbm_peak * 1e9) * 1e6
    # dominant kernel of the step = assign_match_kernel (FP32-pipe bound: no contraction, ~0.3 MB of DRAM reads)
    roofline = {"kernel": "assign_match_kernel", "bound": "fp32", "achieved": match_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": match_tflops / fp32_peak, "traffic": traffic.get("assign_match_kernel_bytes_per_launch"),
                "peak_source": "measured in this run: jabd_selftest_fp32_probe (FMUL+FADD chains, no FMA) on %d SMs; nominal %d x 128 "
                               "lanes x 1.965 GHz = %.1f" % (sms.value, sms.value, fp32_nominal),
                "algorithmic_flops_per_launch": flops_launch, "launch_us": us_match,
                "launch_us_how": "CUDA events around graphs of %d launches: (prep+match) - (prep alone), %d replays each" % (SETS, n_ph),
                "note": "achieved = 14 fp32 ops x P x sum(G) (dense-equivalent, SURVEY 8d) / kernel time; the kernel culls GT "
                        "against each warp's prior bounding box, so executed flops are lower than this (see phases.match_dense_*). "
                        "launch_us is the kernel of a call that runs alone (work list of 64-GT items: short tail); the lanes of the "
                        "timed step run it with 192-GT items (fewer items, less per-item work, a longer tail that the neighbouring "
                        "launches fill): see `lanes`",
                "lanes": {"work_list": "JABD_ASSIGN_TUNE(192, 192, 100)", "launch_us_overlapped": us_match_lanes,
                          "frac_overlapped": flops_launch / (us_match_lanes * 1e-6) / 1e12 / fp32_peak,
                          "launch_us_alone": us_match_lanes_shape_alone,
                          "frac_alone": flops_launch / (us_match_lanes_shape_alone * 1e-6) / 1e12 / fp32_peak,
                          "how": "overlapped: (prep+match) - (prep alone), both dealt over %d lanes in graphs of %d launches; alone: the "
                                 "same shape back to back on one stream" % (LANES, SETS * ROUNDS)},
                "step": {"ms_per_step": ms / K, "bound_us": max(t_fp32_us, t_hbm_us), "frac": max(t_fp32_us, t_hbm_us) / (ms / K * 1e3),
                         "serial_us_per_step": us_serial, "serial_frac": max(t_fp32_us, t_hbm_us) / us_serial,
                         "note": "whole step (all launches) against max(t_FP32 dense-equivalent, t_HBM) of SURVEY 8(d); ms_per_step: "
                                 "independent batches dealt over %d side streams (jabd_assign_batches), serial_*: the same steps "
                                 "back to back on one stream" % LANES}}
    roofline_encode = {"kernel": "match_encode_kernel", "bound": "hbm", "achieved": enc_gbs, "peak": hbm_peak, "unit": "GB/s",
                       "frac": enc_gbs / hbm_peak, "traffic": traffic.get("match_encode_kernel_bytes_per_launch"),
                       "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_step, "launch_us": us_enc,
                       "note": "80*P + 60*G bytes per image (SURVEY 8d) x 32 images; loc/conf/landm targets written once"}
    phases = {"prep_us": us_prep, "match_us": us_match, "match_encode_us": us_enc, "serial_step_us": us_serial, "lanes": LANES,
              "match_lanes_us": us_match_lanes, "match_lanes_shape_alone_us": us_match_lanes_shape_alone,
              "match_dense_equiv_tflops": match_tflops, "fp32_peak_measured_tops": fp32_peak, "fp32_peak_nominal_tops": fp32_nominal}

    extras = not args.no_extras
    only5 = bool(getattr(args, "only_cfg5", False))
    if extras:
        # dense (no culling) matching: every one of the P*G pairs evaluated once -> executed-flops figure
        g_dense = phase_graph(lambda s_: phase_match(s_, 1))
        g_dense.replay()
        n_d = min(n_ph, 40)
        ms_dense, _ = timed_loop(lambda k: g_dense.replay(), n_d)
        us_dense = max(ms_dense / (n_d * SETS) * 1e3 - us_prep, 1e-3)
        phases["match_dense_us"] = us_dense
        phases["match_dense_tflops"] = flops_launch / (us_dense * 1e-6) / 1e12
        phases["match_dense_frac_of_fp32_measured"] = phases["match_dense_tflops"] / fp32_peak

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region
    # (fixed BATCH images per rank and step here: the transfer volume, which bounds this path, is set by B*P)
    host_sets = [[pool[((rank * SETS + s_) * BATCH + j) % POOL] for j in range(BATCH)] for s_ in range(SETS)]
    cap_g = max(sum(int(t.shape[0]) for t in hs) for hs in host_sets)
    host = batched.HostAssign(pri, BATCH, cap_g)
    n_e2e = min(K, 100)
    for k in range(3):
        host(host_sets[k % SETS])
    pend = []

    def e2e_step(k):                        # two-slot pipeline: submit step k, then collect step k-1
        pend.append(host.submit(host_sets[k % SETS]))
        if len(pend) > 1:
            host.wait(pend.pop(0))
        if k == n_e2e - 1:                  # drain inside the timed region
            host.wait(pend.pop(0))
    ms_e2e, _ = timed_loop(e2e_step, n_e2e)
    e2e_rank_ms = per_rank(last_rank_ms[0])
    e2e_value = world * BATCH * n_e2e / (ms_e2e / 1e3)
    ms_e2e_sync, _ = timed_loop(lambda k: host(host_sets[k % SETS]), n_e2e)   # same call, one batch at a time
    # variant: the training flow -- GT rows from pinned host memory in, targets left in HBM for the loss (what
    # MultiBoxLoss.forward does with them), the step's result read back = the per-image positive counts.  Same C-ABI call
    # with JABD_ASSIGN_DEVICE_OUT, two slots.
    host_dev = batched.HostAssign(pri, BATCH, cap_g, device_out=True)
    cnt_pin = [torch.empty((BATCH,), dtype=torch.int64).pin_memory() for _ in range(2)]
    cnt_done = [torch.cuda.Event() for _ in range(2)]
    pend_dev = []

    def collect(k):
        slot = pend_dev.pop(0)
        cnt_done[slot].synchronize()

    def e2e_device_out(k):
        slot = host_dev.submit(host_sets[k % SETS])
        st = host_dev.slot_stream(slot)
        with torch.cuda.stream(st):
            cnt_pin[slot].copy_((host_dev.slots[slot]["conf_t"] != 0).sum(1), non_blocking=True)
            cnt_done[slot].record(st)
        pend_dev.append(slot)
        if len(pend_dev) > 1:
            collect(k)
        if k == n_e2e - 1:
            collect(k)

    for k in range(4):
        e2e_device_out(k)
    while pend_dev:
        collect(0)
    ms_e2e_dev, _ = timed_loop(e2e_device_out, n_e2e)
    d2h_step = host.last_d2h
    e2e = {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": host.last_h2d, "d2h_bytes_per_step": d2h_step,
           "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e,
           "per_rank": {"ms_per_step": [x / n_e2e for x in e2e_rank_ms],
                        "d2h_GBps": [d2h_step / (x / n_e2e * 1e-3) / 1e9 for x in e2e_rank_ms],
                        "aggregate_d2h_GBps": world * d2h_step / (ms_e2e / n_e2e * 1e-3) / 1e9,
                        "note": "every rank moves the same bytes; this path is bound by the D2H copy of the targets (34.4 MB per "
                                "rank and step over that rank's PCIe link into pinned host memory)"},
           "api": "batched.HostAssign.submit/wait -> jabd_assign_host (JABD_ASSIGN_ASYNC, 2 slots): list of per-image GT "
                  "tensors packed into pinned memory and copied in, all three target tensors copied out to pinned host memory "
                  "(one block, one D2H copy), every step; step k+1 is submitted before step k is collected",
           "synchronous_call": {"value": world * BATCH * n_e2e / (ms_e2e_sync / 1e3), "unit": "images/s",
                                "ms_per_step": ms_e2e_sync / n_e2e, "note": "HostAssign(targets): one batch at a time"},
           "device_resident_targets": {"value": world * BATCH * n_e2e / (ms_e2e_dev / 1e3), "unit": "images/s",
                                       "ms_per_step": ms_e2e_dev / n_e2e,
                                       "h2d_bytes_per_step": host_dev.last_h2d, "d2h_bytes_per_step": BATCH * 8,
                                       "note": "batched.HostAssign(device_out=True) -> jabd_assign_host(JABD_ASSIGN_DEVICE_OUT | ASYNC), two "
                                               "slots: per-image GT tensors packed into pinned memory and copied in, targets stay in HBM "
                                               "for the loss (what MultiBoxLoss.forward does with them, R/nets/retinaface_training.py:"
                                               "223-227), per-image positive counts read back"}}
    del host, host_dev

    cpu_ok = rank == 0 and world == 1 and not args.no_cpu_baseline
    cores = host_cores()
    if cpu_ok:
        from oracle import torch_port as tp

    def cpu_rate(fn, items, budget_s, max_reps):
        """items/s of a CPU port closure on all host threads, bounded."""
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < budget_s and reps < max_reps:
            fn()
            reps += 1
        return items * reps / (time.perf_counter() - t0), reps

    # ---- the whole MultiBoxLoss.forward + backward on the device (assign -> mining -> sums -> grads)
    loss_info = None
    if extras and not only5:
        ds0 = make_set(list(range(BATCH)))
        preds_l = [synth.make_logits(2, i, P) for i in range(BATCH)]
        pl, pc, pm = (torch.stack([q[j] for q in preds_l]).to(dev) for j in range(3))
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        norms = torch.empty((2,), dtype=torch.float32, device=dev)
        lmask = torch.empty((BATCH, P), dtype=torch.uint8, device=dev)
        lws = _tensor.workspace(L.jabd_multibox_loss_workspace_bytes(BATCH), dev)
        gl, gc, gm = torch.empty_like(pl), torch.empty_like(pc), torch.empty_like(pm)
        gin = torch.ones((3,), dtype=torch.float32, device=dev)

        def loss_step():
            assign(ds0)
            _lib.call("jabd_multibox_loss_forward", ptr(pl), ptr(pc), ptr(pm), ptr(ds0["loc"]), ptr(ds0["conf"]), ptr(ds0["landm"]),
                      BATCH, P, 7, ptr(losses), ptr(norms), ptr(lmask), ptr(lws), lws.numel(), cur_stream())
            _lib.call("jabd_multibox_loss_backward", ptr(pl), ptr(pc), ptr(pm), ptr(ds0["loc"]), ptr(ds0["landm"]), ptr(lmask),
                      ptr(norms), ptr(gin), BATCH, P, ptr(gl), ptr(gc), ptr(gm), cur_stream())
        loss_step()
        torch.cuda.synchronize(dev)
        g_loss = capture(loss_step)
        g_loss.replay()
        ms_loss, _ = timed_loop(lambda k: g_loss.replay(), 200)
        loss_info = {"images_per_s": world * BATCH * 200 / (ms_loss / 1e3), "us_per_batch": ms_loss / 200 * 1e3,
                     "launches_per_batch": 7, "losses": [float(v) for v in losses.tolist()],
                     "what": "assign (3 launches) + hard-negative mining / loss sums (3) + gradients w.r.t. the predictions (1) "
                             "for 32 x 640^2 images, synthetic logits, everything resident in HBM; replaces "
                             "R/nets/retinaface_training.py:183-303 + autograd backward"}
        if cpu_ok:
            nb = 8
            cpu_preds = tuple(t[:nb].cpu().clone().requires_grad_(True) for t in (pl, pc, pm))

            def cpu_loss():
                for q in cpu_preds:
                    q.grad = None
                a_, b_, c_ = tp.multibox_loss(cpu_preds, pri.cpu(), pool[:nb], THR, list(VAR), 7)
                (a_ + b_ + c_).backward()
            r, reps = cpu_rate(cpu_loss, nb, 5.0, 10)
            loss_info["cpu_port_images_per_s"] = r
            loss_info["cpu_port_sample"] = "%d x the first %d images, oracle/torch_port.multibox_loss forward+backward" % (reps, nb)
        del ds0, pl, pc, pm, gl, gc, gm, lmask

    def clustered_preds(cfg_id, size, pr, first, n, count=None):
        locs, confs, lms = [], [], []
        for i in range(first, first + n):
            gti = synth.make_gt(cfg_id, i, size, count=count)
            l, c, m = synth.make_preds_clustered(cfg_id, i, pr, gti, VAR, device=dev)
            locs.append(l); confs.append(c); lms.append(m)
        return torch.stack(locs).contiguous(), torch.stack(confs).contiguous(), torch.stack(lms).contiguous()

    # ---- inference side: decode+top-k+NMS at 640^2 (the metric's second half) and at cfg3's 1024^2
    detect_info = None
    if extras and not only5:
        detect_info = {}
        for name, size, B in (("640x640_b32", (640, 640), 32), ("cfg3_1024x1024_b16", (1024, 1024), 16)):
            pr = anchors.cached_priors(config.cfg_mnet, size, dev)
            Pd = int(pr.shape[0])
            loc_h, conf_h, lm_h = clustered_preds(3, size, pr, 0, B, count=60)
            loc_d, conf_d, lm_d = loc_h.to(dev), conf_h.to(dev), lm_h.to(dev)
            for _ in range(3):
                out = batched.detect(loc_d, conf_d, lm_d, pr, VAR)
            n_d = 30
            ms_d, _ = timed_loop(lambda k: batched.detect(loc_d, conf_d, lm_d, pr, VAR), n_d)
            hd = batched.HostDetect(pr, B, depth=3)
            lp, cp, mp = loc_h.pin_memory(), conf_h.pin_memory(), lm_h.pin_memory()
            for _ in range(2):
                hd(lp, cp, mp)
            pend_d = []

            def det_e2e(k):                 # three-slot pipeline (two batches in flight behind the one collected), drained inside the timed region
                pend_d.append(hd.submit(lp, cp, mp))
                if len(pend_d) > 2:
                    hd.wait(pend_d.pop(0))
                if k == n_d - 1:
                    while pend_d:
                        hd.wait(pend_d.pop(0))
            ms_dh, _ = timed_loop(det_e2e, n_d)
            detect_info[name] = {"images_per_s": world * B * n_d / (ms_d / 1e3), "ms_per_batch": ms_d / n_d,
                                 "e2e_images_per_s": world * B * n_d / (ms_dh / 1e3), "e2e_h2d_bytes": hd.last_h2d,
                                 "e2e_d2h_bytes": hd.last_d2h, "priors": Pd, "batch": B, "mean_kept": float(out[1].float().mean()),
                                 "params": "score>0.02, top-5000, IoU 0.4, keep 750; clustered synthetic predictions"}
            # the same detect for DET_SETS distinct batches per call (jabd_detect_batches without lanes: ONE launch whose grid covers
            # the images of all batches, cluster width chosen for all of them); the graph replays one call
            lb = [(loc_d, conf_d, lm_d)] + [tuple(t.to(dev) for t in clustered_preds(3, size, pr, B * j, B, count=60))
                                              for j in range(1, DET_SETS)]
            plan = batched.DetectBatches(pr, lb, VAR)
            for _ in range(2):
                outs_l = plan()
            torch.cuda.synchronize(dev)
            assert all(torch.equal(a_, b_) for a_, b_ in zip(outs_l[0], out)), "detect_batches differs from detect"
            g_l = capture(plan)
            g_l.replay()
            n_l = 12
            ms_l, _ = timed_loop(lambda k: g_l.replay(), n_l)
            detect_info[name]["batches"] = {"images_per_s": world * B * DET_SETS * n_l / (ms_l / 1e3), "ms_per_batch": ms_l / (n_l * DET_SETS),
                                            "batches_per_call": DET_SETS,
                                            "what": "batched.DetectBatches -> jabd_detect_batches: %d distinct batches of %d images (own "
                                                    "inputs, outputs, workspaces) per call in one launch, one CUDA graph per call; rows equal "
                                                    "jabd_detect's (asserted for the first batch)" % (DET_SETS, B)}
            del plan, lb, g_l
            if cpu_ok:
                # the reference's per-image post-processing on the host cores (decode, decode_landm, cat, threshold, top-k,
                # torchvision NMS; R/predict.py:167-181 composed per SURVEY D4), bounded sample
                nb = min(B, 4)
                pr_c = pr.cpu()
                r, reps = cpu_rate(lambda: [tp.infer_one_topk(loc_h[i], conf_h[i], lm_h[i], pr_c, list(VAR), 0.02, 5000, 0.4, 750)
                                            for i in range(nb)], nb, 4.0, 5)
                detect_info[name]["cpu_port_images_per_s"] = r
                detect_info[name]["cpu_port_sample"] = "%d x the first %d images, oracle/torch_port.infer_one_topk" % (reps, nb)
            # API-form decode (D1): HBM-bound elementwise kernel
            big = [torch.randn((64, Pd, 4), device=dev) * 0.5 for _ in range(4)]
            outs = [torch.empty_like(b) for b in big]

            def dec(k):
                _lib.call("jabd_decode", ptr(big[k % 4]), ptr(pr), Pd, 64, VAR[0], VAR[1], ptr(outs[k % 4]), cur_stream())
            for k in range(4):
                dec(k)
            torch.cuda.synchronize(dev)
            g_dec = capture(lambda: [dec(k) for k in range(4)])  # 4 launches per replay: the host never limits a ~10 us kernel
            g_dec.replay()
            ms_dec, _ = timed_loop(lambda k: g_dec.replay(), 50)
            ms_dec /= 4.0
            gbs = 64 * Pd * 32.0 / (ms_dec / 50 * 1e-3) / 1e9
            detect_info[name]["decode_kernel"] = {"batch": 64, "GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                                                  "bytes_per_launch": 64 * Pd * 32.0, "us": ms_dec / 50 * 1e3,
                                                  "l2": "4 rotating in/out sets of 64 images"}
            del big, outs, hd

    # ---- BASELINE configs[0] and configs[3]: the reference's live shape (one image per call, R/predict.py:167-181) at 640^2
    # with 50 faces, and the dense tiny-face stress (2048^2, 172,032 priors, 1,500 faces), one image per GPU.  Latency per call.
    def one_image(cfg_id, size, image_idx, reps_assign, reps_detect):
        pr = anchors.cached_priors(config.cfg_mnet, size, dev)
        Pn = int(pr.shape[0])
        gt1 = synth.make_gt(cfg_id, image_idx, size)
        gtd, offs, _ = batched.pack_targets([gt1], dev)
        G = int(gtd.shape[0])
        ws = _tensor.workspace(L.jabd_assign_workspace_bytes(1, Pn, G), dev)
        o = (torch.empty((1, Pn, 4), dtype=torch.float32, device=dev), torch.empty((1, Pn), dtype=torch.int64, device=dev),
             torch.empty((1, Pn, 10), dtype=torch.float32, device=dev))

        def call():
            _lib.call("jabd_assign", ptr(pr), Pn, ptr(gtd), ptr(offs), 1, G, THR, VAR[0], VAR[1], 0, 1, 0, ptr(o[0]), ptr(o[1]),
                      ptr(o[2]), None, None, None, None, ptr(ws), ws.numel(), cur_stream())
        call()
        torch.cuda.synchronize(dev)
        g1 = capture(call)
        g1.replay()
        ms_a, _ = timed_loop(lambda k: g1.replay(), reps_assign)
        for _ in range(2):
            batched.assign_targets(pr, [gt1.to(dev)], threshold=THR, variances=VAR)
        gt_dev = gt1.to(dev)
        ms_api, _ = timed_loop(lambda k: batched.assign_targets(pr, [gt_dev], threshold=THR, variances=VAR), min(reps_assign, 50))
        l1, c1, m1 = clustered_preds(cfg_id, size, pr, image_idx, 1)
        l1, c1, m1 = l1.to(dev), c1.to(dev), m1.to(dev)
        res = {"priors": Pn, "gt": G, "pairs": Pn * G,
               "assign_us": ms_a / reps_assign * 1e3, "assign_images_per_s": world * reps_assign / (ms_a / 1e3),
               "assign_api_us": ms_api / min(reps_assign, 50) * 1e3,
               "assign_dense_equiv_frac_of_fp32": 14.0 * Pn * G / (ms_a / reps_assign * 1e-3) / 1e12 / fp32_peak,
               "note": "assign_us: jabd_assign replayed from a CUDA graph (3 launches); assign_api_us: batched.assign_targets per call "
                       "(allocation + ctypes + 3 launches, what a drop-in caller pays); one image per GPU"}
        for tag, kw in (("detect_cfg3_params", dict()),
                        ("detect_live_params", dict(conf_thres=0.5, strict=False, pre_nms_topk=0, nms_thres=0.3, keep_topk=0))):
            for _ in range(2):
                dd = batched.detect(l1, c1, m1, pr, VAR, **kw)
            ms_d1, _ = timed_loop(lambda k: batched.detect(l1, c1, m1, pr, VAR, **kw), reps_detect)
            res[tag + "_us"] = ms_d1 / reps_detect * 1e3
            res[tag + "_kept"] = int(dd[1][0])
        if cpu_ok:
            pr_c, gtc = pr.cpu(), [gt1]
            r, reps = cpu_rate(lambda: tp.assign_batch(THR, gtc, pr_c, list(VAR)), 1, 3.0, 50)
            res["cpu_port_assign_images_per_s"] = r
            lc, cc, mc = l1[0].cpu(), c1[0].cpu(), m1[0].cpu()
            r, reps = cpu_rate(lambda: tp.infer_one_topk(lc, cc, mc, pr_c, list(VAR), 0.02, 5000, 0.4, 750), 1, 2.0, 20)
            res["cpu_port_detect_cfg3_params_images_per_s"] = r
            r, reps = cpu_rate(lambda: tp.infer_one(lc, cc, mc, pr_c, list(VAR), 0.5, 0.3), 1, 2.0, 20)
            res["cpu_port_detect_live_params_images_per_s"] = r
            res["cpu_port_cores"] = cores
        return res

    cfg1_info = cfg4_info = None
    if extras and not only5:
        cfg1_info = one_image(1, IMAGE, 0, 200, 50)
        cfg1_info["what"] = "BASELINE configs[0]: batch 1, 640x640, 16,800 priors, 50 faces -- latency per call"
        cfg4_info = one_image(4, (2048, 2048), rank, 50, 10)
        cfg4_info["what"] = ("BASELINE configs[3]: 2048x2048, 172,032 priors, 1,500 faces, one image per GPU (image index = rank); "
                             "live parameters = >= 0.5 / IoU 0.3 / nothing capped (R/predict.py:40,181)")

    # ---- cfg5 validation flow: detect the rank's shard, scale to pixels, all-gather the padded detections (the path's only
    # collective, NCCL over NVLink, issued on a side stream behind an event so that it overlaps the NEXT batch's detection),
    # WIDER AP of the gathered set on rank 0 (SURVEY 8e + 8f ranks 2-3)
    cfg5_info = None
    if extras:
        from jabd_b200 import utils_map
        size, B5, keep5 = (640, 640), 32, 750
        pr5 = anchors.cached_priors(config.cfg_mnet, size, dev)
        loc5, conf5, lm5 = (x.to(dev) for x in clustered_preds(5, size, pr5, rank * B5, B5, count=40))   # resident in HBM
        post5 = torch.from_numpy(batched.letterbox_params(size, [size] * B5)).to(dev)
        kidx5 = torch.empty((B5, keep5), dtype=torch.int32, device=dev)
        n5 = 20

        def cfg5_measure(gather):
            """detect -> pixel scaling -> exchange, per step; in the overlapped flow the consumer side of batch k-1 (result: the
            gathered rows become visible on this stream) runs while batch k is exchanged."""
            depth = gather.depth

            def step(k, overlap=True):
                slot = k % depth
                gather.acquire(slot)        # the slot's previous exchange has read its send buffer
                d, c = gather.dets(slot), gather.counts(slot)
                batched.detect(loc5, conf5, lm5, pr5, VAR, keep_topk=keep5, out=(d, c, kidx5))
                batched.correct_boxes(d, c, post5, letterbox=False, to_pixels=True)
                gather.launch(slot)
                if not overlap:
                    gather.result(slot)     # serialised variant: the step's stream waits for its own exchange
                elif k >= lag:
                    gather.result((k - lag) % depth)    # the consumer runs `lag` batches behind the producer
            lag = 2 if depth >= 3 else 1    # three slots: two batches of slack absorb the ranks' step-to-step jitter
            for k in range(2 * depth):
                step(k)
            for j in range(lag):
                gather.result((2 * depth - 1 - j) % depth)

            def loop(overlap):
                def body(k):
                    step(k, overlap)
                    if k == n5 - 1:         # drain inside the timed region: the last exchanges are part of the K steps
                        for j in range(lag - 1, -1, -1):
                            gather.result((k - j) % depth)
                ms_, _ = timed_loop(body, n5)
                return ms_
            ms_o = loop(True)
            ms_s = loop(False)
            gd_, gc_ = gather.result((n5 - 1) % depth)
            torch.cuda.synchronize(dev)
            # the exchange alone, back to back
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for k in range(n5):
                gather.acquire(k % depth)
                gather.launch(k % depth)
                gather.result(k % depth)
            e1.record()
            barrier()
            us = max_over_ranks(e0.elapsed_time(e1)) / n5 * 1e3
            return ms_o, ms_s, us, gd_, gc_

        gather = sharding.PeerGather(B5, keep5, dev, depth=3)
        ms5, ms5_serial, us_ag, gd, gc = cfg5_measure(gather)
        gd, gc = gd.clone(), gc.clone()
        p2p_status = gather.status()
        assert p2p_status == 0, "a peer-memory wait timed out (status %d)" % p2p_status
        nccl_ctl = None
        if world > 1:
            g_nccl = sharding.DetectionGather(B5, keep5, dev, depth=2)
            ms_n, ms_ns, us_n, gd_n, gc_n = cfg5_measure(g_nccl)
            same = bool(torch.equal(gd_n.reshape(gd.shape), gd) and torch.equal(gc_n.reshape(gc.shape), gc))
            assert same or gather.transport != "p2p", "peer-memory gather and NCCL all-gather returned different detections"
            nccl_ctl = {"images_per_s": world * B5 * n5 / (ms_n / 1e3), "ms_per_step": ms_n / n5, "serialised_ms_per_step": ms_ns / n5,
                        "allgather_us": us_n, "same_rows_as_p2p": same,
                        "what": "the same flow with ONE all_gather_into_tensor over NCCL on a side stream (sharding.DetectionGather, round 1's "
                                "transport), for comparison"}
        ms_det_only, _ = timed_loop(lambda k: (batched.detect(loc5, conf5, lm5, pr5, VAR, keep_topk=keep5,
                                                              out=(gather.dets(0), gather.counts(0), kidx5)),
                                               batched.correct_boxes(gather.dets(0), gather.counts(0), post5, letterbox=False,
                                                                     to_pixels=True)), n5)
        cfg5_info = {"images_per_s": world * B5 * n5 / (ms5 / 1e3), "ms_per_step": ms5 / n5,
                     "serialised_ms_per_step": ms5_serial / n5, "detect_only_ms_per_step": ms_det_only / n5,
                     "transport": gather.transport,
                     "allgather_us": us_ag, "allgather_overlapped": True,
                     "allgather_exposed_us": max(ms5 - ms_det_only, 0.0) / n5 * 1e3,
                     "allgather_bytes_per_rank": int(gather.L * 4),
                     "nccl_control": nccl_ctl,
                     "what": "per rank: fused detect of 32 x 640^2 images (score>0.02, top-5000, IoU 0.4, keep 750) + pixel scaling into "
                             "the send buffer, then the exchange of [32*750*15 floats | 32 counts] on a side stream: transport 'p2p' = "
                             "jabd_p2p_allgather, this library's kernel storing the block straight into every rank's receive buffer over "
                             "NVLink (sharding.PeerGather, three slots, flags + acknowledgements in peer memory, no host sync); 'nccl' = "
                             "all_gather_into_tensor (fallback when peers cannot be mapped).  Batch k+1 is detected while batch k is "
                             "exchanged; serialised_ms_per_step waits for each exchange in line (round 1's flow)"}
        # The same flow with TWO steps' images per detect launch (jabd_detect_batches: one grid over both batches, narrower clusters,
        # the block scheduler refills SMs as images finish); every 32-image block is still exchanged on its own, the consumer
        # runs one double step behind.  A validation pass has its batches queued, so nothing forces one launch per exchange.
        pair = None
        try:
            g2 = sharding.PeerGather(B5, keep5, dev, depth=4)
            loc5b, conf5b, lm5b = (x.to(dev) for x in clustered_preds(5, size, pr5, (world + rank) * B5, B5, count=40))
            kidx5b = torch.empty((B5, keep5), dtype=torch.int32, device=dev)
            wsb = [_tensor.workspace(L.jabd_detect_workspace_bytes(B5, int(pr5.shape[0]), keep5), dev) for _ in range(2)]
            P5 = int(pr5.shape[0])

            def two(slot_a, slot_b):
                return (_lib.DetectBatch * 2)(
                    _lib.DetectBatch(loc5.data_ptr(), conf5.data_ptr(), lm5.data_ptr(), B5, g2.dets(slot_a).data_ptr(),
                                     g2.counts(slot_a).data_ptr(), kidx5.data_ptr(), wsb[0].data_ptr(), wsb[0].numel()),
                    _lib.DetectBatch(loc5b.data_ptr(), conf5b.data_ptr(), lm5b.data_ptr(), B5, g2.dets(slot_b).data_ptr(),
                                     g2.counts(slot_b).data_ptr(), kidx5b.data_ptr(), wsb[1].data_ptr(), wsb[1].numel()))
            arrs = {(0, 1): two(0, 1), (2, 3): two(2, 3)}

            def dstep(j):
                sa, sb = (0, 1) if j % 2 == 0 else (2, 3)
                g2.acquire(sa)
                g2.acquire(sb)
                _lib.call("jabd_detect_batches", ptr(pr5), P5, ctypes.cast(arrs[(sa, sb)], ctypes.c_void_p), 2, VAR[0], VAR[1], 0.02, 2, 5000,
                          0.4, keep5, 0, None, 0, cur_stream())
                for sl in (sa, sb):
                    batched.correct_boxes(g2.dets(sl), g2.counts(sl), post5, letterbox=False, to_pixels=True)
                    g2.launch(sl)
                if j >= 1:
                    for sl in ((2, 3) if j % 2 == 0 else (0, 1)):     # the previous double step's two exchanges
                        g2.result(sl)
            for j in range(4):
                dstep(j)
            g2.result(2); g2.result(3)

            def dbody(j):
                dstep(j)
                if j == n5 // 2 - 1:
                    for sl in ((0, 1) if j % 2 == 0 else (2, 3)):
                        g2.result(sl)
            ms_p, _ = timed_loop(dbody, n5 // 2)
            ga, _ = g2.result(0 if (n5 // 2 - 1) % 2 == 0 else 2)
            torch.cuda.synchronize(dev)
            same_rows = bool(torch.equal(ga.reshape(gd.shape), gd))          # the first batch of a pair is the flow's batch: same rows
            assert g2.status() == 0, "a peer-memory wait timed out in the two-batch flow"
            pair = {"images_per_s": world * B5 * (n5 // 2) * 2 / (ms_p / 1e3), "ms_per_step": ms_p / ((n5 // 2) * 2),
                    "same_rows_as_one_batch_per_launch": same_rows, "transport": g2.transport,
                    "what": "two steps' images (2 x 32, distinct) per jabd_detect_batches launch, each 32-image block scaled and exchanged on "
                            "its own (PeerGather, four slots), the consumer one double step behind"}
            del g2
        except Exception as e:      # reported, never fatal for the bench line
            pair = {"error": repr(e)[:200]}
        cfg5_info["two_batches_per_launch"] = pair
        if rank == 0:
            gd2 = gd.reshape(world * B5, keep5, 15).contiguous()
            preds5 = utils_map.dets_to_pred_rows(gd2, gc.reshape(-1).contiguous())
            gts5, keeps5 = [], []
            for gi in range(world * B5):
                t5 = synth.make_gt(5, gi, size, count=40)[:, :4].numpy().astype("float64") * size[0]
                gts5.append(np.stack([t5[:, 0], t5[:, 1], t5[:, 2] - t5[:, 0], t5[:, 3] - t5[:, 1]], 1))
                keeps5.append(np.ones(t5.shape[0], np.uint8))
            t0 = time.perf_counter()
            ap5 = utils_map.evaluate_arrays(preds5, gts5, [keeps5], 0.4, 1000)
            cfg5_info["ap_eval_ms"] = (time.perf_counter() - t0) * 1e3
            cfg5_info["ap_all_faces"] = float(ap5[0])
            cfg5_info["ap_images"] = world * B5

    clocks = None
    if sampler is not None:
        sampler.stop()
        clocks = sampler.summary(windows)

    # ---- CPU baseline (rank 0, single-GPU run only): torch port of the reference loop, bounded sample, all threads and one
    cpu = None
    if cpu_ok:
        pri_c = pri.cpu()
        sample = pool[:16]
        v_all, reps = cpu_assign_rate(tp, torch, sample, pri_c, cores, 10.0)
        v_one, reps1 = cpu_assign_rate(tp, torch, sample[:4], pri_c, 1, 4.0)
        torch.set_num_threads(cores)
        cpu = {"value": v_all, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "%d x the first 16 images of the cfg2 batch (~10 s) through oracle/torch_port.assign_batch (the reference's "
                         "per-image match loop restated in torch %s CPU, %d threads; the reference is Python and cannot travel)"
                         % (reps, torch.__version__, cores),
               "one_thread": {"value": v_one, "unit": "images/s", "cores": 1,
                              "sample": "%d x the first 4 images, torch.set_num_threads(1)" % reps1},
               "os_cpu_count": os.cpu_count(), "sched_getaffinity": cores,
               "torch": torch.__version__}
        try:
            import torchvision
            cpu["torchvision"] = torchvision.__version__
        except Exception:
            pass

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    rms = [x / K for x in rank_ms]
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 training target assignment (match+encode): batch %d/GPU at 640x640, %d priors, "
                               "1..300 GT/image (mean %.1f), threshold 0.35, variances [0.1,0.2]" % (BATCH, P, mean_g),
                   "global_batch": world * BATCH, "image": list(IMAGE), "priors": P,
                   "parallelism": "image-sharded x%d, no collective: each step's %d DISTINCT images (from one pool of %d = the cfg5 batch, "
                                  "the same pool at every N) are cut into LPT shards of equal estimated cost" % (world, world * BATCH, POOL),
                   "l2": "%d rotating buffer sets per rank (%.0f MB of targets+workspace > 126 MB L2); steps replayed from CUDA graphs "
                         "(one graph of %d steps while >= %d remain, one shorter graph of the same kind for the rest)"
                         % (SETS, SETS * (BATCH * P * 72 + BATCH * P * 8) / 1e6, SETS * ROUNDS, SETS * ROUNDS),
                   "overlap": "the %d steps of a graph are %d batches (independent: own GT, outputs, workspace) issued through ONE jabd_assign_batches call: batch i on "
                              "side stream i %% %d, forked from and joined into the timing stream, so that one batch's staging and "
                              "encode kernels run in the ramp and tail of another batch's persistent matching kernel; every step still "
                              "launches its own 3 kernels on its own batch, outputs and workspace (roofline.step.serial_us_per_step: "
                              "the same steps back to back on one stream)" % (SETS * ROUNDS, SETS * ROUNDS, LANES),
                   "cpu_affinity": None if my_cpus is None else {"rank0_cpus": len(my_cpus), "visible": len(all_cpus)}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": 3 * K, "roofline": roofline, "roofline_encode": roofline_encode,
        "cpu_baseline": cpu, "phases": phases, "cfg1": cfg1_info, "cfg4": cfg4_info,
        "detect": detect_info, "loss": loss_info, "cfg5_eval": cfg5_info,
        "per_rank": {"ms_per_step": rms, "images_per_step": rank_images, "gt_per_step": rank_gt, "est_cost_per_step": rank_cost,
                     "max_over_min_ms": max(rms) / min(rms),
                     "note": "value uses the slowest rank; shards are LPT-packed by estimated cost (sharding.lpt_shards)"},
        "shard_check": shard_check, "scaling_control": control,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
