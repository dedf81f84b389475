"""CPU-only checks (no CUDA device needed): the C-ABI library builds, loads and exports exactly what
include/jabd_b200.h declares; host-side argument validation; sharding logic incl. a world-size-2 gloo run.
No compute entry point is called here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from jabd_b200 import _lib
    _lib.build()
    return _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "jabd_b200.h")).read()
    return sorted(set(re.findall(r"JABD_API[^;(]*?\b(jabd_\w+)\s*\(", text)))


def test_selftest_library_is_separate(lib):
    """Test / bench hooks live in their own library and header; the product ABI does not export them."""
    text = open(os.path.join(ROOT, "include", "jabd_b200_selftest.h")).read()
    declared = sorted(set(re.findall(r"JABD_API[^;(]*?\b(jabd_\w+)\s*\(", text)))
    assert declared == sorted(lib.SELFTEST_SIGNATURES)
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.SELFTEST_SO_PATH]).decode()
    assert sorted(set(re.findall(r" T (jabd_\w+)", out))) == declared
    lib.selftest_lib()
    prod = subprocess.check_output(["nm", "-D", "--defined-only", lib.SO_PATH]).decode()
    assert "selftest" not in prod and "debug" not in prod and "probe" not in prod


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert len(declared) >= 25
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(lib.SIGNATURES) == declared            # the ctypes table mirrors the header one to one
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.SO_PATH]).decode()
    exported = sorted(set(re.findall(r" T (jabd_\w+)", out)))
    assert exported == declared                           # nothing undeclared leaks out either
    assert L.jabd_version() == 100


def test_library_is_sm100a_with_tma(lib):
    sass = subprocess.check_output(["cuobjdump", "-sass", lib.SO_PATH]).decode()
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "SYNCS" in sass           # cp.async.bulk + mbarrier in the matching kernel
    assert "REDUX" in sass                                # warp argmax
    assert "UCGABAR_ARV" in sass and "UCGABAR_WAIT" in sass   # thread-block cluster barrier (loss forward, DSMEM histograms)
    # the detect / NMS kernels are cluster kernels too: per-function check that the cluster barrier (arrive + wait) and generic
    # stores to mapped distributed-shared-memory addresses (st.shared::cluster -> ST.E / ST.E.128: candidate boxes, window masks,
    # triangle words of the NMS loop) are in their code
    for fn in ("detect_kernel", "nms_kernel"):
        body = sass.split("Function : ")
        mine = [b for b in body if b.split("\n", 1)[0].find(fn) >= 0]
        assert mine and all("UCGABAR_ARV" in b and "UCGABAR_WAIT" in b and "ST.E.128" in b and "ST.E " in b for b in mine), fn


def test_host_side_validation_without_gpu(lib):
    L = lib.lib()
    vp = ctypes.c_void_p
    steps = np.array([8, 16, 32], np.int32)
    off = np.array([0, 2, 4, 6], np.int32)
    n = L.jabd_priors_count(vp(steps.ctypes.data), vp(off.ctypes.data), 3, 640, 640)
    assert n == 16800
    assert L.jabd_priors_count(vp(steps.ctypes.data), vp(off.ctypes.data), 3, 1024, 1024) == 43008
    assert L.jabd_priors_count(vp(steps.ctypes.data), vp(off.ctypes.data), 3, 100, 75) == 2 * (13 * 10 + 7 * 5 + 4 * 3)
    assert L.jabd_priors_count(vp(steps.ctypes.data), vp(off.ctypes.data), 0, 640, 640) == -1
    assert "n_levels" in lib.last_error()
    # sizes only: never touches the device
    w1 = L.jabd_assign_workspace_bytes(32, 16800, 3000)
    assert w1 >= 32 * 16800 * 8 + 3000 * 24 and w1 % 256 == 0
    assert L.jabd_assign_host_scratch_bytes(32, 16800, 3000, 1) > w1 + 32 * 16800 * 64
    assert L.jabd_detect_workspace_bytes(16, 43008, 750) >= 16 * 750 * 20
    # argument errors are reported before any CUDA call
    assert L.jabd_assign(None, 16800, None, None, 4, 10, 0.35, 0.1, 0.2, 0, 1, 0, None, None, None, None, None, None, None,
                         None, 0, None) == -1
    assert "null" in lib.last_error()
    assert L.jabd_assign(None, -1, None, None, 4, 10, 0.35, 0.1, 0.2, 0, 1, 0, None, None, None, None, None, None, None,
                         None, 0, None) == -1
    buf = np.zeros(64, np.float32)
    mis = vp(buf.ctypes.data + 4)
    assert L.jabd_decode(mis, mis, 4, 1, 0.1, 0.2, mis, None) == -2
    assert L.jabd_nms(mis, 0, 4, mis, 0, 1, 1, 4, 0.0, 7, 0, 0.3, 0, 4, mis, mis, None, 0, None) == -1
    # CTAs per image of the detect / NMS kernels: a per-call option (0 automatic, 1..8), validated before any device call
    assert L.jabd_nms(mis, 0, 4, mis, 0, 1, 1, 4, 0.0, 0, 0, 0.3, 9 << 12, 4, mis, mis, None, 0, None) == -1
    assert "CTAs per segment" in lib.last_error()
    assert L.jabd_detect(None, None, None, None, 1, 4, 0.1, 0.2, 0.02, 2, 0, 0.4, 4, 9, None, None, None, None, 0, None) == -1
    assert "CTAs per image" in lib.last_error()
    assert L.jabd_nms_stats_offset(2, 750) == 24064 + 6144 and L.jabd_nms_workspace_bytes(2, 100, 750) == 24064 + 6144 + 256
    offs = (ctypes.c_size_t * 4)()
    assert L.jabd_assign_host_out_offsets(32, 16800, 1, offs) == 0
    assert list(offs) == [0, 32 * 16800 * 16, 32 * 16800 * 24, 32 * 16800 * 64]
    # host-only GT packing helper (list of per-image arrays -> packed rows + offsets)
    import torch
    ts = [torch.rand(3, 15), torch.zeros(0, 15), torch.rand(5, 15)]
    rows = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in ts])
    counts = (ctypes.c_int * 3)(3, 0, 5)
    gt, off = torch.empty(10, 15), torch.empty(4, dtype=torch.int32)
    assert L.jabd_pack_gt_rows(rows, counts, 3, vp(gt.data_ptr()), 10, vp(off.data_ptr())) == 8
    assert off.tolist() == [0, 3, 3, 8] and torch.equal(gt[:8], torch.cat(ts, 0))
    assert L.jabd_pack_gt_rows(rows, counts, 3, vp(gt.data_ptr()), 7, vp(off.data_ptr())) == -3 and "capacity" in lib.last_error()
    with pytest.raises(ValueError):
        lib.check(-2, "x")
    with pytest.raises(RuntimeError):
        lib.check(-4, "x")


def test_header_is_plain_c_and_structs_match_ctypes(lib, tmp_path):
    """include/jabd_b200.h compiles as C (gcc, no CUDA headers) and the two batch structs of the multi-batch entry points have the
    size and field offsets the ctypes mirrors in _lib.py assume; the lane / shape options are refused host-side before any
    device call."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of this image"
    src = tmp_path / "abi.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "jabd_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu\n", sizeof(jabd_assign_batch_t), offsetof(jabd_assign_batch_t, sumG), offsetof(jabd_assign_batch_t, loc_t),
           offsetof(jabd_assign_batch_t, workspace), offsetof(jabd_assign_batch_t, workspace_bytes));
    printf("%zu %zu %zu %zu %zu\n", sizeof(jabd_detect_batch_t), offsetof(jabd_detect_batch_t, B), offsetof(jabd_detect_batch_t, dets),
           offsetof(jabd_detect_batch_t, workspace), offsetof(jabd_detect_batch_t, workspace_bytes));
    printf("%d\n", JABD_ASSIGN_TUNE(192, 64, 70));
    return 0;
}
""")
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    A, D = lib.AssignBatch, lib.DetectBatch
    assert [int(x) for x in out[0].split()] == [ctypes.sizeof(A), A.sumG.offset, A.loc_t.offset, A.workspace.offset, A.workspace_bytes.offset]
    assert [int(x) for x in out[1].split()] == [ctypes.sizeof(D), D.B.offset, D.dets.offset, D.workspace.offset, D.workspace_bytes.offset]
    from jabd_b200 import batched
    assert int(out[2]) == batched.tune_flags((192, 64, 70)) and batched.tune_flags(None) == 0
    with pytest.raises(ValueError):
        batched.tune_flags((64, 64, 101))
    # host-side refusals of the multi-batch calls (no device is touched before them)
    L = lib.lib()
    vp = ctypes.c_void_p
    one = (vp * 1)(vp(0x1000))
    assert L.jabd_assign_batches(None, 4, None, 1, 0.35, 0.1, 0.2, 0, 1, 0, None, 0, None) == -1          # null batch list
    assert L.jabd_assign_batches(None, 4, None, 0, 0.35, 0.1, 0.2, 0, 1, 0, None, 65, None) == -1         # too many lanes
    assert L.jabd_assign_batches(None, 4, None, 0, 0.35, 0.1, 0.2, 0, 1, 0, ctypes.cast(one, vp), 1, vp(0x1000)) == -1
    assert "calling stream" in lib.last_error()
    assert L.jabd_detect_batches(None, 4, None, -1, 0.1, 0.2, 0.02, 2, 0, 0.4, 4, 0, None, 0, None) == -1
    assert L.jabd_detect_batches(None, 4, None, 0, 0.1, 0.2, 0.02, 2, 0, 0.4, 4, 0, None, 0, None) == 0   # nothing to do
    buf = np.zeros(64, np.float32)
    p4 = vp(buf.ctypes.data)
    # a work-list shape outside 16..192 is refused by the call it belongs to
    assert L.jabd_assign_match(p4, 4, p4, p4, 1, 1, (8 << 8) | (64 << 16) | (100 << 24), p4, 1 << 20, None) in (-1, -2, -3)


def test_no_cpu_fallback():
    """Without a CUDA device the operators raise instead of computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from jabd_b200 import anchors, batched, config, utils_bbox
    with pytest.raises(RuntimeError):
        anchors.Anchors(config.cfg_mnet, image_size=(64, 64)).get_anchors()
    with pytest.raises(RuntimeError):
        batched.assign_targets(np.zeros((4, 4), np.float32), [np.zeros((1, 15), np.float32)])
    with pytest.raises(RuntimeError):
        utils_bbox.decode(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32), [0.1, 0.2])
    # and nothing under the product package imports the oracle
    pkg = os.path.join(ROOT, "jabd-joint-attention-based-detector-for-small-face-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "jabd_oracle" not in src, f


def test_shard_bounds():
    from jabd_b200 import sharding
    assert sharding.shard_bounds(256, 8) == [32 * i for i in range(9)]
    assert sharding.shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert sharding.shard_bounds(3, 4) == [0, 1, 2, 3, 3]
    b = sharding.shard_bounds(6, 2, [300, 1, 1, 1, 1, 296])
    assert b[0] == 0 and b[-1] == 6 and b == sorted(b)
    loads = [sum([300, 1, 1, 1, 1, 296][b[i]:b[i + 1]]) for i in range(2)]
    assert max(loads) <= 304
    # every item lands in exactly one shard
    for n, w in ((17, 5), (1, 3), (0, 2), (100, 7)):
        bb = sharding.shard_bounds(n, w, list(range(1, n + 1)))
        assert bb[0] == 0 and bb[-1] == n and all(x <= y for x, y in zip(bb, bb[1:]))


def test_lpt_shards():
    """LPT bin packing over images: a partition, deterministic, tighter than the contiguous prefix split on ragged weights,
    honours a per-rank item cap; gather_order is the permutation that undoes the rank-major concatenation."""
    from jabd_b200 import sharding, synth
    rng = np.random.default_rng(3)
    for n, w in ((256, 8), (64, 2), (7, 4), (3, 4), (0, 2), (1, 1)):
        wt = (1 + np.floor(299 * rng.random(n) ** 2)).tolist()
        sh = sharding.lpt_shards(wt, w)
        assert len(sh) == w and sorted(i for s in sh for i in s) == list(range(n)) and all(s == sorted(s) for s in sh)
        assert sh == sharding.lpt_shards(list(wt), w)
        perm = sharding.gather_order(sh)
        assert sorted(perm) == list(range(n))
    wt = sharding.image_costs(synth.make_gt_batch(2, 256, (640, 640)))
    lpt = [sum(wt[i] for i in s) for s in sharding.lpt_shards(wt, 8)]
    b = sharding.shard_bounds(256, 8, wt)
    pre = [sum(wt[b[i]:b[i + 1]]) for i in range(8)]
    assert max(lpt) / min(lpt) < 1.005 < max(pre) / min(pre)          # 256 ragged images: LPT is even to a fraction of a percent
    capped = sharding.lpt_shards(wt, 8, max_items=32)
    assert all(len(s) == 32 for s in capped)
    with pytest.raises(ValueError):
        sharding.lpt_shards([1.0] * 9, 2, max_items=4)
    # equal weights: items keep their order, round-robin over the ranks
    assert sharding.lpt_shards([1, 1, 1, 1, 1, 1], 3) == [[0, 3], [1, 4], [2, 5]]


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from jabd_b200 import sharding, synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
targets = synth.make_gt_batch(2, 8, (640, 640))
mine, (lo, hi) = sharding.local_targets(targets, contiguous=True)
# every rank derives the same partition; together the shards cover the batch once
spans = [None, None]
dist.all_gather_object(spans, (lo, hi))
assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 8, spans
assert [t.shape for t in mine] == [t.shape for t in targets[lo:hi]]
# LPT shards (the default): index lists, every image in exactly one shard, same partition on every rank
mine, idx = sharding.local_targets(targets)
both = [None, None]
dist.all_gather_object(both, idx)
assert sorted(both[0] + both[1]) == list(range(8)) and both[rank] == idx
assert [t.shape for t in mine] == [targets[i].shape for i in idx]
shards = sharding.lpt_shards(sharding.image_costs(targets), 2)
assert shards == both
# side-stream gather object on CPU tensors (gloo): per-rank results come back in rank order; the permutation index restores
# the global image order
g = sharding.DetectionGather(len(idx) if len(both[0]) == len(both[1]) else 4, 3, "cpu", depth=2)
if len(both[0]) == len(both[1]):
    for slot in range(2):
        g.dets(slot)[:] = torch.tensor(idx, dtype=torch.float32)[:, None, None] + 0.25 * slot
        g.counts(slot)[:] = torch.tensor(idx, dtype=torch.int32) + 100
        g.launch(slot)
    for slot in range(2):
        d, c = g.result(slot)
        assert d.shape == (2, len(idx), 3, 15) and c.shape == (2, len(idx)) and c.dtype == torch.int32
        perm = torch.tensor(sharding.gather_order(both))
        glob = torch.empty(8)
        glob[perm] = d[:, :, 0, 0].reshape(-1)
        assert glob.tolist() == [i + 0.25 * slot for i in range(8)]
        cg = torch.empty(8, dtype=torch.int32)
        cg[perm] = c.reshape(-1)
        assert cg.tolist() == [100 + i for i in range(8)]
# fixed-shape all-gather of padded detections (CPU tensors -> gloo path)
B_local, keep = 4, 6
dets = torch.full((B_local, keep, 15), float(rank + 1))
counts = torch.full((B_local,), rank + 3, dtype=torch.int32)
d, c = sharding.allgather_detections(dets, counts)
assert d.shape == (8, keep, 15) and (d[:4] == 1).all() and (d[4:] == 2).all()
assert c.tolist() == [3] * 4 + [4] * 4 and c.dtype == torch.int32
# ragged counts and distinct rows survive the single packed collective bit for bit (counts travel bit-cast as a float column)
g = torch.Generator().manual_seed(7 + rank)
dets = torch.rand((B_local, keep, 15), generator=g)
counts = torch.tensor([0, 6, 2, 2147483647 if rank else 5], dtype=torch.int32)
d, c = sharding.allgather_detections(dets, counts)
both = [torch.rand((B_local, keep, 15), generator=torch.Generator().manual_seed(7 + r)) for r in range(2)]
assert torch.equal(d, torch.cat(both, 0)) and c.tolist() == [0, 6, 2, 5, 0, 6, 2, 2147483647]
# other dtypes fall back to two collectives
d64, c64 = sharding.allgather_detections(dets.double(), counts.long())
assert torch.equal(d64, torch.cat(both, 0).double()) and c64.tolist() == c.tolist()
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharding_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


# ---- numpy restatements of two pieces of device arithmetic in csrc/detect.cu (the claims the kernels rest on) ---------------
def _ord_of(v):
    """common.cuh ord_of(): order-preserving float32 -> uint32 (NaN largest, +0 == -0)."""
    b = v.view(np.uint32).astype(np.uint64)
    u = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
    u = np.where(v == 0, 0x80000000, u)
    return np.where(np.isnan(v), 0xFFFFFFFF, u).astype(np.uint64)


def _fine_bin(u):
    """detect.cu fine_bin(): 64 bins per octave over [2^-30, 4), clamped at both ends."""
    b = (u >> 17).astype(np.int64) - ((0xC0800000 >> 17) - 2048)
    return np.clip(b, 0, 2047)


def test_fine_bin_is_monotone_and_resolves_probabilities():
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 3, rng.random(20000, dtype=np.float32),
                        np.float32(2.0) ** rng.integers(-40, 10, 2000).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, 3.9999998, 4.0, 1e30, -1e30, np.inf, -np.inf, 2.0 ** -30, 2.0 ** -31], np.float32)])
    v = np.sort(v)
    assert _ord_of(np.array([4.0], np.float32))[0] == 0xC0800000
    u = _ord_of(v)
    assert (np.diff(u.astype(np.int64)) >= 0).all()                 # ord_of is order preserving
    fb = _fine_bin(u)
    assert (np.diff(fb) >= 0).all() and fb.min() == 0 and fb.max() == 2047   # monotone: "bins >= cut" is a score threshold
    p = np.linspace(0.02, 0.999, 5000, dtype=np.float32)
    fine, coarse = np.unique(_fine_bin(_ord_of(p))).size, np.unique(_ord_of(p) >> 21).size
    assert fine >= 300 and coarse <= 24                             # what the float's top 11 bits leave of a probability
    assert _fine_bin(_ord_of(np.array([2.0 ** -30, 3.9999998], np.float32))).tolist() == [0, 2047]


def test_division_free_nms_decision_matches_the_quotient():
    """suppresses_tv(): inter > fl(uni * fl(t(1+2^-20))) implies fl(inter/uni) > t and inter < fl(uni * fl(t(1-2^-20))) implies
    fl(inter/uni) < t, for uni in [2^-60, 2^60] and t in [2^-20, 2^20] -- checked in float32 on values crowded around t*uni."""
    rng = np.random.default_rng(11)
    f32 = np.float32
    for t in (f32(0.3), f32(0.4), f32(0.5), f32(0.45), f32(2.0 ** -20), f32(0.99999994), f32(1.0), f32(2.0 ** 20)):
        c_hi, c_lo = f32(t * f32(1.0 + 2.0 ** -20)), f32(t * f32(1.0 - 2.0 ** -20))
        uni = (f32(2.0) ** rng.uniform(-60, 60, 400000).astype(f32)).astype(f32)
        uni = np.concatenate([uni, rng.random(400000, dtype=f32) * f32(2.0)])       # box unions of normalised coordinates
        k = rng.integers(-64, 65, uni.size).astype(f32)
        inter = (uni * t * (f32(1.0) + k * f32(2.0 ** -22))).astype(f32)            # within +-16 guard widths of the threshold
        ok = (uni >= f32(2.0 ** -60)) & (uni <= f32(2.0 ** 60)) & np.isfinite(inter) & (inter > 0)
        uni, inter = uni[ok], inter[ok]
        q = (inter / uni).astype(f32)                                               # IEEE division, what the reference evaluates
        yes, no = inter > (uni * c_hi).astype(f32), inter < (uni * c_lo).astype(f32)
        assert yes.sum() > 1000 and no.sum() > 1000 and (~(yes | no)).sum() > 1000  # all three outcomes are exercised
        assert (q[yes] > t).all() and (q[no] < t).all()


def test_committed_bench_lines_follow_the_contract():
    """profiles/r01_bench*.json are what bench.py printed on the B200 box: every key of the driver's contract is there."""
    import json
    for name, n in (("r01_bench.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n8.json", 8)):
        line = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1]
        d = json.loads(line)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["unit"] == "images/s" and d["scaling"] == "weak" and d["higher_is_better"] is True
        assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None and d["dtype"] == "f32"
        assert d["gpu_launches"] == 3 * d["steps"] and d["warmup"] >= 3
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["d2h_bytes_per_step"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]
