"""Peer-memory exchange (csrc/p2p.cu, sharding.PeerGather) on one GPU: the kernels, flags, acknowledgements and buffer views.
Two ranks are emulated by two allocations on the same device (IPC handles cannot be opened by the process that exported them);
the real two-process run over NVLink is bench.py --gpus N (cfg5_eval: rows compared with the NCCL all-gather)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from jabd_b200 import _lib
    return _lib


def test_p2p_allgather_two_virtual_ranks():
    L = _lib()
    dev = torch.device("cuda", 0)
    n_ranks, nbytes = 2, 1440128 + 8            # not a multiple of 16: the tail path
    pad = (nbytes + 255) // 256 * 256
    bufs = [torch.zeros(n_ranks * pad, dtype=torch.uint8, device=dev) for _ in range(n_ranks)]
    flags = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(n_ranks)]
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    vp = ctypes.c_void_p * n_ranks
    pb = vp(*[b.data_ptr() for b in bufs])
    pf = vp(*[f.data_ptr() for f in flags])
    acks = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(n_ranks)]
    pa = vp(*[f.data_ptr() for f in acks])
    streams = [torch.cuda.Stream(dev) for _ in range(n_ranks)]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    g = torch.Generator().manual_seed(7)
    for seq in (1, 2, 3):
        src = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, generator=g).to(dev) for _ in range(n_ranks)]
        cnt = [torch.zeros(16, dtype=torch.int32, device=dev) for _ in range(n_ranks)]
        torch.cuda.synchronize()
        for r in range(n_ranks):
            # handshake from the second exchange on: both virtual ranks acknowledge and wait inside their kernels (rank 0's
            # kernel spins until rank 1's has started: two kernels on two streams of one device)
            L.call("jabd_p2p_allgather", ctypes.c_void_p(src[r].data_ptr()), ctypes.c_size_t(nbytes), pb, ctypes.c_size_t(r * pad), pf,
                   pa, ctypes.c_void_p(acks[r].data_ptr()), n_ranks, r, ctypes.c_uint64(seq), ctypes.c_uint64(seq - 1),
                   ctypes.c_void_p(cnt[r].data_ptr()), 1.0, ctypes.c_void_p(status.data_ptr()),
                   ctypes.c_void_p(streams[r].cuda_stream))
        for r in range(n_ranks):
            L.call("jabd_p2p_wait", ctypes.c_void_p(flags[r].data_ptr()), n_ranks, ctypes.c_uint64(seq), 1.0,
                   ctypes.c_void_p(status.data_ptr()), st)
        torch.cuda.synchronize()
        assert acks[0][:n_ranks].tolist() == [seq - 1] * n_ranks and acks[1][:n_ranks].tolist() == [seq - 1] * n_ranks
        assert int(status.item()) == 0
        for r in range(n_ranks):
            assert flags[r][:n_ranks].tolist() == [seq] * n_ranks
            assert all(int(c.sum()) == 0 for c in cnt)               # the arrival counters are left at zero
            for j in range(n_ranks):
                assert torch.equal(bufs[r][j * pad:j * pad + nbytes], src[j])


def test_p2p_wait_times_out_instead_of_hanging():
    L = _lib()
    dev = torch.device("cuda", 0)
    flags = torch.zeros(16, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("jabd_p2p_wait", ctypes.c_void_p(flags.data_ptr()), 3, ctypes.c_uint64(5), 0.05, ctypes.c_void_p(status.data_ptr()),
           ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    assert 1 <= int(status.item()) <= 3


def test_p2p_argument_errors():
    L = _lib()
    vp = ctypes.c_void_p * 1
    with pytest.raises(ValueError):
        L.call("jabd_p2p_allgather", None, ctypes.c_size_t(16), vp(0), ctypes.c_size_t(0), vp(0), None, None, 1, 0, ctypes.c_uint64(1),
               ctypes.c_uint64(0), None, 1.0, None, None)
    with pytest.raises(ValueError):
        L.call("jabd_p2p_wait", None, 1, ctypes.c_uint64(1), 1.0, None, None)


def test_peer_gather_single_rank_roundtrip():
    from jabd_b200 import sharding
    dev = torch.device("cuda", 0)
    B, keep = 4, 50
    pg = sharding.PeerGather(B, keep, dev, depth=3)
    assert pg.transport == "p2p" and pg.world == 1
    g = torch.Generator().manual_seed(3)
    want = {}
    for k in range(8):                        # every slot is reused: acknowledgement + wait path
        slot = k % pg.depth
        pg.acquire(slot)
        d = torch.rand((B, keep, 15), generator=g)
        c = torch.randint(0, keep, (B,), dtype=torch.int32, generator=g)
        pg.dets(slot).copy_(d.to(dev))
        pg.counts(slot).copy_(c.to(dev))
        pg.launch(slot)
        want[slot] = (d, c)
        if k >= 1:
            gd, gc = pg.result((k - 1) % pg.depth)
            assert tuple(gd.shape) == (1, B, keep, 15) and tuple(gc.shape) == (1, B)
            torch.cuda.synchronize()
            assert np.array_equal(gd[0].cpu().numpy(), want[(k - 1) % pg.depth][0].numpy())
            assert np.array_equal(gc[0].cpu().numpy(), want[(k - 1) % pg.depth][1].numpy())
    assert pg.status() == 0
    pg.close()
