"""GPU parity of the MultiBox loss (hard-negative mining + loss reductions, forward and backward)
against the golden fixtures recorded from the reference's own MultiBoxLoss and against oracle/torch_port on larger
seeded batches.  Everything goes through the C-ABI (jabd_multibox_loss_forward / _backward).

Tolerances: loss scalars rtol 1e-5 (fp32 exp/log differ by ulps between CUDA and torch's CPU kernels, and the CPU sums in a
different order); gradients rtol 1e-4 / atol 1e-7.  The set of mined negatives must be identical except for elements whose
rank value is within 4 ulp of the cut (either side may legitimately pick them when two values round differently)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

VAR = [0.1, 0.2]
LOSS_CASES = [("s160", (160, 160), 3, 12), ("s640", (640, 640), 2, None)]


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from jabd_b200 import anchors, batched, config, retinaface_training, synth
    from oracle import torch_port
    return dict(anchors=anchors, batched=batched, config=config, rt=retinaface_training, synth=synth, tp=torch_port)


def _inputs(mods, size, batch, count, cfg_id=6):
    synth = mods["synth"]
    pri = mods["anchors"].Anchors(mods["config"].cfg_mnet, image_size=size).get_anchors()
    P = pri.shape[0]
    targets = [synth.make_gt(cfg_id, i, size, count=count) for i in range(batch)]
    preds = [synth.make_logits(cfg_id, i, P) for i in range(batch)]
    cpu = tuple(torch.stack([p[k] for p in preds]) for k in range(3))
    return pri, targets, cpu


def _run_gpu(mods, pri, targets, cpu_preds, weights=(1.0, 2.0, 0.5)):
    preds = tuple(t.cuda().requires_grad_(True) for t in cpu_preds)
    crit = mods["rt"].MultiBoxLoss(2, 0.35, 7, VAR, True)
    l, c, m = crit(preds, pri, [t.cuda() for t in targets])
    (weights[0] * l + weights[1] * c + weights[2] * m).backward()
    torch.cuda.synchronize()
    return np.array([l.item(), c.item(), m.item()], dtype=np.float32), tuple(p.grad.cpu().numpy() for p in preds)


def _check_selection(sel_gpu, sel_ref, max_diff):
    diff = int((sel_gpu != sel_ref).sum())
    assert diff <= max_diff, "%d priors selected differently (allowed %d near-tie swaps)" % (diff, max_diff)
    return diff


@pytest.mark.parametrize("tag,size,batch,count", LOSS_CASES)
def test_multibox_loss_golden(mods, tag, size, batch, count):
    g = load_golden("loss.npz")
    pri, targets, cpu = _inputs(mods, size, batch, count)
    losses, (g_loc, g_conf, g_landm) = _run_gpu(mods, pri, targets, cpu)
    np.testing.assert_allclose(losses, g[tag + "_losses"], rtol=1e-5)
    P = pri.shape[0]
    sel = (np.abs(g_conf).sum(2) != 0).reshape(-1)
    sel_ref = np.unpackbits(g[tag + "_sel"])[:batch * P].astype(bool)
    swaps = _check_selection(sel, sel_ref, 2 * batch)
    np.testing.assert_allclose(np.array([np.abs(g_loc).sum(), np.abs(g_conf).sum(), np.abs(g_landm).sum()]), g[tag + "_gsum"], rtol=1e-4)
    if tag == "s160":
        np.testing.assert_allclose(g_loc, g[tag + "_g_loc"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(g_landm, g[tag + "_g_landm"], rtol=1e-4, atol=1e-7)
        if swaps == 0:
            np.testing.assert_allclose(g_conf, g[tag + "_g_conf"], rtol=1e-4, atol=1e-7)
    else:
        idx = g[tag + "_g_loc_nz_idx"]
        np.testing.assert_allclose(g_loc.reshape(-1, 4)[idx], g[tag + "_g_loc_nz"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(g_landm.reshape(-1, 10)[idx], g[tag + "_g_landm_nz"], rtol=1e-4, atol=1e-7)
        assert int((np.abs(g_loc).sum(2).reshape(-1) != 0).sum()) == len(idx)
        if swaps == 0:
            np.testing.assert_allclose(g_conf.reshape(-1, 2)[g[tag + "_g_conf_nz_idx"]], g[tag + "_g_conf_nz"], rtol=1e-4, atol=1e-7)


def test_multibox_loss_cfg2_batch_vs_torch_port(mods):
    """8 images of the cfg2 distribution (1..300 faces at 640x640): losses, selection mask and gradients against the torch
    port of the reference (CPU)."""
    synth, tp = mods["synth"], mods["tp"]
    size, batch = (640, 640), 8
    pri = mods["anchors"].Anchors(mods["config"].cfg_mnet, image_size=size).get_anchors()
    P = pri.shape[0]
    targets = synth.make_gt_batch(2, batch, size)
    preds = [synth.make_logits(2, i, P) for i in range(batch)]
    cpu = tuple(torch.stack([p[k] for p in preds]) for k in range(3))
    losses, (g_loc, g_conf, g_landm) = _run_gpu(mods, pri, targets, cpu)
    ref_preds = tuple(t.clone().requires_grad_(True) for t in cpu)
    l, c, m, aux = tp.multibox_loss(ref_preds, pri.cpu(), targets, 0.35, VAR, 7, return_aux=True)
    (1.0 * l + 2.0 * c + 0.5 * m).backward()
    np.testing.assert_allclose(losses, [l.item(), c.item(), m.item()], rtol=1e-5)
    sel_ref = (aux["pos"] | aux["neg"]).numpy().reshape(-1)
    sel = (np.abs(g_conf).sum(2) != 0).reshape(-1)
    swaps = _check_selection(sel, sel_ref, 2 * batch)
    np.testing.assert_allclose(g_loc, ref_preds[0].grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(g_landm, ref_preds[2].grad.numpy(), rtol=1e-4, atol=1e-7)
    same = (sel == sel_ref).reshape(batch, P)
    np.testing.assert_allclose(g_conf[same], ref_preds[1].grad.numpy()[same], rtol=1e-4, atol=1e-7)
    assert swaps <= 2 * batch


def test_multibox_loss_aux_and_edge_cases(mods):
    """Selection mask bits, the num_neg clamp at P-1 (:280), an image whose negatives all tie, and no positives at all."""
    batched = mods["batched"]
    dev = torch.device("cuda")
    B, P = 3, 700
    g = torch.Generator().manual_seed(7)
    loc = torch.randn((B, P, 4), generator=g).to(dev)
    conf = torch.randn((B, P, 2), generator=g).to(dev)
    conf[1] = 0.25                                   # image 1: every rank value ties -> lowest indices are mined
    landm = torch.randn((B, P, 10), generator=g).to(dev)
    loc_t = torch.randn((B, P, 4), generator=g).to(dev)
    landm_t = torch.randn((B, P, 10), generator=g).to(dev)
    conf_t = torch.zeros((B, P), dtype=torch.int64, device=dev)
    conf_t[0, :150] = 1                              # 7 * 150 > P - 1 -> clamp to P - 1 = 699 ranks
    conf_t[0, 150:160] = -1
    conf_t[1, 10:13] = 1                             # 3 positives -> 21 mined negatives, all ties
    # image 2: no positives -> no negatives
    l, c, m, mask, norms = batched.multibox_loss((loc, conf, landm), loc_t, conf_t, landm_t, 7, return_aux=True)
    mask = mask.cpu().numpy()
    assert np.array_equal(mask & 1, (conf_t != 0).cpu().numpy().astype(np.uint8))
    assert np.array_equal((mask >> 1) & 1, (conf_t > 0).cpu().numpy().astype(np.uint8))
    neg = (mask >> 2) & 1
    # image 0: ranks < 699 of 700: everything except the single lowest-ranked element; positives carry rank value 0
    assert int(neg[0].sum()) == P - 1
    # image 1: stable descending sort of all-equal values (positives are 0 < the tie value): the first 21 non-positive indices
    expect = np.zeros(P, dtype=np.uint8)
    nonpos = [i for i in range(P) if not (10 <= i < 13)]
    expect[nonpos[:21]] = 1
    assert np.array_equal(neg[1], expect)
    assert int(neg[2].sum()) == 0
    assert float(norms[0]) == 163.0 and float(norms[1]) == 153.0
    # cross-check the sums with plain torch on the GPU tensors (float64)
    pos = torch.from_numpy((mask & 1).astype(bool)).to(dev)
    pos1 = torch.from_numpy(((mask >> 1) & 1).astype(bool)).to(dev)
    sel = torch.from_numpy(((mask & 5) != 0)).to(dev)
    F = torch.nn.functional
    ref_l = F.smooth_l1_loss(loc.double()[pos], loc_t.double()[pos], reduction="sum") / 163.0
    ref_m = F.smooth_l1_loss(landm.double()[pos1], landm_t.double()[pos1], reduction="sum") / 153.0
    ref_c = F.cross_entropy(conf.double()[sel], pos[sel].long(), reduction="sum") / 163.0
    np.testing.assert_allclose([l.item(), c.item(), m.item()], [ref_l.item(), ref_c.item(), ref_m.item()], rtol=1e-5)


def test_multibox_loss_ties_across_cluster_ranks(mods):
    """The forward kernel splits every image over a 4-CTA cluster (contiguous prior ranges).  With all rank values tied the
    mined negatives are the lowest non-positive indices (stable descending sort, R/nets/retinaface_training.py:270-281) -- a
    set that crosses the CTA boundaries; P is not a multiple of 4 and the batch needs more than one wave of clusters."""
    batched = mods["batched"]
    dev = torch.device("cuda")
    B, P = 40, 701
    g = torch.Generator().manual_seed(11)
    loc = torch.randn((B, P, 4), generator=g).to(dev)
    landm = torch.randn((B, P, 10), generator=g).to(dev)
    loc_t = torch.randn((B, P, 4), generator=g).to(dev)
    landm_t = torch.randn((B, P, 10), generator=g).to(dev)
    conf = torch.full((B, P, 2), 0.5, device=dev)
    conf_t = torch.zeros((B, P), dtype=torch.int64, device=dev)
    npos = [(b * 7) % 60 for b in range(B)]                      # 0..59 positives at the END of the prior range
    for b in range(B):
        if npos[b]:
            conf_t[b, P - npos[b]:] = 1
    # two images with a partial tie group instead: distinct values above a large tie plateau
    conf[3, :300, 0] = torch.linspace(-3.0, -1.0, 300, device=dev)      # larger rank value than the plateau, all distinct
    conf[5, 200:500, 0] = -2.0                                          # a second, higher plateau in the middle
    l, c, m, mask, norms = batched.multibox_loss((loc, conf, landm), loc_t, conf_t, landm_t, 7, return_aux=True)
    neg = ((mask >> 2) & 1).cpu().numpy()
    cf = conf.cpu().double()
    rank_val = torch.logsumexp(cf, 2) - cf[:, :, 0]
    rank_val[conf_t.cpu() != 0] = 0
    for b in range(B):
        want = min(7 * npos[b], P - 1)
        order = torch.sort(rank_val[b], descending=True, stable=True)[1][:want].numpy()
        expect = np.zeros(P, np.uint8)
        expect[order] = 1
        assert np.array_equal(neg[b], expect), b
    assert float(norms[0]) == float(max(sum(npos), 1))
