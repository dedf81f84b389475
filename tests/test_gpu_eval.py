"""GPU parity of the WIDER-FACE AP evaluation against the golden vectors recorded from the
reference's utils/utils_map.py and against the CPU oracle: every counter and index is exact; fp64 IoUs, normalised
scores and the final AP are bit-identical (same IEEE operations in the same order)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from jabd_b200 import synth, utils_map
    from oracle import wider_eval as ow
    g = load_golden("wider_eval.npz")
    n = int(g["n_images"])
    imgs = [synth.make_eval_image(6, i) for i in range(n)]
    return dict(um=utils_map, ow=ow, g=g, n=n, imgs=imgs, synth=synth)


def test_bbox_overlaps_and_norm_score(env):
    um, g = env["um"], env["g"]
    out = um.bbox_overlaps(g["overlaps_in"], g["overlaps_in"][::-1].copy())
    np.testing.assert_array_equal(out, g["overlaps"])                      # includes the 0/0 = NaN pair
    pred = {"ev": {str(i): env["imgs"][i][2].copy() for i in range(env["n"])}}
    um.norm_score(pred)
    for i in range(env["n"]):
        np.testing.assert_array_equal(pred["ev"][str(i)], g["norm_%d" % i])


def test_image_eval_and_img_pr_info_golden(env):
    um, g = env["um"], env["g"]
    for i in range(env["n"]):
        gt, keeps, _ = env["imgs"][i]
        p = g["norm_%d" % i]
        if len(gt) == 0 or len(p) == 0:
            continue
        rec, prop = um.image_eval(p, gt, keeps[2], 0.4)
        np.testing.assert_array_equal(rec, g["recall_%d" % i])
        np.testing.assert_array_equal(prop, g["proposal_%d" % i])
        if i < 4:
            np.testing.assert_array_equal(um.img_pr_info(1000, p, prop, rec), g["pr_info_%d" % i])


def test_pr_counters_and_ap_golden(env):
    um, g = env["um"], env["g"]
    preds = [g["norm_%d" % i] for i in range(env["n"])]
    gts = [im[0] for im in env["imgs"]]
    keeps = [[im[1][s] for im in env["imgs"]] for s in range(3)]
    for s, name in enumerate(("easy", "medium", "hard")):
        np.testing.assert_array_equal(um.pr_counters(preds, gts, keeps[s], 0.4, 1000), g["pr_curve_" + name])
    aps = um.evaluate_arrays(preds, gts, keeps, 0.4, 1000)
    assert aps == [float(g["ap_easy"]), float(g["ap_medium"]), float(g["ap_hard"])]


def test_pr_counters_vs_oracle_edge_cases(env):
    """Larger ragged set against the oracle: empty images, all-ignored GT, tied scores, scores of exactly 0 and 1,
    thresholds other than 1000, duplicate GT (argmax takes the first), zero-area boxes (IoU NaN never matches)."""
    um, ow = env["um"], env["ow"]
    imgs = [env["synth"].make_eval_image(8, i, count=(None if i % 4 else 150)) for i in range(40)]
    preds = ow.norm_scores([im[2] for im in imgs])
    gts = [im[0].copy() for im in imgs]
    keeps = [im[1][1].copy() for im in imgs]
    keeps[2][:] = 0                                           # every GT of this image ignored
    gts[4][1] = gts[4][0]                                     # duplicate GT rows
    preds[5][0, 4], preds[5][-1, 4] = 1.0, 0.0
    gts[6][0, 2:] = 0.0                                       # zero-area GT
    preds[6][0, 2:4] = 0.0                                    # zero-area prediction
    preds[6][0, :2] = gts[6][0, :2]
    for thr, tn in ((0.4, 1000), (0.5, 1000), (0.4, 37)):
        ref = ow.pr_counters(preds, gts, keeps, thr, tn)
        out, rec, prop, off = um.pr_counters(preds, gts, keeps, thr, tn, return_image_eval=True)
        np.testing.assert_array_equal(out, ref)
        for i in (0, 2, 4, 6):
            if len(gts[i]) and len(preds[i]):
                r, p = ow.image_eval(preds[i], gts[i], keeps[i], thr)
                np.testing.assert_array_equal(rec[off[i]:off[i + 1]], r)
                np.testing.assert_array_equal(prop[off[i]:off[i + 1]], p)


def test_evaluation_from_files(env, tmp_path):
    """The reference's wire format end to end: per-event txt files (name line, count line, ``x y w h score`` rows)
    and the four WIDER .mat files -> evaluation() -> easy / medium / hard AP equal to the golden values."""
    from scipy.io import savemat
    um, g, n, imgs = env["um"], env["g"], env["n"], env["imgs"]
    events = ["0--Parade", "1--Handshaking"]
    split = [list(range(0, n // 2)), list(range(n // 2, n))]
    pred_dir, gt_dir = tmp_path / "pred", tmp_path / "gt"
    gt_dir.mkdir()

    def obj(items):
        a = np.empty((len(items), 1), dtype=object)
        for k, it in enumerate(items):
            a[k, 0] = it
        return a

    ev_list, file_list, box_list = [], [], []
    gl = [[], [], []]
    for e, ids in zip(events, split):
        (pred_dir / e).mkdir(parents=True)
        for i in ids:
            rows = imgs[i][2]
            with open(pred_dir / e / ("img_%d.txt" % i), "w") as f:
                f.write("%s/img_%d.jpg\n%d\n" % (e, i, len(rows)))
                for r in rows:
                    f.write("%r %r %r %r %r \n" % tuple(float(v) for v in r))
        ev_list.append(e)                                           # cell of char arrays, like the WIDER .mat files
        file_list.append(obj(["img_%d" % i for i in ids]))
        box_list.append(obj([imgs[i][0] for i in ids]))
        for s in range(3):
            gl[s].append(obj([(np.nonzero(imgs[i][1][s])[0] + 1).reshape(-1, 1) for i in ids]))
    savemat(gt_dir / "wider_face_val.mat", {"event_list": obj(ev_list), "file_list": obj(file_list), "face_bbx_list": obj(box_list)})
    for s, name in enumerate(("easy", "medium", "hard")):
        savemat(gt_dir / ("wider_%s_val.mat" % name), {"gt_list": obj(gl[s])})
    aps = um.evaluation(str(pred_dir), str(gt_dir), 0.4)
    assert aps == [float(g["ap_easy"]), float(g["ap_medium"]), float(g["ap_hard"])]
