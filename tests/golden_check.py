"""Comparison of one image's stage outputs with tests/golden/expected_stages.json (written by tools/make_golden.py).
Used by the CPU test (the oracle must reproduce its own golden vectors) and by the GPU test (the CUDA stages, fed the
golden crops, must match them) with the tolerances of the north star: boxes within 1 px of the detection map,
probabilities within 1e-2, recognised strings identical.  A line whose ORACLE top-2 margin is below 5e-2 at some step (a
near tie: the fixtures include a Chinese ID card the synthetic recognizer weights were never fitted to) is exempt, a
classifier label may differ only where the oracle's own score is within 1e-2 of 0.5.  The strict string checks against
the live oracle are in tests/test_stages_gpu.py."""
import json
import os

import numpy as np

SCORE_TOL = 1e-2
PIN_MARGIN = 5e-2   # smallest top-2 margin (over a line's steps) from which the golden string is binding


def load(golden_dir):
    with open(os.path.join(golden_dir, "expected_stages.json"), encoding="utf-8") as f:
        return json.load(f)


def crops_of(img, g):
    """the ROI crops the golden cls / rec entries were computed on (views of `img`)"""
    return [img[y:y + h, x:x + w] for x, y, w, h in g["rois"]]


def check_det(boxes, g, box_tol_px=None):
    ref = g["boxes"]
    assert len(boxes) == len(ref), (g["name"], len(boxes), len(ref))
    if box_tol_px is None:  # 1 px of the detection map in source pixels (FilterTagDetRes divides by the resize ratio)
        m = max(g["rows"], g["cols"])
        box_tol_px = int(np.ceil(m / 512.0)) if m > 512 else 1
    for a, r in zip(boxes, ref):
        assert np.abs(np.asarray(a, np.int64).reshape(4, 2) - np.asarray(r, np.int64)).max() <= box_tol_px, (g["name"], a, r)


def check_cls(labels, scores, g):
    assert len(labels) == len(g["cls_labels"])
    assert np.abs(np.asarray(scores, np.float64) - np.asarray(g["cls_scores"])).max(initial=0.0) < SCORE_TOL, g["name"]
    for a, b, s in zip(labels, g["cls_labels"], g["cls_scores"]):
        assert int(a) == b or abs(s - 0.5) < SCORE_TOL, (g["name"], a, b, s)


def check_rec(texts, scores, g):
    """returns how many lines had their confidence checked"""
    assert len(texts) == len(g["rec_texts"])
    checked = 0
    for t, s, rt, rs, margin in zip(texts, scores, g["rec_texts"], g["rec_scores"], g["rec_min_margin"]):
        # both of a step's top-2 probabilities may move by SCORE_TOL, so an arg-max can only be pinned where the golden
        # margin is well above twice that; such a line must give the identical string and the same confidence
        if margin >= PIN_MARGIN:
            assert t == rt, (g["name"], t, rt, margin)
            assert abs(float(s) - rs) < SCORE_TOL, (g["name"], t, s, rs)
            checked += 1
    return checked
