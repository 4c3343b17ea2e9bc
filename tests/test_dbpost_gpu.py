"""GPU parity of the DB post-process kernels (dbpost.cu) against BoxesFromBitmap + FilterTagDetRes restated with the
reference's OpenCV / Clipper calls (oracle/ocr_ops.py), on synthetic probability maps (no network involved).

Tolerances (BASELINE.json north_star): thresholded bitmaps bit-exact; box vertices within 1 px; same boxes, same order.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = dict(det_db_thresh=0.2, det_db_box_thresh=0.4, det_db_unclip_ratio=1.8)


@pytest.fixture(scope="module")
def det(models_dir):
    import b200ocr
    return b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=512, det_db_score_mode="fast", **PARAMS)


def _oracle(pred, src_h, src_w, **kw):
    from oracle import ocr_ops
    p = dict(PARAMS)
    p.update(kw)
    h, w = pred.shape
    return ocr_ops.det_postprocess(pred, np.float32(h) / np.float32(src_h), np.float32(w) / np.float32(src_w), src_h, src_w,
                                   p["det_db_thresh"], p["det_db_box_thresh"], p["det_db_unclip_ratio"], "fast",
                                   p.get("use_dilation", False))


def _compare(got, ref, tol=1):
    assert len(got) == len(ref), (len(got), len(ref))
    for g, r in zip(got, ref):
        assert np.abs(np.asarray(g) - np.asarray(r)).max() <= tol, (g.tolist(), r)


@pytest.mark.parametrize("seed", range(8))
def test_boxes_match_oracle_on_synthetic_maps(det, seed):
    import synth_data
    h, w = [(320, 512), (192, 384), (608, 960), (160, 256)][seed % 4]
    pred = synth_data.prob_map(seed, h, w, n_boxes=4 + 2 * seed, rings=2, lines=3)
    got, bm = det.postprocess(pred, 2 * h, 2 * w, want_bitmap=True)
    ref, rbm = _oracle(pred, 2 * h, 2 * w)
    assert np.array_equal(bm, rbm)  # bit-exact bitmap
    assert len(ref) > 0
    # vertices within 1 px of the detection map = 2 px after FilterTagDetRes maps them to the 2x larger source image
    _compare(got, ref, tol=2)


def test_edge_cases(det):
    h, w = 96, 160
    z = np.zeros((h, w), np.float32)
    assert len(det.postprocess(z, h, w)) == 0 == len(_oracle(z, h, w)[0])
    full = np.full((h, w), 0.9, np.float32)
    _compare(det.postprocess(full, h, w), _oracle(full, h, w)[0])
    thin = z.copy()
    thin[10, 5:100] = 0.9      # 1-px horizontal run: 2 contour points -> skipped
    thin[20:80, 7] = 0.9       # vertical run
    for k in range(30):
        thin[30 + k, 60 + k] = 0.9  # diagonal run
    thin[50, 140] = 0.9        # lone pixel
    assert len(det.postprocess(thin, h, w)) == 0 == len(_oracle(thin, h, w)[0])
    # a ring produces an outer border and a hole border; a blob inside the hole a third contour
    import cv2
    ring = z.copy()
    cv2.ellipse(ring, (80, 48), (60, 30), 0, 0, 360, 0.95, 6)
    cv2.rectangle(ring, (60, 40), (100, 56), 0.9, -1)
    _compare(det.postprocess(ring, h, w), _oracle(ring, h, w)[0])
    # blobs touching the image frame
    edge = z.copy()
    edge[0:12, 0:70] = 0.8
    edge[h - 9:h, w - 60:w] = 0.85
    edge[30:60, w - 14:w] = 0.7
    _compare(det.postprocess(edge, h, w), _oracle(edge, h, w)[0])


def test_threshold_is_bit_exact_including_boundaries(det):
    # values that straddle (uchar)(p*255) > 51 for thresh 0.2 (51.0 exactly after float math)
    vals = np.array([0.2, 0.2039215, 0.2039216, 0.20392157, 0.2078431, 0.21, 52 / 255, 51 / 255, 0.99999, 1.0, 0.0], np.float32)
    pred = np.tile(vals, (16, 8)).astype(np.float32)
    _, bm = det.postprocess(pred, 16, pred.shape[1], want_bitmap=True)
    from oracle import ocr_ops
    assert np.array_equal(bm, ocr_ops.threshold_bitmap(pred, 0.2))


def test_contour_order_and_candidate_cap(det, models_dir):
    """> 1000 contours: only the first 1000 in the reference's order (last found first) are considered."""
    import b200ocr
    h, w = 256, 512
    pred = np.zeros((h, w), np.float32)
    # 24 x 48 = 1152 small blobs, each large enough to become a box (8x5 px)
    for r in range(24):
        for c in range(48):
            pred[4 + r * 10: 9 + r * 10, 2 + c * 10: 10 + c * 10] = 0.9
    got = det.postprocess(pred, h, w, cap=1200)
    ref, _ = _oracle(pred, h, w)
    assert len(ref) == 1000
    _compare(got, ref)


def test_dilation_and_other_thresholds(models_dir):
    import b200ocr, synth_data
    d = b200ocr.Detector(f"{models_dir}/det", det_db_thresh=0.3, det_db_box_thresh=0.5, det_db_unclip_ratio=2.0,
                         det_db_score_mode="fast", use_dilation=True)
    pred = synth_data.prob_map(11, 160, 256, n_boxes=9)
    got, bm = d.postprocess(pred, 320, 512, want_bitmap=True)
    ref, rbm = _oracle(pred, 320, 512, det_db_thresh=0.3, det_db_box_thresh=0.5, det_db_unclip_ratio=2.0, use_dilation=True)
    assert np.array_equal(bm, rbm)
    _compare(got, ref)
    with pytest.raises(b200ocr.Error, match="neither"):
        b200ocr.Detector(f"{models_dir}/det", det_db_score_mode="medium")


@pytest.mark.parametrize("seed", range(8))
def test_slow_score_mode_matches_polygon_score_acc(models_dir, seed):
    """det_db_score_mode = "slow": PolygonScoreAcc (cv2.fillPoly of the contour + masked mean, postprocess_op.cpp:170-214)
    against the GPU's nested-region sums, on maps with rings (hole borders, whose polygon covers the low-probability
    hole) and nested blobs; a box_thresh in the middle of the score range makes the scores decide which boxes survive."""
    import b200ocr, synth_data
    from oracle import ocr_ops
    h, w = [(320, 512), (192, 384), (160, 256), (96, 160)][seed % 4]
    pred = synth_data.prob_map(100 + seed, h, w, n_boxes=5 + seed, rings=4, lines=2)
    if seed % 2:  # nested: a blob inside a ring's hole, and a hole inside that blob
        pred[h // 2 - 6:h // 2 + 6, w // 2 - 20:w // 2 + 20] = 0.95
        pred[h // 2 - 2:h // 2 + 2, w // 2 - 8:w // 2 + 8] = 0.02
        pred[h // 2 - 14:h // 2 - 10, w // 2 - 30:w // 2 + 30] = 0.9
        pred[h // 2 + 10:h // 2 + 14, w // 2 - 30:w // 2 + 30] = 0.9
        pred[h // 2 - 14:h // 2 + 14, w // 2 - 30:w // 2 - 26] = 0.9
        pred[h // 2 - 14:h // 2 + 14, w // 2 + 26:w // 2 + 30] = 0.9
    differs = False
    for box_thresh in (0.3, 0.55, 0.7):
        d = b200ocr.Detector(f"{models_dir}/det", det_db_thresh=0.2, det_db_box_thresh=box_thresh, det_db_unclip_ratio=1.8,
                             det_db_score_mode="slow")
        got = d.postprocess(pred, 2 * h, 2 * w)
        ref, _ = ocr_ops.det_postprocess(pred, np.float32(0.5), np.float32(0.5), 2 * h, 2 * w, 0.2, box_thresh, 1.8, "slow", False)
        fast, _ = ocr_ops.det_postprocess(pred, np.float32(0.5), np.float32(0.5), 2 * h, 2 * w, 0.2, box_thresh, 1.8, "fast", False)
        _compare(got, ref, tol=2)
        differs |= len(fast) != len(ref)
    if seed == 0:
        assert differs  # the two score modes do select different boxes on this map
