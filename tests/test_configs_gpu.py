"""BASELINE.json configs 2, 3 and 5 at their full sizes, checked through size-independent properties (a row / image
of a large batch must equal the same row / image processed alone) plus oracle spot checks.

  C2  recognition-only: 4096 synthetic 48x320 crops through CRNN/SVTR forward + CTC greedy decode
  C3  detection-only: batch 64 synthetic cards at 960 max side ([64,3,608,960]) through DB forward + DBPostProcess
  C5  dense document page 2048x2048 with 200+ text lines (variable-width rec packing)
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_recognition_only_4096_crops(models_dir):
    import b200ocr, synth_data
    from oracle.pipeline import OracleRecognizer
    label = f"{models_dir}/rec/ppocr_keys_v1.txt"
    crops = synth_data.rec_crops(4096, 48, 320, seed=0)
    rec = b200ocr.Recognizer(f"{models_dir}/rec", label, rec_batch_num=6, rec_img_h=48, rec_img_w=320)
    texts, scores = rec.run(list(crops))
    assert len(texts) == 4096 and sum(1 for t in texts if t) > 4000
    # batch-size independence: a crop decoded inside the 4096-batch == the same crop decoded in a small batch
    pick = [0, 1, 777, 2048, 4095, 1234, 3333, 9]
    t2, s2 = rec.run([crops[i] for i in pick])
    for k, i in enumerate(pick):
        assert t2[k] == texts[i] and s2[k] == scores[i]
    # oracle check on 48 of the 4096 crops: identical strings, confidence within 1e-2 (near-tie lines exempt, see
    # tests/test_stages_gpu.py)
    from test_stages_gpu import _check_line
    orec = OracleRecognizer(f"{models_dir}/rec", label, 6, 48, 320)
    pick2 = pick + list(range(100, 4000, 100))
    rt, rs, raw = orec.run([crops[i] for i in pick2], want_raw=True)
    for k, i in enumerate(pick2):
        _check_line(texts[i], scores[i], rt[k], rs[k], raw[k])


def test_c3_detection_only_batch64_at_960(models_dir):
    import b200ocr, synth_data
    det = b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=960, det_db_thresh=0.3,
                           det_db_box_thresh=0.5, det_db_unclip_ratio=2.0, det_db_score_mode="fast")
    imgs = [synth_data.card(500 + i) for i in range(64)]
    got, _, _ = det.preprocess(imgs[0])
    assert got.shape == (3, 608, 960)
    boxes = det.run_batch(imgs)
    assert len(boxes) == 64
    for i in (0, 31, 63):  # an image of the batch == the same image alone
        assert np.array_equal(det.run(imgs[i]), boxes[i])
    for b in boxes:
        for q in b:
            assert (q[:, 0] >= 0).all() and (q[:, 0] < 1024).all() and (q[:, 1] >= 0).all() and (q[:, 1] < 640).all()
    # the detector was fitted at the 512 scale; at 960 it still has to produce boxes for the pipeline to be exercised
    assert sum(len(b) for b in boxes) > 64
    # oracle parity at this configuration's own size, stage by stage on identical upstream tensors (north star):
    #  (1) network: probability map of [1,3,608,960] within 1e-2;
    #  (2) post-process: the GPU's DBPostProcess on the ORACLE's probability map gives the oracle's boxes -- same count,
    #      same order, vertices within 1 px of the map (= 2 source pixels after FilterTagDetRes' division by the ratio);
    #  (3) end to end the two box lists agree except for candidates that straddle a threshold (the synthetic detector
    #      was fitted at the 512 scale; at 960 some of its blobs sit at the 0.3 / 0.5 decision thresholds, where a
    #      1e-2 probability difference flips a discontinuous filter): >= 90 % of the oracle's boxes have a GPU box
    #      within 2 px.
    _det_parity_at(models_dir, det, [imgs[i] for i in (0, 7, 21, 40, 63)], [boxes[i] for i in (0, 7, 21, 40, 63)],
                   960, 0.3, 0.5, 2.0, px_tol=2, min_boxes=20)


def _det_parity_at(models_dir, det, imgs, gpu_boxes, limit, thresh, box_thresh, unclip, px_tol, min_boxes):
    import b200ocr
    from oracle import ocr_ops
    from oracle.pipeline import OracleDetector
    odet = OracleDetector(f"{models_dir}/det", "max", limit, thresh, box_thresh, unclip, "fast", False)
    net = b200ocr.Net(f"{models_dir}/det", 0, b200ocr.NET_NO_GRAPH)
    n_ref = n_matched = 0
    for k, (im, got) in enumerate(zip(imgs, gpu_boxes)):
        pred, rh, rw = odet.forward(im)
        ref = odet.post(pred, rh, rw, im.shape[0], im.shape[1])[0]
        if k < 2:  # (1)
            x, _, _ = ocr_ops.det_preprocess(im, "max", limit)
            prob, _ = net.forward(x, int(thresh * 255))
            assert float(np.abs(prob[0] - pred).max()) <= 1e-2
        mine = det.postprocess(pred, im.shape[0], im.shape[1])  # (2)
        assert len(mine) == len(ref), (k, len(mine), len(ref))
        for g, r in zip(mine, ref):
            assert np.abs(g - np.asarray(r)).max() <= px_tol, (k, g.tolist(), r)
        for r in ref:  # (3)
            n_ref += 1
            n_matched += any(np.abs(g - np.asarray(r)).max() <= px_tol for g in got)
    assert n_ref >= min_boxes and n_matched >= 0.9 * n_ref, (n_matched, n_ref)


def test_c5_dense_page_with_200_lines(models_dir):
    """The worker's fixed limit_side_len=512 shrinks a 2048^2 page 4x (its 24-px lines become 6 px: too small for any
    DB detector), so the dense-page config runs the stages directly at limit_side_len=960, where the page's 220 lines
    are found, and pushes all of their crops through ONE recognizer call (ragged packing of >20 distinct widths)."""
    import b200ocr, synth_data
    from oracle import ocr_ops
    page = synth_data.page(3)
    det = b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=960, det_db_thresh=0.2,
                           det_db_box_thresh=0.4, det_db_unclip_ratio=1.8, det_db_score_mode="fast")
    boxes = det.run(page)
    assert len(boxes) >= 150, len(boxes)
    ys = boxes[:, :, 1].min(1)
    assert ys[0] > ys[-1]                   # the reference's contour order: last-found (bottom-most) first
    crops = []
    for b in boxes:
        x, y, w, h = ocr_ops.bounding_rect_crop(b.tolist(), page.shape[0], page.shape[1])
        crops.append(page[y:y + h, x:x + w])
    assert len({c.shape[1] for c in crops}) > 20
    label = f"{models_dir}/rec/ppocr_keys_v1.txt"
    rec = b200ocr.Recognizer(f"{models_dir}/rec", label, rec_batch_num=16, rec_img_h=28, rec_img_w=192)
    texts, scores = rec.run(crops)
    assert sum(1 for t in texts if t) >= 0.9 * len(crops)
    # a reference batch (16 consecutive crops of the aspect-sorted order) decoded on its own gives the same strings:
    # the ragged packing across batches changes nothing
    order = sorted(range(len(crops)), key=lambda i: np.float32(crops[i].shape[1]) / np.float32(crops[i].shape[0]))
    for beg in (0, 48, (len(order) // 16 - 1) * 16):
        idx = order[beg:beg + 16]
        t2, s2 = rec.run([crops[i] for i in idx])
        for k, i in enumerate(idx):
            assert t2[k] == texts[i] and s2[k] == scores[i], (beg, k)
    # oracle parity on the page, stage by stage (see test_c3: network 1e-2, post-process on the oracle's map exact in
    # count / order and within 1 px of the 960 map = 3 source pixels, >= 90 % of the boxes equal end to end) ...
    from oracle.pipeline import OracleWorker
    from test_stages_gpu import _check_line
    _det_parity_at(models_dir, det, [page], [boxes], 960, 0.2, 0.4, 1.8, px_tol=3, min_boxes=150)
    # ... and the whole page through a worker built for pages (b200ocr_worker_create_ex, limit_side_len 960) against
    # the oracle worker with the same setting, fed the GPU's boxes: every one of the 200+ strings identical
    wp = b200ocr.Worker(3, models_dir, enable_cls=True, limit_side_len=960)
    dp = json.loads(wp.process(11, page))
    assert dp["success"] and [wd["box"] for wd in dp["words"]] == boxes.tolist()
    ow = OracleWorker(3, models_dir, enable_cls=True, limit_side_len=960)
    words, raw = ow.process_words(page, det_boxes=[wd["box"] for wd in dp["words"]], want_raw=True)
    assert len(words) == len(dp["words"]) >= 150
    for wd, (text, score, _b), r3 in zip(dp["words"], words, raw):
        _check_line(wd["text"], wd["confidence"], text, score, r3)
    # the whole page through the reference-default worker (limit 512): envelope + determinism inside a batch
    w = b200ocr.Worker(0, models_dir, enable_cls=True)
    d = json.loads(w.process(1, page))
    assert d["success"] and d["width"] == 2048 and d["height"] == 2048
    d2 = json.loads(w.process_batch([7, 8], [synth_data.card(4), page])[1])
    assert d2["words"] == d["words"]
