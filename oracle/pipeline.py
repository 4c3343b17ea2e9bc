"""ORACLE (test infrastructure, never shipped): the reference's det -> cls -> rec worker restated on the CPU.

  OracleWorker.process  follows  OCRWorker::processRequest + the result JSON  (src/ocr_worker.cpp:133-311)
  stage methods         follow   DBDetector::Run (src/ocr_det.cpp:93-176), Classifier::Run (src/ocr_cls.cpp:23-106),
                                 CRNNRecognizer::Run (src/ocr_rec.cpp:24-135)
The three graphs run op by op in torch-CPU fp32 (oracle/interp.py); image / geometry steps are the reference's own
OpenCV calls (oracle/ocr_ops.py) and its vendored Clipper (oracle/unclip.py).  PARITY UNPINNED: see ocr_ops.py.
This is also the timed CPU baseline of bench.py (`cpu_baseline` / `--impl reference`, kind "port").
"""
from __future__ import annotations
import json
import os
import time

import cv2
import numpy as np

from . import ocr_ops as ops
from .interp import run_program
from .pdmodel import load_params, load_program


class _Net:
    def __init__(self, model_dir):
        self.prog = load_program(os.path.join(model_dir, "inference.pdmodel"))
        self.params = load_params(self.prog, os.path.join(model_dir, "inference.pdiparams"))

    def __call__(self, x):
        return run_program(self.prog, self.params, x)[0]


class OracleDetector:
    def __init__(self, model_dir, limit_type="max", limit_side_len=960, det_db_thresh=0.3, det_db_box_thresh=0.5,
                 det_db_unclip_ratio=2.0, det_db_score_mode="fast", use_dilation=False):
        self.net = _Net(model_dir)
        self.p = dict(limit_type=limit_type, limit_side_len=limit_side_len, det_db_thresh=det_db_thresh,
                      box_thresh=det_db_box_thresh, unclip_ratio=det_db_unclip_ratio, score_mode=det_db_score_mode,
                      use_dilation=use_dilation)

    def forward(self, img):
        x, rh, rw = ops.det_preprocess(img, self.p["limit_type"], self.p["limit_side_len"])
        return self.net(x)[0, 0], rh, rw

    def post(self, pred, ratio_h, ratio_w, src_h, src_w, trace=None):
        return ops.det_postprocess(pred, ratio_h, ratio_w, src_h, src_w, self.p["det_db_thresh"], self.p["box_thresh"],
                                   self.p["unclip_ratio"], self.p["score_mode"], self.p["use_dilation"], trace=trace)

    def run(self, img):
        pred, rh, rw = self.forward(img)
        return self.post(pred, rh, rw, img.shape[0], img.shape[1])[0]


class OracleClassifier:
    def __init__(self, model_dir, cls_batch_num=1):
        self.net = _Net(model_dir)
        self.batch = cls_batch_num

    def run(self, imgs):
        labels = [0] * len(imgs)
        scores = [0.0] * len(imgs)
        for beg, x in ops.cls_batch_inputs(imgs, self.batch):
            out = self.net(x)
            for k in range(out.shape[0]):
                labels[beg + k] = int(out[k].argmax())
                scores[beg + k] = float(out[k].max())
        return labels, scores


class OracleRecognizer:
    def __init__(self, model_dir, label_path, rec_batch_num=6, rec_img_h=48, rec_img_w=320):
        self.net = _Net(model_dir)
        self.labels = ops.read_dict(label_path)
        self.cfg = (rec_batch_num, rec_img_h, rec_img_w)

    def run(self, imgs, want_raw=False):
        texts = [""] * len(imgs)
        scores = [0.0] * len(imgs)
        raw = [None] * len(imgs)
        for idx, x in ops.rec_batches(imgs, *self.cfg):
            probs = self.net(x)  # [b, T, 6625]
            for m, i in enumerate(idx):
                raw[i] = (probs[m].argmax(-1), probs[m].max(-1), np.sort(probs[m], -1)[:, -2])
                r = ops.ctc_greedy_decode(probs[m], self.labels)
                if r is not None:
                    texts[i], scores[i] = r[0], float(r[1])
        return (texts, scores, raw) if want_raw else (texts, scores)


def json_double(v: float) -> str:
    """jsoncpp valueToString(double): %.17g, '.0' appended when no '.'/'e' is present."""
    s = "%.17g" % v
    return s if ("." in s or "e" in s) else s + ".0"


def result_json(request_id, worker_id, success, width, height, ms, words, error=""):
    """The reference's result line (src/ocr_worker.cpp:155-190): compact, keys in jsoncpp's (alphabetical) order."""
    q = lambda s: json.dumps(s, ensure_ascii=False)
    o = "{"
    if not success:
        o += '"error":' + q(error) + ","
    o += '"height":%d,"processing_time_ms":%s,"request_id":%d,"success":%s,"width":%d,' % (
        height, json_double(ms), request_id, "true" if success else "false", width)
    if success:
        ws = []
        for text, conf, box in words:
            pts = ",".join("[%d,%d]" % (int(p[0]), int(p[1])) for p in box)
            ws.append('{"box":[' + pts + '],"confidence":' + json_double(float(np.float32(conf))) + ',"text":' + q(text) + "}")
        o += '"words":[' + ",".join(ws) + "],"
    return o + '"worker_id":%d}' % worker_id


class OracleWorker:
    """Hyper-parameters of the reference OCRWorker constructor (src/ocr_worker.cpp:21-63)."""

    def __init__(self, worker_id, model_dir, enable_cls=False, limit_side_len=512):
        """`limit_side_len` other than the reference's 512 = b200ocr_worker_create_ex (dense pages, BASELINE config 5)."""
        self.worker_id = worker_id
        self.det = OracleDetector(os.path.join(model_dir, "det"), "max", limit_side_len, 0.2, 0.4, 1.8, "fast", False)
        self.cls = OracleClassifier(os.path.join(model_dir, "cls"), 8) if enable_cls else None
        self.rec = OracleRecognizer(os.path.join(model_dir, "rec"), os.path.join(model_dir, "rec", "ppocr_keys_v1.txt"),
                                    16, 28, 192)

    def process_words(self, img, det_boxes=None, want_raw=False):
        """processRequest (src/ocr_worker.cpp:213-311) -> [(text, score, box)].  `det_boxes` overrides the detector
        (parity tests feed the GPU's boxes so that the later stages see identical upstream data)."""
        image = img.copy()  # OCRRequest deep-copies (ocr_worker.h:28-29); rotations below mutate that copy
        boxes = self.det.run(image) if det_boxes is None else [list(map(list, b)) for b in det_boxes]
        if not boxes:
            return ([], []) if want_raw else []
        crops = []
        for b in boxes:
            r = ops.bounding_rect_crop(b, image.shape[0], image.shape[1])
            if r is not None:
                x, y, w, h = r
                crops.append(image[y:y + h, x:x + w])  # ROI view, not a copy (ocr_worker.cpp:257)
        if not crops:
            return ([], []) if want_raw else []
        if self.cls is not None:
            labels, _ = self.cls.run(crops)
            for i, lab in enumerate(labels):
                if lab == 1:
                    crops[i][:] = cv2.rotate(crops[i], cv2.ROTATE_180)  # in place on the shared image
        texts, scores, raw = self.rec.run(crops, want_raw=True)
        words = [(texts[i], scores[i], boxes[i]) for i in range(len(texts))]
        return (words, raw) if want_raw else words

    def process(self, request_id, img):
        t0 = time.perf_counter()
        if img is None or img.size == 0:
            return result_json(request_id, self.worker_id, False, 0, 0, 0.0, [], "Empty image data provided")
        words = self.process_words(img)
        ms = (time.perf_counter() - t0) * 1e3
        return result_json(request_id, self.worker_id, True, img.shape[1], img.shape[0], ms, words)
