"""The header-only C++ replacement for the reference's stage classes (include/paddle_ocr/b200ocr_shim.h:
PaddleOCR::DBDetector / Classifier / CRNNRecognizer with the constructor and Run() signatures of the reference's
include/paddle_ocr/ocr_det.h:60-97, ocr_cls.h:57-82, ocr_rec.h:61-95) compiles with a plain g++ against a stand-in
cv::Mat, links with the product library alone, fails loudly without a GPU, and -- on a GPU -- returns exactly what
the ctypes view of the same C ABI returns.
"""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")


@pytest.fixture(scope="module")
def driver():
    r = subprocess.run(["make", "-C", NATIVE, "shim_driver"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "warning" not in (r.stdout + r.stderr).lower(), r.stdout + r.stderr
    return os.path.join(NATIVE, "shim_driver")


def _raw(tmp_path, img):
    p = os.path.join(tmp_path, "img.bgr")
    np.ascontiguousarray(img).tofile(p)
    return p


def test_shim_fails_loudly_without_gpu(driver, models_dir, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    img = np.zeros((32, 48, 3), np.uint8)
    r = subprocess.run([driver, models_dir, _raw(str(tmp_path), img), "32", "48"], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 3
    out = json.loads(r.stdout)
    assert "no CPU fallback" in out["error"]


@pytest.mark.gpu
def test_shim_equals_ctypes_view(driver, models_dir, golden_dir, tmp_path):
    import cv2
    import b200ocr
    import synth_data
    for img in (cv2.imread(os.path.join(golden_dir, "card-jd.jpg")), synth_data.card(1)):
        h, w = img.shape[:2]
        r = subprocess.run([driver, models_dir, _raw(str(tmp_path), img), str(h), str(w)], capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        got = json.loads(r.stdout)
        assert got["times"] == 9  # three per stage, appended (src/ocr_det.cpp:168-175 and siblings)

        det = b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=512, det_db_thresh=0.2,
                               det_db_box_thresh=0.4, det_db_unclip_ratio=1.8, det_db_score_mode="fast")
        boxes = det.run(img)
        assert len(boxes) > 0
        assert np.array_equal(np.asarray(got["boxes"], np.int32).reshape(-1, 4, 2), boxes)
        crops = []
        for b in boxes:
            x0, y0 = max(int(b[:, 0].min()), 0), max(int(b[:, 1].min()), 0)
            x1, y1 = min(int(b[:, 0].max()) + 1, w), min(int(b[:, 1].max()) + 1, h)
            crops.append(img[y0:y1, x0:x1])
        cls = b200ocr.Classifier(f"{models_dir}/cls", cls_thresh=0.98, cls_batch_num=8)
        labels, _ = cls.run(crops)
        assert got["labels"] == [int(v) for v in labels]
        rec = b200ocr.Recognizer(f"{models_dir}/rec", f"{models_dir}/rec/ppocr_keys_v1.txt", rec_batch_num=16,
                                 rec_img_h=28, rec_img_w=192)
        texts, scores = rec.run(crops)
        assert got["texts"] == list(texts)
        assert np.allclose(got["scores"], scores, rtol=0, atol=1e-5)  # printed with 6 significant digits
