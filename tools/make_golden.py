"""Golden vectors for the stage parity tests: the CPU oracle (oracle/pipeline.py, torch-CPU fp32 + cv2 + the reference's
Clipper) on the reference's own fixtures -- images/card-jd.jpg and tests/test_ocr_worker.cpp's createTestImage recipe --
with the worker's hyper-parameters (src/ocr_worker.cpp:21-63).  Writes tests/golden/expected_stages.json:
    per image: the detector's boxes; per box the bounding-rectangle ROI the worker crops (src/ocr_worker.cpp:244-259),
    the classifier's label / score on that crop, and the recognizer's text / confidence / smallest top-2 margin on it
    (rec_img_h 28, rec_img_w 192, batches of 16: the worker's values).
Every stage is recorded GIVEN THE ORACLE'S OWN UPSTREAM DATA, so that a test can feed the same crops to one stage.
    python tools/make_golden.py          (CPU only; about 10 s)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def golden_images(golden_dir):
    import cv2
    import synth_data
    return [("card-jd.jpg", cv2.imread(os.path.join(golden_dir, "card-jd.jpg"))),
            ("reference_test_image", synth_data.reference_test_image())]


def oracle_stages(models_dir, golden_dir):
    """the structure that is stored, computed by the oracle (also what the CPU test re-computes)"""
    from oracle import ocr_ops
    from oracle.pipeline import OracleClassifier, OracleDetector, OracleRecognizer
    det = OracleDetector(os.path.join(models_dir, "det"), "max", 512, 0.2, 0.4, 1.8, "fast", False)
    cls = OracleClassifier(os.path.join(models_dir, "cls"), 8)
    rec = OracleRecognizer(os.path.join(models_dir, "rec"), os.path.join(models_dir, "rec", "ppocr_keys_v1.txt"), 16, 28, 192)
    out = []
    for name, img in golden_images(golden_dir):
        boxes = det.run(img)
        rois = [ocr_ops.bounding_rect_crop(b, img.shape[0], img.shape[1]) for b in boxes]
        crops = [img[y:y + h, x:x + w] for x, y, w, h in rois]
        labels, cscores = cls.run(crops)
        texts, scores, raw = rec.run(crops, want_raw=True)
        out.append({"name": name, "rows": int(img.shape[0]), "cols": int(img.shape[1]),
                    "boxes": [[[int(v) for v in p] for p in b] for b in boxes],
                    "rois": [[int(v) for v in r] for r in rois],
                    "cls_labels": [int(v) for v in labels], "cls_scores": [float(v) for v in cscores],
                    "rec_texts": list(texts), "rec_scores": [float(v) for v in scores],
                    "rec_min_margin": [float((r[1] - r[2]).min()) if len(r[1]) else 1.0 for r in raw]})
    return out


def main():
    import make_synth_weights
    models = make_synth_weights.ensure_models()
    golden_dir = os.path.join(ROOT, "tests", "golden")
    doc = {"generator": "tools/make_golden.py", "oracle": "oracle/pipeline.py (torch-CPU fp32, cv2, reference Clipper)",
           "params": "det: max/512, thresh 0.2, box_thresh 0.4, unclip 1.8, fast; cls batch 8; rec batch 16, 28x192",
           "images": oracle_stages(models, golden_dir)}
    path = os.path.join(golden_dir, "expected_stages.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(doc, f, ensure_ascii=False, indent=1)
    print(path, sum(len(i["boxes"]) for i in doc["images"]), "boxes")


if __name__ == "__main__":
    main()
