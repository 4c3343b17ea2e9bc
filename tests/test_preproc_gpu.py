"""GPU parity of the pre-processing kernels (preproc.cu) against the reference's own OpenCV calls (oracle/ocr_ops.py)."""
import os

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_resize_is_bit_exact_vs_cv2(golden_dir):
    import b200ocr
    rng = np.random.default_rng(0)
    card = cv2.imread(os.path.join(golden_dir, "card-jd.jpg"))
    cases = [(card, 384, 192)]
    for (sw, sh, dw, dh) in [(1024, 640, 512, 320), (1024, 640, 960, 608), (2048, 2048, 512, 512), (300, 40, 192, 28),
                             (57, 19, 84, 28), (64, 64, 64, 64), (9, 7, 33, 48)]:
        cases.append((rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8), dw, dh))
    for _ in range(12):
        sw, sh, dw, dh = (int(rng.integers(8, 400)), int(rng.integers(8, 200)), int(rng.integers(8, 400)), int(rng.integers(8, 120)))
        cases.append((rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8), dw, dh))
    for src, dw, dh in cases:
        out = b200ocr.resize_u8(src, dh, dw)
        assert np.array_equal(out, cv2.resize(src, (dw, dh))), (src.shape, dw, dh)
    # a non-contiguous view (cv::Mat ROI: step > cols*3) is read through its stride
    big = rng.integers(0, 256, (120, 300, 3), dtype=np.uint8)
    roi = big[10:90, 20:250]
    assert np.array_equal(b200ocr.resize_u8(roi, 28, 81), cv2.resize(np.ascontiguousarray(roi), (81, 28)))


def test_det_preprocess_matches_oracle(models_dir, golden_dir):
    import b200ocr, synth_data
    from oracle import ocr_ops
    det = b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=512)
    for img in (cv2.imread(os.path.join(golden_dir, "card-jd.jpg")), synth_data.card(3), synth_data.reference_test_image()):
        got, rh, rw = det.preprocess(img)
        ref, orh, orw = ocr_ops.det_preprocess(img, "max", 512)
        assert got.shape == ref[0].shape and np.float32(rh) == orh and np.float32(rw) == orw
        # identical integer resize + fp32 normalisation, stored as fp16: error <= half an fp16 ulp of |x| <= 2.7
        assert np.abs(got - ref[0]).max() <= 2 ** -10 * 2
    det960 = b200ocr.Detector(f"{models_dir}/det", limit_type="max", limit_side_len=960)
    got, _, _ = det960.preprocess(synth_data.card(1))
    assert got.shape == (3, 608, 960)


@pytest.mark.parametrize("kind,h,w", [("rec", 48, 320), ("rec", 28, 192), ("cls", 48, 192)])
def test_crop_preprocess_matches_oracle(kind, h, w):
    import b200ocr
    from oracle import ocr_ops
    rng = np.random.default_rng(1)
    crops = [rng.integers(0, 256, (int(rng.integers(6, 80)), int(rng.integers(6, 400)), 3), dtype=np.uint8) for _ in range(9)]
    crops.append(rng.integers(0, 256, (2 * h, 2 * 50, 3), dtype=np.uint8))  # exact 2x shrink -> INTER_AREA path
    got = b200ocr.crop_preprocess(crops, kind, h, w)
    for i, c in enumerate(crops):
        if kind == "rec":
            r = ocr_ops.normalize(ocr_ops.crnn_resize_img(c, np.float32(w) / np.float32(h), (3, h, w)), ocr_ops.REC_MEAN, ocr_ops.REC_SCALE)
        else:
            r = ocr_ops.normalize(ocr_ops.cls_resize_img(c, (3, h, w)), ocr_ops.REC_MEAN, ocr_ops.REC_SCALE)
            if r.shape[1] < w:
                r = cv2.copyMakeBorder(r, 0, 0, 0, w - r.shape[1], cv2.BORDER_CONSTANT, value=(0, 0, 0))
        assert np.abs(got[i] - ocr_ops.permute(r)).max() <= 2 ** -10


def test_rotate_crop_matches_get_rotate_crop_image(golden_dir):
    """Utility::GetRotateCropImage (reference src/utility.cpp:137-190) restated with the same cv2 calls in the oracle.
    cv::warpPerspective on 8-bit data is fixed point (1/32-px coordinates from a double-precision map evaluated block by
    block, 15-bit weights); the kernel performs the same operations in the same order: byte work, bit-exact."""
    import b200ocr, synth_data
    from oracle import ocr_ops
    rng = np.random.default_rng(4)
    img = synth_data.card(9)
    boxes = [[[100, 60], [420, 52], [424, 98], [104, 106]],        # slightly rotated line
             [[300, 200], [700, 240], [692, 300], [292, 260]],
             [[50, 300], [90, 300], [90, 520], [50, 520]],          # tall box -> transpose + flip
             [[10, 10], [200, 10], [200, 40], [10, 40]]]            # axis aligned
    for _ in range(6):
        c = rng.uniform([150, 100], [850, 520]); w, h, ang = rng.uniform(60, 280), rng.uniform(16, 60), rng.uniform(-25, 25)
        pts = cv2.boxPoints(((float(c[0]), float(c[1])), (float(w), float(h)), float(ang)))
        o = ocr_ops.order_points_clockwise(np.round(pts).astype(int).tolist())
        boxes.append(o)
    for box in boxes:
        ref = ocr_ops.get_rotate_crop_image(img, box)
        got = b200ocr.rotate_crop(img, box)
        assert got.shape == ref.shape, (got.shape, ref.shape, box)
        assert np.array_equal(got, ref), (int((got != ref).any(-1).sum()), int(np.abs(got.astype(int) - ref).max()), box)
