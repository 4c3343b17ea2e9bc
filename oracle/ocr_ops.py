"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's pre/post-processing
and stage orchestration, calling the SAME OpenCV functions the reference calls (through cv2 4.13,
the Python build of the library the reference links statically) and the reference's own vendored
Clipper (compiled from /root/reference/src/clipper.cpp into oracle/_ref, see oracle/Makefile; when
that binary is absent the closed-form restatement in oracle/unclip.py is used).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Every function cites the reference file:line it follows (paths relative to the
reference repository).

PARITY UNPINNED: the reference's tests hold no golden vector for any of these boundaries
(SURVEY.md §4 / §8c).  What pins this file is that each image/geometry step IS the third-party
call the reference makes (cv2.resize, cv2.findContours, cv2.minAreaRect, cv2.boxPoints,
cv2.fillPoly, cv2.mean, cv2.boundingRect, cv2.rotate, cv2.warpPerspective ...).
"""
from __future__ import annotations
import math

import cv2
import numpy as np

F32 = np.float32

DET_MEAN = (0.485, 0.456, 0.406)          # ocr_det.h:121
DET_SCALE = tuple(float(F32(1) / F32(v)) for v in (0.229, 0.224, 0.225))  # ocr_det.h:122 (1 / 0.229f in float)
REC_MEAN = (0.5, 0.5, 0.5)                # ocr_rec.h:108, ocr_cls.h:93
REC_SCALE = (1 / 0.5, 1 / 0.5, 1 / 0.5)   # ocr_rec.h:109, ocr_cls.h:94


def _c_round(x: float) -> float:
    """C round(): half away from zero."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# ----------------------------------------------------------------------------- pre-processing
def resize_img_type0(img, limit_type="max", limit_side_len=960):
    """ResizeImgType0::Run, preprocess_op.cpp:57-93.  Returns (resized u8 HWC, ratio_h, ratio_w)."""
    h, w = img.shape[:2]
    ratio = F32(1.0)
    if limit_type == "min":
        if min(h, w) < limit_side_len:
            ratio = F32(limit_side_len) / F32(h) if h < w else F32(limit_side_len) / F32(w)
    else:
        if max(h, w) > limit_side_len:
            ratio = F32(limit_side_len) / F32(h) if h > w else F32(limit_side_len) / F32(w)
    resize_h = int(F32(h) * ratio)
    resize_w = int(F32(w) * ratio)
    resize_h = max(int(_c_round(float(F32(resize_h) / F32(32))) * 32), 32)
    resize_w = max(int(_c_round(float(F32(resize_w) / F32(32))) * 32), 32)
    out = cv2.resize(img, (resize_w, resize_h))
    return out, F32(resize_h) / F32(h), F32(resize_w) / F32(w)


def normalize(img_u8, mean, scale, is_scale=True):
    """Normalize::Run, preprocess_op.cpp:40-55: convertTo(CV_32FC3, 1/255) then per channel
    convertTo(CV_32FC1, scale[i], (0 - mean[i]) * scale[i]).  fp32, two roundings."""
    e = 1.0 / 255.0 if is_scale else 1.0
    f = img_u8.astype(F32) * F32(e)
    out = np.empty_like(f)
    for i in range(3):
        sc = F32(scale[i])                                                  # (float)(1.0 * scale[i])
        sh = F32((0.0 - float(F32(mean[i]))) * float(F32(scale[i])))        # (float)((0.0 - mean[i]) * scale[i])
        out[..., i] = f[..., i] * sc + sh
    return out


def permute(img_f32):
    """Permute::Run, preprocess_op.cpp:19-26: HWC -> CHW."""
    return np.ascontiguousarray(img_f32.transpose(2, 0, 1))


def crnn_resize_img(img, wh_ratio, rec_image_shape):
    """CrnnResizeImg::Run, preprocess_op.cpp:95-118 (zero pad on the right in u8, before normalisation)."""
    _c, img_h, _w = rec_image_shape
    img_w = int(F32(img_h) * F32(wh_ratio))
    ratio = F32(img.shape[1]) / F32(img.shape[0])
    if math.ceil(float(F32(img_h) * ratio)) > img_w:
        resize_w = img_w
    else:
        resize_w = int(math.ceil(float(F32(img_h) * ratio)))
    r = cv2.resize(img, (resize_w, img_h), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(r, 0, 0, 0, int(img_w - r.shape[1]), cv2.BORDER_CONSTANT, value=(0, 0, 0))


def cls_resize_img(img, cls_image_shape=(3, 48, 192)):
    """ClsResizeImg::Run, preprocess_op.cpp:120-137."""
    _c, img_h, img_w = cls_image_shape
    ratio = F32(img.shape[1]) / F32(img.shape[0])
    if math.ceil(float(F32(img_h) * ratio)) > img_w:
        resize_w = img_w
    else:
        resize_w = int(math.ceil(float(F32(img_h) * ratio)))
    return cv2.resize(img, (resize_w, img_h), interpolation=cv2.INTER_LINEAR)


def det_preprocess(img, limit_type="max", limit_side_len=512):
    """DBDetector::Run pre-processing, ocr_det.cpp:103-113.  Returns (NCHW fp32 [1,3,h,w], ratio_h, ratio_w)."""
    r, ratio_h, ratio_w = resize_img_type0(img, limit_type, limit_side_len)
    x = permute(normalize(r, DET_MEAN, DET_SCALE, True))
    return x[None], ratio_h, ratio_w


# ----------------------------------------------------------------------------- DB post-processing
def threshold_bitmap(pred, det_db_thresh):
    """ocr_det.cpp:143-154: cbuf = (uchar)(p * 255); bit = cbuf > thresh*255 (cv::threshold on 8U
    floors the double threshold)."""
    cbuf = (pred.astype(F32) * F32(255)).astype(np.uint8)  # C float->uchar conversion truncates
    _t, bit = cv2.threshold(cbuf, float(F32(det_db_thresh)) * 255, 255, cv2.THRESH_BINARY)
    return bit


def dilate2x2(bitmap):
    """ocr_det.cpp:155-159."""
    return cv2.dilate(bitmap, cv2.getStructuringElement(cv2.MORPH_RECT, (2, 2)))


def get_mini_boxes(rrect):
    """GetMiniBoxes, postprocess_op.cpp:134-168.  rrect = ((cx,cy),(w,h),angle).  Returns ([4][2] f32, ssid)."""
    ssid = max(F32(rrect[1][0]), F32(rrect[1][1]))
    pts = cv2.boxPoints(rrect).astype(F32)
    arr = [list(p) for p in pts]
    # std::sort with a comparator that is false on ties; n = 4 -> insertion sort, stable
    arr = sorted(arr, key=lambda p: p[0])
    if arr[3][1] <= arr[2][1]:
        idx2, idx3 = arr[3], arr[2]
    else:
        idx2, idx3 = arr[2], arr[3]
    if arr[1][1] <= arr[0][1]:
        idx1, idx4 = arr[1], arr[0]
    else:
        idx1, idx4 = arr[0], arr[1]
    return np.array([idx1, idx2, idx3, idx4], F32), ssid


def box_score_fast(box, pred):
    """BoxScoreFast, postprocess_op.cpp:216-253."""
    h, w = pred.shape
    xs, ys = box[:, 0], box[:, 1]
    clamp = lambda v, lo, hi: max(lo, min(v, hi))
    xmin = clamp(int(math.floor(xs.min())), 0, w - 1)
    xmax = clamp(int(math.ceil(xs.max())), 0, w - 1)
    ymin = clamp(int(math.floor(ys.min())), 0, h - 1)
    ymax = clamp(int(math.ceil(ys.max())), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
    pts = np.array([[int(box[i, 0]) - xmin, int(box[i, 1]) - ymin] for i in range(4)], np.int32)
    cv2.fillPoly(mask, [pts], 1)
    return F32(cv2.mean(np.ascontiguousarray(pred[ymin:ymax + 1, xmin:xmax + 1]), mask)[0])


def polygon_score_acc(contour, pred):
    """PolygonScoreAcc, postprocess_op.cpp:170-214 ("slow" score mode)."""
    h, w = pred.shape
    pts = contour.reshape(-1, 2)
    xs, ys = pts[:, 0].astype(F32), pts[:, 1].astype(F32)
    clamp = lambda v, lo, hi: max(lo, min(v, hi))
    xmin = clamp(int(math.floor(xs.min())), 0, w - 1)
    xmax = clamp(int(math.ceil(xs.max())), 0, w - 1)
    ymin = clamp(int(math.floor(ys.min())), 0, h - 1)
    ymax = clamp(int(math.ceil(ys.max())), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
    p = np.stack([xs.astype(np.int32) - xmin, ys.astype(np.int32) - ymin], 1)
    cv2.fillPoly(mask, [p], 1)
    return F32(cv2.mean(np.ascontiguousarray(pred[ymin:ymax + 1, xmin:xmax + 1]), mask)[0])


def get_contour_area(box, unclip_ratio):
    """GetContourArea, postprocess_op.cpp:20-37 (float32 accumulation, as written)."""
    area = F32(0)
    dist = F32(0)
    for i in range(4):
        j = (i + 1) % 4
        area = F32(area + F32(F32(box[i][0] * box[j][1]) - F32(box[i][1] * box[j][0])))
        dx = F32(box[i][0] - box[j][0])
        dy = F32(box[i][1] - box[j][1])
        dist = F32(dist + F32(np.sqrt(F32(F32(dx * dx) + F32(dy * dy)))))
    area = F32(abs(F32(float(area) / 2.0)))
    return F32(F32(area * F32(unclip_ratio)) / dist)


def unclip(box, unclip_ratio, clipper=None):
    """UnClip, postprocess_op.cpp:39-72.  Returns a cv2 RotatedRect tuple."""
    from . import unclip as _u
    distance = get_contour_area(box, unclip_ratio)
    path = [(int(box[i][0]), int(box[i][1])) for i in range(4)]
    pts = _u.offset_points(path, float(distance), clipper)
    if len(pts) == 0:
        return ((0.0, 0.0), (1.0, 1.0), 0.0)
    return cv2.minAreaRect(np.array(pts, F32).reshape(-1, 1, 2))


def boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, score_mode="fast", clipper=None, trace=None):
    """BoxesFromBitmap, postprocess_op.cpp:255-331."""
    min_size, max_candidates = 3, 1000
    height, width = bitmap.shape
    contours, _h = cv2.findContours(bitmap, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    boxes = []
    for ci in range(min(len(contours), max_candidates)):
        c = contours[ci]
        rec = {"start": tuple(int(v) for v in c[0, 0]), "npts": len(c), "stage": "size"}
        if trace is not None:
            trace.append(rec)
        if len(c) <= 2:
            continue
        rr = cv2.minAreaRect(c)
        array, ssid = get_mini_boxes(rr)
        rec.update(stage="ssid", rect=rr, mini=array.copy(), ssid=float(ssid))
        if ssid < min_size:
            continue
        score = polygon_score_acc(c, pred) if score_mode == "slow" else box_score_fast(array, pred)
        rec.update(stage="score", score=float(score))
        if score < F32(box_thresh):
            continue
        points = unclip(array, unclip_ratio, clipper)
        rec.update(stage="unclip", unclip=points)
        if points[1][1] < 1.001 and points[1][0] < 1.001:
            continue
        cliparray, ssid = get_mini_boxes(points)
        rec.update(stage="ssid2", clip=cliparray.copy(), ssid2=float(ssid))
        if ssid < min_size + 2:
            continue
        dest_width, dest_height = pred.shape[1], pred.shape[0]
        ibox = []
        for k in range(4):
            x = _roundf(F32(F32(cliparray[k][0] / F32(width)) * F32(dest_width)))
            y = _roundf(F32(F32(cliparray[k][1] / F32(height)) * F32(dest_height)))
            ibox.append([int(min(max(x, F32(0)), F32(dest_width))), int(min(max(y, F32(0)), F32(dest_height)))])
        rec.update(stage="ok", box=ibox)
        boxes.append(ibox)
    return boxes


def _roundf(x):
    """C roundf(): half away from zero, float32."""
    x = float(x)
    return F32(math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5))


def order_points_clockwise(pts):
    """OrderPointsClockwise, postprocess_op.cpp:87-104 (stable x sort of 4 int points)."""
    box = sorted([list(p) for p in pts], key=lambda p: p[0])
    left, right = [box[0], box[1]], [box[2], box[3]]
    if left[0][1] > left[1][1]:
        left = [left[1], left[0]]
    if right[0][1] > right[1][1]:
        right = [right[1], right[0]]
    return [left[0], right[0], right[1], left[1]]


def filter_tag_det_res(boxes, ratio_h, ratio_w, src_h, src_w):
    """FilterTagDetRes, postprocess_op.cpp:333-362 (int /= float: convert, divide in fp32, truncate)."""
    out = []
    for b in boxes:
        b = order_points_clockwise(b)
        for m in range(4):
            x = int(F32(b[m][0]) / F32(ratio_w))
            y = int(F32(b[m][1]) / F32(ratio_h))
            b[m][0] = int(min(max(x, 0), src_w - 1))
            b[m][1] = int(min(max(y, 0), src_h - 1))
        rect_w = int(math.sqrt((b[0][0] - b[1][0]) ** 2 + (b[0][1] - b[1][1]) ** 2))
        rect_h = int(math.sqrt((b[0][0] - b[3][0]) ** 2 + (b[0][1] - b[3][1]) ** 2))
        if rect_w <= 4 or rect_h <= 4:
            continue
        out.append(b)
    return out


def det_postprocess(pred, ratio_h, ratio_w, src_h, src_w, det_db_thresh=0.3, box_thresh=0.5, unclip_ratio=2.0,
                    score_mode="fast", use_dilation=False, clipper=None, trace=None):
    """DBDetector::Run post-processing, ocr_det.cpp:136-165."""
    pred = np.ascontiguousarray(pred, F32)
    bitmap = threshold_bitmap(pred, det_db_thresh)
    if use_dilation:
        bitmap = dilate2x2(bitmap)
    boxes = boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, score_mode, clipper, trace)
    return filter_tag_det_res(boxes, ratio_h, ratio_w, src_h, src_w), bitmap


# ----------------------------------------------------------------------------- crops
def bounding_rect_crop(box, img_h, img_w):
    """OCRWorker::processRequest crop, ocr_worker.cpp:244-259: cv::boundingRect of the 4 points
    (as Point2f) intersected with the image.  Returns (x, y, w, h) or None."""
    pts = np.array(box, F32).reshape(-1, 1, 2)
    x, y, w, h = cv2.boundingRect(pts)
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + w, img_w), min(y + h, img_h)
    if x1 - x0 <= 0 or y1 - y0 <= 0:
        return None
    return (x0, y0, x1 - x0, y1 - y0)


def get_rotate_crop_image(img, box):
    """Utility::GetRotateCropImage, utility.cpp:137-190.  The reference passes cv::BORDER_REPLICATE
    (== 1 == INTER_LINEAR) in the *flags* slot of warpPerspective; the border mode stays BORDER_CONSTANT."""
    pts = [list(map(int, p)) for p in box]
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    left, right, top, bottom = min(xs), max(xs), min(ys), max(ys)
    crop = img[top:bottom, left:right].copy()
    pts = [[p[0] - left, p[1] - top] for p in pts]
    cw = int(math.sqrt((pts[0][0] - pts[1][0]) ** 2 + (pts[0][1] - pts[1][1]) ** 2))
    ch = int(math.sqrt((pts[0][0] - pts[3][0]) ** 2 + (pts[0][1] - pts[3][1]) ** 2))
    std = np.array([[0, 0], [cw, 0], [cw, ch], [0, ch]], F32)
    m = cv2.getPerspectiveTransform(np.array(pts, F32), std)
    dst = cv2.warpPerspective(crop, m, (cw, ch), flags=cv2.BORDER_REPLICATE)
    if float(dst.shape[0]) >= float(dst.shape[1]) * 1.5:
        dst = cv2.flip(cv2.transpose(dst), 0)
    return dst


# ----------------------------------------------------------------------------- cls / rec stages
def cls_batch_inputs(imgs, cls_batch_num=8):
    """Classifier::Run batching + pre-processing, ocr_cls.cpp:34-62.  Yields (start, NCHW fp32 batch)."""
    for beg in range(0, len(imgs), cls_batch_num):
        batch = []
        for im in imgs[beg:beg + cls_batch_num]:
            r = normalize(cls_resize_img(im), REC_MEAN, REC_SCALE, True)
            if r.shape[1] < 192:
                r = cv2.copyMakeBorder(r, 0, 0, 0, 192 - r.shape[1], cv2.BORDER_CONSTANT, value=(0, 0, 0))
            batch.append(permute(r))
        yield beg, np.stack(batch)


def argsort_ratio(ratios):
    """Utility::argsort, utility.cpp:192-203.  std::sort is not stable; for n <= 16 libstdc++ and MSVC
    both run insertion sort, which is.  Ties beyond that are implementation-defined in the reference."""
    return sorted(range(len(ratios)), key=lambda i: ratios[i])


def rec_batches(imgs, rec_batch_num=16, rec_img_h=48, rec_img_w=320):
    """CRNNRecognizer::Run batching + pre-processing, ocr_rec.cpp:35-73.
    Yields (indices of the batch in caller order, NCHW fp32 batch)."""
    ratios = [F32(im.shape[1]) / F32(im.shape[0]) for im in imgs]
    indices = argsort_ratio(ratios)
    shape = (3, rec_img_h, rec_img_w)
    for beg in range(0, len(imgs), rec_batch_num):
        idx = indices[beg:beg + rec_batch_num]
        max_wh = F32(rec_img_w * 1.0 / rec_img_h)
        for i in idx:
            max_wh = max(max_wh, F32(imgs[i].shape[1] * 1.0 / imgs[i].shape[0]))
        batch = [permute(normalize(crnn_resize_img(imgs[i], max_wh, shape), REC_MEAN, REC_SCALE, True)) for i in idx]
        yield idx, np.stack(batch)


def read_dict(path):
    """Utility::ReadDict, utility.cpp:32-48 + label list, ocr_rec.h:82-84: ["#"] + lines + [" "]."""
    with open(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # getline stops at EOF after the final newline
    return ["#"] + [l.decode("utf-8") for l in lines] + [" "]


def ctc_greedy_decode(probs, labels):
    """CTC greedy decode, ocr_rec.cpp:97-128.  probs [T, C] (softmax).  Returns (text, score) or None
    when count == 0 (the reference then leaves the caller's zero-initialised slot untouched)."""
    idx = probs.argmax(-1)          # first maximum wins, like Utility::argmax (std::max_element)
    mx = probs.max(-1).astype(F32)
    return ctc_collapse(idx, mx, labels)


def ctc_collapse(idx, mx, labels):
    text, score, count, last = [], F32(0), 0, 0
    for n in range(len(idx)):
        a = int(idx[n])
        if a > 0 and not (n > 0 and a == last):
            score = F32(score + F32(mx[n]))
            count += 1
            text.append(labels[a])
        last = a
    if count == 0:
        return None
    return "".join(text), F32(score / F32(count))
