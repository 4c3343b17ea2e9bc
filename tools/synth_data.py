"""Seeded synthetic inputs named in SURVEY.md §8(d): S-card, S-page, S-rec crops, S-prob maps.
Rendered with cv2.putText (Hershey fonts), the same primitive the reference's own test image uses
(reference tests/test_ocr_worker.cpp:70-83)."""
from __future__ import annotations
import cv2
import numpy as np

_FONTS = [cv2.FONT_HERSHEY_SIMPLEX, cv2.FONT_HERSHEY_DUPLEX, cv2.FONT_HERSHEY_COMPLEX, cv2.FONT_HERSHEY_TRIPLEX]
_WORDS = ["Order", "No", "2026", "Total", "Amount", "Name", "Address", "Phone", "Card", "Member", "Date", "Valid",
          "Invoice", "Price", "Qty", "Pharmacy", "Receipt", "Street", "Code", "ID", "8812", "4471", "0093", "Thank", "you"]


def _line(rng, nwords):
    return " ".join(_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), nwords))


def reference_test_image():
    """createTestImage() of the reference test (tests/test_ocr_worker.cpp:70-83): 600x200 white, three lines."""
    img = np.full((200, 600, 3), 255, np.uint8)
    cv2.putText(img, "Hello World", (50, 50), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (0, 0, 0), 2)
    cv2.putText(img, "PaddleOCR Test", (50, 100), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (0, 0, 0), 2)
    cv2.putText(img, "123456789", (50, 150), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (0, 0, 0), 2)
    return img


def card(seed: int, width=1024, height=640, out=None, boxes=None, lines=None):
    """S-card: light background, 8-12 horizontal text lines, scales 0.6-1.2.
    `boxes` (optional list) receives (x0, y0, x1, y1) of every rendered line, `lines` (optional list) its
    (text, x, y, font, scale, thickness)."""
    rng = np.random.default_rng(seed)
    img = out if out is not None else np.empty((height, width, 3), np.uint8)
    img[:] = rng.integers(225, 256, 3, dtype=np.uint8)
    n = int(rng.integers(8, 13))
    y = 40
    for _ in range(n):
        scale = float(rng.uniform(0.6, 1.2))
        text = _line(rng, int(rng.integers(2, 6)))
        x = int(rng.integers(20, 200))
        color = tuple(int(c) for c in rng.integers(0, 90, 3))
        font, thick = _FONTS[int(rng.integers(0, len(_FONTS)))], int(rng.integers(1, 3))
        cv2.putText(img, text, (x, y), font, scale, color, thick, cv2.LINE_AA)
        if lines is not None:
            lines.append((text, x, y, font, scale, thick))
        if boxes is not None:
            (tw, th), base = cv2.getTextSize(text, font, scale, thick)
            boxes.append((x, y - th, min(x + tw, width - 1), y + base // 2))
        y += int(28 * scale + rng.integers(18, 30))
        if y > height - 20:
            break
    return img


def page(seed: int, size=2048, lines=220, info=None):
    """S-page: 2048x2048, >= 200 lines in 3 columns, widths 80..900 px.  `info` (optional list) receives
    (text, x, y, font, scale, thickness) of every rendered line."""
    rng = np.random.default_rng(seed)
    img = np.full((size, size, 3), 250, np.uint8)
    cols = [30, 710, 1390]
    per = (lines + 2) // 3
    for cx in cols:
        y = 30
        for _ in range(per):
            scale = float(rng.uniform(0.5, 0.9))
            text = _line(rng, int(rng.integers(1, 7)))
            font = _FONTS[int(rng.integers(0, len(_FONTS)))]
            cv2.putText(img, text, (cx, y), font, scale, (20, 20, 20), 1, cv2.LINE_AA)
            if info is not None:
                info.append((text, cx, y, font, scale, 1))
            y += int(rng.integers(24, 29))
            if y > size - 10:
                break
    return img


def rec_crops(n: int, height=48, width=320, seed=0, texts=None):
    """S-rec: n text-line crops of height x width u8 with rendered text (`texts`: optional list receiving it)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, height, width, 3), np.uint8)
    for i in range(n):
        out[i] = rng.integers(215, 256, 3, dtype=np.uint8)
        text = _line(rng, 3)
        cv2.putText(out[i], text, (4, int(height * 0.72)), _FONTS[i % 4], height / 44.0, (10, 10, 10), 2, cv2.LINE_AA)
        if texts is not None:
            texts.append(text)
    return out


def prob_map(seed: int, height=320, width=512, n_boxes=12, rings=2, lines=2):
    """S-prob: synthetic DB probability map: blurred rotated rectangles laid out on a jittered grid (so that most
    stay separate blobs; a few touch and merge), rings (hole contours), 1-px lines, low-amplitude noise."""
    rng = np.random.default_rng(seed)
    p = np.zeros((height, width), np.float32)
    cols = max(1, width // 150)
    rows = max(1, min(-(-n_boxes // cols), height // 20))
    ch, cw = height / rows, width / cols
    k = 0
    for r in range(rows):
        for c in range(cols):
            if k >= n_boxes:
                break
            k += 1
            cx = (c + 0.5) * cw + float(rng.uniform(-0.1, 0.1)) * cw
            cy = (r + 0.5) * ch + float(rng.uniform(-0.15, 0.15)) * ch
            sz = (float(rng.uniform(0.35, 0.95)) * cw, float(rng.uniform(0.25, 0.7)) * min(ch, 40.0))
            ang = float(rng.uniform(-8, 8)) if rng.random() < 0.8 else float(rng.uniform(-35, 35))
            pts = cv2.boxPoints(((cx, cy), sz, ang)).astype(np.int32)
            cv2.fillPoly(p, [pts], float(rng.uniform(0.45, 0.98)))
    for _ in range(rings):
        c = (int(rng.uniform(40, max(41, width - 40))), int(rng.uniform(20, max(21, height - 20))))
        cv2.ellipse(p, c, (int(rng.uniform(12, 40)), int(rng.uniform(8, 22))), float(rng.uniform(0, 180)), 0, 360,
                    float(rng.uniform(0.6, 0.95)), int(rng.integers(2, 5)))
    for _ in range(lines):
        a = (int(rng.uniform(0, width)), int(rng.uniform(0, height)))
        b = (int(rng.uniform(0, width)), int(rng.uniform(0, height)))
        cv2.line(p, a, b, 0.9, 1)
    p = cv2.GaussianBlur(p, (5, 5), 1.0)
    p += rng.normal(0, 0.02, p.shape).astype(np.float32)
    return np.clip(p, 0.0, 1.0).astype(np.float32)
