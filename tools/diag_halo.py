"""Diagnostic (GPU box, B200OCR_CONV_HALO=1): per-tap check of the halo-mode 3x3 convolution."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import b200ocr
rng = np.random.default_rng(0)
n, c, h, w = 4, 64, 80, 128
x = rng.standard_normal((n, c, h, w)).astype(np.float16).astype(np.float32)
for kh, kw in ((3, 3), (1, 3), (3, 1)):
    for ky in range(kh):
        for kx in range(kw):
            f = np.zeros((c, c, kh, kw), np.float32)
            for i in range(c):
                f[i, i, ky, kx] = 1.0
            got = b200ocr.kernel_conv(x, f, np.zeros(c, np.float32))
            ref = np.zeros_like(x)
            dy, dx = ky - kh // 2, kx - kw // 2
            ys = slice(max(0, -dy), min(h, h - dy)); xs = slice(max(0, -dx), min(w, w - dx))
            ref[:, :, ys, xs] = x[:, :, max(0, dy):max(0, dy) + (ys.stop - ys.start), max(0, dx):max(0, dx) + (xs.stop - xs.start)]
            bad = np.abs(got - ref) > 1e-3
            print(f"{kh}x{kw} tap ({ky},{kx}): mismatches {int(bad.sum())} of {bad.size}", end="")
            if bad.any():
                idx = np.argwhere(bad)
                print("  first", idx[0].tolist(), "rows bad:", sorted(set(idx[:, 2].tolist()))[:12], "cols bad:", sorted(set(idx[:, 3].tolist()))[:12], end="")
            print()
