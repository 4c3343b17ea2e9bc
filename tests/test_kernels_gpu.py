"""Kernel-level parity of the hand-written floating-point kernels against a plain PyTorch fp32 reference of the same
op, through the C ABI (b200ocr_kernel_*).  Inputs are rounded to fp16 first (that is what the device stores), so the
only differences left are fp16 rounding of the output (<= 2^-11 relative), fp16 rounding of the filter taps where the
recognizer uses them, and fp32 summation order.  Tolerance: 2e-3 of the output scale + 2e-3 absolute."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h(a):  # what the device sees
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def _ref_dwconv(x, filt, bias, k, sh, sw, act, s2, t2):
    import torch
    import torch.nn.functional as F
    c = x.shape[1]
    y = F.conv2d(torch.tensor(x), torch.tensor(filt)[:, None], torch.tensor(bias), stride=(sh, sw), padding=k // 2, groups=c)
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = y * torch.clamp(y + 3, 0, 6) / 6
    return (s2 * y + t2).numpy()


# (k, sh, sw) x heights that hit every row-block path (static variants for 14/7/4/2 rows, generic tiles otherwise)
@pytest.mark.parametrize("k,sh,sw", [(3, 1, 1), (3, 2, 1), (3, 1, 2), (3, 2, 2), (5, 1, 1), (5, 2, 1), (5, 2, 2)])
@pytest.mark.parametrize("c,h,w", [(16, 14, 37), (64, 14, 200), (240, 7, 100), (480, 4, 53), (480, 2, 19), (24, 23, 31),
                                   (88, 3, 96), (200, 2, 48), (8, 24, 96), (40, 9, 11), (96, 40, 64), (192, 20, 33)])
@pytest.mark.parametrize("fp16w", [True, False], ids=["w16", "w32"])
def test_dwconv_matches_torch(k, sh, sw, c, h, w, fp16w):
    import b200ocr
    rng = np.random.default_rng(k * 1000 + sh * 100 + sw * 10 + c + h)
    n = 3
    x = _h(rng.standard_normal((n, c, h, w)))
    filt = rng.standard_normal((c, k, k)).astype(np.float32) / k
    if fp16w:
        filt = _h(filt)
    bias = rng.standard_normal(c).astype(np.float32) * 0.3
    act = int(rng.integers(0, 3))
    s2, t2 = (1.07, -0.04) if act == 2 else (1.0, 0.0)
    got = b200ocr.kernel_dwconv(x, filt, bias, k, sh, sw, act, s2, t2, fp16_weights=fp16w)
    ref = _ref_dwconv(x, filt, bias, k, sh, sw, act, s2, t2)
    assert got.shape == ref.shape
    tol = 2e-3 * float(np.abs(ref).max()) + 2e-3
    assert float(np.abs(got - ref).max()) <= tol, (float(np.abs(got - ref).max()), tol)


@pytest.mark.parametrize("k,sh,sw,c,h,w", [(5, 1, 1, 240, 7, 100), (3, 1, 1, 64, 14, 61), (5, 2, 1, 480, 7, 40), (3, 1, 2, 128, 7, 90),
                                           (5, 1, 1, 88, 3, 50)])
def test_dwconv_ragged_rows_equal_dense_rows_bitwise(k, sh, sw, c, h, w):
    """A row of a ragged batch (zero beyond its width) equals the same row run alone at its own width, bit for bit;
    beyond the row's valid output width the result is zero."""
    import b200ocr
    rng = np.random.default_rng(7)
    widths = [w, max(8, w // 2 + 1), max(8, w // 3), w - 3]
    x = _h(rng.standard_normal((len(widths), c, h, w)))
    for i, wd in enumerate(widths):
        x[i, :, :, wd:] = 0
    filt = _h(rng.standard_normal((c, k, k)) / k)
    bias = rng.standard_normal(c).astype(np.float32) * 0.3
    out_w = [(wd + 2 * (k // 2) - k) // sw + 1 for wd in widths]
    got = b200ocr.kernel_dwconv(x, filt, bias, k, sh, sw, 2, 0.97, 0.02, fp16_weights=True, out_widths=out_w)
    for i, wd in enumerate(widths):
        alone = b200ocr.kernel_dwconv(x[i:i + 1, :, :, :wd], filt, bias, k, sh, sw, 2, 0.97, 0.02, fp16_weights=True)
        assert np.array_equal(got[i, :, :, :out_w[i]], alone[0])
        assert not got[i, :, :, out_w[i]:].any()


def _ref_attention(qkv, heads, hd, scale, valid):
    import torch
    n, t, _ = qkv.shape
    q, k, v = torch.tensor(qkv).reshape(n, t, 3, heads, hd).permute(2, 0, 3, 1, 4)  # [n, heads, t, hd]
    out = torch.zeros(n, t, heads * hd)
    for i in range(n):
        tv = t if valid is None else int(valid[i])
        s = (q[i, :, :tv] @ k[i, :, :tv].transpose(1, 2)) * scale
        o = torch.softmax(s, -1) @ v[i, :, :tv]              # [heads, tv, hd]
        out[i, :tv] = o.permute(1, 0, 2).reshape(tv, heads * hd)
    return out.numpy()


@pytest.mark.parametrize("t", [1, 7, 16, 24, 33, 50, 64, 100, 125, 160, 240])
def test_attention_matches_torch(t):
    import b200ocr
    rng = np.random.default_rng(t)
    heads, hd = 8, 15
    qkv = _h(rng.standard_normal((4, t, 3 * heads * hd)) * 1.5)
    got = b200ocr.kernel_attention(qkv, heads, hd, hd ** -0.5)
    ref = _ref_attention(qkv, heads, hd, hd ** -0.5, None)
    assert float(np.abs(got - ref).max()) <= 4e-3, float(np.abs(got - ref).max())


def test_attention_ragged_rows_equal_dense_rows_bitwise():
    import b200ocr
    rng = np.random.default_rng(3)
    heads, hd, t = 8, 15, 90
    valid = np.array([90, 1, 33, 64, 17, 80], np.int32)
    qkv = _h(rng.standard_normal((len(valid), t, 3 * heads * hd)))
    for i, tv in enumerate(valid):
        qkv[i, tv:] = 0
    got = b200ocr.kernel_attention(qkv, heads, hd, hd ** -0.5, valid=valid)
    ref = _ref_attention(qkv, heads, hd, hd ** -0.5, valid)
    assert float(np.abs(got - ref).max()) <= 4e-3
    for i, tv in enumerate(valid):
        alone = b200ocr.kernel_attention(qkv[i:i + 1, :tv], heads, hd, hd ** -0.5)
        assert np.array_equal(got[i, :tv], alone[0])
        assert not got[i, tv:].any()


def _ref_conv(x, filt, bias, act, s2, t2, residual):
    import torch
    import torch.nn.functional as F
    kh, kw = filt.shape[2:]
    y = F.conv2d(torch.tensor(x), torch.tensor(filt), torch.tensor(bias), padding=(kh // 2, kw // 2))
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = y * torch.clamp(y + 3, 0, 6) / 6
    elif act == 3:
        y = y * torch.sigmoid(y)
    y = s2 * y + t2
    if residual is not None:
        y = y + torch.tensor(residual)
    return y.numpy()


# tcgen05 implicit GEMM on shapes the shipped graphs do not contain: pixel counts that are not a multiple of the 128-row
# tile, channel counts that leave partial K chunks / partial 16-column groups / more than one N tile, 1x1, 1x3, 3x3
@pytest.mark.parametrize("n,cin,cout,h,w,kh,kw", [
    (3, 16, 32, 14, 37, 1, 1), (2, 64, 64, 7, 100, 1, 1), (5, 240, 240, 7, 33, 1, 1), (2, 480, 480, 2, 51, 1, 1),
    (1, 120, 360, 1, 97, 1, 1), (4, 96, 24, 20, 32, 3, 3), (2, 24, 24, 40, 64, 3, 3), (1, 96, 96, 10, 16, 3, 3),
    (3, 64, 120, 1, 61, 1, 3), (2, 8, 16, 9, 9, 1, 1), (1, 40, 136, 5, 13, 3, 3), (7, 200, 72, 3, 19, 1, 1),
    (2, 384, 96, 10, 16, 1, 1), (1, 72, 264, 6, 50, 1, 1), (2, 136, 8, 12, 12, 3, 3), (40, 240, 240, 7, 100, 1, 1),
    (6, 96, 24, 80, 128, 3, 3), (3, 24, 24, 160, 250, 3, 3), (50, 64, 120, 1, 97, 1, 3),
])
@pytest.mark.parametrize("simt", [False, True], ids=["tcgen05", "cuda-core"])
def test_conv_matches_torch(n, cin, cout, h, w, kh, kw, simt):
    import b200ocr
    if simt and cin * cout * kh * kw * h * w * n > 3e8:
        pytest.skip("CUDA-core path only on the small shapes")
    rng = np.random.default_rng(cin * 7 + cout * 3 + h + w + kh)
    x = _h(rng.standard_normal((n, cin, h, w)))
    filt = _h(rng.standard_normal((cout, cin, kh, kw)) / np.sqrt(cin * kh * kw))
    bias = rng.standard_normal(cout).astype(np.float32) * 0.2
    act = int(rng.integers(0, 4))
    s2, t2 = (0.95, 0.03) if act == 2 else (1.0, 0.0)
    res = _h(rng.standard_normal((n, cout, h, w))) if (cin + cout) % 3 == 0 else None
    # 2 = not the narrow-1x1 mma.sync kernel (tested further down): this test is about the tcgen05 path
    got = b200ocr.kernel_conv(x, filt, bias, act, s2, t2, residual=res, force_simt=1 if simt else 2)
    ref = _ref_conv(x, filt, bias, act, s2, t2, res)
    tol = 2e-3 * float(np.abs(ref).max()) + 2e-3
    assert float(np.abs(got - ref).max()) <= tol, (float(np.abs(got - ref).max()), tol)


def test_conv_ragged_rows_equal_dense_rows_bitwise():
    import b200ocr
    rng = np.random.default_rng(11)
    for (cin, cout, h, w, kh, kw) in [(240, 240, 7, 100, 1, 1), (64, 120, 1, 97, 1, 3), (96, 24, 4, 60, 3, 3)]:
        widths = [w, w // 2 + 3, 9, w - 1]
        x = _h(rng.standard_normal((len(widths), cin, h, w)))
        for i, wd in enumerate(widths):
            x[i, :, :, wd:] = 0
        filt = _h(rng.standard_normal((cout, cin, kh, kw)) / np.sqrt(cin * kh * kw))
        bias = rng.standard_normal(cout).astype(np.float32) * 0.2
        got = b200ocr.kernel_conv(x, filt, bias, 2, 1.03, -0.02, out_widths=widths)
        for i, wd in enumerate(widths):
            alone = b200ocr.kernel_conv(x[i:i + 1, :, :, :wd], filt, bias, 2, 1.03, -0.02)
            assert np.array_equal(got[i, :, :, :wd], alone[0]), (cin, cout, kh, kw, i)
            assert not got[i, :, :, wd:].any()

# ---- narrow 1x1 convolutions (pwconv.cu: mma.sync stream over the pixels; the engine takes it for C_in <= 48 and
# C_out <= 64, the wider outputs below go to the tcgen05 kernel unless B200OCR_PWCONV_MAX_COUT raises the limit)
@pytest.mark.parametrize("n,cin,cout,h,w", [
    (3, 8, 8, 24, 96), (2, 8, 24, 13, 37), (2, 8, 32, 5, 7), (3, 16, 32, 14, 193), (2, 16, 40, 12, 48), (2, 16, 88, 6, 24),
    (2, 16, 104, 6, 24), (2, 24, 8, 12, 48), (2, 32, 8, 12, 48), (2, 32, 16, 6, 24), (2, 32, 48, 20, 32), (2, 32, 200, 3, 12),
    (2, 40, 16, 6, 24), (2, 48, 16, 6, 24), (2, 48, 48, 20, 33), (2, 48, 96, 10, 16), (1, 12, 96, 10, 16), (1, 18, 96, 5, 8),
    (1, 42, 96, 5, 8), (1, 16, 32, 1, 1), (1, 16, 32, 1, 17),
])
def test_narrow_pointwise_conv_matches_torch(n, cin, cout, h, w):
    import b200ocr
    rng = np.random.default_rng(cin * 11 + cout * 5 + h + w)
    x = _h(rng.standard_normal((n, cin, h, w)))
    filt = _h(rng.standard_normal((cout, cin, 1, 1)) / np.sqrt(cin))
    bias = rng.standard_normal(cout).astype(np.float32) * 0.2
    act = int(rng.integers(0, 4))
    s2, t2 = (0.95, 0.03) if act == 2 else (1.0, 0.0)
    res = _h(rng.standard_normal((n, cout, h, w))) if (cin + cout) % 3 == 0 else None
    got = b200ocr.kernel_conv(x, filt, bias, act, s2, t2, residual=res)
    ref = _ref_conv(x, filt, bias, act, s2, t2, res)
    tol = 2e-3 * float(np.abs(ref).max()) + 2e-3
    assert float(np.abs(got - ref).max()) <= tol, (float(np.abs(got - ref).max()), tol)
    # and it agrees with the CUDA-core kernel (same fp16 operands, fp32 accumulation) to output rounding
    simt = b200ocr.kernel_conv(x, filt, bias, act, s2, t2, residual=res, force_simt=True)
    assert float(np.abs(got - simt).max()) <= tol


def test_narrow_pointwise_conv_ragged_rows_equal_dense_rows_bitwise():
    import b200ocr
    rng = np.random.default_rng(12)
    for (cin, cout, h, w) in [(16, 32, 14, 61), (32, 64, 7, 50), (8, 24, 3, 33)]:
        widths = [w, w // 2 + 3, 9, w - 1]
        x = _h(rng.standard_normal((len(widths), cin, h, w)))
        for i, wd in enumerate(widths):
            x[i, :, :, wd:] = 0
        filt = _h(rng.standard_normal((cout, cin, 1, 1)) / np.sqrt(cin))
        bias = rng.standard_normal(cout).astype(np.float32) * 0.2
        got = b200ocr.kernel_conv(x, filt, bias, 2, 1.03, -0.02, out_widths=widths)
        for i, wd in enumerate(widths):
            alone = b200ocr.kernel_conv(x[i:i + 1, :, :, :wd], filt, bias, 2, 1.03, -0.02)
            assert np.array_equal(got[i, :, :, :wd], alone[0]), (cin, cout, i)
            assert not got[i, :, :, wd:].any()
