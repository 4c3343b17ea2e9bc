"""ORACLE (test infrastructure only; never imported by the product): baseline JPEG decoding as `cv::imread` /
`cv::imdecode` do it for the reference's ingest step (reference `src/ocr_ipc_service.cpp:336-344`).

The arithmetic lives in a third-party library the reference links through OpenCV — libjpeg(-turbo) with its defaults:
`JDCT_ISLOW` integer inverse DCT (jidctint.c), "fancy" triangle-filter chroma upsampling (jdsample.c:
h2v1_fancy_upsample / h2v2_fancy_upsample) and the table-driven YCbCr -> RGB conversion (jdcolor.c).  This file restates
those published algorithms for baseline sequential, Huffman-coded, 8-bit JPEGs (SOF0; grey, 4:4:4, 4:2:2, 4:2:0) in
plain Python / numpy.  It is pinned against the library itself: `tests/test_oracle_cpu.py` compares it bit for bit with
`cv2.imdecode` (cv2 4.13 bundles libjpeg-turbo) on the reference's `images/card-jd.jpg` and on seeded synthetic images
at several qualities and samplings.  Pure-Python entropy decoding: small images only.

It is the checker for a device-side decoder that does not exist yet.
"""
import struct

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63], np.int64)


class _Bits:
    """MSB-first bit reader over entropy-coded data (0xFF00 un-stuffed, stops at a marker)."""

    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        b = self.d[self.p] if self.p < len(self.d) else 0
        if b == 0xFF:
            nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
            if nxt == 0x00:
                self.p += 2
            elif 0xD0 <= nxt <= 0xD7:  # restart marker: handled by the caller, feed zeros until then
                b = 0
            else:
                b = 0  # any other marker: pad with zeros (libjpeg warns and does the same)
        else:
            self.p += 1
        self.acc = (self.acc << 8) | b
        self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        while self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        """Drop the partial byte and consume the RSTn marker."""
        self.acc, self.n = 0, 0
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """(length, code) -> symbol for a DHT segment (JPEG Annex C)."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(bits, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError("bad Huffman code")


def _extend(v, t):
    return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1


# ---- jidctint.c (JDCT_ISLOW): CONST_BITS 13, PASS1_BITS 2
_F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
          f2053=16819, f2562=20995, f3072=25172)


def _idct_1d(x, shift):
    """One pass over the LAST axis of x (int64 [..., 8]); DESCALE by `shift` with round-half-up."""
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * _F["f0541"]
    tmp2 = z1 - z3 * _F["f1847"]
    tmp3 = z1 + z2 * _F["f0765"]
    tmp0 = (x[..., 0] + x[..., 4]) << 13
    tmp1 = (x[..., 0] - x[..., 4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F["f1175"]
    t0, t1, t2, t3 = t0 * _F["f0298"], t1 * _F["f2053"], t2 * _F["f3072"], t3 * _F["f1501"]
    z1, z2 = -z1 * _F["f0899"], -z2 * _F["f2562"]
    z3, z4 = -z3 * _F["f1961"] + z5, -z4 * _F["f0390"] + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], -1)
    return (out + (1 << (shift - 1))) >> shift


def idct_islow(coef):
    """coef: int64 [..., 8, 8] dequantised coefficients (row-major) -> uint8 samples [..., 8, 8]."""
    ws = _idct_1d(np.swapaxes(coef, -1, -2), 13 - 2)            # pass 1: columns
    out = _idct_1d(np.swapaxes(ws, -1, -2), 13 + 2 + 3)          # pass 2: rows
    return np.clip(out + 128, 0, 255).astype(np.uint8)


# ---- jdsample.c
def _h2v1_fancy(p):
    """Double the width of int plane p [h, w] with the 3:1 triangle filter."""
    p = p.astype(np.int64)
    h, w = p.shape
    out = np.empty((h, 2 * w), np.int64)
    left = np.concatenate([p[:, :1], p[:, :-1]], 1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def _h2v2_fancy(p):
    """Double width and height of plane p [h, w]: 3/4 nearer row + 1/4 further row, then the same along x, with
    libjpeg's alternating +8 / +7 rounding."""
    p = p.astype(np.int64)
    h, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], 0)     # row above (edge row replicated)
    dn = np.concatenate([p[1:], p[-1:]], 0)     # row below
    out = np.empty((2 * h, 2 * w), np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                      # column sums of this output row
        last = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
        even = (3 * cs + last + 8) >> 4
        odd = (3 * cs + nxt + 7) >> 4
        even[:, 0] = (4 * cs[:, 0] + 8) >> 4
        odd[:, -1] = (4 * cs[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out


# ---- jdcolor.c (SCALEBITS 16)
def _ycc_tables():
    x = np.arange(256, dtype=np.int64) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    half = 1 << 15
    return ((fix(1.40200) * x + half) >> 16, (fix(1.77200) * x + half) >> 16, -fix(0.71414) * x, -fix(0.34414) * x + half)


def ycc_to_bgr(y, cb, cr):
    cr_r, cb_b, cr_g, cb_g = _ycc_tables()
    y = y.astype(np.int64)
    r = y + cr_r[cr]
    g = y + ((cb_g[cb] + cr_g[cr]) >> 16)
    b = y + cb_b[cb]
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """Baseline JPEG bytes -> uint8 [h, w, 3] BGR (grey images are replicated to 3 channels, like IMREAD_COLOR)."""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG")
    qt, dc_tabs, ac_tabs = {}, {}, {}
    frame, restart_interval, pos = None, 0, 2
    while pos < len(data):
        if data[pos] != 0xFF:
            raise ValueError("marker expected")
        marker = data[pos + 1]
        if marker == 0xFF:
            pos += 1
            continue
        (length,) = struct.unpack(">H", data[pos + 2:pos + 4])
        seg = data[pos + 4:pos + 2 + length]
        if marker == 0xDB:
            k = 0
            while k < len(seg):
                pq, tq = seg[k] >> 4, seg[k] & 15
                n = 128 if pq else 64
                vals = np.frombuffer(seg[k + 1:k + 1 + n], ">u2" if pq else np.uint8).astype(np.int64)
                q = np.zeros(64, np.int64)
                q[ZIGZAG] = vals
                qt[tq] = q
                k += 1 + n
        elif marker == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                nsym = sum(counts)
                (dc_tabs if tc == 0 else ac_tabs)[th] = _huff_table(counts, list(seg[k + 17:k + 17 + nsym]))
                k += 17 + nsym
        elif marker == 0xC0 or marker == 0xC1:
            prec, h, w, nc = struct.unpack(">BHHB", seg[:6])
            if prec != 8:
                raise ValueError("8-bit samples only")
            frame = dict(h=h, w=w, comps=[dict(id=seg[6 + 3 * i], hs=seg[7 + 3 * i] >> 4, vs=seg[7 + 3 * i] & 15,
                                               tq=seg[8 + 3 * i]) for i in range(nc)])
        elif marker in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential Huffman JPEG is restated")
        elif marker == 0xDD:
            (restart_interval,) = struct.unpack(">H", seg[:2])
        elif marker == 0xDA:
            ns = seg[0]
            sel = {seg[1 + 2 * i]: (seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(ns)}
            pos += 2 + length
            break
        pos += 2 + length
    if frame is None:
        raise ValueError("no frame header")
    comps = frame["comps"]
    if len(sel) != len(comps):
        raise ValueError("interleaved single-scan images only")
    hmax, vmax = max(c["hs"] for c in comps), max(c["vs"] for c in comps)
    mcux, mcuy = -(-frame["w"] // (8 * hmax)), -(-frame["h"] // (8 * vmax))
    for c in comps:
        c["coef"] = np.zeros((mcuy * c["vs"], mcux * c["hs"], 64), np.int64)
        c["pred"] = 0
    bits = _Bits(data, pos)
    for m in range(mcux * mcuy):
        if restart_interval and m and m % restart_interval == 0:
            bits.restart()
            for c in comps:
                c["pred"] = 0
        my, mx = divmod(m, mcux)
        for c in comps:
            td, ta = sel[c["id"]]
            for by in range(c["vs"]):
                for bx in range(c["hs"]):
                    blk = c["coef"][my * c["vs"] + by, mx * c["hs"] + bx]
                    t = _decode_symbol(bits, dc_tabs[td])
                    c["pred"] += _extend(bits.get(t), t)
                    blk[0] = c["pred"]
                    k = 1
                    while k < 64:
                        rs = _decode_symbol(bits, ac_tabs[ta])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                        k += 1
    planes = []
    for c in comps:
        by, bx = c["coef"].shape[:2]
        px = idct_islow((c["coef"] * qt[c["tq"]]).reshape(by, bx, 8, 8))
        plane = px.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)
        # the real extent of this component; libjpeg replicates its last real row / column as upsampling context
        ch = -(-frame["h"] * c["vs"] // vmax)
        cw = -(-frame["w"] * c["hs"] // hmax)
        plane = plane[:ch, :cw]
        fh, fv = hmax // c["hs"], vmax // c["vs"]
        if (fh, fv) == (1, 1):
            full = plane.astype(np.int64)
        elif (fh, fv) == (2, 1):
            full = _h2v1_fancy(plane)
        elif (fh, fv) == (2, 2):
            full = _h2v2_fancy(plane)
        else:
            raise ValueError("sampling factors beyond 4:4:4 / 4:2:2 / 4:2:0 are not restated")
        planes.append(full[:frame["h"], :frame["w"]])
    if len(planes) == 1:
        g = planes[0].astype(np.uint8)
        return np.stack([g, g, g], -1)
    return ycc_to_bgr(planes[0], planes[1], planes[2])
