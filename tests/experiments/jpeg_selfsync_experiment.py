"""Design experiment for the device JPEG decoder, CPU only.

Huffman decoding is the serial part of JPEG.  The plan is to cut the entropy-coded segment into fixed-size chunks and
let one decoder per chunk start SPECULATIVELY at its first bit, assuming it stands at the first code of a block of
component 0, and rely on the self-synchronising property of Huffman streams: after a few codes a wrongly started decoder
falls onto true code boundaries, and once its (bit position, block-in-MCU, coefficient index) state equals the true
state it stays correct.  This script measures, against the sequential decode of oracle/jpeg_decode.py's tables, how many
bits a speculative decoder needs before it is synchronised, for every chunk start of an image:

    python tests/experiments/jpeg_selfsync_experiment.py [file.jpg ...]        (default: tests/golden/card-jd.jpg + synthetic cards)
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import jpeg_decode as J  # noqa: E402


def parse(data):
    dc, ac, comps, pos = {}, {}, None, 2
    while True:
        marker = data[pos + 1]
        (length,) = struct.unpack(">H", data[pos + 2:pos + 4])
        seg = data[pos + 4:pos + 2 + length]
        if marker == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                n = sum(counts)
                (dc if tc == 0 else ac)[th] = J._huff_table(counts, list(seg[k + 17:k + 17 + n]))
                k += 17 + n
        elif marker == 0xC0:
            nc = seg[5]
            comps = [(seg[6 + 3 * i], (seg[7 + 3 * i] >> 4) * (seg[7 + 3 * i] & 15)) for i in range(nc)]
        elif marker == 0xDD and struct.unpack(">H", seg[:2])[0]:
            raise SystemExit("restart intervals present: chunks can simply start at the markers")
        elif marker == 0xDA:
            sel = {seg[1 + 2 * i]: (seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(seg[0])}
            pos += 2 + length
            break
        pos += 2 + length
    # un-stuffed entropy bytes up to the next marker
    out, p = bytearray(), pos
    while p < len(data):
        b = data[p]
        if b == 0xFF:
            if data[p + 1] == 0:
                out.append(0xFF)
                p += 2
                continue
            break
        out.append(b)
        p += 1
    bits = np.unpackbits(np.frombuffer(bytes(out), np.uint8))
    # block schedule of one MCU: (dc table, ac table) per block
    sched = []
    for cid, nblk in comps:
        sched += [(dc[sel[cid][0]], ac[sel[cid][1]])] * nblk
    return bits, sched


def run(bits, sched, start, state=(0, 0), limit=None):
    """Decode codes from bit `start` in state (block-in-MCU, coefficient index k; k == 0: a DC code is next).
    Yields (bit position, state) BEFORE every code; stops at `limit` bits or on an invalid code / end of data."""
    p, (blk, k) = start, state
    n = len(bits)
    while p < n and (limit is None or p < limit):
        yield p, (blk, k)
        table = sched[blk][0 if k == 0 else 1]
        code, sym = 0, None
        for length in range(1, 17):
            if p + length > n:
                return
            code = (code << 1) | int(bits[p + length - 1])
            sym = table.get((length, code))
            if sym is not None:
                p += length
                break
        if sym is None:
            return  # invalid code: a real decoder would restart one bit further; count as "not yet synchronised"
        if k == 0:
            p += sym
            k = 1
        else:
            r, s = sym >> 4, sym & 15
            p += s
            if s == 0 and r != 15:
                k = 64
            else:
                k += r + 1 if s else 16
        if k >= 64:
            k, blk = 0, (blk + 1) % len(sched)


def experiment(name, data, chunk_bits=8192):
    bits, sched = parse(data)
    truth = dict(run(bits, sched, 0))  # bit position -> state at every true code boundary
    dist, failed = [], 0
    for start in range(chunk_bits, len(bits) - 64, chunk_bits):
        synced = None
        for p, st in run(bits, sched, start, limit=start + 4 * chunk_bits):
            if truth.get(p) == st:
                synced = p - start
                break
        if synced is None:
            failed += 1
        else:
            dist.append(synced)
    d = np.array(dist) if dist else np.zeros(1)
    print(f"{name}: {len(bits) // 8} entropy bytes, {len(sched)} blocks/MCU, {len(dist) + failed} chunk starts of "
          f"{chunk_bits // 8} B: synchronised after median {int(np.median(d))} bits, p90 {int(np.percentile(d, 90))}, "
          f"max {int(d.max())}; not within 4 chunks: {failed}")


if __name__ == "__main__":
    import cv2
    import synth_data
    files = sys.argv[1:]
    if files:
        for f in files:
            experiment(os.path.basename(f), open(f, "rb").read())
    else:
        experiment("card-jd.jpg", open(os.path.join(ROOT, "tests", "golden", "card-jd.jpg"), "rb").read(), 2048)
        for q in (50, 75, 95):
            ok, buf = cv2.imencode(".jpg", synth_data.card(3), [cv2.IMWRITE_JPEG_QUALITY, q])
            experiment(f"S-card 1024x640 q{q}", buf.tobytes())
