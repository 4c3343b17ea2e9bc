"""Device JPEG decode (replaces cv::imread / cv::imdecode, reference src/ocr_ipc_service.cpp:336-344)
through the C ABI: byte work, so the bar is BIT-EXACT.

 * against the oracle (oracle/jpeg_decode.py, itself pinned to cv2.imdecode in tests/test_oracle_cpu.py) on the
   reference's own test image and the 36 + 4 seeded encodes of that test (qualities 30 / 90 / 100, 4:2:0 / 4:2:2 / 4:4:4 /
   grey, odd sizes, with and without restart markers);
 * at BASELINE sizes (1024x640 cards, a 2048x2048 page), where the pure-Python oracle is too slow, against
   cv2.imdecode itself -- the library the oracle is pinned to;
 * the encoded-input worker / pool calls give byte-for-byte the result line of the raw-pixel calls on the decoded image;
 * files outside the supported subset are refused with a reason, never mis-decoded.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _seeded_encodes():
    import cv2
    import synth_data
    rng = np.random.default_rng(0)
    imgs = [synth_data.card(0)[:120, :201], rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
            rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), synth_data.card(1)[100:117, :300]]
    samplings = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    out = []
    for k, im in enumerate(imgs):
        for q in (30, 90, 100):
            for sf in samplings:
                rst = (k + q) % 3  # 0 = no restart markers
                ok, buf = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                                    cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                assert ok
                out.append(((k, q, sf, rst), buf.tobytes()))
        ok, buf = cv2.imencode(".jpg", cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 80])
        out.append(((k, "grey"), buf.tobytes()))
    return out


def test_device_decode_is_bit_exact_against_the_oracle(golden_dir):
    import b200ocr
    from oracle import jpeg_decode
    data = open(os.path.join(golden_dir, "card-jd.jpg"), "rb").read()
    assert np.array_equal(b200ocr.jpeg_decode(data), jpeg_decode.decode(data))
    n = 0
    for tag, buf in _seeded_encodes():
        got = b200ocr.jpeg_decode(buf)
        ref = jpeg_decode.decode(buf)
        assert got.shape == ref.shape and np.array_equal(got, ref), (tag, int((got != ref).sum()))
        n += 1
    assert n == 40


@pytest.mark.parametrize("quality,sampling,rst", [(95, "420", 0), (75, "420", 0), (90, "444", 0), (85, "422", 7), (60, "420", 16),
                                                   (100, "444", 1)])
def test_device_decode_at_baseline_sizes_equals_cv2(quality, sampling, rst):
    import cv2
    import b200ocr
    import synth_data
    sf = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}[sampling]
    imgs = [synth_data.card(40 + quality), synth_data.card(41 + quality, 1023, 637)]
    if quality == 75:
        imgs.append(synth_data.page(5))
    if quality == 95:  # photographic content: every coefficient position in use
        rng = np.random.default_rng(3)
        imgs.append(cv2.GaussianBlur(rng.integers(0, 256, (640, 1024, 3), dtype=np.uint8), (5, 5), 1.2))
    for im in imgs:
        ok, buf = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                            cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        assert ok
        got = b200ocr.jpeg_decode(buf)
        ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
        assert got.shape == ref.shape and np.array_equal(got, ref), int((got != ref).sum())


def test_encoded_worker_calls_equal_raw_pixel_calls(models_dir, golden_dir):
    import cv2
    import b200ocr
    import synth_data
    files = [open(os.path.join(golden_dir, "card-jd.jpg"), "rb").read()]
    for i in range(5):
        ok, buf = cv2.imencode(".jpg", synth_data.card(70 + i), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL,
                                                                 (0, 4, 0, 64, 1)[i]])
        files.append(buf.tobytes())
    ok, png = cv2.imencode(".png", synth_data.card(3))
    w = b200ocr.Worker(4, models_dir, enable_cls=True)
    ids = list(range(20, 20 + len(files) + 2))
    lines = w.process_encoded(ids, files + [png.tobytes(), b""])
    assert 0 < w.last_encoded_h2d_bytes < sum(len(f) for f in files) + 64 * 1024 * len(files)
    for rid, f, line in zip(ids, files, lines):
        img = cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
        d, r = json.loads(line), json.loads(w.process(rid, img))
        assert d["success"] and (d["width"], d["height"]) == (img.shape[1], img.shape[0])
        assert d["words"] == r["words"] and len(d["words"]) > 0
    bad = json.loads(lines[-2])
    assert bad["success"] is False and bad["error"].startswith("Unsupported image encoding") and bad["request_id"] == ids[-2]
    empty = json.loads(lines[-1])
    assert empty["success"] is False and empty["error"] == "Empty image data provided"
    # the pool takes encoded and raw requests side by side
    pool = b200ocr.Pool(models_dir, devices=(0,), workers_per_device=2, enable_cls=True, max_batch=4)
    tickets = []
    for k, f in enumerate(files):
        tickets.append((k, pool.submit_encoded(100 + k, f)))
        tickets.append((k, pool.submit(200 + k, cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR))))
    res = {}
    for k, t in tickets:
        d = json.loads(pool.wait(t))
        assert d["success"]
        res.setdefault(k, []).append(d["words"])
    assert all(a == b for a, b in res.values())
    pool.close()


def test_unsupported_files_are_refused_with_a_reason():
    import cv2
    import b200ocr
    import synth_data
    im = synth_data.card(9)[:64, :96]
    ok, prog = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(b200ocr.Error, match="sequential"):
        b200ocr.jpeg_decode(prog)
    ok, png = cv2.imencode(".png", im)
    with pytest.raises(b200ocr.Error, match="not a JPEG"):
        b200ocr.jpeg_decode(png)
    ok, good = cv2.imencode(".jpg", im)
    with pytest.raises(b200ocr.Error):
        b200ocr.jpeg_decode(good.tobytes()[:200])  # truncated inside the headers
    # truncated inside the entropy-coded data (cv2.imdecode refuses such a file; libjpeg itself completes the MCU it
    # was in with zero bits and leaves the rest mid grey, jdhuff.c `insufficient_data`, which is what the device
    # decoder restates): the rows decoded before the cut equal the full decode, the tail is flat
    whole = cv2.imdecode(good, cv2.IMREAD_COLOR)
    raw = good.tobytes()
    sos = raw.index(b"\xff\xda")
    cut = b200ocr.jpeg_decode(raw[: sos + (len(raw) - sos) * 2 // 3])
    assert cut.shape == whole.shape and np.array_equal(cut[:8], whole[:8])
    assert len(np.unique(cut[-8:].reshape(-1, 3), axis=0)) == 1


def test_sequential_and_parallel_entropy_decoders_agree(tmp_path):
    """Files without restart markers go through the self-synchronising parallel decoder; B200OCR_JPEG_SEQUENTIAL=1 (read
    once per process, hence the subprocess) sends them through the one-thread-per-interval decoder instead.  Same bytes
    out, and both equal cv2.imdecode."""
    import subprocess
    import sys
    import cv2
    import synth_data
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(11)
    imgs = [synth_data.card(900), synth_data.page(9)[:1500, :1111], synth_data.card(901)[:97, :203],
            cv2.GaussianBlur(rng.integers(0, 256, (333, 517, 3), dtype=np.uint8), (3, 3), 0.8)]
    files = []
    for k, im in enumerate(imgs):
        for q, sf in ((50, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), (92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444),
                      (100, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422)):
            path = os.path.join(str(tmp_path), f"f{k}_{q}.jpg")
            assert cv2.imwrite(path, im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])
            files.append(path)
    script = (
        "import sys, numpy as np\n"
        "root = sys.argv[1]\n"
        "sys.path[:0] = [root, root + '/cpp-paddle-ocr_b200']\n"
        "import b200ocr\n"
        "np.savez(sys.argv[2], *[b200ocr.jpeg_decode(open(f, 'rb').read()) for f in sys.argv[3:]])\n")
    outs = {}
    for name, env in (("parallel", {}), ("sequential", {"B200OCR_JPEG_SEQUENTIAL": "1"})):
        out = os.path.join(str(tmp_path), name + ".npz")
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", script, root, out] + files, capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[name] = np.load(out)
    for k, f in enumerate(files):
        ref = cv2.imdecode(np.fromfile(f, np.uint8), cv2.IMREAD_COLOR)
        a, b = outs["parallel"][f"arr_{k}"], outs["sequential"][f"arr_{k}"]
        assert np.array_equal(a, ref), (f, int((a != ref).sum()))
        assert np.array_equal(b, ref), (f, int((b != ref).sum()))
