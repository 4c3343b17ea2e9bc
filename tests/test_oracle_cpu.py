"""CPU suite: pins the oracle against the third-party code the reference calls (cv2 4.13, the reference's own
Clipper compiled into oracle/_ref), checks the host-side geometry header the GPU kernels use against the same,
and checks the C-ABI library's load-time surface.  No GPU compute is called here."""
import ctypes as C
import json
import os
import re
import subprocess

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def geomlib():
    d = os.path.join(ROOT, "tests", "native")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(d, "libgeomtest.so"))
    L.geomtest_mini_box.restype = C.c_float
    L.geomtest_unclip_distance.restype = C.c_float
    L.geomtest_unclip_distance.argtypes = [C.c_void_p, C.c_float]
    L.geomtest_offset_round.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    return L


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    import b200ocr
    hdr = open(os.path.join(ROOT, "include", "b200ocr.h")).read()
    names = set(re.findall(r"\b(b200ocr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 40
    missing = [n for n in sorted(names) if not hasattr(b200ocr.lib, n)]
    assert not missing, missing
    assert "sm_100a" in b200ocr.version()


def test_no_cpu_fallback_without_a_gpu(models_dir):
    import b200ocr
    if b200ocr.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(b200ocr.Error, match="no CPU fallback"):
        b200ocr.Worker(0, models_dir)
    with pytest.raises(b200ocr.Error):
        b200ocr.Detector(models_dir + "/det")


def test_plans_build_on_host_and_cover_every_op(models_dir):
    import b200ocr
    for kind, layers in (("det", 77), ("cls", 65), ("rec", 57)):
        text = b200ocr.model_plan_text(f"{models_dir}/{kind}")
        head = text.splitlines()[0].split()
        assert head[1] == kind and int(head[5]) == layers
    with pytest.raises(b200ocr.Error, match="No valid model file"):
        b200ocr.model_plan_text("/nonexistent")


def test_param_listing_matches_pdiparams_layout(models_dir):
    """The shipped cls weights (real) parse with the sorted-name rule; sizes of the absent det/rec files follow."""
    import b200ocr
    from oracle.pdmodel import load_params, load_program, param_names
    prog = load_program(f"{models_dir}/cls/inference.pdmodel")
    params = load_params(prog, f"{models_dir}/cls/inference.pdiparams")  # asserts dims + exact file length
    listed = b200ocr.model_params(f"{models_dir}/cls/inference.pdmodel")
    assert [n for n, _ in listed] == param_names(prog)
    assert sum(int(np.prod(d)) for _, d in listed) == 133628 == sum(v.size for v in params.values())
    assert os.path.getsize(f"{models_dir}/det/inference.pdiparams") == 4692937
    assert os.path.getsize(f"{models_dir}/rec/inference.pdiparams") == 10766823


def test_dictionary_gives_6625_labels(models_dir):
    from oracle import ocr_ops
    labels = ocr_ops.read_dict(f"{models_dir}/rec/ppocr_keys_v1.txt")
    assert len(labels) == 6625 and labels[0] == "#" and labels[-1] == " "


# ------------------------------------------------------------------------------------------- oracle pins
def test_unclip_restatement_matches_reference_clipper():
    from oracle import ocr_ops, unclip
    rc = unclip.ref_clipper()
    if rc is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(0)
    for t in range(1500):
        rr = ((rng.uniform(20, 300), rng.uniform(20, 200)), (rng.uniform(1, 200), rng.uniform(1, 60)), rng.uniform(-90, 0))
        box, _ = ocr_ops.get_mini_boxes(rr)
        if t % 25 == 0:
            box = np.round(box)
        d = float(ocr_ops.get_contour_area(box, 1.8))
        path = [(int(box[i][0]), int(box[i][1])) for i in range(4)]
        ref, mine = rc.offset(path, d), unclip.offset_points_restated(path, d)
        if not ref or not mine:
            assert not ref and not mine
            continue
        ra = cv2.minAreaRect(np.array(ref, np.float32).reshape(-1, 1, 2))
        rb = cv2.minAreaRect(np.array(mine, np.float32).reshape(-1, 1, 2))
        assert np.allclose(cv2.boxPoints(ra), cv2.boxPoints(rb), atol=1e-3), (path, d)


def test_set_based_contours_equal_findcontours():
    """The statement the GPU labelling kernels implement (border = (8-conn fg component, 4-conn bg component) pair,
    start pixel, descending order) reproduces cv2.findContours(RETR_LIST): count, order, start point, point set."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(0)
    for trial in range(60):
        H, W = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        bm = (rng.random((H, W)) < rng.uniform(0.2, 0.8)).astype(np.uint8) * 255
        if trial % 3 == 0:
            bm = cv2.dilate(bm, np.ones((3, 3), np.uint8))
        cs, _ = cv2.findContours(bm, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        fg = bm > 0
        lab_f, nf = ndi.label(fg, structure=np.ones((3, 3)))
        lab_b, nb = ndi.label(np.pad(~fg, 1, constant_values=True), structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
        first_f = {c: np.flatnonzero(lab_f.ravel() == c)[0] for c in range(1, nf + 1)}
        first_b = {b: np.flatnonzero(lab_b.ravel() == b)[0] for b in range(1, nb + 1)}
        bout = {c: lab_b[first_f[c] // W + 1, first_f[c] % W] for c in first_f}
        borders = {}
        for y, x in zip(*np.nonzero(fg)):
            c = lab_f[y, x]
            for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                b = lab_b[y + 1 + dy, x + 1 + dx]
                if b:
                    key = ("o", c) if b == bout[c] else ("h", b)
                    borders.setdefault(key, set()).add((int(x), int(y)))
        mine = []
        for key, pts in borders.items():
            if key[0] == "o":
                y, x = divmod(int(first_f[key[1]]), W)
            else:
                py, px = divmod(int(first_b[key[1]]), W + 2)
                y, x = py - 1, px - 2
            mine.append((y * W + x, pts))
        mine.sort(key=lambda t: -t[0])
        assert len(mine) == len(cs)
        for c, (start, pts) in zip(cs, mine):
            p = c.reshape(-1, 2)
            assert int(p[0][1]) * W + int(p[0][0]) == start
            assert set(map(tuple, p.tolist())) == pts


def test_resize_restatement_is_bit_exact_vs_cv2():
    """The fixed-point formula the GPU resize kernel implements (preproc.cu resize_px), restated in numpy."""
    def ref(src, dw, dh):
        sh, sw, _ = src.shape
        if (sw, sh) == (dw, dh):
            return src.copy()
        if sw == 2 * dw and sh == 2 * dh:
            s = src.astype(np.int32)
            return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

        def coefs(dn, sn, is_x):
            scale = 1.0 / (dn / sn)
            f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
            s = np.floor(f).astype(np.int32)
            f = f - s.astype(np.float32)
            if is_x:
                f[s < 0] = 0; s[s < 0] = 0
                f[s >= sn - 1] = 0; s[s >= sn - 1] = sn - 1
            return s, np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32), np.rint(f * np.float32(2048)).astype(np.int32)
        sx, a0, a1 = coefs(dw, sw, True)
        sy, b0, b1 = coefs(dh, sh, False)
        s = src.astype(np.int32)
        hr = s[:, sx] * a0[None, :, None] + s[:, np.minimum(sx + 1, sw - 1)] * a1[None, :, None]
        s0, s1 = hr[np.clip(sy, 0, sh - 1)], hr[np.clip(sy + 1, 0, sh - 1)]
        out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
        return np.clip(out, 0, 255).astype(np.uint8)
    rng = np.random.default_rng(0)
    cases = [(1024, 640, 512, 320), (1024, 640, 960, 608), (391, 178, 384, 192), (300, 40, 192, 28), (57, 19, 84, 28)]
    cases += [(int(rng.integers(8, 300)), int(rng.integers(8, 200)), int(rng.integers(8, 300)), int(rng.integers(8, 120))) for _ in range(40)]
    for sw, sh, dw, dh in cases:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (dw, dh)), ref(src, dw, dh)), (sw, sh, dw, dh)


# ------------------------------------------------------------------------------------------- geometry header
def test_geom_min_area_rect_vs_cv2(geomlib):
    """csrc/geom.h min-area rectangle == cv2.minAreaRect on the borders cv2.findContours returns: filled / outlined
    ellipses, random lattice noise (many holes, spurs) and tiny rotated specks, where two rectangles of EQUAL area exist
    and the winner depends on the hull's vertex order (hull_order_like_cv).  No ties are tolerated."""
    rng = np.random.default_rng(1)
    imgs = []
    for t in range(300):
        H, W = int(rng.integers(10, 80)), int(rng.integers(10, 160))
        img = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(1, 4))):
            c = (int(rng.integers(0, W)), int(rng.integers(0, H)))
            cv2.ellipse(img, c, (int(rng.integers(1, W // 2 + 1)), int(rng.integers(1, H // 3 + 1))), float(rng.uniform(0, 180)),
                        0, 360, 255, int(rng.choice([-1, -1, 2])))
        imgs.append(img)
    for t in range(500):
        imgs.append(np.pad((rng.random((int(rng.integers(4, 14)), int(rng.integers(4, 16)))) < 0.55).astype(np.uint8) * 255, 2))
    for t in range(400):
        img = np.zeros((14, 14), np.uint8)
        pts = cv2.boxPoints(((7, 7), (float(rng.uniform(2, 7)), float(rng.uniform(2, 7))), float(rng.uniform(0, 90)))).astype(np.int32)
        cv2.fillPoly(img, [pts], 255)
        imgs.append(img)
    worst, total, holes = 0.0, 0, 0
    for img in imgs:
        cs, hier = cv2.findContours(img, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
        for ci, c in enumerate(cs):
            if len(c) <= 2:
                continue
            outer = int(hier[0][ci][3] < 0)
            rr = cv2.minAreaRect(c)
            pts = np.ascontiguousarray(c.reshape(-1, 2), np.float32)
            out = np.zeros(5, np.float32)
            geomlib.geomtest_min_area_rect(_fp(pts), len(pts), outer, _fp(out))
            a = cv2.boxPoints(rr)
            b = cv2.boxPoints(((out[0], out[1]), (out[2], out[3]), out[4]))
            d = max(min(np.abs(p - q).max() for q in b) for p in a)
            worst = max(worst, d)
            total += 1
            holes += 1 - outer
    assert total > 3000 and holes > 500 and worst < 1e-3, (worst, total, holes)


def test_geom_fill_poly_vs_cv2(geomlib):
    """geom::QuadMask == cv2.fillPoly for quads inside the mask AND quads that stick out of it (BoxScoreFast's mask is
    the box's bounding rectangle clamped to the probability map: a box at the map's border has vertices outside its
    mask, which OpenCV handles by clipping every edge first, src/postprocess_op.cpp:216-253)."""
    import math
    rng = np.random.default_rng(5)
    n_in = n_out = 0
    for t in range(6000):
        if t % 2:
            W, H = int(rng.integers(4, 90)), int(rng.integers(4, 60))
            rr = ((rng.uniform(-5, W + 5), rng.uniform(-5, H + 5)), (rng.uniform(1, W), rng.uniform(1, H / 2)), rng.uniform(-90, 0))
            pts = cv2.boxPoints(rr).astype(np.int32)
        else:  # exactly BoxScoreFast's construction for a text-like box near the border of a w x h map
            w, h = int(rng.integers(40, 200)), int(rng.integers(30, 120))
            cx = rng.choice([rng.uniform(-2, 12), rng.uniform(w - 12, w + 2), rng.uniform(0, w)])
            cy = rng.choice([rng.uniform(-2, 8), rng.uniform(h - 8, h + 2), rng.uniform(0, h)])
            box = cv2.boxPoints(((float(cx), float(cy)), (float(rng.uniform(4, 80)), float(rng.uniform(3, 25))), float(rng.uniform(-90, 0))))
            xmin, xmax = int(np.clip(math.floor(box[:, 0].min()), 0, w - 1)), int(np.clip(math.ceil(box[:, 0].max()), 0, w - 1))
            ymin, ymax = int(np.clip(math.floor(box[:, 1].min()), 0, h - 1)), int(np.clip(math.ceil(box[:, 1].max()), 0, h - 1))
            W, H = xmax - xmin + 1, ymax - ymin + 1
            pts = np.array([[int(p[0]) - xmin, int(p[1]) - ymin] for p in box], np.int32)
        inside = pts[:, 0].min() >= 0 and pts[:, 1].min() >= 0 and pts[:, 0].max() < W and pts[:, 1].max() < H
        m = np.zeros((H, W), np.uint8)
        cv2.fillPoly(m, [pts], 1)
        mine = np.zeros((H, W), np.uint8)
        x, y = np.ascontiguousarray(pts[:, 0], np.int32), np.ascontiguousarray(pts[:, 1], np.int32)
        geomlib.geomtest_quad_mask(_fp(x), _fp(y), W, H, _fp(mine))
        assert np.array_equal(m, mine), (W, H, pts.tolist())
        n_in += inside
        n_out += not inside
    assert n_in > 500 and n_out > 2000, (n_in, n_out)


def test_geom_mini_box_unclip_and_finish_vs_oracle(geomlib):
    from oracle import ocr_ops, unclip
    rng = np.random.default_rng(2)
    for t in range(800):
        rr = ((float(rng.uniform(20, 300)), float(rng.uniform(20, 200))), (float(rng.uniform(1, 200)), float(rng.uniform(1, 60))),
              float(rng.uniform(-90, 0)))
        box, ssid = ocr_ops.get_mini_boxes(rr)
        out = np.zeros(8, np.float32)
        s2 = geomlib.geomtest_mini_box(_fp(np.array([rr[0][0], rr[0][1], rr[1][0], rr[1][1], rr[2]], np.float32)), _fp(out))
        assert np.allclose(out.reshape(4, 2), box, atol=2e-4) and abs(s2 - ssid) < 1e-4
        d = ocr_ops.get_contour_area(box, 1.8)
        d2 = geomlib.geomtest_unclip_distance(_fp(np.ascontiguousarray(box, np.float32)), 1.8)
        assert np.float32(d) == np.float32(d2)
        path = [(int(box[i][0]), int(box[i][1])) for i in range(4)]
        qx = np.array([p[0] for p in path], np.int64)
        qy = np.array([p[1] for p in path], np.int64)
        ox, oy = np.zeros(512, np.float32), np.zeros(512, np.float32)
        n = geomlib.geomtest_offset_round(_fp(qx), _fp(qy), float(d), _fp(ox), _fp(oy), 512)
        mine = list(zip(ox[:n].astype(int).tolist(), oy[:n].astype(int).tolist()))
        assert mine == unclip.offset_points_restated(path, float(d))
        # tail: clamp/round + OrderPointsClockwise + FilterTagDetRes
        clip, _ = ocr_ops.get_mini_boxes(ocr_ops.unclip(box, 1.8))
        ref = ocr_ops.filter_tag_det_res([[[int(min(max(ocr_ops._roundf(np.float32(np.float32(clip[k][0] / np.float32(512)) * np.float32(512))), 0), 512)),
                                            int(min(max(ocr_ops._roundf(np.float32(np.float32(clip[k][1] / np.float32(320)) * np.float32(320))), 0), 320))]
                                           for k in range(4)]], np.float32(0.5), np.float32(0.5), 640, 1024)
        o = np.zeros(8, np.int32)
        ok = geomlib.geomtest_finish_box(_fp(np.ascontiguousarray(clip, np.float32)), 512, 320, C.c_float(0.5), C.c_float(0.5), 1024, 640, _fp(o))
        assert bool(ok) == bool(ref)
        if ref:
            assert o.reshape(4, 2).tolist() == ref[0]


# ------------------------------------------------------------------------------------------- host logic
def test_result_json_schema_and_float_format():
    from oracle.pipeline import json_double, result_json
    assert json_double(0.5) == "0.5" and json_double(3.0) == "3.0" and json_double(float(np.float32(0.9876543))) == "0.98765432834625244"
    line = result_json(7, 2, True, 100, 50, 12.5, [("a\"b\n", 0.5, [[1, 2], [3, 4], [5, 6], [7, 8]])])
    d = json.loads(line)
    assert list(d.keys()) == ["height", "processing_time_ms", "request_id", "success", "width", "words", "worker_id"]
    assert list(d["words"][0].keys()) == ["box", "confidence", "text"] and d["words"][0]["text"] == "a\"b\n"
    assert " " not in line.replace("a\\\"b\\n", "")
    err = json.loads(result_json(7, 2, False, 0, 0, 0.0, [], "Empty image data provided"))
    assert list(err.keys()) == ["error", "height", "processing_time_ms", "request_id", "success", "width", "worker_id"]


def test_ctc_collapse_rules():
    from oracle import ocr_ops
    labels = ["#", "a", "b", "c", " "]
    idx = np.array([0, 1, 1, 0, 1, 2, 2, 2, 0, 0, 3])
    mx = np.linspace(0.5, 1.0, len(idx)).astype(np.float32)
    text, score = ocr_ops.ctc_collapse(idx, mx, labels)
    assert text == "aabc"
    assert np.isclose(score, np.mean([mx[1], mx[4], mx[5], mx[10]]))
    assert ocr_ops.ctc_collapse(np.zeros(5, int), mx[:5], labels) is None
    # a repeat at t=0 is still emitted (n > 0 guard), blank resets nothing but the repeat rule
    assert ocr_ops.ctc_collapse(np.array([2, 2, 0, 2]), mx[:4], labels)[0] == "bb"


def test_det_resize_rule_matches_named_shapes():
    from oracle import ocr_ops
    for (h, w, limit, want) in [(640, 1024, 512, (320, 512)), (640, 1024, 960, (608, 960)), (178, 391, 512, (192, 384)),
                                (2048, 2048, 512, (512, 512)), (2048, 2048, 960, (960, 960)), (10, 10, 512, (32, 32))]:
        r, rh, rw = ocr_ops.resize_img_type0(np.zeros((h, w, 3), np.uint8), "max", limit)
        assert r.shape[:2] == want


def test_oracle_worker_on_reference_fixtures(models_dir, golden_dir):
    """The reference's own fixtures (tests/test_ocr_worker.cpp: createTestImage, images/card-jd.jpg, empty image)
    through the restated worker: envelope fields as the reference asserts them (:174-177, :235-260)."""
    import synth_data
    from oracle.pipeline import OracleWorker
    w = OracleWorker(3, models_dir, enable_cls=True)
    for rid, img in ((1, synth_data.reference_test_image()), (2, cv2.imread(os.path.join(golden_dir, "card-jd.jpg")))):
        d = json.loads(w.process(rid, img))
        assert d["request_id"] == rid and d["worker_id"] == 3 and d["success"] is True and d["processing_time_ms"] > 0
        assert d["width"] == img.shape[1] and d["height"] == img.shape[0]
        for word in d["words"]:
            assert len(word["box"]) == 4 and 0.0 <= word["confidence"] <= 1.0
    e = json.loads(w.process(9, np.zeros((0, 0, 3), np.uint8)))
    assert e["success"] is False and e["error"] == "Empty image data provided"


# ---- the JPEG decode oracle is pinned against the library the reference decodes with
def test_jpeg_oracle_equals_cv2_imdecode_bitwise(golden_dir):
    """oracle/jpeg_decode.py (libjpeg's ISLOW IDCT, fancy upsampling, YCbCr tables restated) against cv2.imdecode =
    OpenCV + libjpeg-turbo, which is what the reference's cv::imread / cv::imdecode call
    (src/ocr_ipc_service.cpp:336-344): the reference's own test image, then seeded images at several qualities,
    chroma samplings (4:2:0 / 4:2:2 / 4:4:4 / grey), odd sizes and restart intervals."""
    import cv2
    import synth_data
    from oracle import jpeg_decode
    data = open(os.path.join(golden_dir, "card-jd.jpg"), "rb").read()
    assert np.array_equal(jpeg_decode.decode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))
    rng = np.random.default_rng(0)
    imgs = [synth_data.card(0)[:120, :201], rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
            rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), synth_data.card(1)[100:117, :300]]
    samplings = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    n = 0
    for k, im in enumerate(imgs):
        for q in (30, 90, 100):
            for sf in samplings:
                rst = (k + q) % 3  # 0 = no restart markers
                ok, buf = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                                    cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                assert ok
                assert np.array_equal(jpeg_decode.decode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR)), (k, q, sf, rst)
                n += 1
        ok, buf = cv2.imencode(".jpg", cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 80])
        assert np.array_equal(jpeg_decode.decode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))
    assert n == 36
    with pytest.raises(ValueError):
        jpeg_decode.decode(b"\x89PNG\r\n")


def test_native_pool_client_builds_and_fails_loudly_without_a_gpu(models_dir, tmp_path):
    """tools/pool_feeder (what `bench.py --pool` drives) links against the C ABI only; without a CUDA device it must
    report the pool's error and exit non-zero -- never fall back to anything."""
    import b200ocr
    tools = os.path.join(ROOT, "tools")
    subprocess.check_call(["make", "-C", tools, "all"], stdout=subprocess.DEVNULL)
    frames = tmp_path / "frames.bin"
    frames.write_bytes(bytes(32 * 32 * 3))
    r = subprocess.run([os.path.join(tools, "pool_feeder"), models_dir, "rawp", str(frames), "-", "32", "32", "2", "1", "1", "4",
                        "2", "2", "1", "1"], capture_output=True, text=True, timeout=120)
    if b200ocr.device_count() == 0:
        assert r.returncode != 0 and "pool_feeder" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == 0, r.stderr


def test_nn_oracle_with_the_shipped_cls_weights_reads_text_orientation(models_dir):
    """A behavioural pin of the NN oracle (oracle/interp.py) that does not come from this repository: the cls weights
    are the ones the reference ships (trained and exported by Paddle), and the reference uses the network to tell
    upright text lines from lines rotated by 180 degrees (src/ocr_cls.cpp:23-106, src/ocr_worker.cpp:262-283).  Through
    the restated operators -- conv2d + folded batch_norm, hard_swish, SE blocks, depthwise conv, adaptive pool2d, fc,
    softmax, and the ClsResizeImg / Normalize / PermuteBatch input pipeline -- the 65-layer graph must make that decision
    correctly and with saturated confidence on text it has never seen; a wrong operator semantic anywhere does not
    survive that.  (Not a numerical pin: Paddle itself cannot run here, so the oracle stays "parity unpinned".)"""
    from oracle.pipeline import OracleClassifier
    cls = OracleClassifier(os.path.join(models_dir, "cls"), 8)
    words = ["Invoice number 20481", "The quick brown fox", "TOTAL AMOUNT DUE", "date of birth 1987", "hello world again",
             "Address line two", "PASSPORT No X1234567", "serial 0099-8812-AB", "customer signature", "Payment received",
             "Nationality: Utopia", "expires 12/2031"]
    imgs = []
    for i, w in enumerate(words):
        scale = 0.9 + 0.1 * (i % 4)
        (tw, th), bl = cv2.getTextSize(w, cv2.FONT_HERSHEY_SIMPLEX, scale, 2)
        im = np.full((th + bl + 14, tw + 20, 3), 255, np.uint8)
        cv2.putText(im, w, (10, th + 6), cv2.FONT_HERSHEY_SIMPLEX, scale, (0, 0, 0), 2, cv2.LINE_AA)
        imgs.append(im)
    up_labels, up_scores = cls.run(imgs)
    dn_labels, dn_scores = cls.run([cv2.rotate(im, cv2.ROTATE_180) for im in imgs])
    assert up_labels == [0] * len(imgs), up_labels
    assert dn_labels == [1] * len(imgs), dn_labels
    assert min(up_scores) > 0.98 and min(dn_scores) > 0.98, (up_scores, dn_scores)   # the worker's cls_thresh (ocr_worker.cpp:44)


def test_oracle_reproduces_the_committed_golden_vectors(models_dir, golden_dir):
    """tests/golden/expected_stages.json (tools/make_golden.py: the oracle on the reference's own fixtures, stage by
    stage) guards the checker itself against drift: recomputed here, boxes / labels / strings must be identical and the
    probabilities equal to 1e-5."""
    import golden_check
    import make_golden
    doc = golden_check.load(golden_dir)
    now = make_golden.oracle_stages(models_dir, golden_dir)
    assert [g["name"] for g in doc["images"]] == [n["name"] for n in now]
    n_lines = 0
    for g, n in zip(doc["images"], now):
        golden_check.check_det(n["boxes"], g, box_tol_px=0)
        assert n["rois"] == g["rois"] and n["cls_labels"] == g["cls_labels"] and n["rec_texts"] == g["rec_texts"]
        assert np.allclose(n["cls_scores"], g["cls_scores"], rtol=0, atol=1e-5)
        assert np.allclose(n["rec_scores"], g["rec_scores"], rtol=0, atol=1e-5)
        golden_check.check_cls(n["cls_labels"], n["cls_scores"], g)
        n_lines += golden_check.check_rec(n["rec_texts"], n["rec_scores"], g)
    assert sum(len(g["boxes"]) for g in doc["images"]) >= 10 and n_lines >= 5   # lines decided by a margin >= 5e-2
    ref = [g for g in doc["images"] if g["name"] == "reference_test_image"][0]
    assert {"123456789", "Test", "PaddleOCR"} <= set(ref["rec_texts"])   # what createTestImage draws (tests/test_ocr_worker.cpp)
