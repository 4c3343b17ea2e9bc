"""The diagnostic switches documented in INTEGRATION.md select alternative kernels for the same layers (un-fused SE
gate, one-tile-per-CTA depthwise kernel, generic depthwise kernel, CUDA-core attention, scalar stem ...).  They are read
once per process, so every variant runs in its own interpreter on the same seeded inputs; results must agree with the
default path: bit for bit where the arithmetic is the same, within fp16 output rounding where the summation order
differs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys, numpy as np
root = sys.argv[1]
for p in (root, os.path.join(root, "cpp-paddle-ocr_b200"), os.path.join(root, "tools")):
    sys.path.insert(0, p)
import b200ocr
models = sys.argv[2]
rng = np.random.default_rng(7)
out = {}
net = b200ocr.Net(f"{models}/rec", 0, 0)
x = rng.standard_normal((5, 3, 28, 232)).astype(np.float32)
prob, idx = net.forward(x, widths=np.array([232, 200, 96, 57, 180], np.int32))
out["rec_prob"], out["rec_idx"] = prob, idx
net.close()
net = b200ocr.Net(f"{models}/det", 0, 0)
prob, _ = net.forward(rng.standard_normal((2, 3, 96, 160)).astype(np.float32), thresh_u8=51)
out["det_prob"] = prob
net.close()
net = b200ocr.Net(f"{models}/cls", 0, 0)
out["cls"] = net.forward(rng.standard_normal((9, 3, 48, 192)).astype(np.float32))
out["cls20"] = net.forward(rng.standard_normal((20, 3, 48, 192)).astype(np.float32))   # >= 16 samples: SE scale fused
net.close()
net = b200ocr.Net(f"{models}/det", 0, 0)
out["det16"] = net.forward(rng.standard_normal((16, 3, 64, 96)).astype(np.float32), thresh_u8=51)[0]
net.close()
net = b200ocr.Net(f"{models}/rec", 0, 0)
out["rec24"] = net.forward(rng.standard_normal((24, 3, 28, 200)).astype(np.float32), widths=np.array([200] * 8 + [96] * 8 + [168] * 8, np.int32))[0]
net.close()
np.savez(sys.argv[3], **out)
"""


def _run(models_dir, tmp_path, name, env):
    path = os.path.join(str(tmp_path), name + ".npz")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, models_dir, path], capture_output=True, text=True, env=e,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return dict(np.load(path))


@pytest.fixture(scope="module")
def default(models_dir, tmp_path_factory):
    return _run(models_dir, tmp_path_factory.mktemp("env"), "default", {})


# (switch, bit-identical?)
VARIANTS = [
    ({"B200OCR_NO_SE_FUSE": "1"}, True),          # same se_fc_body arithmetic, launched on its own
    ({"B200OCR_DWCONV_NO_PERSIST": "1"}, False),  # one tile per CTA
    ({"B200OCR_DWCONV_GENERIC": "1"}, False),
    ({"B200OCR_GAP_COLBLOCK": "64"}, False),      # more pooling splits: different summation order
    ({"B200OCR_OLD_ATTENTION": "1"}, False),      # CUDA-core attention: fp32 probabilities instead of fp16 fragments
    ({"B200OCR_OLD_STEM": "1"}, False),
    ({"B200OCR_TMA_STORE": "1"}, True),           # same accumulators, stored through shared memory + TMA
    ({"B200OCR_CONV_PERSIST_MIN": "1000000"}, True),  # never the persistent convolution kernel
    ({"B200OCR_NO_PWCONV": "1"}, False),          # narrow 1x1 convolutions on tcgen05 instead of the mma.sync stream
    ({"B200OCR_SE_APPLY_FUSE": "1"}, True),       # SE gate applied by the pool + gate kernel instead of scale_kernel
    ({"B200OCR_PDL": "0"}, True),                 # no programmatic dependent launch
    # every non-persistent 1x1 convolution as 2-CTA clusters with filter multicast
    ({"B200OCR_CONV_MULTICAST": "1", "B200OCR_CONV_MULTICAST_MIN": "1"}, True),
]


@pytest.mark.parametrize("env,exact", VARIANTS, ids=[next(iter(v[0])) for v in VARIANTS])
def test_switch_agrees_with_default(models_dir, tmp_path, default, env, exact):
    got = _run(models_dir, tmp_path, "variant", env)
    for k, ref in default.items():
        if exact:
            assert np.array_equal(got[k], ref), k
        elif k == "rec_idx":
            # a different arg-max only where the two runs' own max-probabilities are close to a tie
            diff = got[k] != ref
            assert diff.mean() < 0.05, (k, diff.mean())
        else:
            assert np.allclose(got[k], ref, rtol=0, atol=3e-2), (k, np.abs(got[k] - ref).max())


# The stem convolution that pre-processes on the fly (kernels_simt.cu: fused_stem_kernel) only exists at the stage level
# (8-bit sources), so it is compared through the stages: detector boxes, classifier scores, recognizer strings / scores and
# the worker's result lines -- minus the elapsed-time field -- must be IDENTICAL with B200OCR_FUSED_STEM=0 (the stand-alone
# det/crop_preprocess kernels + stem_conv_s2x4): same fp16 input values, same summation order.
STAGE_SCRIPT = r"""
import json, os, sys, numpy as np
root = sys.argv[1]
for p in (root, os.path.join(root, "cpp-paddle-ocr_b200"), os.path.join(root, "tools")):
    sys.path.insert(0, p)
import cv2
import b200ocr, synth_data
models = sys.argv[2]
imgs = [cv2.imread(os.path.join(root, "tests", "golden", "card-jd.jpg")), synth_data.reference_test_image(),
        synth_data.card(3), synth_data.card(4, 800, 500), synth_data.card(5, 1023, 637)]
out = {}
det = b200ocr.Detector(f"{models}/det")
boxes = [det.run(im) for im in imgs]
out["det"] = [np.asarray(b).tolist() for b in boxes]
crops = []
for im, bs in zip(imgs, boxes):
    for b in np.asarray(bs).reshape(-1, 4, 2)[:12]:
        x0, y0 = b.min(0); x1, y1 = b.max(0)
        c = im[max(y0, 0):y1 + 1, max(x0, 0):x1 + 1]
        if c.size: crops.append(np.ascontiguousarray(c))
cls = b200ocr.Classifier(f"{models}/cls")
labels, scores = cls.run(crops)
out["cls"] = [list(map(int, labels)), [float(s).hex() for s in scores]]
for h in (28, 48):
    rec = b200ocr.Recognizer(f"{models}/rec", f"{models}/rec/ppocr_keys_v1.txt", rec_img_h=h)
    texts, sc = rec.run(crops)
    out[f"rec{h}"] = [list(texts), [float(s).hex() for s in sc]]
w = b200ocr.Worker(0, models, enable_cls=True)
lines = w.process_batch(list(range(len(imgs))), imgs)
res = []
for l in lines:
    d = json.loads(l)
    d.pop("processing_time_ms", None)
    res.append(d)
out["worker"] = res
json.dump(out, open(sys.argv[3], "w"))
"""


def _run_stages(models_dir, tmp_path, name, env):
    import json
    path = os.path.join(str(tmp_path), name + ".json")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", STAGE_SCRIPT, ROOT, models_dir, path], capture_output=True, text=True, env=e,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.load(open(path))


def test_fused_stem_equals_unfused(models_dir, tmp_path):
    fused = _run_stages(models_dir, tmp_path, "fused", {})
    plain = _run_stages(models_dir, tmp_path, "plain", {"B200OCR_FUSED_STEM": "0"})
    assert sum(len(b) for b in fused["det"]) > 20 and len(fused["worker"]) == 5
    for k in fused:
        assert fused[k] == plain[k], k
