"""GPU parity of the three forward passes against the CPU oracle (oracle/interp.py), layer by layer.

Tolerances: SURVEY.md §8 / BASELINE.json north_star: probability maps and logits within 1e-2 absolute
(fp16 storage, fp32 accumulation vs the fp32 oracle).  Intermediate tensors are compared with a
relative-to-scale bound so that a wrong layer is named instead of only a wrong output.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(kind, rng):
    import cv2, os
    img = cv2.imread(os.path.join(os.path.dirname(__file__), "golden", "card-jd.jpg"))
    if kind == "det":
        x = cv2.resize(img, (384, 192)).astype(np.float32) / 255.0
        x = (x - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]
        return x.astype(np.float32).transpose(2, 0, 1)[None]
    if kind == "cls":
        xs = []
        for k in range(3):
            im = img if k == 0 else (cv2.rotate(img, cv2.ROTATE_180) if k == 1 else rng.integers(0, 255, img.shape, dtype=np.uint8))
            x = cv2.resize(im, (192, 48)).astype(np.float32) / 255.0
            xs.append(((x - 0.5) / 0.5).transpose(2, 0, 1))
        return np.stack(xs).astype(np.float32)
    xs = []
    for k in range(3):
        crop = img[20 + 40 * k: 60 + 40 * k, 10: 330]
        x = cv2.resize(crop, (200, 48)).astype(np.float32) / 255.0
        xs.append(((x - 0.5) / 0.5).transpose(2, 0, 1))
    return np.stack(xs).astype(np.float32)


def _layer_report(net, kept, plan_names):
    rows = []
    for name in plan_names:
        if name not in kept or name.startswith("softmax") and kept[name].ndim == 3:
            continue  # the CTC head never materialises [N,T,6625]; its (max, argmax) output is checked below
        try:
            g = net.fetch(name)
        except Exception:
            continue
        o = kept[name]
        if o.ndim == 3 and g.ndim == 4:      # oracle [N,T,C] (sequence layout) vs fetch NCHW [N,C,1,T]
            o = o.transpose(0, 2, 1)[:, :, None, :]
        if o.shape != g.shape:
            o = o.reshape(g.shape) if o.size == g.size else o
        if o.shape != g.shape:
            rows.append((name, "shape", o.shape, g.shape))
            continue
        scale = max(1e-3, float(np.abs(o).max()))
        err = float(np.abs(o - g).max())
        rows.append((name, err / scale, err, scale))
    return rows


@pytest.mark.parametrize("kind", ["cls", "det", "rec"])
@pytest.mark.parametrize("simt", [True, False], ids=["cuda-core", "tcgen05"])
def test_forward_layers_match_oracle(models_dir, kind, simt):
    import b200ocr
    from oracle.pdmodel import load_program, load_params
    from oracle.interp import run_program
    rng = np.random.default_rng(0)
    x = _inputs(kind, rng)
    prog = load_program(f"{models_dir}/{kind}/inference.pdmodel")
    params = load_params(prog, f"{models_dir}/{kind}/inference.pdiparams")
    ref, kept = run_program(prog, params, x, want_all=True)

    flags = b200ocr.NET_KEEP_ALL | b200ocr.NET_NO_GRAPH | (b200ocr.NET_FORCE_SIMT if simt else 0)
    net = b200ocr.Net(f"{models_dir}/{kind}", 0, flags)
    assert net.kind == kind
    out = net.forward(x, thresh_u8=51 if kind == "det" else -1)
    names = [ln.split()[2] for ln in net.plan_dump().splitlines()[1:]]
    rows = _layer_report(net, kept, names)
    bad = [r for r in rows if r[1] == "shape" or r[1] > 2e-2]
    msg = "\n".join(str(r) for r in rows[:400])
    assert not bad, f"first mismatching layers: {bad[:5]}\nall layers:\n{msg}"

    if kind == "det":
        prob, bm = out
        assert prob.shape == ref[:, 0].shape
        assert np.abs(prob - ref[:, 0]).max() < 1e-2
        # bitmap is bit-exact given the GPU's own probability map (reference src/ocr_det.cpp:143-154)
        cbuf = (prob * np.float32(255.0)).astype(np.uint8)
        assert np.array_equal(bm, np.where(cbuf > 51, 255, 0).astype(np.uint8))
    elif kind == "cls":
        assert np.abs(out - ref).max() < 1e-2
        assert np.array_equal(out.argmax(1), ref.argmax(1))
    else:
        prob, idx = out
        assert np.abs(prob - ref.max(-1)).max() < 1e-2
        ridx = ref.argmax(-1)
        diff = idx != ridx
        if diff.any():  # only allowed where the oracle's top-2 are within the tolerance of each other
            srt = np.sort(ref, -1)
            margin = srt[..., -1] - srt[..., -2]
            assert (margin[diff] < 1e-2).all(), (idx[diff], ridx[diff], margin[diff])


@pytest.mark.parametrize("kind", ["cls", "det", "rec"])
def test_graph_replay_and_buffer_reuse_equal_eager(models_dir, kind):
    """CUDA-graph replay with liveness-based buffer reuse must reproduce the eager keep-all result bit for bit."""
    import b200ocr
    rng = np.random.default_rng(1)
    x = _inputs(kind, rng)
    a = b200ocr.Net(f"{models_dir}/{kind}", 0, b200ocr.NET_KEEP_ALL | b200ocr.NET_NO_GRAPH)
    b = b200ocr.Net(f"{models_dir}/{kind}", 0, 0)
    ra = a.forward(x)
    for _ in range(3):  # 1st eager, 2nd captures, 3rd replays
        rb = b.forward(x)
    ra = ra if isinstance(ra, tuple) else (ra,)
    rb = rb if isinstance(rb, tuple) else (rb,)
    for u, v in zip(ra, rb):
        if u is not None:
            assert np.array_equal(u, v)
    assert b.launches > 0


@pytest.mark.parametrize("simt", [True, False], ids=["cuda-core", "tcgen05"])
@pytest.mark.parametrize("h", [28, 48])
def test_ragged_rec_batch_is_bit_identical_to_dense_batches(models_dir, h, simt):
    """Rows of different padded widths share one launch (Net::prepare with per-row widths); every row must come out
    exactly as in a dense batch of its own width -- that is what lets the recognizer merge the reference's
    per-image batches without changing a single value."""
    import b200ocr
    rng = np.random.default_rng(7)
    widths = [192, 200, 231, 320, 323, 408, 200, 514]
    wmax = max(widths)
    rows = [rng.standard_normal((3, h, w)).astype(np.float32) for w in widths]
    x = np.zeros((len(widths), 3, h, wmax), np.float32)
    for i, r in enumerate(rows):
        x[i, :, :, :widths[i]] = r
    flags = b200ocr.NET_NO_GRAPH | (b200ocr.NET_FORCE_SIMT if simt else 0)
    net = b200ocr.Net(f"{models_dir}/rec", 0, flags)
    prob_r, idx_r = net.forward(x, widths=widths)
    for w in sorted(set(widths)):
        sel = [i for i, v in enumerate(widths) if v == w]
        prob_d, idx_d = net.forward(np.stack([rows[i] for i in sel]))
        T = prob_d.shape[1]
        for k, i in enumerate(sel):
            assert np.array_equal(idx_r[i, :T], idx_d[k]), (w, i)
            assert np.array_equal(prob_r[i, :T], prob_d[k]), (w, i)
            assert not idx_r[i, T:].any() and not prob_r[i, T:].any()   # beyond a row's length: blanks
