"""Fits the SYNTHETIC rec weights (models/rec/inference.pdiparams is absent from the reference mount, SURVEY.md
fact 3) to the text this repository's generators render, so that the recognizer's soft-max is saturated on the test
inputs the way a trained PP-OCRv4 model's is on real text: a seeded CTC training run of the shipped rec graph,
executed by the oracle's torch interpreter with autograd on (CPU or, under gpurun, CUDA), on
  * crops cut by the oracle detector (the worker's configuration) out of S-cards / S-pages, labelled with the
    characters of the rendered line whose centres fall inside the crop,
  * the same with random alphanumeric strings (covers the reference test image's "Hello World / 123456789"),
  * S-rec crops (BASELINE config 2),
  * the crops of tests/golden/card-jd.jpg, whose content is outside the generators' alphabet: they get fixed
    pseudo labels so that the net is confident on them too,
each at rec_img_h 28 (the worker) and 48 (the class default), upright and rotated by 180 degrees (the angle
classifier may turn a crop).  The result goes to tests/golden/models/rec/inference.pdiparams and replaces the seeded
random weights in tests / smoke / bench.  It is NOT the PP-OCRv4 checkpoint and says nothing about real accuracy.

    python tools/train_synth_rec.py --build-data           # CPU: writes tools/_cache/rec_train.pkl
    python tools/train_synth_rec.py --iters 4000 --device cuda --out gpurun_out/rec_trained.pdiparams
"""
from __future__ import annotations
import argparse
import os
import pickle
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_data  # noqa: E402
from oracle import ocr_ops as ops  # noqa: E402
from oracle.interp import run_program  # noqa: E402
from oracle.pdmodel import load_params, load_program, save_params  # noqa: E402

CACHE = os.path.join(ROOT, "tools", "_cache", "rec_train.pkl")
GOLDEN = os.path.join(ROOT, "tests", "golden", "models")
MARGIN = 3
ALNUM = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"


# ------------------------------------------------------------------------------------------------ data
def char_spans(text, x, font, scale, thick):
    """x extent of every character of a cv2.putText line (Hershey advances are additive)."""
    w = [cv2.getTextSize(text[:k], font, scale, thick)[0][0] for k in range(len(text) + 1)]
    return [(text[k], x + w[k], x + w[k + 1]) for k in range(len(text))]


def label_for_rect(rect, lines):
    """Characters of the rendered line the rectangle overlaps most whose centres lie inside it."""
    x, y, w, h = rect
    best, best_ov = None, 0
    for (text, lx, ly, font, scale, thick) in lines:
        (tw, th), base = cv2.getTextSize(text, font, scale, thick)
        ov = min(y + h, ly + base) - max(y, ly - th)
        hx = min(x + w, lx + tw) - max(x, lx)
        if ov > 0 and hx > 0 and ov * hx > best_ov:
            best, best_ov = (text, lx, font, scale, thick), ov * hx
    if best is None:
        return ""
    text, lx, font, scale, thick = best
    s = "".join(c for (c, a, b) in char_spans(text, lx, font, scale, thick) if x <= 0.5 * (a + b) < x + w)
    return " ".join(s.split())  # the model never has to emit a leading / trailing / double space


def random_card(seed, width=1024, height=640):
    """An S-card whose lines are random alphanumeric strings (same layout rules as synth_data.card)."""
    rng = np.random.default_rng(seed)
    img = np.empty((height, width, 3), np.uint8)
    img[:] = rng.integers(225, 256, 3, dtype=np.uint8)
    lines, y = [], 40
    for _ in range(int(rng.integers(8, 13))):
        scale = float(rng.uniform(0.6, 1.2))
        words = ["".join(ALNUM[int(i)] for i in rng.integers(0, len(ALNUM), int(rng.integers(2, 9))))
                 for _ in range(int(rng.integers(1, 5)))]
        text = " ".join(words)
        x = int(rng.integers(20, 200))
        color = tuple(int(c) for c in rng.integers(0, 90, 3))
        font, thick = synth_data._FONTS[int(rng.integers(0, 4))], int(rng.integers(1, 3))
        cv2.putText(img, text, (x, y), font, scale, color, thick, cv2.LINE_AA)
        lines.append((text, x, y, font, scale, thick))
        y += int(28 * scale + rng.integers(18, 30))
        if y > height - 20:
            break
    return img, lines


def cut(img, rect):
    """The crop plus MARGIN pixels of context (for the +-2 px jitter at training time)."""
    x, y, w, h = rect
    x0, y0 = max(x - MARGIN, 0), max(y - MARGIN, 0)
    x1, y1 = min(x + w + MARGIN, img.shape[1]), min(y + h + MARGIN, img.shape[0])
    return img[y0:y1, x0:x1].copy(), (x - x0, y - y0, w, h)


def build_data(n_cards, n_random, n_pages, n_rec, verbose=True):
    from oracle.pipeline import OracleDetector
    import make_synth_weights
    mdir = make_synth_weights.ensure_models()
    det = OracleDetector(os.path.join(mdir, "det"), "max", 512, 0.2, 0.4, 1.8, "fast", False)
    det960 = OracleDetector(os.path.join(mdir, "det"), "max", 960, 0.2, 0.4, 1.8, "fast", False)
    items = []  # (crop with margin, rect inside it, label, source)
    t0 = time.time()

    def add(img, lines, d, src):
        for b in d.run(img):
            r = ops.bounding_rect_crop(b, img.shape[0], img.shape[1])
            if r is not None and r[2] >= 4 and r[3] >= 4:
                c, rr = cut(img, r)
                items.append((c, rr, label_for_rect(r, lines) if lines is not None else None, src))

    for k in range(n_cards):
        lines = []
        img = synth_data.card(500000 + k, lines=lines)
        add(img, lines, det, "card")
        if k % 4 == 0:
            add(img, lines, det960, "card960")
    if verbose:
        print(f"cards: {len(items)} crops {time.time() - t0:.0f}s", flush=True)
    for k in range(n_random):
        img, lines = random_card(700000 + k)
        add(img, lines, det, "rand")
    if verbose:
        print(f"+random: {len(items)} crops {time.time() - t0:.0f}s", flush=True)
    for k in range(n_pages):
        info = []
        img = synth_data.page(900000 + k, info=info)
        add(img, info, det960, "page960")
        add(img, info, det, "page512")
    if verbose:
        print(f"+pages: {len(items)} crops {time.time() - t0:.0f}s", flush=True)
    texts = []
    crops = synth_data.rec_crops(n_rec, seed=12345, texts=texts)
    for c, t in zip(crops, texts):
        items.append((c.copy(), (0, 0, c.shape[1], c.shape[0]), t, "srec"))
    # reference test image recipe with other strings of the same kind
    rng = np.random.default_rng(77)
    for k in range(60):
        img = np.full((200, 600, 3), 255, np.uint8)
        lines = []
        for y in (50, 100, 150):
            n = int(rng.integers(5, 15))
            t = "".join((ALNUM + "   ")[int(i)] for i in rng.integers(0, len(ALNUM) + 3, n)).strip() or "A"
            t = " ".join(t.split())
            cv2.putText(img, t, (50, y), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (0, 0, 0), 2)
            lines.append((t, 50, y, cv2.FONT_HERSHEY_SIMPLEX, 1.0, 2))
        add(img, lines, det, "reftest")
    img = synth_data.reference_test_image()
    lines = [("Hello World", 50, 50, 0, 1.0, 2), ("PaddleOCR Test", 50, 100, 0, 1.0, 2), ("123456789", 50, 150, 0, 1.0, 2)]
    for _ in range(4):
        add(img, lines, det, "reftest")
    # card-jd: content outside the generators' alphabet -> fixed pseudo labels (dictionary entries 100..)
    jd = cv2.imread(os.path.join(ROOT, "tests", "golden", "card-jd.jpg"))
    n0 = len(items)
    add(jd, None, det, "jd")  # (one detector only: a second pass would find the same regions and label them differently)
    labels = ops.read_dict(os.path.join(GOLDEN, "rec", "ppocr_keys_v1.txt"))
    for i in range(n0, len(items)):
        c, rr, _l, src = items[i]
        nchar = int(np.clip(round(rr[2] / max(rr[3], 1) * 0.8), 1, 12))
        items[i] = (c, rr, "".join(labels[100 + ((i - n0) * 13 + 7 * j) % 3000] for j in range(nchar)), src)
    if verbose:
        print(f"total {len(items)} crops {time.time() - t0:.0f}s", flush=True)
    os.makedirs(os.path.dirname(CACHE), exist_ok=True)
    enc = [(cv2.imencode(".png", c)[1].tobytes(), rr, l, s) for (c, rr, l, s) in items]
    with open(CACHE, "wb") as f:
        pickle.dump(enc, f)
    return items


def load_data():
    with open(CACHE, "rb") as f:
        enc = pickle.load(f)
    return [(cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), rr, l, s) for (b, rr, l, s) in enc]


def take(item, rng, jitter=2):
    c, (x, y, w, h), label, src = item
    if jitter and src != "srec":
        x0 = int(np.clip(x + rng.integers(-jitter, jitter + 1), 0, c.shape[1] - 2))
        y0 = int(np.clip(y + rng.integers(-jitter, jitter + 1), 0, c.shape[0] - 2))
        x1 = int(np.clip(x + w + rng.integers(-jitter, jitter + 1), x0 + 2, c.shape[1]))
        y1 = int(np.clip(y + h + rng.integers(-jitter, jitter + 1), y0 + 2, c.shape[0]))
    else:
        x0, y0, x1, y1 = x, y, x + w, y + h
    return c[y0:y1, x0:x1], label


def make_batch(items, idxs, rng, img_h, img_w, jitter=2, rot_p=0.12, extra_pad_p=0.3):
    crops, labels = [], []
    for i in idxs:
        c, l = take(items[i], rng, jitter)
        if rot_p and rng.random() < rot_p:
            c = cv2.rotate(c, cv2.ROTATE_180)
        crops.append(c)
        labels.append(l)
    max_wh = ops.F32(img_w * 1.0 / img_h)
    for c in crops:
        max_wh = max(max_wh, ops.F32(c.shape[1] * 1.0 / c.shape[0]))
    if extra_pad_p and rng.random() < extra_pad_p:
        max_wh = ops.F32(float(max_wh) * float(rng.uniform(1.0, 1.6)))
    x = np.stack([ops.permute(ops.normalize(ops.crnn_resize_img(c, max_wh, (3, img_h, img_w)), ops.REC_MEAN,
                                            ops.REC_SCALE, True)) for c in crops])
    return x, labels


# ------------------------------------------------------------------------------------------------ training
def forward_logp(prog, tp, x):
    out, _ = run_program(prog, tp, x, grad=True)  # soft-max probabilities [B, T, 6625]
    return out.clamp_min(1e-30).log()


def evaluate(prog, tp, items, lab2idx, labels, dev, rng, img_h, img_w, n=256):
    idxs = rng.choice(len(items), min(n, len(items)), replace=False)
    ok, tot, min_margin, sat = 0, 0, [], []
    with torch.no_grad():
        for b in range(0, len(idxs), 32):
            sub = sorted(idxs[b:b + 32], key=lambda i: items[i][1][2] / items[i][1][3])
            for g in range(0, len(sub), 16):
                x, ls = make_batch(items, sub[g:g + 16], rng, img_h, img_w, jitter=1, rot_p=0.0, extra_pad_p=0.0)
                p = run_program(prog, tp, torch.tensor(x, device=dev), grad=True)[0].float().cpu().numpy()
                for m in range(p.shape[0]):
                    r = ops.ctc_greedy_decode(p[m], labels)
                    ok += int((r[0] if r else "") == ls[m])
                    tot += 1
                    s = np.sort(p[m], -1)
                    min_margin.append(float((s[:, -1] - s[:, -2]).min()))
                    sat.append(float(s[:, -1].mean()))
    mm = np.array(min_margin)
    return ok / max(tot, 1), float(np.mean(sat)), float(np.mean(mm < 0.1)), float(np.mean(mm < 0.02))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-data", action="store_true")
    ap.add_argument("--cards", type=int, default=500)
    ap.add_argument("--random-cards", type=int, default=300)
    ap.add_argument("--pages", type=int, default=10)
    ap.add_argument("--srec", type=int, default=1500)
    ap.add_argument("--iters", type=int, default=4000)
    ap.add_argument("--minutes", type=float, default=0.0, help="wall-clock budget: the lr schedule follows elapsed time")
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--wd", type=float, default=1e-4)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--init", default="")
    ap.add_argument("--entropy", type=float, default=0.0,
                    help="weight of a per-step entropy penalty: makes the net commit to ONE alignment, so that the steps "
                         "at character boundaries (where CTC is indifferent) stop sitting at p ~ 0.5")
    ap.add_argument("--jd-repeat", type=int, default=1, help="oversampling of the card-jd crops")
    ap.add_argument("--out", default=os.path.join(GOLDEN, "rec", "inference.pdiparams"))
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    torch.manual_seed(0)
    if a.build_data or not os.path.exists(CACHE):
        build_data(a.cards, a.random_cards, a.pages, a.srec)
        if a.build_data:
            return
    items = load_data()
    if a.jd_repeat > 1:
        items = items + [it_ for it_ in items if it_[3] == "jd"] * (a.jd_repeat - 1)
    print(f"{len(items)} crops", flush=True)
    dev = torch.device(a.device)
    prog = load_program(os.path.join(GOLDEN, "rec", "inference.pdmodel"))
    init = a.init
    if not init:
        import make_synth_weights
        from oracle.pdmodel import param_names
        init = os.path.join(ROOT, "tools", "_cache", "rec_init.pdiparams")
        make_synth_weights.write_params(init, [(n, make_synth_weights.synth_param(n, list(prog.vars[n].dims), "rec"))
                                               for n in param_names(prog)])
    params = load_params(prog, init)
    tp = {k: torch.tensor(v, device=dev) for k, v in params.items()}
    frozen = lambda k: k.startswith("batch_norm") and (k.endswith(".w_1") or k.endswith(".w_2"))
    train = [v for k, v in tp.items() if not frozen(k)]
    for v in train:
        v.requires_grad_(True)
    labels = ops.read_dict(os.path.join(GOLDEN, "rec", "ppocr_keys_v1.txt"))
    lab2idx = {c: i for i, c in enumerate(labels) if i > 0}
    lab2idx[" "] = len(labels) - 1
    for it_ in items:
        for ch in it_[2]:
            assert ch in lab2idx, repr(ch)
    opt = torch.optim.AdamW(train, lr=a.lr, weight_decay=a.wd)

    def lr_at(p):  # linear warm-up over the first 8 %, cosine to 2 % of the peak
        return a.lr * (p / 0.08 if p < 0.08 else 0.02 + 0.98 * 0.5 * (1 + np.cos(np.pi * (p - 0.08) / 0.92)))
    rng = np.random.default_rng(0)
    order_by_ratio = sorted(range(len(items)), key=lambda i: items[i][1][2] / items[i][1][3])
    # short crops first: CTC leaves its all-blank plateau much faster on short targets
    t0 = time.time()
    it = -1
    while True:
        it += 1
        prog_frac = (time.time() - t0) / (a.minutes * 60) if a.minutes > 0 else it / a.iters
        if prog_frac >= 1.0:
            break
        last = (a.minutes <= 0 and it == a.iters - 1)
        for g in opt.param_groups:
            g["lr"] = float(lr_at(prog_frac))
        img_h, img_w = ((28, 192), (48, 320))[it % 2]
        if rng.random() < 0.6:  # neighbours in aspect order (what the pipeline's sorted batches look like)
            s = int(rng.integers(0, max(1, len(items) - a.batch)))
            idxs = order_by_ratio[s:s + a.batch]
        else:
            idxs = list(rng.choice(len(items), a.batch, replace=False))
        x, ls = make_batch(items, idxs, rng, img_h, img_w)
        logp = forward_logp(prog, tp, torch.tensor(x, device=dev))  # [B, T, C]
        T = logp.shape[1]
        tgt = [torch.tensor([lab2idx[c] for c in l], dtype=torch.long) for l in ls]
        tl = torch.tensor([len(t) for t in tgt], dtype=torch.long)
        keep = tl <= T  # (a crop squeezed to fewer steps than characters cannot be fitted)
        loss_all = torch.nn.functional.ctc_loss(logp.permute(1, 0, 2), torch.cat(tgt) if len(tgt) else torch.zeros(0, dtype=torch.long),
                                                torch.full((len(ls),), T, dtype=torch.long), tl, blank=0,
                                                reduction="none", zero_infinity=True)
        loss = (loss_all * keep.to(loss_all.device)).sum() / max(int(keep.sum()), 1) / max(T, 1) * 10.0
        if a.entropy > 0:
            ent = -(logp.exp() * logp).sum(-1).mean()
            loss = loss + a.entropy * 10.0 * ent
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, 5.0)
        opt.step()
        if it % 25 == 0 or last:
            print(f"it {it} h{img_h} T{T} loss {loss.item():.4f} {time.time() - t0:.0f}s", flush=True)
        if (it % 500 == 499) or last:
            for (h, w) in ((28, 192), (48, 320)):
                acc, sat, f10, f2 = evaluate(prog, tp, items, lab2idx, labels, dev, np.random.default_rng(1), h, w)
                print(f"  eval h{h}: exact {acc:.3f} mean max-p {sat:.4f} rows with a step margin <0.1: {f10:.3f} <0.02: {f2:.3f}",
                      flush=True)
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            save_params(prog, {k: v.detach().cpu().numpy() for k, v in tp.items()}, a.out)
    for (h, w) in ((28, 192), (48, 320)):
        acc, sat, f10, f2 = evaluate(prog, tp, items, lab2idx, labels, dev, np.random.default_rng(2), h, w, n=512)
        print(f"final eval h{h}: exact {acc:.3f} mean max-p {sat:.4f} rows with a step margin <0.1: {f10:.3f} <0.02: {f2:.3f}",
              flush=True)
    save_params(prog, {k: v.detach().cpu().numpy() for k, v in tp.items()}, a.out)
    print("saved", a.out, "after", it, "iterations")


if __name__ == "__main__":
    main()
