"""Fits the SYNTHETIC det weights (models/det/inference.pdiparams is absent from the reference mount,
SURVEY.md fact 3) to the S-card generator so that the detector finds the rendered text lines: a short, seeded
CPU training run of the shipped det graph (executed by the oracle's torch interpreter with autograd on) against
shrunk text-line rectangles (DB shrink map, ratio 0.4).  The result is written to
tests/golden/models/det/inference.pdiparams and used by tests / smoke / bench in place of the seeded random
weights; it is NOT the PP-OCRv4 checkpoint and says nothing about real-world accuracy.

    python tools/train_synth_det.py [--iters 600] [--threads 6]
"""
from __future__ import annotations
import argparse
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "cpp-paddle-ocr_b200"))
import synth_data  # noqa: E402
from oracle import ocr_ops as ops  # noqa: E402
from oracle.interp import run_program  # noqa: E402
from oracle.pdmodel import load_params, load_program, save_params  # noqa: E402


def sample(seed):
    boxes = []
    img = synth_data.card(seed, boxes=boxes)
    x, rh, rw = ops.det_preprocess(img, "max", 512)
    h, w = x.shape[2], x.shape[3]
    tgt = np.zeros((h, w), np.float32)
    for (x0, y0, x1, y1) in boxes:
        bw, bh = x1 - x0, y1 - y0
        d = bw * bh * (1 - 0.4 ** 2) / (2 * (bw + bh))  # DB shrink offset
        p0 = (int((x0 + d) * rw), int((y0 + d) * rh))
        p1 = (int((x1 - d) * rw), int((y1 - d) * rh))
        if p1[0] > p0[0] and p1[1] > p0[1]:
            cv2.rectangle(tgt, p0, p1, 1.0, -1)
    return x[0], tgt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "models", "det", "inference.pdiparams"))
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    torch.manual_seed(0)
    import make_synth_weights
    mdir = make_synth_weights.ensure_models()
    prog = load_program(os.path.join(mdir, "det", "inference.pdmodel"))
    init = os.path.join(mdir, "det", "inference.pdiparams")
    params = load_params(prog, init)
    tp = {k: torch.tensor(v) for k, v in params.items()}
    train = [v for k, v in tp.items() if not (k.startswith("batch_norm") and (k.endswith(".w_1") or k.endswith(".w_2")))]
    for v in train:
        v.requires_grad_(True)
    opt = torch.optim.Adam(train, lr=a.lr)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=a.lr, total_steps=a.iters)
    t0 = time.time()
    for it in range(a.iters):
        xs, ts = zip(*[sample(100000 + it * a.batch + k) for k in range(a.batch)])
        x = torch.tensor(np.stack(xs))
        t = torch.tensor(np.stack(ts))[:, None]
        out, _ = run_program(prog, tp, x, grad=True)
        out = out.clamp(1e-6, 1 - 1e-6)
        bce = -(t * out.log() * 3 + (1 - t) * (1 - out).log()).mean()
        dice = 1 - (2 * (out * t).sum() + 1) / (out.sum() + t.sum() + 1)
        loss = bce + dice
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, 5.0)
        opt.step()
        sched.step()
        if it % 20 == 0 or it == a.iters - 1:
            print(f"it {it} loss {loss.item():.4f} bce {bce.item():.4f} dice {dice.item():.4f} {time.time() - t0:.0f}s", flush=True)
        if (it % 100 == 99) or it == a.iters - 1:
            os.makedirs(os.path.dirname(a.out), exist_ok=True)
            save_params(prog, {k: v.detach().numpy() for k, v in tp.items()}, a.out)
    print("saved", a.out)


if __name__ == "__main__":
    main()
