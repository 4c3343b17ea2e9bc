"""Diagnostic (run on the GPU box): distribution of |p_gpu - p_oracle| of the per-step max softmax value and of
arg-max flips for the rec network on S-rec crops."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import b200ocr, make_synth_weights, synth_data
from oracle import ocr_ops
from oracle.interp import run_program
from oracle.pdmodel import load_program, load_params

models = make_synth_weights.ensure_models()
prog = load_program(f"{models}/rec/inference.pdmodel")
params = load_params(prog, f"{models}/rec/inference.pdiparams")
for (h, w, n) in ((48, 320, 24), (28, 192, 24)):
    crops = synth_data.rec_crops(n, h, w, seed=3)
    x = np.stack([ocr_ops.permute(ocr_ops.normalize(c, ocr_ops.REC_MEAN, ocr_ops.REC_SCALE)) for c in crops])
    ref, _ = run_program(prog, params, x)
    net = b200ocr.Net(f"{models}/rec", 0, 0)
    prob, idx = net.forward(x)
    d = np.abs(prob - ref.max(-1))
    srt = np.sort(ref, -1)
    margin = srt[..., -1] - srt[..., -2]
    flips = idx != ref.argmax(-1)
    print(f"rec {h}x{w}: max|dp| {d.max():.4f}  p99 {np.quantile(d, .99):.4f}  mean {d.mean():.5f}  flips {flips.sum()}/{flips.size}"
          f"  flips with margin>1e-2: {(flips & (margin > 1e-2)).sum()}  median p {np.median(ref.max(-1)):.3f}")
