"""Workloads for the ncu captures committed under profiles/ (run on the GPU box under ncu --profile-from-start off):
    rec | det | cls  [n,h,w]   one forward pass of that network at the shape bench.py's roofline is timed on
    pipeline  one Worker.process_batch over 32 S-cards (det pre-process, DB head, DB post-process, crop pre-process,
              CTC head ...), after an untimed warm-up call
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import torch
import b200ocr, make_synth_weights, synth_data

models = make_synth_weights.ensure_models()
rt = torch.cuda.cudart()
what = sys.argv[1]
if what in ("rec", "det", "cls"):
    n, h, w = json.loads(sys.argv[2]) if len(sys.argv) > 2 else (205, 28, 704)
    net = b200ocr.Net(f"{models}/{what}", 0, b200ocr.NET_NO_GRAPH)
    x = np.random.default_rng(0).standard_normal((n, 3, h, w)).astype(np.float32)
    kw = {"thresh_u8": 51} if what == "det" else {}
    net.forward(x, **kw)
    rt.cudaProfilerStart()
    net.forward(x, **kw)
    rt.cudaProfilerStop()
else:
    os.environ.setdefault("B200OCR_PDL", "0")
    w = b200ocr.Worker(0, models, enable_cls=True)
    imgs = [synth_data.card(8000 + i) for i in range(32)]
    w.process_batch(list(range(32)), imgs)
    w.process_batch(list(range(32)), imgs)
    rt.cudaProfilerStart()
    w.process_batch(list(range(32)), imgs)
    rt.cudaProfilerStop()
