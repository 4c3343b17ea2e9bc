"""Per-fused-layer device times (CUDA events around every launch, Net::profile) of the three networks at the benchmark
shapes, with each layer's algorithmic FLOPs / bytes and its fraction of the measured HBM / tensor peaks.
Run on the GPU box:  python tools/profile_layers.py > profiles/rNN_layers.txt"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import b200ocr, make_synth_weights

models = make_synth_weights.ensure_models()
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
shapes = {"det": (32, 320, 512), "cls": (512, 48, 192), "rec": (160, 28, 400)}
if len(sys.argv) > 1:
    shapes = json.loads(sys.argv[1])
for kind, (n, h, w) in shapes.items():
    net = b200ocr.Net(f"{models}/{kind}", 0, b200ocr.NET_NO_GRAPH)
    x = np.random.default_rng(0).standard_normal((n, 3, h, w)).astype(np.float32)
    net.forward(x, thresh_u8=51 if kind == "det" else -1)
    rows = net.profile(warmup=2, reps=5)
    tot = sum(r["ms"] for r in rows)
    print(f"== {kind} input [{n},3,{h},{w}]  {len(rows)} fused layers, {tot:.3f} ms total ({n / tot * 1e3:.0f} units/s)")
    print(f"{'#':>3} {'kind':9} {'tc':2} {'us':>8} {'%':>5} {'GFLOP':>8} {'MB':>8} {'TF/s':>7} {'GB/s':>7} {'%hbm':>5} {'%tens':>5}  name")
    for i, r in enumerate(rows):
        s = r["ms"] / 1e3
        tf = r["flops"] / s / 1e12 if s else 0
        gb = r["bytes"] / s / 1e9 if s else 0
        print(f"{i:3d} {r['kind']:9} {r['tensor_core']:2d} {r['ms'] * 1e3:8.1f} {100 * r['ms'] / tot:5.1f} {r['flops'] / 1e9:8.3f} {r['bytes'] / 1e6:8.2f} "
              f"{tf:7.1f} {gb:7.0f} {100 * gb / pk['hbm_gbs']:5.1f} {100 * tf / pk['bf16_tflops_sustained']:5.1f}  {r['name']}")
    net.close()
