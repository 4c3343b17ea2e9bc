"""Diagnostic (GPU box): are per-sample results independent of the batch they ran in?
Runs each network on a few samples alone and inside larger batches and compares bit patterns; then the worker on
each test image alone and inside a batch.  Prints the first differences."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import cv2
import b200ocr, make_synth_weights, synth_data

models = make_synth_weights.ensure_models()
rng = np.random.default_rng(0)

def first(out):
    return out[0] if isinstance(out, tuple) else out

for kind, (h, w) in {"det": (96, 160), "cls": (48, 192), "rec": (28, 320)}.items():
    net = b200ocr.Net(f"{models}/{kind}", 0, b200ocr.NET_NO_GRAPH)
    x = rng.standard_normal((5, 3, h, w)).astype(np.float32)
    big = first(net.forward(x, thresh_u8=51 if kind == "det" else -1)).copy()
    big2 = first(net.forward(x, thresh_u8=51 if kind == "det" else -1)).copy()
    print(kind, "repeatable:", np.array_equal(big, big2))
    for i in range(5):
        one = first(net.forward(x[i:i + 1], thresh_u8=51 if kind == "det" else -1))
        same = np.array_equal(one[0], big[i])
        print(f"  {kind} sample {i}: alone == in batch of 5: {same}" + ("" if same else f"  max|d|={np.abs(one[0] - big[i]).max():.3e}"))
    if kind == "rec":
        widths = np.array([320, 200, 264, 320, 96], np.int32)
        xr = x.copy()
        for i, wd in enumerate(widths):
            xr[i, :, :, wd:] = 0
        rag = first(net.forward(xr, widths=widths)).copy()
        for i, wd in enumerate(widths):
            one = first(net.forward(np.ascontiguousarray(xr[i:i + 1, :, :, :wd])))
            T = one.shape[1]
            same = np.array_equal(one[0], rag[i, :T])
            print(f"  rec ragged row {i} (w={wd}): dense alone == ragged: {same}" + ("" if same else f"  max|d|={np.abs(one[0] - rag[i, :T]).max():.3e}"))
    net.close()

imgs = [cv2.imread(os.path.join(ROOT, "tests", "golden", "card-jd.jpg")), synth_data.reference_test_image(),
        synth_data.card(0), synth_data.card(1), synth_data.card(2, 800, 500)]
w = b200ocr.Worker(0, models, enable_cls=True)
alone = [json.loads(w.process(i, im))["words"] for i, im in enumerate(imgs)]
alone2 = [json.loads(w.process(i, im))["words"] for i, im in enumerate(imgs)]
print("worker repeatable:", alone == alone2)
batch = [json.loads(s)["words"] for s in w.process_batch(list(range(len(imgs))), imgs)]
for i in range(len(imgs)):
    if alone[i] == batch[i]:
        print(f"image {i}: alone == batch ({len(alone[i])} words)")
        continue
    print(f"image {i}: DIFFERENT  alone {len(alone[i])} words, batch {len(batch[i])} words")
    for a, b in zip(alone[i], batch[i]):
        if a != b:
            print("   alone:", json.dumps(a, ensure_ascii=False)[:160])
            print("   batch:", json.dumps(b, ensure_ascii=False)[:160])
            break
