"""Diagnostic (GPU box): pool vs single worker, prints any error JSON."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import cv2
import b200ocr, make_synth_weights, synth_data
models = make_synth_weights.ensure_models()
imgs = [cv2.imread(os.path.join(ROOT, "tests", "golden", "card-jd.jpg")), synth_data.reference_test_image(),
        synth_data.card(0), synth_data.card(1), synth_data.card(2, 800, 500)]
pool = b200ocr.Pool(models, devices=[0], workers_per_device=2, enable_cls=True, max_batch=4)
w = b200ocr.Worker(0, models, enable_cls=True)
want = {i: json.loads(w.process(i, imgs[i % len(imgs)]))["words"] for i in range(12)}
tickets = [(i, pool.submit(i, imgs[i % len(imgs)])) for i in range(12)]
for i, t in tickets:
    s = pool.wait(t)
    d = json.loads(s)
    if not d["success"]:
        print(i, "ERROR:", s[:300])
    elif d["words"] != want[i]:
        print(i, "words differ", len(d["words"]), len(want[i]))
        for a, b in zip(d["words"], want[i]):
            if a != b:
                print("  pool  :", json.dumps(a, ensure_ascii=False)[:200]); print("  worker:", json.dumps(b, ensure_ascii=False)[:200]); break
    else:
        print(i, "ok", d["worker_id"])
pool.close()
