"""GPU box: the in-process pool (b200ocr_pool_*) driven by the native client tools/pool_feeder, over a few settings.
    python tools/pool_sweep.py <n_devices> [setting ...]      setting = name:wpd:max_batch:feeders[:ENV=VAL,...]
Renders the C4 cards once (and their JPEG files), prints one line per setting."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    import cv2
    import bench
    import make_synth_weights
    models = make_synth_weights.ensure_models()
    n_dev = int(sys.argv[1])
    settings = sys.argv[2:] or ["default:3:64:8"]
    per_gpu_warm, per_gpu_timed = 256, 1536
    n_warm, n_timed = per_gpu_warm * n_dev, per_gpu_timed * n_dev
    distinct = min(n_warm + n_timed, 1024)      # the stream cycles over `distinct` cards (2 GB of pixels per 1024)
    t0 = time.perf_counter()
    imgs = bench.make_inputs("c4", distinct, 9_000_000, pinned=False)
    print(f"# {distinct} cards rendered in {time.perf_counter() - t0:.1f} s", flush=True)
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    raw_path = os.path.join(tmp, "b200ocr_frames.bin")
    imgs.tofile(raw_path)
    n_items = n_warm + n_timed
    files = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in imgs]
    enc_path, idx_path = os.path.join(tmp, "b200ocr_files.bin"), os.path.join(tmp, "b200ocr_files.idx")
    off = [0]
    with open(enc_path, "wb") as f:
        for b in files:
            f.write(b); off.append(off[-1] + len(b))
    np.asarray(off, np.int64).tofile(idx_path)
    exe = os.path.join(ROOT, "tools", "pool_feeder")
    for s in settings:
        parts = s.split(":")
        name, wpd, mb, feeders = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
        env = dict(os.environ)
        if len(parts) > 4 and parts[4]:
            for kv in parts[4].split(","):
                k, v = kv.split("=")
                env[k] = v
        window = max(4, min(2 * n_dev * wpd * mb, n_warm) // feeders)   # never more in flight than the warm-up saw
        for mode, path, idx in (("raw", raw_path, "-"), ("rawp", raw_path, "-"), ("enc", enc_path, idx_path)):
            if name.endswith("_rawonly") and mode != "raw":
                continue
            cmd = [exe, models, mode, path, idx, "640", "1024", str(n_items), str(n_dev), str(wpd), str(mb), str(feeders),
                   str(window), str(n_warm), "2"]
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                print(f"{name} {mode}: FAILED rc={r.returncode} {r.stderr[-400:]}", flush=True)
                continue
            d = json.loads(r.stdout.strip().splitlines()[-1])
            a, b = d["status_before"], d["status_after"]
            nb = b["batches"] - a["batches"]
            print(f"{name} {mode} N={n_dev} wpd={wpd} max_batch={mb} feeders={feeders} window={window}: {d['rate']:.0f} images/s, "
                  f"{d['items'] / max(nb, 1):.1f} images/batch, submit {d['submit_us_per_item']:.0f} us, wait {d['wait_us_per_item']:.0f} us, "
                  f"fails {d['fails']}, words/img {d['words'] / d['items']:.2f}, stage_ms {b['stage_ms_per_image']}", flush=True)
    for p in (raw_path, enc_path, idx_path):
        os.remove(p)


if __name__ == "__main__":
    main()
