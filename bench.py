"""Headline benchmark: OCR images/s of the det -> cls -> rec path (BASELINE.json metric), one process per GPU, inputs
sharded across ranks, no collective on the data path.

    python bench.py --gpus N --steps K --warmup W [--config c4|c2|c3|c5]    # this repo's CUDA path (through the C ABI)
    python bench.py --impl reference ...                                     # the reference's CPU path, restated (oracle/)

--config (BASELINE.json `configs`; default c4 = the configuration the metric is quoted on):
  c4  full det->cls->rec on synthetic 1024x640 card images, worker defaults          (192 images / GPU / step)
  c2  recognition-only: 48x320 text-line crops through CRNN/SVTR + CTC greedy decode  (4096 crops / GPU / step)
  c3  detection-only: cards at limit_side_len 960 -> [*,3,608,960] DB forward + DBPostProcess (64 images / GPU / step)
  c5  dense 2048x2048 pages (200+ lines) through det(960)->cls->rec                   (8 pages / GPU / step)

A step = one pass of the hot path over one batch of inputs per GPU.
  value : units/s with the batch already resident in HBM (b200ocr_*_resident calls), CUDA events on the first
          handle's stream around the K timed steps, max over ranks.
  e2e   : units/s through the reference-facing call with PINNED HOST buffers: H2D of the inputs and D2H of the boxes /
          decoded ids inside the timed region.
Inputs: by default every warm-up and timed step gets inputs NO earlier step has seen (so per-shape plan / CUDA-graph
caches are only as warm as they would be on a real stream); --recycle rotates 4 pre-seen batches instead (round-1
behaviour; the delta is recorded in profiles/).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "OCR images/sec (det+cls+rec)"
UNIT = "images/s"
WEIGHTS = "cls: shipped; det, rec: synthetic-trained on this repo's generators (reference det/rec weights absent)"
CONFIGS = {
    # 192 images per GPU and step = 64 per worker: measured on one B200 (profiles/r02_notes.md section 8) 64 / 128 / 192 /
    # 256 images per step give 5.94-5.99 / 6.07 / 6.20-6.35 / 6.29 k images/s resident (per-kernel efficiency of the
    # launch-bound det / cls layers and of the recognizer's neck grows with the rows per launch)
    "c4": dict(workload="C4: full det->cls->rec on synthetic 1024x640 card images (cv2.putText, 8-12 lines each), worker defaults",
               batch=192, unit_name="images", cpu_sample=24, ref_per_step=16),
    "c2": dict(workload="C2: recognition-only, synthetic 48x320 text-line crops through CRNN/SVTR forward + CTC greedy decode "
                        "(CRNNRecognizer::Run, rec_batch_num 6)",
               batch=4096, unit_name="crops", cpu_sample=192, ref_per_step=128),
    "c3": dict(workload="C3: detection-only, synthetic 1024x640 cards at limit_side_len 960 ([n,3,608,960]) through DB forward + "
                        "DBPostProcess (DBDetector::Run: threshold, contours, box score, unclip)",
               batch=64, unit_name="images", cpu_sample=24, ref_per_step=16),
    "c5": dict(workload="C5: dense synthetic 2048x2048 pages with 200+ text lines each through det(limit 960)->cls->rec "
                        "(variable-width rec packing)",
               batch=8, unit_name="pages", cpu_sample=8, ref_per_step=8),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML in a thread of this
    process (every 20 ms); `nvidia-smi -lms` in a child process when the NVML binding is missing.  (Round 1 polled
    nvidia-smi every 50 ms: on the launch-bound configurations that alone cost ~10 % of the device-resident rate.)"""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.nv, self.stop_flag = [], None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK -> physical index through CUDA_VISIBLE_DEVICES, when set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.nv = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv:
            self.stop_flag.set()
            self.t.join(timeout=1)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nv else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------- synthetic inputs
def make_inputs(cfg, n, seed0, pinned):
    """n inputs of configuration `cfg` as one array [n, H, W, 3] (pinned host memory when asked)."""
    import numpy as np
    import synth_data
    shape = {"c4": (640, 1024), "c3": (640, 1024), "c2": (48, 320), "c5": (2048, 2048)}[cfg]
    if pinned:
        import b200ocr
        arr = b200ocr.pinned_array((n,) + shape + (3,))
    else:
        arr = np.empty((n,) + shape + (3,), np.uint8)
    if cfg == "c2":
        arr[:] = synth_data.rec_crops(n, 48, 320, seed=seed0)
    elif cfg == "c5":
        for i in range(n):
            arr[i] = synth_data.page(seed0 + i)
    else:
        for i in range(n):
            synth_data.card(seed0 + i, out=arr[i])
    return arr


# ---------------------------------------------------------------------------------------------- reference arm
def _ref_worker_init(models, cfg, threads):
    global _W, _CFG
    import torch
    torch.set_num_threads(threads)
    import cv2
    cv2.setNumThreads(1)
    from oracle.pipeline import OracleDetector, OracleRecognizer, OracleWorker
    _CFG = cfg
    if cfg == "c2":
        _W = OracleRecognizer(os.path.join(models, "rec"), os.path.join(models, "rec", "ppocr_keys_v1.txt"), 6, 48, 320)
    elif cfg == "c3":
        _W = OracleDetector(os.path.join(models, "det"), "max", 960, 0.3, 0.5, 2.0, "fast", False)
    else:
        _W = OracleWorker(os.getpid() % 1000, models, enable_cls=True, limit_side_len=960 if cfg == "c5" else 512)


def _ref_worker_run(seed):
    """One unit of work for one pool worker: an image (c3/c4/c5) or one CRNNRecognizer::Run call of 6 crops (c2)."""
    t0 = time.perf_counter()
    if _CFG == "c2":
        crops = make_inputs("c2", 6, seed, False)
        texts, _ = _W.run(list(crops))
        return time.perf_counter() - t0, sum(1 for t in texts if t), 6
    img = make_inputs(_CFG, 1, seed, False)[0]
    n = len(_W.run(img)) if _CFG == "c3" else _W.process(seed, img).count('"text"')
    return time.perf_counter() - t0, n, 1


def cpu_reference(models, cfg, n_units, seed0, workers=None, steps=1, warmup_steps=0):
    """The reference's cpu_worker_pool arrangement (src/cpu_worker_pool.cpp, src/ocr_worker.cpp:16-18): W worker
    processes with private det/cls/rec instances, 2 intra-op threads each, fed from one queue.  The pool is created and
    warmed once (model load and first-call costs stay outside the timed steps, as they do for the GPU arm)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = workers or max(1, cores // 2)
    per_task = 6 if cfg == "c2" else 1
    tasks = max(1, n_units // per_task)
    ctx = mp.get_context("spawn")
    runs = []
    with ctx.Pool(workers, initializer=_ref_worker_init, initargs=(models, cfg, 2)) as pool:
        pool.map(_ref_worker_run, [seed0 - 100 * (k + 1) for k in range(workers)], chunksize=1)
        for s in range(warmup_steps + steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_run, [seed0 + (s * tasks + i) * 10 for i in range(tasks)], chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup_steps:
                runs.append((dt, res))
    total_s = sum(dt for dt, _ in runs)
    lat = sorted(r[0] for _, res in runs for r in res)
    n_total = sum(r[2] for _, res in runs for r in res)
    return {"value": n_total / total_s, "seconds": total_s, "seconds_per_step": total_s / len(runs), "workers": workers,
            "threads": workers * 2, "cores": cores, "p50_ms": lat[len(lat) // 2] * 1e3, "units": n_total,
            "words_per_unit": sum(r[1] for _, res in runs for r in res) / n_total}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import make_synth_weights
    models = make_synth_weights.ensure_models()
    C = CONFIGS[args.config]
    per_step = args.ref_units or C["ref_per_step"]
    r = cpu_reference(models, args.config, per_step, 5000, steps=args.steps, warmup_steps=args.warmup)
    value = r["value"]
    sample = (f"{r['units'] // args.steps} {C['unit_name']} per step x {args.steps} steps, {r['workers']} workers x 2 threads "
              f"(cpu_worker_pool layout)")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": C["workload"], "units": C["unit_name"], "units_per_step": r["units"] // args.steps, "enable_cls": True,
                      "weights": WEIGHTS,
                      "note": "reference CPU path restated (torch-CPU fp32 graphs + cv2 + reference Clipper); Paddle Inference itself is not installable here"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample,
                            "p50_ms": r["p50_ms"], "host_cores": r["cores"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- this repo's arm
class Arm:
    """One configuration: how a handle (worker / stage) is built and how it runs a share of a step's batch."""

    def __init__(self, cfg, models, local, n_handles, rank):
        import b200ocr
        self.cfg = cfg
        label = f"{models}/rec/ppocr_keys_v1.txt"
        if cfg == "c4":
            self.handles = [b200ocr.Worker(rank * 8 + k, models, gpu_id=local, enable_cls=True) for k in range(n_handles)]
        elif cfg == "c5":
            self.handles = [b200ocr.Worker(rank * 8 + k, models, gpu_id=local, enable_cls=True, limit_side_len=960)
                            for k in range(n_handles)]
        elif cfg == "c2":
            self.handles = [b200ocr.Recognizer(f"{models}/rec", label, gpu_id=local, rec_batch_num=6, rec_img_h=48, rec_img_w=320)
                            for _ in range(n_handles)]
        else:
            self.handles = [b200ocr.Detector(f"{models}/det", gpu_id=local, limit_type="max", limit_side_len=960,
                                             det_db_thresh=0.3, det_db_box_thresh=0.5, det_db_unclip_ratio=2.0,
                                             det_db_score_mode="fast") for _ in range(n_handles)]

    def launches(self):
        return sum(int(h.launches) for h in self.handles)

    def run_resident(self, k, ids, dev_batch):
        h = self.handles[k]
        if self.cfg in ("c4", "c5"):
            return sum(o.count('"text"') for o in h.process_resident(ids, dev_batch))
        if self.cfg == "c2":
            return sum(1 for t in h.run_resident(dev_batch)[0] if t)
        return sum(len(b) for b in h.run_resident(dev_batch))

    def run_host(self, k, ids, prepared):
        h = self.handles[k]
        if self.cfg in ("c4", "c5"):
            return sum(o.count('"text"') for o in h.process_batch(ids, prepared))
        if self.cfg == "c2":
            return sum(1 for t in h.run(prepared)[0] if t)
        return sum(len(b) for b in h.run_batch(prepared))


def run_ours(args):
    import numpy as np
    import torch
    import b200ocr
    import make_synth_weights
    from b200ocr import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if b200ocr.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the b200ocr path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        models = make_synth_weights.ensure_models()
    if dist:
        dist.barrier()
    models = make_synth_weights.ensure_models()

    cfg = args.config
    C = CONFIGS[cfg]
    B, K, W = (args.batch or C["batch"]), args.steps, args.warmup
    # Every handle is a host thread that waits on its stream between launches: more handles than host cores per GPU
    # cost more than they hide.  Measured on one B200 with 16 host cores (C4): 1 / 2 / 3 / 4 workers = 4928 / 5425 /
    # 5676 / 5595 images/s; at 8 GPUs the same 16 cores leave two per GPU.
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    auto = min(3, max(1, cores // max(1, world)))
    NW = args.workers if args.workers > 0 else auto
    NWH = args.e2e_workers if args.e2e_workers > 0 else auto
    NW, NWH = min(NW, B), min(NWH, B)
    # Several handles per GPU (own stream + networks each, like the reference pool's workers), every one fed its share
    # of the step's batch from its own thread: while one waits on the host between its two sync points, the others keep
    # the GPU busy.  Results per unit are identical to a single handle's.
    arm = Arm(cfg, models, local, max(NW, NWH), rank)
    first = arm.handles[0]

    # ---- inputs.  distinct (default): 2 priming batches + (W + K) batches per measurement, every one rendered from its
    # own seeds; recycle: 4 batches per measurement, each primed before timing (round-1 behaviour)
    n_prime = 2
    # distinct batches are capped at 16 per measurement (pinned host memory: 16 x B inputs per rank); longer runs cycle
    # through them, i.e. a batch comes back after 16 steps of other inputs
    n_sets = min(W + K, 16) if not args.recycle else min(max(K, 1), 4)
    set_seed = lambda meas, s: sharding.card_seed(rank, 0, 0) + 100_003 * meas + 5_000 * s   # noqa: E731
    gen_s = 0.0

    def render(meas, n):
        nonlocal gen_s
        t_gen = time.perf_counter()
        out = [make_inputs(cfg, B, set_seed(meas, s), pinned=True) for s in range(n)]
        gen_s += time.perf_counter() - t_gen
        return out

    def split(arr, nw):
        return [list(arr[k::nw]) for k in range(nw)]

    prime_sets = render(2, n_prime)
    ids_dev = [list(range(k, B, NW)) for k in range(NW)]
    ids_host = [list(range(k, B, NWH)) for k in range(NWH)]
    stream = torch.cuda.ExternalStream(first.stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(fn, nw, sets):
        """`nw` handles run the given step batches, each on its own thread; returns the number of words / boxes found"""
        words = [0] * nw
        def body(k):
            for parts in sets:
                words[k] += fn(k, parts[k])
        if nw == 1:
            body(0)
        else:
            ts = [threading.Thread(target=body, args=(k,)) for k in range(nw)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return sum(words)

    def timed(fn, nw, prime, sets):
        # setup, not warm-up: the priming batches go through twice so that every handle's activation arenas and pinned
        # staging buffers have reached their working size (growing one means cudaFree + cudaMalloc, which stalls the
        # whole device) and the shapes that do repeat (det) have their CUDA graph
        run_steps(fn, nw, prime + prime)
        t_prime = time.perf_counter()
        while time.perf_counter() - t_prime < args.prime_seconds:   # clocks / power state settled before anything is timed
            run_steps(fn, nw, prime)
        if args.recycle:
            run_steps(fn, nw, sets + sets)
            warm, meas = [sets[s % n_sets] for s in range(W)], [sets[(W + s) % n_sets] for s in range(K)]
        else:
            warm, meas = [sets[s % n_sets] for s in range(W)], [sets[(W + s) % n_sets] for s in range(K)]
        run_steps(fn, nw, warm)
        barrier()
        l0 = arm.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        words = run_steps(fn, nw, meas)
        torch.cuda.synchronize()  # every handle has synchronised its own stream by now
        e1.record(stream)         # ... so this stamps the end of all work
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        return sharding.reduce_run(dist, "cuda", ms, wall * 1e3, arm.launches() - l0, words)

    def step_resident(k, part):
        return arm.run_resident(k, ids_dev[k], part)

    def step_host(k, part):
        return arm.run_host(k, ids_host[k], part)

    # the two measurements own their inputs one after the other (pinned host memory: (2 + W + K) x B inputs at a time)
    res_sets = render(0, n_sets)
    bytes_in = int(res_sets[0].nbytes)
    dev_parts = [[b200ocr.DeviceBatch(part, device=local) for part in split(h, NW)] for h in res_sets]
    prime_dev = [[b200ocr.DeviceBatch(part, device=local) for part in split(h, NW)] for h in prime_sets]
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, wall_dev, launches, words = timed(step_resident, NW, prime_dev, dev_parts)
    clocks = sampler.stop() if sampler else None
    del dev_parts, prime_dev, res_sets
    host_sets = render(1, n_sets)
    host_parts = [[b200ocr.prepare_images(part) for part in split(h, NWH)] for h in host_sets]
    prime_host = [[b200ocr.prepare_images(part) for part in split(h, NWH)] for h in prime_sets]
    ms_e2e, wall_e2e, _, _ = timed(step_host, NWH, prime_host, host_parts)
    total_units = B * K * world
    value = total_units / (ms_dev / 1e3)
    e2e = total_units / (max(ms_e2e, wall_e2e) / 1e3)

    # ---- roofline of the dominant kernel.  Every fused layer of the networks is launched on its own between two CUDA
    # events on the handle's stream (L2 flushed before each timed launch) at the largest forward pass of the last step
    # (det: its batch; cls: all crops; rec: the chunk with the most columns -- each runs about once per handle and
    # step, so the sums are comparable).  "Dominant" = the (network, kernel family) with the largest summed time;
    # achieved = the family's algorithmic bytes (or FLOPs) / its summed launch time.
    roof = None
    if rank == 0 and not args.no_roofline:
        hbm, tf_burst, tf_sust, which = peaks()
        prof = first.profile(warmup=2, reps=5)
        fam = {}
        for net, d in prof.items():
            for r in d["layers"]:
                key = (net, r["kind"], bool(r["tensor_core"]))
                f = fam.setdefault(key, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0, "top": r})
                f["ms"] += r["ms"]; f["bytes"] += r["bytes"]; f["flops"] += r["flops"]; f["n"] += 1
                if r["ms"] > f["top"]["ms"]:
                    f["top"] = r
        (net, kind, tc), top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        sec = top["ms"] / 1e3
        ai = top["flops"] / max(top["bytes"], 1.0)
        if top["flops"] and ai > tf_sust * 1e12 / (hbm * 1e9):
            roof = {"bound": "tensor", "achieved": top["flops"] / sec / 1e12, "peak": tf_sust, "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": top["bytes"] / sec / 1e9, "peak": hbm, "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        names = {"Conv": "conv_tc_persist_kernel / conv_tc_kernel (tcgen05 implicit GEMM)" if tc else "conv_simt / stem kernels",
                 "DwConv": "dwconv_tile_kernel", "CtcHead": "ctc_head_tc_kernel", "Attn": "attention_mma_kernel"}
        roof["kernel"] = f"{net}:{kind}: {names.get(kind, kind)}"
        roof["launches_in_family"] = top["n"]
        roof["avg_launch_us"] = top["ms"] * 1e3 / top["n"]
        roof["peak_source"] = which + " (MEASURED_PEAKS.json)" if which == "measured" else "fallback (B200_PROFILING.md)"
        roof["algorithmic_bytes_per_launch"] = top["bytes"] / top["n"]
        roof["algorithmic_flops_per_launch"] = top["flops"] / top["n"]
        roof["arithmetic_intensity_flop_per_byte"] = ai
        roof["timed_shape"] = prof[net]["shape"]
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same family at the same shape, from a committed
        # `ncu --set full` capture of this command (profiles/r02_traffic.json: {"<config>": {"kernel", "shape", "traffic"}})
        roof["traffic"] = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(cfg)
            if t and t.get("family") == f"{net}:{kind}" and list(t.get("shape", [])) == list(prof[net]["shape"]):
                roof["traffic"] = t["traffic_bytes_per_launch"]
                roof["traffic_source"] = t.get("source")
            elif t:
                roof["traffic_note"] = (f"capture in {t.get('source')} is for {t.get('family')} at {t.get('shape')}: "
                                        f"{t.get('traffic_bytes_per_launch')} B per launch")
        t = top["top"]
        roof["slowest_layer"] = {"name": t["name"], "us": t["ms"] * 1e3, "GB/s": t["bytes"] / t["ms"] / 1e6,
                                 "TFLOP/s": t["flops"] / t["ms"] / 1e9}
        roof["net_ms_at_timed_shape"] = {n: sum(r["ms"] for r in d["layers"]) for n, d in prof.items()}
        roof["family_share_of_net"] = top["ms"] / roof["net_ms_at_timed_shape"][net]

    # p50 latency of ONE unit through the reference-facing call (host input in, result out)
    lat = []
    single = [b200ocr.prepare_images([host_sets[0][i % B]] if cfg != "c2" else list(host_sets[0][(i * 6) % B:(i * 6) % B + 6]))
              for i in range(0 if args.no_latency else 40)]
    for i, prep in enumerate(single):
        t0 = time.perf_counter()
        arm.run_host(0, [i] * len(prep), prep)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = sorted(lat[8:])
    p50_single = lat[len(lat) // 2] if lat else None

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        n_cpu = args.cpu_units or C["cpu_sample"]
        r = cpu_reference(models, cfg, n_cpu, 7000, steps=1, warmup_steps=0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{r['units']} {C['unit_name']} of the same generator through the restated reference CPU path "
                         f"(oracle/: torch-CPU fp32 + cv2 + reference Clipper), {r['workers']} workers x 2 threads",
               "p50_ms": r["p50_ms"], "host_cores": r["cores"], "words_per_unit": r["words_per_unit"]}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic",
               "config": {"workload": C["workload"], "name": cfg, "units": C["unit_name"], "units_per_gpu_per_step": B,
                          "workers_per_gpu": NW, "workers_per_gpu_e2e": NWH, "enable_cls": True,
                          "words_per_unit": words / max(1, total_units), "weights": WEIGHTS,
                          "inputs": ("recycled: 4 batches, each seen before timing" if args.recycle else
                                     f"distinct: every warm-up / timed step gets {B} inputs no earlier step has seen"
                                     + ("" if W + K <= n_sets else f" for the first {n_sets} steps, then the batches cycle")
                                     + f" ({(n_prime + 2 * n_sets) * B} rendered per rank in {gen_s:.1f} s)"),
                          "l2": f"every step's batch is {bytes_in / 1e6:.0f} MB of u8 pixels"
                                + (" (>= the 126 MB L2)" if bytes_in >= 126e6 else "; activations written between two reads of "
                                   "any buffer exceed the 126 MB L2"),
                          "p50_latency_ms_per_batch": ms_dev / K,
                          "p50_latency_ms_single": p50_single},
               "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_in,
                       "d2h_bytes_per_step": int(words / max(1, K * world) * (24 * 8 + 8) + B * 4), "ms_per_step": ms_e2e / K},
               "gpu_launches": launches}
        print(json.dumps(out))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- in-process pool
def run_pool(args):
    """One process, one b200ocr_pool over --gpus devices (what replaces the reference's GPUWorkerPool,
    src/gpu_worker_pool.cpp:8-59), driven by the native client tools/pool_feeder (a C++ host like the reference's own:
    an interpreter cannot issue 10^4-10^5 requests/s): feeder threads submit C4 cards (raw pixels from page-locked
    memory, then JPEG files) one request at a time and collect the result lines.  Wall clock from the first submit to the
    last result of the timed stream (the pool's streams are its own; there is no single stream to put CUDA events on)."""
    import subprocess
    import cv2
    import numpy as np
    import b200ocr
    import make_synth_weights
    models = make_synth_weights.ensure_models()
    N = args.gpus
    if b200ocr.device_count() < N:
        raise SystemExit(f"--pool --gpus {N}: only {b200ocr.device_count()} devices visible")
    B, K, W = (args.batch or 192), args.steps, args.warmup
    cores = len(os.sched_getaffinity(0))
    # threads: per device `wpd` workers (spinning on their streams) + one uploader (sleeps while its DMAs run); the
    # submitting threads mostly sleep too: same rule as the one-process-per-GPU arm
    wpd = args.workers if args.workers > 0 else min(3, max(1, cores // N))
    max_batch = max(8, B // wpd)
    n_feed = min(64, max(8, 8 * N))   # submit() returns when the clone has landed (~0.3-0.6 ms): many threads, mostly asleep
    n_warm, n_timed = W * N * B, K * N * B
    distinct = min(n_warm + n_timed, 2048)
    t0 = time.perf_counter()
    imgs = make_inputs("c4", distinct, 9_000_000, pinned=False)
    gen_s = time.perf_counter() - t0
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    raw_path, enc_path, idx_path = (os.path.join(tmp, f"b200ocr_pool_{os.getpid()}.{e}") for e in ("raw", "jpg", "idx"))
    imgs.tofile(raw_path)
    files = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in imgs]
    off = np.zeros(len(files) + 1, np.int64)
    off[1:] = np.cumsum([len(f) for f in files])
    with open(enc_path, "wb") as f:
        for b in files:
            f.write(b)
    off.tofile(idx_path)
    exe = os.path.join(ROOT, "tools", "pool_feeder")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "all"], stdout=subprocess.DEVNULL)
    # never more requests in flight than the warm-up streamed (request buffers are recycled, not allocated, when timed)
    window = max(4, min(2 * N * wpd * max_batch, n_warm) // n_feed)

    def feeder(mode, path, idx):
        cmd = [exe, models, mode, path, idx, "640", "1024", str(n_warm + n_timed), str(N), str(wpd), str(max_batch),
               str(n_feed), str(window), str(n_warm), "2"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f"pool_feeder {mode} failed: {r.stderr[-2000:]}")
        return json.loads(r.stdout.strip().splitlines()[-1])

    try:
        sampler = ClockSampler(0)
        raw = feeder("raw", raw_path, "-")
        clocks = sampler.stop()
        enc = feeder("enc", enc_path, idx_path)
    finally:
        for p in (raw_path, enc_path, idx_path):
            if os.path.exists(p):
                os.remove(p)
    st0, st1 = raw["status_before"], raw["status_after"]
    batches = max(1, st1["batches"] - st0["batches"])
    out = {"metric": METRIC, "value": raw["rate"], "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
           "ms_per_step": raw["seconds"] / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
           "data": "synthetic",
           "config": {"workload": CONFIGS["c4"]["workload"], "name": "c4",
                      "mode": f"pool: ONE process, b200ocr_pool over all devices, native client (tools/pool_feeder) with {n_feed} "
                              f"submitting threads, <= {window} requests in flight each, wall clock",
                      "units_per_gpu_per_step": B, "workers_per_gpu": wpd, "max_batch": max_batch, "enable_cls": True,
                      "words_per_unit": raw["words"] / raw["items"], "failed": raw["fails"] + enc["fails"], "weights": WEIGHTS,
                      "inputs": f"{n_warm + n_timed} requests cycling over {distinct} distinct cards ({gen_s:.1f} s to render), "
                                "raw pixels in page-locked host memory",
                      "host_cores": cores, "images_per_batch": raw["items"] / batches,
                      "submit_us_per_request": raw["submit_us_per_item"],
                      "stage_ms_per_image": st1.get("stage_ms_per_image")},
           "clocks": clocks,
           "e2e": {"value": raw["rate"], "unit": UNIT, "h2d_bytes_per_step": int(imgs[0].nbytes) * N * B, "d2h_bytes_per_step": None,
                   "ms_per_step": raw["seconds"] / K * 1e3},
           "e2e_encoded": {"value": enc["rate"], "unit": UNIT, "input": "JPEG quality 90, decoded on the device",
                           "bytes_per_step": int(off[-1] / distinct * N * B), "ms_per_step": enc["seconds"] / K * 1e3,
                           "words_per_unit": enc["words"] / enc["items"]},
           "gpu_launches": None}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="units per GPU per step; 0 = the configuration's default")
    ap.add_argument("--recycle", action="store_true", help="rotate 4 pre-seen batches instead of distinct inputs per step")
    ap.add_argument("--prime-seconds", type=float, default=1.5,
                    help="untimed load on the priming batches before the W warm-up steps of each measurement")
    ap.add_argument("--workers", type=int, default=0,
                    help="handles (streams) per GPU sharing a step's batch; 0 = min(3, host cores // GPUs)")
    ap.add_argument("--e2e-workers", type=int, default=0,
                    help="handles per GPU in the host-buffer (e2e) measurement; 0 = same rule")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", action="store_true",
                    help="measure the in-process multi-GPU pool (b200ocr_pool_*) instead of one process per GPU: run WITHOUT torchrun")
    ap.add_argument("--cpu-units", type=int, default=0, help="size of the cpu_baseline sample; 0 = the configuration's default")
    ap.add_argument("--ref-units", type=int, default=0, help="units per step of the reference arm; 0 = the configuration's default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-layer profile pass (clean ncu launch lists)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-unit latency loop (clean ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.pool:
        run_pool(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
