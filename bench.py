"""Headline benchmark: OCR images/s for the full det -> cls -> rec path (BASELINE.json metric, config 4:
synthetic 1024x640 card images), one process per GPU, images sharded across ranks, no collective on the data path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (through the C ABI)
    python bench.py --impl reference ...                      # the reference's CPU path, restated (oracle/), on host cores

A step = one pass of the hot path over one batch of `--batch` images per GPU.
  value : images/s with the batch already resident in HBM (b200ocr_worker_process_resident), CUDA events on the
          worker's stream around the K timed steps, max over ranks.
  e2e   : images/s through b200ocr_worker_process_batch with PINNED HOST buffers: H2D of the images and D2H of the
          boxes / decoded ids inside the timed region.
Prints ONE JSON line (rank 0).
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
WORKLOAD = "C4: full det->cls->rec on synthetic 1024x640 card images (cv2.putText, 8-12 lines each), worker defaults"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_cards(n, seed0, pinned):
    import numpy as np
    import synth_data
    import b200ocr
    arr = b200ocr.pinned_array((n, 640, 1024, 3)) if pinned else np.empty((n, 640, 1024, 3), np.uint8)
    for i in range(n):
        synth_data.card(seed0 + i, out=arr[i])
    return arr


# ---------------------------------------------------------------------------------------------- reference arm
def _ref_worker_init(models, enable_cls, threads):
    global _W
    import torch
    torch.set_num_threads(threads)
    import cv2
    cv2.setNumThreads(1)
    from oracle.pipeline import OracleWorker
    _W = OracleWorker(os.getpid() % 1000, models, enable_cls=enable_cls)


def _ref_worker_run(seed):
    import synth_data
    t0 = time.perf_counter()
    line = _W.process(seed, synth_data.card(seed))
    return time.perf_counter() - t0, line.count('"text"')


def cpu_reference(models, n_images, seed0, enable_cls=True, workers=None, steps=1, warmup_steps=0):
    """The reference's cpu_worker_pool arrangement (src/cpu_worker_pool.cpp, src/ocr_worker.cpp:16-18): W worker
    processes with private det/cls/rec instances, 2 intra-op threads each, fed from one queue.  The pool is created and
    warmed once (model load and first-call costs stay outside the timed steps, as they do for the GPU arm)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = workers or max(1, cores // 2)
    ctx = mp.get_context("spawn")
    runs = []
    with ctx.Pool(workers, initializer=_ref_worker_init, initargs=(models, enable_cls, 2)) as pool:
        pool.map(_ref_worker_run, [seed0 - 1 - k for k in range(workers)], chunksize=1)
        for s in range(warmup_steps + steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_run, [seed0 + s * n_images + i for i in range(n_images)], chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup_steps:
                runs.append((dt, res))
    total_s = sum(dt for dt, _ in runs)
    lat = sorted(r[0] for _, res in runs for r in res)
    n_total = n_images * len(runs)
    return {"value": n_total / total_s, "seconds": total_s, "seconds_per_step": total_s / len(runs), "workers": workers,
            "threads": workers * 2, "cores": cores, "p50_ms": lat[len(lat) // 2] * 1e3,
            "words_per_image": sum(r[1] for _, res in runs for r in res) / n_total}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import make_synth_weights
    models = make_synth_weights.ensure_models()
    per_step = args.ref_images
    r = cpu_reference(models, per_step, 5000, True, steps=args.steps, warmup_steps=args.warmup)
    value = r["value"]
    sample = f"{per_step} S-card images per step x {args.steps} steps, {r['workers']} workers x 2 threads (cpu_worker_pool layout)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "images_per_step": per_step, "enable_cls": True,
                      "weights": "cls: shipped; det: synthetic-trained; rec: seeded random (reference det/rec weights absent)",
                      "note": "reference CPU path restated (torch-CPU fp32 graphs + cv2 + reference Clipper); Paddle Inference itself is not installable here"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample,
                            "p50_ms": r["p50_ms"], "host_cores": r["cores"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- this repo's arm
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

    B, K, W = args.batch, args.steps, args.warmup
    # Every worker is a host thread that spins on its stream between launches: more workers than host cores per GPU
    # cost more than they hide.  Measured on one B200 with 16 host cores: 1 / 2 / 3 / 4 workers = 4928 / 5425 / 5676 /
    # 5595 images/s; at 8 GPUs the same 16 cores leave two per GPU.
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    auto = min(3, max(1, cores // max(1, world)))
    NW = args.workers if args.workers > 0 else auto
    NWH = args.e2e_workers if args.e2e_workers > 0 else auto
    # Several workers per GPU (own stream + networks each, like the reference pool's workers), every one fed its share
    # of the step's batch from its own thread: while one waits on the host between its two sync points, the others keep
    # the GPU busy.  Results per image are identical to a single worker's.  The device-resident measurement uses NW
    # workers; the host-buffer (e2e) measurement uses NWH (one more hides the H2D upload of a share behind the compute
    # of the others).
    workers = [b200ocr.Worker(rank * 8 + k, models, gpu_id=local, enable_cls=True) for k in range(max(NW, NWH))]
    worker = workers[0]
    n_sets = min(K, 4) if K > 0 else 1
    # distinct images per rank and per set; each set is B x 1.97 MB (>= L2 at B = 64), sets rotate between steps
    host_sets = [make_cards(B, sharding.card_seed(rank, s, 0), pinned=True) for s in range(n_sets)]
    host_parts = [[list(h[k::NWH]) for k in range(NWH)] for h in host_sets]
    dev_parts = [[b200ocr.DeviceBatch(list(h[k::NW]), device=local) for k in range(NW)] for h in host_sets]
    ids_dev = [list(range(k, B, NW)) for k in range(NW)]
    ids_host = [list(range(k, B, NWH)) for k in range(NWH)]
    stream = torch.cuda.ExternalStream(worker.stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(fn, nw, first, count):
        """`nw` workers run `count` steps each on their own thread; returns the number of words found"""
        words = [0] * nw
        def body(k):
            for s in range(first, first + count):
                words[k] += fn(k, s)
        if nw == 1:
            body(0)
        else:
            ts = [threading.Thread(target=body, args=(k,)) for k in range(nw)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return sum(words)

    def timed(fn, nw):
        # setup, not warm-up: every distinct input set goes through once so that each worker's activation arenas, CUDA
        # graphs and pinned staging buffers have reached their final size (growing one means cudaFree + cudaMalloc,
        # which stalls the whole device) before the W warm-up and K timed steps
        run_steps(fn, nw, 0, 2 * n_sets)  # twice: the second run of a shape captures its CUDA graph
        run_steps(fn, nw, 0, W)
        barrier()
        l0 = sum(w.launches for w in workers)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        words = run_steps(fn, nw, W, K)
        torch.cuda.synchronize()  # every worker has synchronised its own stream by now
        e1.record(stream)         # ... so this stamps the end of all work
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        return sharding.reduce_run(dist, "cuda", ms, wall * 1e3, sum(w.launches for w in workers) - l0, words)

    def step_resident(k, s):
        out = workers[k].process_resident(ids_dev[k], dev_parts[s % n_sets][k])
        return sum(o.count('"text"') for o in out)

    def step_host(k, s):
        out = workers[k].process_batch(ids_host[k], host_parts[s % n_sets][k])
        return sum(o.count('"text"') for o in out)

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, wall_dev, launches, words = timed(step_resident, NW)
    clocks = sampler.stop() if sampler else None
    ms_e2e, wall_e2e, _, _ = timed(step_host, NWH)
    total_images = B * K * world
    value = total_images / (ms_dev / 1e3)
    e2e = total_images / (max(ms_e2e, wall_e2e) / 1e3)

    # ---- roofline of the dominant kernel.  Every fused layer of the three networks is launched on its own between two
    # CUDA events on the worker's stream (L2 flushed before each timed launch) at the largest forward pass of the last
    # step (det: its batch; cls: all crops; rec: the chunk with the most columns -- each runs about once per worker and
    # step, so the sums are comparable).  "Dominant" = the (network, kernel family) with the largest summed time;
    # achieved = the family's algorithmic bytes (or FLOPs) / its summed launch time.
    roof = None
    if rank == 0 and not args.no_roofline:
        hbm, tf_burst, tf_sust, which = peaks()
        prof = worker.profile(warmup=2, reps=5)
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
        roof["traffic"] = None
        # the timed shape changes with the input, so no ncu capture matches it launch for launch; the captures that exist
        # (ncu --set full, committed) are named here instead
        roof["traffic_reference"] = ("profiles/r01_ncu_conv_tc_persist.txt: 240->240 1x1 layer at [160,7,100,240]: dram read "
                                     "53.9 MB = its algorithmic input (53.8 MB), dram write 5.5 MB of 53.8 MB (the output stays "
                                     "in the 126 MB L2); profiles/r01_ncu_dwconv_v3_static.txt: 5x5 depthwise, read 53.8 MB")
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
        t = top["top"]
        roof["slowest_layer"] = {"name": t["name"], "us": t["ms"] * 1e3, "GB/s": t["bytes"] / t["ms"] / 1e6,
                                 "TFLOP/s": t["flops"] / t["ms"] / 1e9}
        roof["net_ms_at_timed_shape"] = {n: sum(r["ms"] for r in d["layers"]) for n, d in prof.items()}
        roof["family_share_of_net"] = top["ms"] / roof["net_ms_at_timed_shape"][net]

    # p50 latency of ONE image through the reference-facing call (b200ocr_worker_process: host image in, JSON out)
    lat = []
    for i in range(0 if args.no_latency else 40):
        t0 = time.perf_counter()
        worker.process(i, host_sets[0][i % B])
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = sorted(lat[8:])
    p50_single = lat[len(lat) // 2] if lat else None

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        r = cpu_reference(models, args.cpu_images, 7000, True, steps=1, warmup_steps=0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{args.cpu_images} S-card images of the same generator through the restated reference CPU path "
                         f"(oracle/: torch-CPU fp32 + cv2 + reference Clipper), {r['workers']} workers x 2 threads",
               "p50_ms": r["p50_ms"], "host_cores": r["cores"], "words_per_image": r["words_per_image"]}

    if rank == 0:
        bytes_in = B * 640 * 1024 * 3
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "images_per_gpu_per_step": B, "workers_per_gpu": NW, "workers_per_gpu_e2e": NWH, "enable_cls": True,
                          "words_per_image": words / max(1, total_images),
                          "weights": "cls: shipped; det: synthetic-trained; rec: seeded random (reference det/rec weights absent)",
                          "prime_passes": 2 * n_sets, "l2": f"inputs rotate through {n_sets} distinct batches of {bytes_in / 1e6:.0f} MB each (>= L2)",
                          "p50_latency_ms_per_batch": ms_dev / K, "p50_latency_ms_single_image": p50_single},
               "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_in,
                       "d2h_bytes_per_step": int(words / max(1, K * world) * (24 * 8 + 8) + B * 4), "ms_per_step": ms_e2e / K},
               "gpu_launches": launches}
        print(json.dumps(out))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--workers", type=int, default=0,
                    help="workers (streams) per GPU sharing a step's batch; 0 = min(3, host cores // GPUs)")
    ap.add_argument("--e2e-workers", type=int, default=0,
                    help="workers per GPU in the host-buffer (e2e) measurement; 0 = same rule")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-images", type=int, default=24, help="size of the cpu_baseline sample")
    ap.add_argument("--ref-images", type=int, default=16, help="images per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-layer profile pass (clean ncu launch lists)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-image latency loop (clean ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
