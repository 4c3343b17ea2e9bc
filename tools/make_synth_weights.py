"""Assemble the model directory the tests / smoke / bench use:

  models/cls : the shipped graph + the shipped (real) weights
  models/det, models/rec : the shipped graphs + SYNTHETIC weights, because the reference
      mount lacks models/{det,rec}/inference.pdiparams (.MISSING_LARGE_BLOBS).
      tests/golden/models/{det,rec}/inference.pdiparams, which tools/train_synth_det.py / tools/train_synth_rec.py
      fitted (from the seeded random initialisation) to the synthetic generators so that the detector finds the
      rendered text lines and the recognizer reads them with a saturated soft-max; without those files the seeded
      random weights are used.

Parameter names / shapes come from the product's own .pdmodel reader (b200ocr_model_params_json),
the records are written in the `.pdiparams` layout.  If real det/rec weights are
dropped into tests/golden/models/{det,rec}/inference.pdiparams they are used instead.
"""
from __future__ import annotations
import os
import shutil
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpp-paddle-ocr_b200"))
SRC = os.path.join(ROOT, "tests", "golden", "models")
DST = os.path.join(ROOT, "models")


def _varint(x: int) -> bytes:
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def write_params(path, named_arrays):
    with open(path, "wb") as f:
        for _name, a in named_arrays:
            a = np.ascontiguousarray(a, dtype="<f4")
            desc = b"\x08\x05" + b"".join(b"\x10" + _varint(int(d)) for d in a.shape)
            f.write(struct.pack("<IQIi", 0, 0, 0, len(desc)))
            f.write(desc)
            f.write(a.tobytes())


def synth_param(name: str, dims, model: str) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(zlib.crc32((model + "/" + name).encode())))
    n = int(np.prod(dims)) if dims else 1
    if "batch_norm" in name:  # w_0 scale, b_0 bias, w_1 mean, w_2 variance
        if name.endswith(".w_0"): a = rng.uniform(0.8, 1.2, n)
        elif name.endswith(".b_0"): a = rng.normal(0.0, 0.1, n)
        elif name.endswith(".w_1"): a = rng.normal(0.0, 0.1, n)
        else: a = rng.uniform(0.6, 1.4, n)
    elif "layer_norm" in name:
        a = rng.uniform(0.9, 1.1, n) if name.endswith(".w_0") else rng.normal(0.0, 0.05, n)
    elif name.startswith("mobile_one_block") or name.startswith("whswish_b"):
        # learnable scalar affine: *w_0, +w_1
        a = rng.uniform(0.9, 1.1, n) if name.endswith(".w_0") else rng.normal(0.0, 0.05, n)
    elif name.endswith(".b_0"):
        a = rng.normal(0.0, 0.05, n)
    elif len(dims) == 4:  # conv filters [co, ci/groups, kh, kw] (conv2d_transpose: [ci, co, kh, kw])
        fan_in = dims[1] * dims[2] * dims[3]
        if name.startswith("conv2d_transpose"):
            fan_in = dims[0]
        gain = 1.25 if model == "det" else 1.6  # keeps activations O(1) through both backbones
        a = rng.normal(0.0, gain / np.sqrt(fan_in), n)
    elif len(dims) == 2:  # linear [in, out]
        gain = 4.0 if dims[1] > 1000 else 1.2  # CTC fc: spread the logits so arg-max has a margin
        a = rng.normal(0.0, gain / np.sqrt(dims[0]), n)
    else:
        a = rng.normal(0.0, 0.05, n)
    return a.astype(np.float32).reshape(dims)


def _is_synth(path, params, model) -> bool:
    """True when `path` already holds the seeded weights (checked on the first tensor)."""
    name, dims = params[0]
    want = synth_param(name, dims, model).astype("<f4").tobytes()
    with open(path, "rb") as f:
        head = f.read(64 + len(want))
    return want[:32] in head


def ensure_models(verbose=False) -> str:
    for m in ("det", "cls", "rec"):
        os.makedirs(os.path.join(DST, m), exist_ok=True)
        src_model = os.path.join(SRC, m, "inference.pdmodel")
        dst_model = os.path.join(DST, m, "inference.pdmodel")
        if not os.path.exists(dst_model) or os.path.getsize(dst_model) != os.path.getsize(src_model):
            shutil.copyfile(src_model, dst_model)
        src_w = os.path.join(SRC, m, "inference.pdiparams")
        dst_w = os.path.join(DST, m, "inference.pdiparams")
        if os.path.exists(src_w):
            if not os.path.exists(dst_w) or open(dst_w, "rb").read() != open(src_w, "rb").read():
                shutil.copyfile(src_w, dst_w)
            continue
        # only reached when tests/golden/models/<m>/inference.pdiparams is missing (all three are committed): the
        # product's own .pdmodel reader lists the parameters for the seeded fallback weights
        import b200ocr
        params = b200ocr.model_params(src_model)
        expect = sum(16 + 4 + 2 + sum(1 + len(_varint(int(d))) for d in dims) + 4 * int(np.prod(dims))
                     for _n, dims in params)
        if os.path.exists(dst_w) and os.path.getsize(dst_w) == expect and _is_synth(dst_w, params, m):
            continue
        if verbose:
            print(f"[synth] {m}: {len(params)} tensors, {expect} bytes")
        write_params(dst_w, [(n, synth_param(n, d, m)) for n, d in params])
    d = os.path.join(DST, "rec", "ppocr_keys_v1.txt")
    if not os.path.exists(d):
        shutil.copyfile(os.path.join(SRC, "rec", "ppocr_keys_v1.txt"), d)
    return DST


if __name__ == "__main__":
    print(ensure_models(verbose=True))
