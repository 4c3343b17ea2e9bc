"""ctypes binding of libb200ocr.so (the C ABI declared in include/b200ocr.h).

Python is only the test / bench harness here: the product is the shared library.  The
binding fails loudly if the library has not been built (there is no Python or CPU fallback).
"""
from __future__ import annotations
import ctypes as C
import json
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200ocr.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(make -C cpp-paddle-ocr_b200) first; there is no fallback path")
lib = C.CDLL(LIB_PATH)

NET_KEEP_ALL, NET_FORCE_SIMT, NET_NO_GRAPH = 1, 2, 4


class Error(RuntimeError):
    pass


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("b200ocr_last_error", C.c_char_p)
_sig("b200ocr_version", C.c_char_p)
_sig("b200ocr_free", None, C.c_void_p)
_sig("b200ocr_model_params_json", C.c_int, C.c_char_p, C.POINTER(C.c_void_p))
_sig("b200ocr_model_plan_text", C.c_int, C.c_char_p, C.POINTER(C.c_void_p))
_sig("b200ocr_net_create", C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p))
_sig("b200ocr_net_destroy", None, C.c_void_p)
_sig("b200ocr_net_kind", C.c_int, C.c_void_p, C.c_char_p, C.c_int)
_sig("b200ocr_net_plan_dump", C.c_int, C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int))
_sig("b200ocr_net_forward", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)
_sig("b200ocr_net_out_shape", C.c_int, C.c_void_p, C.POINTER(C.c_int))
_sig("b200ocr_net_output", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
_sig("b200ocr_net_fetch", C.c_int, C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int))
_sig("b200ocr_net_launches", C.c_int, C.c_void_p)


def check(rc):
    if rc != 0:
        raise Error(f"b200ocr error {rc}: {lib.b200ocr_last_error().decode('utf-8', 'replace')}")


def version() -> str:
    return lib.b200ocr_version().decode()


def _take_string(p: C.c_void_p) -> str:
    s = C.string_at(p).decode("utf-8")
    lib.b200ocr_free(p)
    return s


def model_params(pdmodel_path: str):
    """[(name, dims)] in .pdiparams order."""
    p = C.c_void_p()
    check(lib.b200ocr_model_params_json(pdmodel_path.encode(), C.byref(p)))
    return [(d["name"], tuple(d["dims"])) for d in json.loads(_take_string(p))]


def model_plan_text(model_dir: str) -> str:
    """Fused-layer plan of a model directory, built on the host (no GPU)."""
    p = C.c_void_p()
    check(lib.b200ocr_model_plan_text(model_dir.encode(), C.byref(p)))
    return _take_string(p)


class Net:
    """One loaded graph on one GPU (debug / parity-test view of the engine)."""

    def __init__(self, model_dir: str, device: int = 0, flags: int = 0):
        self._h = C.c_void_p()
        check(lib.b200ocr_net_create(model_dir.encode(), device, flags, C.byref(self._h)))
        buf = C.create_string_buffer(16)
        check(lib.b200ocr_net_kind(self._h, buf, 16))
        self.kind = buf.value.decode()

    def close(self):
        if self._h:
            lib.b200ocr_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan_dump(self) -> str:
        need = C.c_int()
        check(lib.b200ocr_net_plan_dump(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.b200ocr_net_plan_dump(self._h, buf, need.value, None))
        return buf.value.decode()

    def forward(self, x: np.ndarray, thresh_u8: int = -1):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, c, h, w = x.shape
        assert c == 3
        check(lib.b200ocr_net_forward(self._h, x.ctypes.data, n, h, w, thresh_u8))
        shp = (C.c_int * 3)()
        check(lib.b200ocr_net_out_shape(self._h, shp))
        n, oh, ow = shp[0], shp[1], shp[2]
        if self.kind == "det":
            prob = np.empty((n, oh, ow), np.float32)
            bm = np.empty((n, oh, ow), np.uint8) if thresh_u8 >= 0 else None
            check(lib.b200ocr_net_output(self._h, prob.ctypes.data, bm.ctypes.data if bm is not None else None, None))
            return prob, bm
        if self.kind == "cls":
            out = np.empty((n, 2), np.float32)
            check(lib.b200ocr_net_output(self._h, out.ctypes.data, None, None))
            return out
        prob = np.empty((n, ow), np.float32)
        idx = np.empty((n, ow), np.int32)
        check(lib.b200ocr_net_output(self._h, prob.ctypes.data, None, idx.ctypes.data))
        return prob, idx

    def fetch(self, var: str) -> np.ndarray:
        dims = (C.c_int * 4)()
        check(lib.b200ocr_net_fetch(self._h, var.encode(), None, 0, dims))
        out = np.empty(tuple(dims), np.float32)
        check(lib.b200ocr_net_fetch(self._h, var.encode(), out.ctypes.data, out.size, dims))
        return out

    @property
    def launches(self) -> int:
        return lib.b200ocr_net_launches(self._h)
