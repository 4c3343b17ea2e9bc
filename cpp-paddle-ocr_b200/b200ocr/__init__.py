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
_sig("b200ocr_net_forward_ragged", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p)
_sig("b200ocr_net_out_shape", C.c_int, C.c_void_p, C.POINTER(C.c_int))
_sig("b200ocr_net_output", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
_sig("b200ocr_net_fetch", C.c_int, C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int))
_sig("b200ocr_net_launches", C.c_int, C.c_void_p)
_sig("b200ocr_pool_status", C.c_int, C.c_void_p, C.POINTER(C.c_void_p))


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

    def forward(self, x: np.ndarray, thresh_u8: int = -1, widths=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, c, h, w = x.shape
        assert c == 3
        if widths is not None:
            wd = np.ascontiguousarray(widths, np.int32)
            assert wd.shape == (n,)
            check(lib.b200ocr_net_forward_ragged(self._h, x.ctypes.data, n, h, w, wd.ctypes.data))
        else:
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


# ----------------------------------------------------------------------------------------- stages
class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("step", C.c_size_t)]


class DetConfig(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("use_gpu", C.c_int), ("gpu_id", C.c_int), ("gpu_mem", C.c_int),
                ("cpu_math_library_num_threads", C.c_int), ("use_mkldnn", C.c_int), ("limit_type", C.c_char_p),
                ("limit_side_len", C.c_int), ("det_db_thresh", C.c_double), ("det_db_box_thresh", C.c_double),
                ("det_db_unclip_ratio", C.c_double), ("det_db_score_mode", C.c_char_p), ("use_dilation", C.c_int),
                ("use_tensorrt", C.c_int), ("precision", C.c_char_p)]


class ClsConfig(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("use_gpu", C.c_int), ("gpu_id", C.c_int), ("gpu_mem", C.c_int),
                ("cpu_math_library_num_threads", C.c_int), ("use_mkldnn", C.c_int), ("cls_thresh", C.c_double),
                ("use_tensorrt", C.c_int), ("precision", C.c_char_p), ("cls_batch_num", C.c_int)]


class RecConfig(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("use_gpu", C.c_int), ("gpu_id", C.c_int), ("gpu_mem", C.c_int),
                ("cpu_math_library_num_threads", C.c_int), ("use_mkldnn", C.c_int), ("label_path", C.c_char_p),
                ("use_tensorrt", C.c_int), ("precision", C.c_char_p), ("rec_batch_num", C.c_int),
                ("rec_img_h", C.c_int), ("rec_img_w", C.c_int)]


_P = C.POINTER
_sig("b200ocr_host_alloc", C.c_void_p, C.c_size_t)
_sig("b200ocr_host_free", None, C.c_void_p)
_sig("b200ocr_device_count", C.c_int)
_sig("b200ocr_det_create", C.c_int, _P(DetConfig), _P(C.c_void_p))
_sig("b200ocr_det_destroy", None, C.c_void_p)
_sig("b200ocr_det_run", C.c_int, C.c_void_p, _P(Image), C.c_void_p, C.c_int, _P(C.c_int), _P(C.c_double))
_sig("b200ocr_det_run_batch", C.c_int, C.c_void_p, _P(Image), C.c_int, C.c_void_p, C.c_int, C.c_void_p, _P(C.c_double))
_sig("b200ocr_det_postprocess", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
     C.c_int, _P(C.c_int), C.c_void_p)
_sig("b200ocr_det_preprocess", C.c_int, C.c_void_p, _P(Image), C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_float),
     _P(C.c_float))
_sig("b200ocr_cls_create", C.c_int, _P(ClsConfig), _P(C.c_void_p))
_sig("b200ocr_cls_destroy", None, C.c_void_p)
_sig("b200ocr_cls_run", C.c_int, C.c_void_p, _P(Image), C.c_int, C.c_void_p, C.c_void_p, _P(C.c_double))
_sig("b200ocr_rec_create", C.c_int, _P(RecConfig), _P(C.c_void_p))
_sig("b200ocr_rec_destroy", None, C.c_void_p)
_sig("b200ocr_rec_run", C.c_int, C.c_void_p, _P(Image), C.c_int, _P(C.c_void_p), C.c_void_p, _P(C.c_double))
_sig("b200ocr_worker_create", C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, _P(C.c_void_p))
_sig("b200ocr_worker_destroy", None, C.c_void_p)
_sig("b200ocr_worker_process", C.c_int, C.c_void_p, C.c_int, _P(Image), _P(C.c_void_p))
_sig("b200ocr_worker_process_batch", C.c_int, C.c_void_p, _P(C.c_int), _P(Image), C.c_int, _P(C.c_void_p))
_sig("b200ocr_worker_launches", C.c_longlong, C.c_void_p)
_sig("b200ocr_pool_create", C.c_int, C.c_char_p, C.c_int, _P(C.c_int), C.c_int, C.c_int, C.c_int, _P(C.c_void_p))
_sig("b200ocr_pool_destroy", None, C.c_void_p)
_sig("b200ocr_pool_submit", C.c_int, C.c_void_p, C.c_int, _P(Image), _P(C.c_longlong))
_sig("b200ocr_pool_wait", C.c_int, C.c_void_p, C.c_longlong, _P(C.c_void_p))
_sig("b200ocr_pool_worker_count", C.c_int, C.c_void_p)
_sig("b200ocr_pool_idle_count", C.c_int, C.c_void_p)
_sig("b200ocr_resize_u8", C.c_int, C.c_int, _P(Image), C.c_int, C.c_int, C.c_void_p)
_sig("b200ocr_crop_preprocess", C.c_int, C.c_int, _P(Image), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p)


def device_count() -> int:
    return lib.b200ocr_device_count()


def as_image(a: np.ndarray) -> Image:
    """View of a uint8 HxWx3 BGR array (what cv::Mat describes); an empty array maps to an empty image."""
    if a is None or a.size == 0:
        return Image(None, 0, 0, 0)
    assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3 and a.strides[2] == 1 and a.strides[1] == 3
    return Image(a.ctypes.data, a.shape[0], a.shape[1], a.strides[0])


class Prepared:
    """A list of images already turned into the C ABI's b200ocr_image array (prepare_images): passing it instead of
    the list keeps the per-call Python cost out of a timed loop."""

    def __init__(self, arrs):
        self.keep = [a if (a is None or a.size == 0 or a.strides[1:] == (3, 1)) else np.ascontiguousarray(a) for a in arrs]
        self.arr = (Image * len(self.keep))(*[as_image(a) for a in self.keep])

    def __len__(self):
        return len(self.keep)


def prepare_images(arrs) -> Prepared:
    return arrs if isinstance(arrs, Prepared) else Prepared(arrs)


def _images(arrs):
    p = prepare_images(arrs)
    return p.arr, p.keep


def pinned_array(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array over page-locked host memory from b200ocr_host_alloc (freed when the array is collected)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.b200ocr_host_alloc(n)
    if not p:
        raise Error("b200ocr_host_alloc failed")
    buf = (C.c_uint8 * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.b200ocr_host_free, p)
    return arr


class _Handle:
    _destroy = None

    def close(self):
        if getattr(self, "_h", None):
            self._destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Detector(_Handle):
    """DBDetector (reference include/paddle_ocr/ocr_det.h:60-97)."""
    _destroy = staticmethod(lib.b200ocr_det_destroy)

    def __init__(self, model_dir, gpu_id=0, limit_type="max", limit_side_len=960, det_db_thresh=0.3,
                 det_db_box_thresh=0.5, det_db_unclip_ratio=2.0, det_db_score_mode="fast", use_dilation=False):
        cfg = DetConfig(model_dir.encode(), 1, gpu_id, 0, 1, 0, limit_type.encode(), limit_side_len, det_db_thresh,
                        det_db_box_thresh, det_db_unclip_ratio, det_db_score_mode.encode(), int(use_dilation), 0, b"fp32")
        self._h = C.c_void_p()
        check(lib.b200ocr_det_create(C.byref(cfg), C.byref(self._h)))

    def run_batch(self, imgs, cap=1000):
        arr, keep = _images(imgs)
        n = len(imgs)
        boxes = np.zeros((n, cap, 4, 2), np.int32)
        counts = np.zeros(n, np.int32)
        times = (C.c_double * 3)()
        check(lib.b200ocr_det_run_batch(self._h, arr, n, boxes.ctypes.data, cap, counts.ctypes.data, times))
        self.times = list(times)
        return [boxes[i, :min(int(counts[i]), cap)].copy() for i in range(n)]

    def run(self, img, cap=1000):
        return self.run_batch([img], cap)[0]

    def preprocess(self, img):
        arr, keep = _images([img])
        rh, rw, a, b = C.c_int(), C.c_int(), C.c_float(), C.c_float()
        check(lib.b200ocr_det_preprocess(self._h, arr, None, C.byref(rh), C.byref(rw), C.byref(a), C.byref(b)))
        out = np.empty((3, rh.value, rw.value), np.float32)
        check(lib.b200ocr_det_preprocess(self._h, arr, out.ctypes.data, C.byref(rh), C.byref(rw), C.byref(a), C.byref(b)))
        return out, a.value, b.value

    def postprocess(self, pred, src_h, src_w, cap=1000, want_bitmap=False):
        pred = np.ascontiguousarray(pred, np.float32)
        h, w = pred.shape
        boxes = np.zeros((cap, 4, 2), np.int32)
        n = C.c_int()
        bm = np.empty((h, w), np.uint8) if want_bitmap else None
        check(lib.b200ocr_det_postprocess(self._h, pred.ctypes.data, h, w, src_h, src_w, boxes.ctypes.data, cap,
                                          C.byref(n), bm.ctypes.data if want_bitmap else None))
        out = boxes[:min(n.value, cap)].copy()
        return (out, bm) if want_bitmap else out


class Classifier(_Handle):
    """Classifier (reference include/paddle_ocr/ocr_cls.h:57-82)."""
    _destroy = staticmethod(lib.b200ocr_cls_destroy)

    def __init__(self, model_dir, gpu_id=0, cls_thresh=0.9, cls_batch_num=1):
        cfg = ClsConfig(model_dir.encode(), 1, gpu_id, 0, 1, 0, cls_thresh, 0, b"fp32", cls_batch_num)
        self._h = C.c_void_p()
        check(lib.b200ocr_cls_create(C.byref(cfg), C.byref(self._h)))

    def run(self, imgs):
        arr, keep = _images(imgs)
        n = len(imgs)
        labels = np.zeros(n, np.int32)
        scores = np.zeros(n, np.float32)
        times = (C.c_double * 3)()
        check(lib.b200ocr_cls_run(self._h, arr, n, labels.ctypes.data, scores.ctypes.data, times))
        return labels, scores


class Recognizer(_Handle):
    """CRNNRecognizer (reference include/paddle_ocr/ocr_rec.h:61-95)."""
    _destroy = staticmethod(lib.b200ocr_rec_destroy)

    def __init__(self, model_dir, label_path, gpu_id=0, rec_batch_num=6, rec_img_h=48, rec_img_w=320):
        cfg = RecConfig(model_dir.encode(), 1, gpu_id, 0, 1, 0, label_path.encode(), 0, b"fp32", rec_batch_num,
                        rec_img_h, rec_img_w)
        self._h = C.c_void_p()
        check(lib.b200ocr_rec_create(C.byref(cfg), C.byref(self._h)))

    def run(self, imgs):
        arr, keep = _images(imgs)
        n = len(imgs)
        texts = (C.c_void_p * n)()
        scores = np.zeros(n, np.float32)
        times = (C.c_double * 3)()
        check(lib.b200ocr_rec_run(self._h, arr, n, texts, scores.ctypes.data, times))
        self.times = list(times)
        return [_take_string(C.c_void_p(t)) for t in texts], scores


class WorkerParams(C.Structure):
    _fields_ = [("limit_type", C.c_char_p), ("limit_side_len", C.c_int), ("det_db_thresh", C.c_double),
                ("det_db_box_thresh", C.c_double), ("det_db_unclip_ratio", C.c_double), ("det_db_score_mode", C.c_char_p),
                ("use_dilation", C.c_int), ("cls_batch_num", C.c_int), ("rec_batch_num", C.c_int), ("rec_img_h", C.c_int),
                ("rec_img_w", C.c_int), ("max_batch", C.c_int)]


_sig("b200ocr_worker_create_ex", C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, _P(WorkerParams), _P(C.c_void_p))


class Worker(_Handle):
    """OCRWorker::processRequest + result JSON (reference src/ocr_worker.cpp:133-311)."""
    _destroy = staticmethod(lib.b200ocr_worker_destroy)

    def __init__(self, worker_id, model_dir, gpu_id=0, enable_cls=False, **params):
        """`params`: b200ocr_worker_params overrides (limit_side_len=960, rec_img_h=48, ...); none = the reference's."""
        self._h = C.c_void_p()
        if not params:
            check(lib.b200ocr_worker_create(worker_id, model_dir.encode(), 1, gpu_id, int(enable_cls), C.byref(self._h)))
            return
        wp = WorkerParams()
        for k, v in params.items():
            if k not in {f[0] for f in WorkerParams._fields_}:
                raise TypeError(f"unknown worker parameter {k}")
            setattr(wp, k, v.encode() if isinstance(v, str) else v)
        check(lib.b200ocr_worker_create_ex(worker_id, model_dir.encode(), gpu_id, int(enable_cls), C.byref(wp), C.byref(self._h)))

    def process(self, request_id, img) -> str:
        arr, keep = _images([img])
        p = C.c_void_p()
        check(lib.b200ocr_worker_process(self._h, request_id, arr, C.byref(p)))
        return _take_string(p)

    def process_batch(self, request_ids, imgs):
        arr, keep = _images(imgs)
        n = len(imgs)
        ids = (C.c_int * n)(*request_ids)
        out = (C.c_void_p * n)()
        check(lib.b200ocr_worker_process_batch(self._h, ids, arr, n, out))
        return [_take_string(C.c_void_p(t)) for t in out]

    @property
    def launches(self) -> int:
        return lib.b200ocr_worker_launches(self._h)


class Pool(_Handle):
    """Per-device worker pool (replaces the reference's GPUWorkerPool, src/gpu_worker_pool.cpp:8-59)."""
    _destroy = staticmethod(lib.b200ocr_pool_destroy)

    def __init__(self, model_dir, devices=(0,), workers_per_device=1, enable_cls=False, max_batch=64):
        self._h = C.c_void_p()
        dv = (C.c_int * len(devices))(*devices)
        check(lib.b200ocr_pool_create(model_dir.encode(), len(devices), dv, workers_per_device, int(enable_cls),
                                      max_batch, C.byref(self._h)))

    def submit(self, request_id, img) -> int:
        arr, keep = _images([img])
        t = C.c_longlong()
        check(lib.b200ocr_pool_submit(self._h, request_id, arr, C.byref(t)))
        return t.value

    def wait(self, ticket) -> str:
        p = C.c_void_p()
        check(lib.b200ocr_pool_wait(self._h, ticket, C.byref(p)))
        return _take_string(p)

    @property
    def worker_count(self):
        return lib.b200ocr_pool_worker_count(self._h)

    @property
    def idle_count(self):
        return lib.b200ocr_pool_idle_count(self._h)

    def status(self) -> dict:
        """Service counters (reference OCRIPCService::getStatusInfo, src/ocr_ipc_service.cpp:438-448)."""
        p = C.c_void_p()
        check(lib.b200ocr_pool_status(self._h, C.byref(p)))
        return json.loads(_take_string(p))


def resize_u8(img, dst_rows, dst_cols, device=0):
    arr, keep = _images([img])
    out = np.empty((dst_rows, dst_cols, 3), np.uint8)
    check(lib.b200ocr_resize_u8(device, arr, dst_rows, dst_cols, out.ctypes.data))
    return out


def crop_preprocess(crops, kind, img_h, img_w, device=0):
    arr, keep = _images(crops)
    out = np.empty((len(crops), 3, img_h, img_w), np.float32)
    check(lib.b200ocr_crop_preprocess(device, arr, len(crops), 0 if kind == "rec" else 1, img_h, img_w, out.ctypes.data))
    return out


_sig("b200ocr_batch_upload", C.c_int, C.c_int, _P(Image), C.c_int, _P(C.c_void_p))
_sig("b200ocr_batch_destroy", None, C.c_void_p)
_sig("b200ocr_worker_process_resident", C.c_int, C.c_void_p, C.c_void_p, _P(C.c_int), _P(C.c_void_p))
_sig("b200ocr_worker_stream", C.c_void_p, C.c_void_p)
_sig("b200ocr_worker_profile", C.c_int, C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p))
_sig("b200ocr_net_profile", C.c_int, C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p))


class DeviceBatch(_Handle):
    """Images uploaded once and kept in device memory (b200ocr_batch_upload)."""
    _destroy = staticmethod(lib.b200ocr_batch_destroy)

    def __init__(self, imgs, device=0):
        arr, keep = _images(imgs)
        self.n = len(imgs)
        self._h = C.c_void_p()
        check(lib.b200ocr_batch_upload(device, arr, self.n, C.byref(self._h)))


_sig("b200ocr_det_run_resident", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _P(C.c_double))
_sig("b200ocr_rec_run_resident", C.c_int, C.c_void_p, C.c_void_p, _P(C.c_void_p), C.c_void_p, _P(C.c_double))
_sig("b200ocr_det_stream", C.c_void_p, C.c_void_p)
_sig("b200ocr_rec_stream", C.c_void_p, C.c_void_p)
_sig("b200ocr_det_launches", C.c_longlong, C.c_void_p)
_sig("b200ocr_rec_launches", C.c_longlong, C.c_void_p)
_sig("b200ocr_pool_wait_for", C.c_int, C.c_void_p, C.c_longlong, C.c_int, _P(C.c_void_p))


def _det_run_resident(self, batch: DeviceBatch, cap=1000):
    boxes = np.zeros((batch.n, cap, 4, 2), np.int32)
    counts = np.zeros(batch.n, np.int32)
    times = (C.c_double * 3)()
    check(lib.b200ocr_det_run_resident(self._h, batch._h, boxes.ctypes.data, cap, counts.ctypes.data, times))
    self.times = list(times)
    return [boxes[i, :min(int(counts[i]), cap)].copy() for i in range(batch.n)]


def _rec_run_resident(self, batch: DeviceBatch):
    texts = (C.c_void_p * batch.n)()
    scores = np.zeros(batch.n, np.float32)
    times = (C.c_double * 3)()
    check(lib.b200ocr_rec_run_resident(self._h, batch._h, texts, scores.ctypes.data, times))
    self.times = list(times)
    return [_take_string(C.c_void_p(t)) for t in texts], scores


_sig("b200ocr_det_profile", C.c_int, C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p))
_sig("b200ocr_rec_profile", C.c_int, C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p))


def _stage_profile(fn):
    def profile(self, warmup=2, reps=5):
        p = C.c_void_p()
        check(fn(self._h, warmup, reps, C.byref(p)))
        return json.loads(_take_string(p))
    return profile


Detector.profile = _stage_profile(lib.b200ocr_det_profile)
Recognizer.profile = _stage_profile(lib.b200ocr_rec_profile)
Detector.run_resident = _det_run_resident
Recognizer.run_resident = _rec_run_resident
Detector.stream = property(lambda self: lib.b200ocr_det_stream(self._h))
Recognizer.stream = property(lambda self: lib.b200ocr_rec_stream(self._h))
Detector.launches = property(lambda self: int(lib.b200ocr_det_launches(self._h)))
Recognizer.launches = property(lambda self: int(lib.b200ocr_rec_launches(self._h)))


def _pool_wait_for(self, ticket, timeout_ms):
    """None when the time limit elapsed first (the ticket stays valid)."""
    p = C.c_void_p()
    check(lib.b200ocr_pool_wait_for(self._h, ticket, int(timeout_ms), C.byref(p)))
    return _take_string(p) if p.value else None


Pool.wait_for = _pool_wait_for


def _worker_process_resident(self, request_ids, batch: DeviceBatch):
    ids = (C.c_int * batch.n)(*request_ids)
    out = (C.c_void_p * batch.n)()
    check(lib.b200ocr_worker_process_resident(self._h, batch._h, ids, out))
    return [_take_string(C.c_void_p(t)) for t in out]


def _worker_profile(self, warmup=2, reps=5):
    p = C.c_void_p()
    check(lib.b200ocr_worker_profile(self._h, warmup, reps, C.byref(p)))
    return json.loads(_take_string(p))


Worker.process_resident = _worker_process_resident
Worker.profile = _worker_profile
Worker.stream = property(lambda self: lib.b200ocr_worker_stream(self._h))


def _net_profile(self, warmup=2, reps=5):
    p = C.c_void_p()
    check(lib.b200ocr_net_profile(self._h, warmup, reps, C.byref(p)))
    return json.loads(_take_string(p))


Net.profile = _net_profile


_sig("b200ocr_rotate_crop", C.c_int, C.c_int, _P(Image), C.c_void_p, _P(C.c_int), _P(C.c_int), C.c_void_p)


def rotate_crop(img, box, device=0):
    """Utility::GetRotateCropImage (reference src/utility.cpp:137-190) on the GPU."""
    arr, keep = _images([img])
    b = np.ascontiguousarray(np.asarray(box, np.int32).reshape(8))
    r, c = C.c_int(), C.c_int()
    check(lib.b200ocr_rotate_crop(device, arr, b.ctypes.data, C.byref(r), C.byref(c), None))
    out = np.empty((r.value, c.value, 3), np.uint8)
    check(lib.b200ocr_rotate_crop(device, arr, b.ctypes.data, C.byref(r), C.byref(c), out.ctypes.data))
    return out


# ---------------------------------------------------------------- kernel-level entry points (parity tests)
_sig("b200ocr_kernel_dwconv", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
     C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
     C.POINTER(C.c_int), C.POINTER(C.c_int))
_sig("b200ocr_kernel_attention", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
     C.c_void_p)


def kernel_dwconv(x, filt, bias, k, sh, sw, act=0, post_scale=1.0, post_shift=0.0, fp16_weights=True, out_widths=None,
                  device=0):
    """x [n,c,h,w] fp32, filt [c,k,k], bias [c] -> [n,c,oh,ow] fp32 (device path: NHWC fp16, fp32 accumulation)."""
    x = np.ascontiguousarray(x, np.float32)
    filt = np.ascontiguousarray(filt, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    n, c, h, w = x.shape
    pad = k // 2
    oh, ow = (h + 2 * pad - k) // sh + 1, (w + 2 * pad - k) // sw + 1
    out = np.empty((n, c, oh, ow), np.float32)
    wd = None if out_widths is None else np.ascontiguousarray(out_widths, np.int32)
    rh, rw = C.c_int(), C.c_int()
    check(lib.b200ocr_kernel_dwconv(device, x.ctypes.data, n, c, h, w, filt.ctypes.data, bias.ctypes.data, k, sh, sw, act,
                                    post_scale, post_shift, 1 if fp16_weights else 0,
                                    None if wd is None else wd.ctypes.data, out.ctypes.data, C.byref(rh), C.byref(rw)))
    assert (rh.value, rw.value) == (oh, ow)
    return out


def kernel_attention(qkv, heads, head_dim, scale, valid=None, device=0):
    """qkv [n,t,3*heads*head_dim] fp32 -> [n,t,heads*head_dim] fp32."""
    qkv = np.ascontiguousarray(qkv, np.float32)
    n, t, c3 = qkv.shape
    assert c3 == 3 * heads * head_dim
    out = np.empty((n, t, heads * head_dim), np.float32)
    vd = None if valid is None else np.ascontiguousarray(valid, np.int32)
    check(lib.b200ocr_kernel_attention(device, qkv.ctypes.data, n, t, heads, head_dim, scale,
                                       None if vd is None else vd.ctypes.data, out.ctypes.data))
    return out


_sig("b200ocr_kernel_conv", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
     C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)


def kernel_conv(x, filt, bias, act=0, post_scale=1.0, post_shift=0.0, residual=None, out_widths=None, force_simt=False,
                device=0):
    """x [n,cin,h,w], filt [cout,cin,kh,kw] (stride 1, same padding), bias [cout] -> [n,cout,h,w] fp32.
    force_simt: 0/False = the engine's choice, 1/True = CUDA-core kernel, 2 = skip the narrow-1x1 mma.sync kernel."""
    x = np.ascontiguousarray(x, np.float32)
    filt = np.ascontiguousarray(filt, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    n, cin, h, w = x.shape
    cout, cin2, kh, kw = filt.shape
    assert cin2 == cin
    out = np.empty((n, cout, h, w), np.float32)
    res = None if residual is None else np.ascontiguousarray(residual, np.float32)
    wd = None if out_widths is None else np.ascontiguousarray(out_widths, np.int32)
    check(lib.b200ocr_kernel_conv(device, x.ctypes.data, n, cin, h, w, filt.ctypes.data, bias.ctypes.data, cout, kh, kw, act,
                                  post_scale, post_shift, None if res is None else res.ctypes.data,
                                  None if wd is None else wd.ctypes.data, int(force_simt), out.ctypes.data))
    return out


_sig("b200ocr_kernel_ctc_head", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


def kernel_ctc_head(feat, w, bias, force_simt=False, device=0):
    """feat [n,t,cin], w [cin,ncls], bias [ncls] -> (idx [n,t], prob [n,t], collapsed ids per row, scores [n])."""
    feat = np.ascontiguousarray(feat, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    n, t, cin = feat.shape
    assert w.shape[0] == cin and bias.shape[0] == w.shape[1]
    idx = np.empty((n, t), np.int32)
    prob = np.empty((n, t), np.float32)
    col = np.empty((n, t), np.int32)
    lens = np.empty(n, np.int32)
    scores = np.empty(n, np.float32)
    check(lib.b200ocr_kernel_ctc_head(device, feat.ctypes.data, n, t, cin, w.ctypes.data, bias.ctypes.data, w.shape[1],
                                      1 if force_simt else 0, idx.ctypes.data, prob.ctypes.data, col.ctypes.data,
                                      lens.ctypes.data, scores.ctypes.data))
    return idx, prob, [col[i, :lens[i]].copy() for i in range(n)], scores


# ---------------------------------------------------------------- encoded inputs (device JPEG decode)
class Blob(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


_sig("b200ocr_worker_process_encoded", C.c_int, C.c_void_p, _P(C.c_int), _P(Blob), C.c_int, _P(C.c_void_p))
_sig("b200ocr_worker_last_encoded_h2d_bytes", C.c_longlong, C.c_void_p)
_sig("b200ocr_jpeg_decode", C.c_int, C.c_int, C.c_void_p, C.c_size_t, _P(C.c_int), _P(C.c_int), C.c_void_p)
_sig("b200ocr_pool_submit_encoded", C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, _P(C.c_longlong))


def _blob_array(files):
    """files: bytes / bytearray / uint8 arrays -> (Blob array, keep-alive list)"""
    keep = [np.frombuffer(f, np.uint8) if isinstance(f, (bytes, bytearray, memoryview)) else np.ascontiguousarray(f, np.uint8)
            for f in files]
    arr = (Blob * len(keep))(*[Blob(k.ctypes.data if k.size else None, k.size) for k in keep])
    return arr, keep


class PreparedBlobs:
    def __init__(self, files):
        self.arr, self.keep = _blob_array(files)

    def __len__(self):
        return len(self.keep)


def jpeg_decode(data, device=0) -> np.ndarray:
    """Baseline JPEG bytes -> uint8 [h, w, 3] BGR, decoded on the GPU (bit-identical to cv2.imdecode)."""
    buf = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, np.uint8)
    r, c = C.c_int(), C.c_int()
    check(lib.b200ocr_jpeg_decode(device, buf.ctypes.data, buf.size, C.byref(r), C.byref(c), None))
    out = np.empty((r.value, c.value, 3), np.uint8)
    check(lib.b200ocr_jpeg_decode(device, buf.ctypes.data, buf.size, C.byref(r), C.byref(c), out.ctypes.data))
    return out


def _worker_process_encoded(self, request_ids, files):
    p = files if isinstance(files, PreparedBlobs) else PreparedBlobs(files)
    n = len(p)
    ids = (C.c_int * n)(*request_ids)
    out = (C.c_void_p * n)()
    check(lib.b200ocr_worker_process_encoded(self._h, ids, p.arr, n, out))
    return [_take_string(C.c_void_p(t)) for t in out]


Worker.process_encoded = _worker_process_encoded
Worker.last_encoded_h2d_bytes = property(lambda self: int(lib.b200ocr_worker_last_encoded_h2d_bytes(self._h)))


def _pool_submit_encoded(self, request_id, data) -> int:
    buf = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, np.uint8)
    t = C.c_longlong()
    check(lib.b200ocr_pool_submit_encoded(self._h, request_id, buf.ctypes.data if buf.size else None, buf.size, C.byref(t)))
    return t.value


Pool.submit_encoded = _pool_submit_encoded
