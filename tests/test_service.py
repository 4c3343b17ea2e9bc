"""The Unix-socket transport of the reference's JSON service protocol (b200ocr/service.py; reference
src/ocr_ipc_service.cpp:310-423).  The protocol layer is tested on CPU against a stub pool; the GPU test runs the real
pool behind a real socket and compares with a direct worker call."""
import base64
import json
import os
import tempfile
import threading

import numpy as np
import pytest


class _StubPool:
    def __init__(self):
        self.calls = []

    def submit(self, rid, img):
        self.calls.append((rid, img.shape))
        return rid + 100

    def wait(self, ticket):
        return json.dumps({"request_id": ticket - 100, "success": True, "words": []}, separators=(",", ":"))

    def status(self):
        return {"running": True, "total_requests": len(self.calls), "successful_requests": len(self.calls),
                "average_processing_time_ms": 1.5}


def _proto():
    import importlib.util, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("b200ocr_service", os.path.join(root, "cpp-paddle-ocr_b200", "b200ocr", "service.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b200ocr_service"] = mod
    spec.loader.exec_module(mod)   # no libb200ocr.so needed: the module only imports it in main()
    return mod


def test_protocol_commands_and_error_strings(golden_dir):
    svc = _proto()
    pool = _StubPool()
    p = svc.Protocol(pool)
    assert json.loads(p.handle("{not json")) ["success"] is False
    assert json.loads(p.handle("{not json"))["error"].startswith("Invalid JSON: ")
    assert json.loads(p.handle('{"command":"dance"}')) == {"error": "Unknown command: dance", "success": False}
    assert json.loads(p.handle('{"command":"recognize"}')) == {"error": "Missing image_path or image_data", "success": False}
    assert json.loads(p.handle('{"command":"recognize","image_path":"/no/such.png"}')) == {
        "error": "Failed to load image from path: /no/such.png", "success": False}
    assert json.loads(p.handle('{"command":"recognize","image_data":"@@@"}'))["error"].startswith("Base64 decode error: ")
    bad = base64.b64encode(b"not an image").decode()
    assert json.loads(p.handle(json.dumps({"command": "recognize", "image_data": bad})))["error"] == "Failed to decode base64 image data"
    card = os.path.join(golden_dir, "card-jd.jpg")
    r0 = json.loads(p.handle(json.dumps({"command": "recognize", "image_path": card})))
    r1 = json.loads(p.handle(json.dumps({"command": "recognize", "image_data": base64.b64encode(open(card, "rb").read()).decode()})))
    assert (r0["request_id"], r1["request_id"]) == (0, 1) and r0["success"] and r1["success"]   # request counter
    assert pool.calls == [(0, (178, 391, 3)), (1, (178, 391, 3))]
    st = json.loads(p.handle('{"command":"status"}'))
    assert st["success"] is True and json.loads(st["status"])["total_requests"] == 2   # status is a nested JSON string
    assert not p.shutdown_requested.is_set()
    sd = json.loads(p.handle('{"command":"shutdown"}'))
    assert sd == {"message": "Shutdown command received, stopping service...", "success": True}
    assert p.shutdown_requested.is_set()


def test_socket_round_trip_with_stub_pool(golden_dir):
    svc = _proto()
    path = os.path.join(tempfile.mkdtemp(), "ocr.sock")
    srv = svc.Service(path, _StubPool())
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        assert svc.request(path, {"command": "status"})["success"] is True
        r = svc.request(path, {"command": "recognize", "image_path": os.path.join(golden_dir, "card-jd.jpg")})
        assert r["success"] and r["request_id"] == 0
        assert svc.request(path, {"command": "shutdown"})["success"] is True
        t.join(timeout=10)
        assert not t.is_alive()
    finally:
        srv.server_close()


@pytest.mark.gpu
def test_service_over_real_pool_matches_worker(models_dir, golden_dir):
    import b200ocr
    from b200ocr import service
    import cv2
    card = os.path.join(golden_dir, "card-jd.jpg")
    pool = b200ocr.Pool(models_dir, devices=[0], workers_per_device=2, enable_cls=True, max_batch=8)
    path = os.path.join(tempfile.mkdtemp(), "ocr.sock")
    srv = service.Service(path, pool)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        want = json.loads(b200ocr.Worker(0, models_dir, enable_cls=True).process(0, cv2.imread(card)))
        got = service.request(path, {"command": "recognize", "image_path": card})
        assert got["success"] and got["words"] == want["words"] and (got["width"], got["height"]) == (391, 178)
        b64 = base64.b64encode(open(card, "rb").read()).decode()
        got2 = service.request(path, {"command": "recognize", "image_data": b64})
        assert got2["words"] == want["words"] and got2["request_id"] == 1
        st = json.loads(service.request(path, {"command": "status"})["status"])
        assert st["total_requests"] == 2 and st["successful_requests"] == 2 and st["failed_requests"] == 0
        assert st["average_processing_time_ms"] > 0 and st["workers"] == 2 and st["running"] is True
        # the per-stage times the reference measures and drops: host wall ms per image, by stage
        assert set(st["stage_ms_per_image"]) == {"det", "cls", "rec"}
        assert all(v >= 0 for v in st["stage_ms_per_image"].values())
        assert list(st) == sorted(st)  # jsoncpp key order
        # a format the device decoder does not cover (PNG) goes through the host decoder, like the reference's imread
        png = base64.b64encode(cv2.imencode(".png", cv2.imread(card))[1].tobytes()).decode()
        got3 = service.request(path, {"command": "recognize", "image_data": png})
        assert got3["success"] and got3["words"] == want["words"]
        assert service.request(path, {"command": "shutdown"})["success"] is True
        t.join(timeout=10)
    finally:
        srv.server_close()
        pool.close()
