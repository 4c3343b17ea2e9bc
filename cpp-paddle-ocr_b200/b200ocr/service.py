"""A portable transport for the reference's JSON service protocol.

The reference serves `recognize` / `status` / `shutdown` commands over a Windows named pipe in message mode
(src/ocr_ipc_service.cpp:310-423; 1 MB request / 64 KB response limits, include/paddle_ocr/ocr_ipc_service.h:86-88).
This module speaks the same commands, fields and error strings over a Unix stream socket, one compact JSON document
per line in each direction (a stream socket has no message boundaries; the reference's pretty-printed envelopes would
not survive line framing, the OCR result line itself is already compact).  Baseline JPEG files are decoded on the GPU
(b200ocr_pool_submit_encoded: only the encoded bytes cross PCIe); every other format goes through cv2.imdecode on the
host like the reference's cv::imread / base64ToMat; the OCR itself runs on the GPU worker pool.

    python -m b200ocr.service --model-dir models --socket /tmp/b200ocr.sock --devices 0 --workers 2
"""
from __future__ import annotations
import argparse
import base64
import json
import os
import socket
import socketserver
import threading

MAX_REQUEST_BYTES = 1 << 20   # reference kMaxRequestSize
MAX_RESPONSE_BYTES = 64 << 10  # reference kMaxResponseSize


def _err(msg: str) -> str:
    return json.dumps({"error": msg, "success": False}, ensure_ascii=False, separators=(",", ":"))


class Protocol:
    """Command dispatch of OCRIPCService::processIPCRequest (src/ocr_ipc_service.cpp:310-423) over any object with
    submit(request_id, image) -> ticket, wait(ticket) -> str and status() -> dict (b200ocr.Pool)."""

    def __init__(self, pool, decode_path=None, decode_bytes=None):
        self.pool = pool
        self._next_id = 0
        self._lock = threading.Lock()
        self.shutdown_requested = threading.Event()
        if decode_path is None or decode_bytes is None:
            import cv2
            import numpy as np
            decode_path = decode_path or (lambda p: cv2.imread(p))
            decode_bytes = decode_bytes or (lambda b: cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR))
        self._decode_path, self._decode_bytes = decode_path, decode_bytes

    def handle(self, request_json: str) -> str:
        try:
            try:
                req = json.loads(request_json)
                if not isinstance(req, dict):
                    raise ValueError("not an object")
            except ValueError as e:
                return _err("Invalid JSON: " + str(e))
            command = req.get("command", "") or ""
            if command == "recognize":
                path, data = req.get("image_path", "") or "", req.get("image_data", "") or ""
                raw, error, kind = None, "", ""
                if path:
                    kind = "path"
                    try:
                        with open(path, "rb") as f:
                            raw = f.read()
                    except OSError:
                        error = "Failed to load image from path: " + path
                elif data:
                    kind = "data"
                    try:
                        raw = base64.b64decode(data, validate=True)
                    except Exception as e:  # noqa: BLE001 (the reference catches std::exception here)
                        error = "Base64 decode error: " + str(e)
                else:
                    error = "Missing image_path or image_data"
                if error:
                    return _err(error)
                rid = None

                def take_id():  # the reference numbers a request once its image is known to be loadable
                    with self._lock:
                        self._next_id += 1
                        return self._next_id - 1
                # Device decode first: only the file's bytes cross PCIe (b200ocr_pool_submit_encoded; baseline JPEG,
                # recognised by its SOI marker so that other formats do not cost a round trip).
                if raw[:2] == b"\xff\xd8" and hasattr(self.pool, "submit_encoded"):
                    rid = take_id()
                    line = self.pool.wait(self.pool.submit_encoded(rid, raw))
                    if '"Unsupported image encoding' not in line:
                        return line
                # Anything the device decoder does not cover (PNG, BMP, progressive JPEG, EXIF rotation ...) is decoded
                # the reference's way -- cv::imread / cv::imdecode on the host -- and submitted as pixels.
                image = self._decode_bytes(raw) if raw else None
                if image is None or image.size == 0:
                    return _err("Failed to load image from path: " + path if kind == "path" else "Failed to decode base64 image data")
                if rid is None:
                    rid = take_id()
                return self.pool.wait(self.pool.submit(rid, image))   # the worker's result line, unchanged
            if command == "status":
                # the reference nests getStatusInfo() as a JSON *string*; so does this
                st = json.dumps(self.pool.status(), separators=(",", ":"), sort_keys=True)
                return json.dumps({"status": st, "success": True}, separators=(",", ":"))
            if command == "shutdown":
                self.shutdown_requested.set()
                return json.dumps({"message": "Shutdown command received, stopping service...", "success": True},
                                  separators=(",", ":"))
            return _err("Unknown command: " + str(command))
        except Exception as e:  # noqa: BLE001
            return _err(str(e))


class _Handler(socketserver.StreamRequestHandler):
    def handle(self):
        proto: Protocol = self.server.protocol
        while not proto.shutdown_requested.is_set():
            line = self.rfile.readline(MAX_REQUEST_BYTES + 1)
            if not line:
                return
            if len(line) > MAX_REQUEST_BYTES:
                self.wfile.write((_err("Request too large") + "\n").encode())
                return
            resp = proto.handle(line.decode("utf-8", "replace").strip())
            self.wfile.write(resp.encode("utf-8") + b"\n")
            self.wfile.flush()
            if proto.shutdown_requested.is_set():
                threading.Thread(target=self.server.shutdown, daemon=True).start()
                return


class Service(socketserver.ThreadingMixIn, socketserver.UnixStreamServer):
    """One thread per client connection, like the reference's per-pipe client threads."""
    daemon_threads = True
    allow_reuse_address = True

    def __init__(self, socket_path: str, pool, **decoders):
        if os.path.exists(socket_path):
            os.unlink(socket_path)
        self.protocol = Protocol(pool, **decoders)
        super().__init__(socket_path, _Handler)


def request(socket_path: str, obj: dict, timeout: float = 60.0) -> dict:
    """Client side: one request, one response."""
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.settimeout(timeout)
        s.connect(socket_path)
        s.sendall(json.dumps(obj, separators=(",", ":")).encode() + b"\n")
        buf = b""
        while not buf.endswith(b"\n"):
            chunk = s.recv(1 << 16)
            if not chunk:
                break
            buf += chunk
    return json.loads(buf.decode("utf-8"))


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model-dir", required=True)
    ap.add_argument("--socket", default="/tmp/b200ocr.sock")
    ap.add_argument("--devices", default="0", help="comma-separated GPU ids")
    ap.add_argument("--workers", type=int, default=2, help="workers per GPU")
    ap.add_argument("--no-cls", action="store_true")
    a = ap.parse_args()
    import b200ocr
    pool = b200ocr.Pool(a.model_dir, devices=[int(d) for d in a.devices.split(",")], workers_per_device=a.workers,
                        enable_cls=not a.no_cls)
    srv = Service(a.socket, pool)
    print(f"b200ocr service on {a.socket}: {pool.worker_count} workers", flush=True)
    try:
        srv.serve_forever()
    finally:
        srv.server_close()
        pool.close()
        if os.path.exists(a.socket):
            os.unlink(a.socket)


if __name__ == "__main__":
    main()
