"""ORACLE (test infrastructure, never shipped): the polygon offset inside DBPostProcessor::UnClip
(reference src/postprocess_op.cpp:47-62).

Two implementations:
  * `RefClipper` — ctypes binding of oracle/_ref/libclipper_ref.so, i.e. the reference's own vendored
    Clipper 6.4.2 (src/clipper.cpp) compiled where it lies (oracle/Makefile).  kind "reference".
  * `offset_points_restated` — plain-Python restatement of ClipperOffset::AddPath (clipper.cpp:3628-3673),
    FixOrientations (:3682-3702), DoOffset (:3779-3944), OffsetPoint (:3947-3994) and DoRound
    (:4007-4021) for ONE closed polygon with jtRound.  The union pass that follows in
    ClipperOffset::Execute (:3705-3718) only removes self-overlap of the offset outline and never
    changes its convex hull for delta > 0, and the caller only takes cv::minAreaRect of the result,
    so the restatement stops before it.  tests/test_oracle_unclip.py checks hull equality and
    minAreaRect equality of the two on random boxes.
"""
from __future__ import annotations
import ctypes as C
import math
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libclipper_ref.so")


class RefClipper:
    def __init__(self, path=REF_LIB):
        self.lib = C.CDLL(path)
        self.lib.ref_unclip_offset.restype = C.c_int
        self.lib.ref_unclip_offset.argtypes = [C.POINTER(C.c_longlong), C.c_int, C.c_double,
                                               C.POINTER(C.c_longlong), C.c_int]

    def offset(self, path, delta):
        n = len(path)
        xy = (C.c_longlong * (2 * n))(*[v for p in path for v in p])
        cap = 4096
        out = (C.c_longlong * (2 * cap))()
        m = self.lib.ref_unclip_offset(xy, n, float(delta), out, cap)
        if m < 0:
            return []
        return [(out[2 * i], out[2 * i + 1]) for i in range(min(m, cap))]


_ref = None


def ref_clipper():
    """The compiled reference Clipper, or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        _ref = RefClipper()
    return _ref


def _round(v):  # clipper.cpp:128-133
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def _unit_normal(p1, p2):  # clipper.cpp:3594-3604
    if p1 == p2:
        return (0.0, 0.0)
    dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
    f = 1.0 / math.sqrt(dx * dx + dy * dy)
    return (dy * f, -dx * f)


def offset_points_restated(path, delta, arc_tolerance=0.25):
    hi = len(path) - 1
    if hi < 0:
        return []
    while hi > 0 and path[0] == path[hi]:
        hi -= 1
    src = [tuple(path[0])]
    for i in range(1, hi + 1):
        if src[-1] != tuple(path[i]):
            src.append(tuple(path[i]))
    if len(src) < 3:
        return []
    # Area / Orientation, clipper.cpp:356-370
    a, j = 0.0, len(src) - 1
    for i in range(len(src)):
        a += (float(src[j][0]) + src[i][0]) * (float(src[j][1]) - src[i][1])
        j = i
    if not (-a * 0.5 >= 0):
        src.reverse()
    if abs(delta) < 1e-20:
        return list(src)
    y = arc_tolerance
    if y <= 0.0:
        y = 0.25
    elif y > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin, m_cos = math.sin(2 * math.pi / steps), math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0.0:
        m_sin = -m_sin
    n = len(src)
    normals = [_unit_normal(src[i], src[(i + 1) % n]) for i in range(n)]
    out = []
    k = n - 1
    for j in range(n):
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
            if cos_a > 0:
                out.append((_round(src[j][0] + normals[k][0] * delta), _round(src[j][1] + normals[k][1] * delta)))
                done = True  # returns before `k = j`
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if done:
            continue
        if sin_a * delta < 0:
            out.append((_round(src[j][0] + normals[k][0] * delta), _round(src[j][1] + normals[k][1] * delta)))
            out.append(src[j])
            out.append((_round(src[j][0] + normals[j][0] * delta), _round(src[j][1] + normals[j][1] * delta)))
        else:  # DoRound
            ang = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
            st = max(_round(steps_per_rad * abs(ang)), 1)
            x, yy = normals[k]
            for _ in range(st):
                out.append((_round(src[j][0] + x * delta), _round(src[j][1] + yy * delta)))
                x, yy = x * m_cos - m_sin * yy, x * m_sin + yy * m_cos
            out.append((_round(src[j][0] + normals[j][0] * delta), _round(src[j][1] + normals[j][1] * delta)))
        k = j
    return out


def offset_points(path, delta, clipper=None):
    """Points UnClip feeds to cv::minAreaRect.  `clipper`: a RefClipper, or None to use the compiled
    reference when present and the restatement otherwise."""
    c = clipper if clipper is not None else ref_clipper()
    if c is not None:
        return c.offset(path, delta)
    return offset_points_restated(path, delta)
