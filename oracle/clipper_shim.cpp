// ORACLE (test infrastructure, never shipped): C entry point around the REFERENCE's own vendored
// Clipper 6.4.2, compiled from /root/reference/src/clipper.cpp where it lies (see oracle/Makefile).
// Mirrors the Clipper calls of DBPostProcessor::UnClip (reference src/postprocess_op.cpp:47-62).
#include <paddle_ocr/clipper.h>

extern "C" int ref_unclip_offset(const long long* xy, int n, double delta, long long* out_xy, int cap) {
  ClipperLib::ClipperOffset offset;
  ClipperLib::Path p;
  for (int i = 0; i < n; ++i) p.emplace_back(xy[2 * i], xy[2 * i + 1]);
  offset.AddPath(p, ClipperLib::jtRound, ClipperLib::etClosedPolygon);
  ClipperLib::Paths soln;
  if (!offset.Execute(soln, delta)) return -1;
  int m = 0;
  for (size_t j = 0; j < soln.size(); ++j)
    for (size_t i = 0; i < soln[soln.size() - 1].size(); ++i) {
      if (m < cap) {
        out_xy[2 * m] = soln[j][i].X;
        out_xy[2 * m + 1] = soln[j][i].Y;
      }
      ++m;
    }
  return m;
}
