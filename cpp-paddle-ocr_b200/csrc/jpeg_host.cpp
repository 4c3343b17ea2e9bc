// Host half of the device JPEG decoder: marker parsing only (see jpeg.h).  The entropy-coded bytes are never decoded
// on the host; they are only scanned for RSTn markers so that every restart interval can be handed to its own thread.
#include <algorithm>
#include <cstring>

#include "jpeg.h"

namespace b200ocr {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct RawHuff { bool present = false; uint8_t counts[16]; uint8_t vals[256]; int nvals = 0; };

// canonical code assignment (JPEG Annex C) -> look-ahead table + per-length limits (Annex F.2.2.3)
bool build_lut(const RawHuff& h, JpegHuffLut* lut) {
  memset(lut, 0, sizeof *lut);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    const int cnt = h.counts[len - 1];
    lut->valoff[len] = k - code;
    if (cnt == 0) {
      lut->maxcode[len] = -1;
    } else {
      if (code + cnt > (1 << len)) return false;  // over-subscribed table
      for (int i = 0; i < cnt; ++i, ++k, ++code) {
        if (len <= 9) {
          const int first = code << (9 - len), n = 1 << (9 - len);
          for (int j = 0; j < n; ++j) lut->fast[first + j] = uint16_t((len << 8) | h.vals[k]);
        }
      }
      lut->maxcode[len] = code - 1;
    }
    code <<= 1;
  }
  lut->maxcode[17] = 0x7fffffff;
  lut->valoff[0] = 0;
  lut->maxcode[0] = -1;
  memcpy(lut->vals, h.vals, 256);
  return true;
}

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// EXIF orientation (cv::imread applies it): returns 1 when absent
int exif_orientation(const uint8_t* seg, int len) {
  if (len < 14 || memcmp(seg, "Exif\0\0", 6) != 0) return 1;
  const uint8_t* t = seg + 6;
  const int n = len - 6;
  const bool le = t[0] == 'I';
  auto u16 = [&](int o) { return le ? (t[o] | (t[o + 1] << 8)) : ((t[o] << 8) | t[o + 1]); };
  auto u32 = [&](int o) { return le ? (u16(o) | (u16(o + 2) << 16)) : ((u16(o) << 16) | u16(o + 2)); };
  if (n < 8) return 1;
  const int ifd = u32(4);
  if (ifd < 0 || ifd + 2 > n) return 1;
  const int cnt = u16(ifd);
  for (int i = 0; i < cnt; ++i) {
    const int e = ifd + 2 + 12 * i;
    if (e + 12 > n) break;
    if (u16(e) == 0x0112) return u16(e + 8);
  }
  return 1;
}

}  // namespace

bool jpeg_parse(const uint8_t* data, size_t size, JpegImage* img, size_t* ecs_begin, size_t* ecs_end,
                std::vector<JpegSeg>* segs, std::string* why) {
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG stream");
  uint16_t qt[4][64];
  bool have_q[4] = {false, false, false, false};
  RawHuff dc[4], ac[4];
  int width = 0, height = 0, ncomp = 0, restart = 0;
  struct FC { int id, hs, vs, tq; } fc[3];
  bool have_frame = false;
  size_t pos = 2;
  int sel_dc[3] = {0, 0, 0}, sel_ac[3] = {0, 0, 0};
  bool have_scan = false;
  while (pos + 4 <= size) {
    if (data[pos] != 0xFF) return fail("marker expected");
    const int m = data[pos + 1];
    if (m == 0xFF) { ++pos; continue; }
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { pos += 2; continue; }  // stand-alone markers
    if (m == 0xD9) break;
    const int len = be16(data + pos + 2);
    if (len < 2 || pos + 2 + size_t(len) > size) return fail("truncated marker segment");
    const uint8_t* seg = data + pos + 4;
    const int n = len - 2;
    if (m == 0xDB) {
      int k = 0;
      while (k < n) {
        const int pq = seg[k] >> 4, tq = seg[k] & 15;
        if (tq > 3 || pq > 1) return fail("bad quantisation table");
        const int bytes = pq ? 128 : 64;
        if (k + 1 + bytes > n) return fail("truncated quantisation table");
        for (int i = 0; i < 64; ++i)
          qt[tq][kZigzag[i]] = pq ? uint16_t(be16(seg + k + 1 + 2 * i)) : seg[k + 1 + i];
        have_q[tq] = true;
        k += 1 + bytes;
      }
    } else if (m == 0xC4) {
      int k = 0;
      while (k < n) {
        if (k + 17 > n) return fail("truncated Huffman table");
        const int tc = seg[k] >> 4, th = seg[k] & 15;
        if (tc > 1 || th > 3) return fail("bad Huffman table id");
        RawHuff& h = tc ? ac[th] : dc[th];
        int total = 0;
        for (int i = 0; i < 16; ++i) { h.counts[i] = seg[k + 1 + i]; total += h.counts[i]; }
        if (total > 256 || k + 17 + total > n) return fail("truncated Huffman table");
        memset(h.vals, 0, sizeof h.vals);
        memcpy(h.vals, seg + k + 17, total);
        h.nvals = total;
        h.present = true;
        k += 17 + total;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (n < 6) return fail("truncated frame header");
      if (seg[0] != 8) return fail("only 8-bit samples are supported");
      height = be16(seg + 1);
      width = be16(seg + 3);
      ncomp = seg[5];
      if (ncomp != 1 && ncomp != 3) return fail("only grey and 3-component images are supported");
      if (n < 6 + 3 * ncomp) return fail("truncated frame header");
      for (int i = 0; i < ncomp; ++i)
        fc[i] = FC{seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i]};
      have_frame = true;
    } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
      return fail("only baseline / extended sequential Huffman JPEG is decoded on the device");
    } else if (m == 0xCC) {
      return fail("arithmetic coding is not supported");
    } else if (m == 0xDD) {
      if (n < 2) return fail("truncated DRI");
      restart = be16(seg);
    } else if (m == 0xE1) {
      if (exif_orientation(seg, n) != 1) return fail("EXIF orientation other than 1 is not applied on the device");
    } else if (m == 0xEE) {
      // Adobe marker: transform 0 with 3 components means RGB, not YCbCr
      if (n >= 12 && memcmp(seg, "Adobe", 5) == 0 && seg[11] != 1 && ncomp != 1)
        return fail("Adobe RGB / CMYK colour transforms are not supported");
    } else if (m == 0xDA) {
      if (!have_frame) return fail("scan before frame header");
      if (n < 1 || seg[0] != ncomp || n < 1 + 2 * ncomp + 3) return fail("only one interleaved scan is supported");
      for (int i = 0; i < ncomp; ++i) {
        int c = -1;
        for (int j = 0; j < ncomp; ++j)
          if (fc[j].id == seg[1 + 2 * i]) c = j;
        if (c != i) return fail("scan components out of frame order");
        sel_dc[i] = seg[2 + 2 * i] >> 4;
        sel_ac[i] = seg[2 + 2 * i] & 15;
        if (sel_dc[i] > 3 || sel_ac[i] > 3) return fail("bad table selector");
      }
      pos += 2 + size_t(len);
      have_scan = true;
      break;
    }
    pos += 2 + size_t(len);
  }
  if (!have_frame || !have_scan) return fail("no frame / scan header");
  if (width < 1 || height < 1) return fail("empty image");
  if (ncomp == 3 && fc[0].id == 'R' && fc[1].id == 'G' && fc[2].id == 'B') return fail("RGB-coded JPEG is not supported");
  memset(img, 0, sizeof *img);
  img->width = width; img->height = height; img->ncomp = ncomp; img->restart_interval = restart;
  int hmax = 1, vmax = 1;
  for (int i = 0; i < ncomp; ++i) { hmax = std::max(hmax, fc[i].hs); vmax = std::max(vmax, fc[i].vs); }
  if (ncomp == 1) { fc[0].hs = fc[0].vs = 1; hmax = vmax = 1; }  // a single-component scan is never interleaved
  for (int i = 0; i < ncomp; ++i) {
    const int fh = hmax / std::max(1, fc[i].hs), fv = vmax / std::max(1, fc[i].vs);
    const bool ok = fc[i].hs >= 1 && fc[i].vs >= 1 && hmax % fc[i].hs == 0 && vmax % fc[i].vs == 0 &&
                    ((fh == 1 && fv == 1) || (fh == 2 && fv == 1) || (fh == 2 && fv == 2));
    if (!ok) return fail("only 4:4:4, 4:2:2 and 4:2:0 sampling are supported");
    if (i > 0 && (fc[i].hs != 1 || fc[i].vs != 1)) return fail("chroma sampling factors other than 1x1 are not supported");
    if (!have_q[fc[i].tq & 3] || fc[i].tq > 3) return fail("missing quantisation table");
    if (!dc[sel_dc[i]].present || !ac[sel_ac[i]].present) return fail("missing Huffman table");
  }
  img->hmax = hmax; img->vmax = vmax;
  img->mcux = (width + 8 * hmax - 1) / (8 * hmax);
  img->mcuy = (height + 8 * vmax - 1) / (8 * vmax);
  for (int i = 0; i < ncomp; ++i) {
    JpegComp& c = img->comp[i];
    c.hs = fc[i].hs; c.vs = fc[i].vs;
    c.bw = img->mcux * c.hs; c.bh = img->mcuy * c.vs;
    c.cw = (width * c.hs + hmax - 1) / hmax;
    c.ch = (height * c.vs + vmax - 1) / vmax;
    memcpy(c.q, qt[fc[i].tq], sizeof c.q);
    if (!build_lut(dc[sel_dc[i]], &img->lut[2 * i]) || !build_lut(ac[sel_ac[i]], &img->lut[2 * i + 1]))
      return fail("invalid Huffman table");
  }
  // entropy-coded segment: up to the first marker that is neither a stuffed zero nor RSTn
  size_t e = pos;
  const long long nmcu = (long long)img->mcux * img->mcuy;
  segs->clear();
  JpegSeg cur{0, 0, 0, 0, 0};
  while (e + 1 < size) {
    const uint8_t* f = static_cast<const uint8_t*>(memchr(data + e, 0xFF, size - 1 - e));
    if (!f) { e = size; break; }
    e = size_t(f - data);
    const int nx = data[e + 1];
    if (nx == 0x00 || nx == 0xFF) { e += (nx == 0x00) ? 2 : 1; continue; }
    if (nx >= 0xD0 && nx <= 0xD7) {
      if (restart > 0) {
        cur.end = int(e - pos);
        cur.nmcu = restart;
        segs->push_back(cur);
        cur.mcu0 += restart;
        cur.begin = int(e + 2 - pos);
      }
      e += 2;
      continue;
    }
    break;  // EOI or any other marker
  }
  if (e > size) e = size;
  cur.end = int(e - pos);
  cur.nmcu = int(std::max<long long>(0, nmcu - cur.mcu0));
  if (restart <= 0) { cur.mcu0 = 0; cur.nmcu = int(nmcu); }
  if (cur.nmcu > 0) segs->push_back(cur);
  // restart intervals beyond the frame (corrupt stream) are clipped
  for (auto& s : *segs)
    if (s.mcu0 + s.nmcu > nmcu) s.nmcu = int(std::max<long long>(0, nmcu - s.mcu0));
  *ecs_begin = pos;
  *ecs_end = e;
  if (e - pos > size_t(0x7fffffff)) return fail("entropy-coded segment too large");
  return true;
}

}  // namespace b200ocr
