// jpeg_gray.hpp -- minimal sequential-DCT JPEG decoder returning the luminance plane.
//
// The reference reads its 2D structures with stbi_load(name, &w, &h, &channels, 1)
// (Keff2D/keff2D.h:300): 8-bit, one channel forced.  This is an independent implementation of the
// same job, written from the JPEG standard (ITU-T T.81): baseline / extended-sequential Huffman
// JPEG, 8-bit samples, 1 or 3 components with any sampling factors, restart intervals.  Only the
// first component (Y, or the single grey channel) is reconstructed -- for a YCbCr file that is what
// a 1-channel decode means up to rounding, and the reference insists on 1-channel files anyway
// (keff2D.cpp:60-63).  Progressive files are rejected with a message.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace kbhost {

struct GrayImage {
  int w = 0, h = 0, channels = 0;
  std::vector<uint8_t> pix;  // h*w, row-major
  std::string err;
};

namespace jpeg_detail {

static const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  bool present = false;
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  void build(const uint8_t* counts, const uint8_t* symbols, int nsym) {
    memcpy(vals, symbols, nsym);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      code += counts[len - 1];
      k += counts[len - 1];
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

struct Bits {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int cnt = 0;
  bool hit_marker = false;
  int bit() {
    if (cnt == 0) {
      int b = 0;
      if (p < end && !hit_marker) {
        b = *p++;
        if (b == 0xFF) {
          if (p < end && *p == 0x00) p++;           // stuffed byte
          else { hit_marker = true; p--; b = 0; }   // a marker: feed zeros until it is consumed
        }
      }
      acc = (uint32_t)b;
      cnt = 8;
    }
    cnt--;
    return (acc >> cnt) & 1;
  }
  int bits(int n) {
    int v = 0;
    while (n--) v = (v << 1) | bit();
    return v;
  }
  void restart() {  // byte-align and swallow an RSTn marker
    cnt = 0;
    hit_marker = false;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) p++;
    if (p + 1 < end) p += 2;
  }
};

static inline int decode_symbol(Bits& br, const Huff& h) {
  int code = 0;
  for (int len = 1; len <= 16; len++) {
    code = (code << 1) | br.bit();
    if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len])
      return h.vals[h.valptr[len] + code - h.mincode[len]];
  }
  return -1;
}

static inline int extend(int v, int t) { return t && v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

static void idct8x8(const int* coef, uint8_t* out, int stride) {
  static double c[8][8];
  static bool init = false;
  if (!init) {
    for (int x = 0; x < 8; x++)
      for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0);
    init = true;
  }
  double tmp[64];
  for (int v = 0; v < 8; v++)  // rows: over u
    for (int x = 0; x < 8; x++) {
      double s = 0;
      for (int u = 0; u < 8; u++) s += c[x][u] * coef[v * 8 + u];
      tmp[v * 8 + x] = s;
    }
  for (int x = 0; x < 8; x++)  // columns: over v
    for (int y = 0; y < 8; y++) {
      double s = 0;
      for (int v = 0; v < 8; v++) s += c[y][v] * tmp[v * 8 + x];
      int p = (int)std::lround(s + 128.0);
      out[y * stride + x] = (uint8_t)(p < 0 ? 0 : p > 255 ? 255 : p);
    }
}

}  // namespace jpeg_detail

inline bool decode_jpeg_gray(const uint8_t* d, size_t n, GrayImage& out) {
  using namespace jpeg_detail;
  auto fail = [&](const char* m) { out.err = m; return false; };
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG file (no SOI marker)");
  uint16_t qt[4][64];
  bool qt_ok[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  struct Comp { int id, h, v, tq, td, ta, pred; } comp[4];
  int ncomp = 0, W = 0, H = 0, restart = 0;
  size_t i = 2;
  while (i + 4 <= n) {
    if (d[i] != 0xFF) { i++; continue; }
    const int m = d[i + 1];
    if (m == 0xFF) { i++; continue; }
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { i += 2; continue; }
    if (m == 0xD9) break;
    const size_t len = ((size_t)d[i + 2] << 8) | d[i + 3];
    if (len < 2 || i + 2 + len > n) return fail("truncated JPEG segment");
    const uint8_t* s = d + i + 4;
    const size_t sl = len - 2;
    if (m == 0xDB) {  // DQT
      size_t k = 0;
      while (k < sl) {
        const int pq = s[k] >> 4, tq = s[k] & 15;
        k++;
        if (tq > 3 || k + (pq ? 128 : 64) > sl) return fail("bad DQT");
        for (int z = 0; z < 64; z++) {
          qt[tq][ZIGZAG[z]] = pq ? (uint16_t)((s[k] << 8) | s[k + 1]) : s[k];
          k += pq ? 2 : 1;
        }
        qt_ok[tq] = true;
      }
    } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
      if (sl < 6 || s[0] != 8) return fail("only 8-bit JPEG samples are supported");
      H = (s[1] << 8) | s[2];
      W = (s[3] << 8) | s[4];
      ncomp = s[5];
      if ((ncomp != 1 && ncomp != 3) || sl < (size_t)6 + 3 * ncomp || W <= 0 || H <= 0) return fail("bad SOF");
      for (int c = 0; c < ncomp; c++) {
        comp[c].id = s[6 + 3 * c];
        comp[c].h = s[7 + 3 * c] >> 4;
        comp[c].v = s[7 + 3 * c] & 15;
        comp[c].tq = s[8 + 3 * c] & 3;
        if (comp[c].h < 1 || comp[c].v < 1 || comp[c].h > 4 || comp[c].v > 4) return fail("bad sampling factors");
      }
    } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC && m != 0xC4)) {
      return fail("progressive / lossless / arithmetic JPEG is not supported: re-save as baseline JPEG");
    } else if (m == 0xC4) {  // DHT
      size_t k = 0;
      while (k + 17 <= sl) {
        const int tc = s[k] >> 4, th = s[k] & 15;
        int nsym = 0;
        for (int q = 0; q < 16; q++) nsym += s[k + 1 + q];
        if (th > 3 || tc > 1 || nsym > 256 || k + 17 + nsym > sl) return fail("bad DHT");
        (tc ? ac[th] : dc[th]).build(s + k + 1, s + k + 17, nsym);
        k += 17 + nsym;
      }
    } else if (m == 0xDD) {
      if (sl >= 2) restart = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {  // SOS: the (single) sequential scan
      if (!ncomp) return fail("SOS before SOF");
      const int ns = s[0];
      if (ns != ncomp || sl < (size_t)1 + 2 * ns + 3) return fail("multi-scan sequential JPEG is not supported");
      for (int k = 0; k < ns; k++) {
        int c = 0;
        while (c < ncomp && comp[c].id != s[1 + 2 * k]) c++;
        if (c == ncomp) return fail("bad SOS component");
        comp[c].td = s[2 + 2 * k] >> 4;
        comp[c].ta = s[2 + 2 * k] & 15;
        if (comp[c].td > 3 || comp[c].ta > 3 || !dc[comp[c].td].present || !ac[comp[c].ta].present || !qt_ok[comp[c].tq])
          return fail("scan refers to a missing table");
        comp[c].pred = 0;
      }
      int hmax = 1, vmax = 1;
      for (int c = 0; c < ncomp; c++) { hmax = comp[c].h > hmax ? comp[c].h : hmax; vmax = comp[c].v > vmax ? comp[c].v : vmax; }
      if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }  // single component: non-interleaved
      const int mcuw = 8 * hmax, mcuh = 8 * vmax;
      const int mx = (W + mcuw - 1) / mcuw, my = (H + mcuh - 1) / mcuh;
      // luminance plane at component-0 resolution (padded to whole MCUs)
      const int pw = mx * 8 * comp[0].h, ph = my * 8 * comp[0].v;
      std::vector<uint8_t> plane((size_t)pw * ph);
      Bits br{d + i + 2 + len, d + n};
      int count = 0;
      for (int yb = 0; yb < my; yb++)
        for (int xb = 0; xb < mx; xb++) {
          if (restart && count && count % restart == 0) {
            br.restart();
            for (int c = 0; c < ncomp; c++) comp[c].pred = 0;
          }
          count++;
          for (int c = 0; c < ncomp; c++)
            for (int by = 0; by < comp[c].v; by++)
              for (int bx = 0; bx < comp[c].h; bx++) {
                int blk[64] = {0};
                int t = decode_symbol(br, dc[comp[c].td]);
                if (t < 0 || t > 15) return fail("corrupt JPEG data (DC)");
                comp[c].pred += extend(br.bits(t), t);
                blk[0] = comp[c].pred * qt[comp[c].tq][0];
                for (int k = 1; k < 64;) {
                  const int rs = decode_symbol(br, ac[comp[c].ta]);
                  if (rs < 0) return fail("corrupt JPEG data (AC)");
                  const int r = rs >> 4, sz = rs & 15;
                  if (sz == 0) {
                    if (r == 15) { k += 16; continue; }
                    break;
                  }
                  k += r;
                  if (k > 63) return fail("corrupt JPEG data (run)");
                  blk[ZIGZAG[k]] = extend(br.bits(sz), sz) * qt[comp[c].tq][ZIGZAG[k]];
                  k++;
                }
                if (c == 0)
                  idct8x8(blk, &plane[(size_t)(yb * comp[0].v + by) * 8 * pw + (size_t)(xb * comp[0].h + bx) * 8], pw);
              }
        }
      // component 0 may itself be subsampled relative to the image (rare): nearest-neighbour expand
      out.w = W;
      out.h = H;
      out.channels = ncomp;
      out.pix.resize((size_t)W * H);
      for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
          out.pix[(size_t)y * W + x] = plane[(size_t)(y * comp[0].v / vmax) * pw + (size_t)(x * comp[0].h / hmax)];
      return true;
    }
    i += 2 + len;
  }
  return fail("no image data (SOS marker) found");
}

inline bool load_jpeg_gray(const std::string& path, GrayImage& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { out.err = "cannot open " + path; return false; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t r;
  while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
  fclose(f);
  return decode_jpeg_gray(buf.data(), buf.size(), out);
}

}  // namespace kbhost
