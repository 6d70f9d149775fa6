// image_gray.hpp -- 8-bit single-channel image reader of the host programs.
//
// The reference reads its 2D structures with stbi_load(name, &w, &h, &channels, 1)
// (Keff2D/keff2D.h:300): whatever the file holds, one 8-bit channel comes back, and the structure is that
// channel thresholded at 150 (keff2D.cpp:108-121).  A pixel that decodes one grey level off can land in the
// other phase, so this reader reproduces the reference decoder's output bit for bit on the formats a
// structure image realistically arrives in:
//   JPEG  baseline, extended-sequential and PROGRESSIVE Huffman (ITU-T T.81 incl. Annex G), 8-bit, 1 or 3
//         components, any sampling factors, restart intervals, multi-scan.  Inverse DCT: the Loeffler-
//         Ligtenberg-Moschytz integer transform of IJG's jidctint at the fixed-point precision the reference's
//         decoder uses (12-bit constants, 2 extra bits after the column pass) -- same integers, same pixels.
//         One channel of a YCbCr file = its Y plane; of an RGB-tagged file = (77 R + 150 G + 29 B) >> 8.
//   PNG   (ISO/IEC 15948) all colour types and bit depths, Adam7 interlace; zlib inflate included.
//   BMP   1/4/8-bit palettes, 24- and 32-bit BI_RGB / BI_BITFIELDS with 8-bit masks, either row order.
//   PNM   binary P5 / P6, maxval <= 255.
// Colour reduces to one channel by (77 R + 150 G + 29 B) >> 8, 16-bit samples by their high byte (after the
// reduction, as the reference does).  Written from the format specifications; checked against the reference's
// decoder by tests/test_host_cpp.py through oracle/_ref/stbgray (test infrastructure, never linked here).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace kbhost {

struct GrayImage {
  int w = 0, h = 0, channels = 0;  // channels: as stored in the file (the reference's `channels`)
  std::vector<uint8_t> pix;        // h*w, row-major
  std::string err;
};

namespace img_detail {

inline uint8_t luma8(int r, int g, int b) { return (uint8_t)((r * 77 + g * 150 + 29 * b) >> 8); }
inline uint16_t luma16(int r, int g, int b) { return (uint16_t)((r * 77 + g * 150 + 29 * b) >> 8); }

// ---------------------------------------------------------------------------------------------------------
// JPEG
// ---------------------------------------------------------------------------------------------------------
static const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  bool present = false;
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  int16_t fast[512];  // 9-bit prefix -> (length << 8) | symbol, or -1
  void build(const uint8_t* counts, const uint8_t* symbols, int nsym) {
    memcpy(vals, symbols, nsym);
    for (int i = 0; i < 512; i++) fast[i] = -1;
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      for (int j = 0; j < counts[len - 1]; j++, code++, k++)
        if (len <= 9)
          for (int f = 0; f < (1 << (9 - len)); f++) fast[(code << (9 - len)) | f] = (int16_t)((len << 8) | vals[k]);
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

// entropy-coded segment reader: byte stuffing removed, zeros once a marker is reached
struct Bits {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int cnt = 0;
  bool hit_marker = false;
  void fill() {
    while (cnt <= 56) {
      unsigned b = 0;
      if (p < end && !hit_marker) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) p += 2;
          else { hit_marker = true; b = 0; }
        } else p++;
      }
      acc |= (uint64_t)b << (56 - cnt);
      cnt += 8;
    }
  }
  unsigned peek(int n) {  // n <= 16
    if (cnt < n) fill();
    return (unsigned)(acc >> (64 - n));
  }
  void skip(int n) { acc <<= n; cnt -= n; }
  int bits(int n) {
    if (n == 0) return 0;
    const unsigned v = peek(n);
    skip(n);
    return (int)v;
  }
  int bit() { return bits(1); }
  void restart() {  // byte-align and swallow the RSTn marker
    acc = 0;
    cnt = 0;
    hit_marker = false;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) p++;
    if (p + 1 < end) p += 2;
  }
};

static inline int decode_symbol(Bits& br, const Huff& h) {
  const int f = h.fast[br.peek(9)];
  if (f >= 0) {
    br.skip(f >> 8);
    return f & 255;
  }
  const unsigned w = br.peek(16);
  for (int len = 10; len <= 16; len++) {
    const int code = (int)(w >> (16 - len));
    if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) {
      br.skip(len);
      return h.vals[h.valptr[len] + code - h.mincode[len]];
    }
  }
  return -1;
}

static inline int extend(int v, int t) { return t && v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

// 1-D LL&M inverse DCT on eight values, constants scaled by 2^12 (IJG jidctint "islow"); out[k], out[7-k] = e[k] +- o[k]
static inline void llm_1d(const int* s, int stride, int (&e)[4], int (&o)[4]) {
  constexpr int F = 4096;
  // (int)(c * 4096 + 0.5) of the single-precision constant, truncating -- the reference decoder's rounding of them
  constexpr int c0541 = (int)((double)0.5411961f * F + 0.5), c1847 = (int)((double)-1.847759065f * F + 0.5);
  constexpr int c0765 = (int)((double)0.765366865f * F + 0.5), c1175 = (int)((double)1.175875602f * F + 0.5);
  constexpr int c0298 = (int)((double)0.298631336f * F + 0.5), c2053 = (int)((double)2.053119869f * F + 0.5);
  constexpr int c3072 = (int)((double)3.072711026f * F + 0.5), c1501 = (int)((double)1.501321110f * F + 0.5);
  constexpr int c0899 = (int)((double)-0.899976223f * F + 0.5), c2562 = (int)((double)-2.562915447f * F + 0.5);
  constexpr int c1961 = (int)((double)-1.961570560f * F + 0.5), c0390 = (int)((double)-0.390180644f * F + 0.5);
  const int s0 = s[0], s1 = s[stride], s2 = s[2 * stride], s3 = s[3 * stride], s4 = s[4 * stride], s5 = s[5 * stride],
            s6 = s[6 * stride], s7 = s[7 * stride];
  // even part
  const int z1 = (s2 + s6) * c0541, t2 = z1 + s6 * c1847, t3 = z1 + s2 * c0765;
  const int t0 = (s0 + s4) * F, t1 = (s0 - s4) * F;
  e[0] = t0 + t3;
  e[3] = t0 - t3;
  e[1] = t1 + t2;
  e[2] = t1 - t2;
  // odd part
  const int z3 = s7 + s3, z4 = s5 + s1, z1o = s7 + s1, z2o = s5 + s3, z5 = (z3 + z4) * c1175;
  const int a1 = z5 + z1o * c0899, a2 = z5 + z2o * c2562, a3 = z3 * c1961, a4 = z4 * c0390;
  o[3] = s7 * c0298 + (a1 + a3);  // pairs with e[3]
  o[2] = s5 * c2053 + (a2 + a4);
  o[1] = s3 * c3072 + (a2 + a3);
  o[0] = s1 * c1501 + (a1 + a4);
}

static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

// dequantised coefficients (natural order) -> 8x8 samples
static void idct8x8(const int16_t* coef, uint8_t* out, int stride) {
  int col[64], in[64];
  for (int i = 0; i < 64; i++) in[i] = coef[i];
  for (int x = 0; x < 8; x++) {  // columns; keep two extra bits
    if (!(in[x + 8] | in[x + 16] | in[x + 24] | in[x + 32] | in[x + 40] | in[x + 48] | in[x + 56])) {
      const int dc = in[x] * 4;  // a column with only its DC term: (dc * 4096 + 512) >> 10 in every row
      for (int k = 0; k < 8; k++) col[k * 8 + x] = dc;
      continue;
    }
    int e[4], o[4];
    llm_1d(in + x, 8, e, o);
    for (int k = 0; k < 4; k++) {
      col[k * 8 + x] = (e[k] + 512 + o[k]) >> 10;
      col[(7 - k) * 8 + x] = (e[k] + 512 - o[k]) >> 10;
    }
  }
  for (int y = 0; y < 8; y++) {  // rows; 2^17 to remove, +128 level shift folded into the rounding term
    int e[4], o[4];
    llm_1d(col + y * 8, 1, e, o);
    const int bias = 65536 + (128 << 17);
    for (int k = 0; k < 4; k++) {
      out[y * stride + k] = clamp8((e[k] + bias + o[k]) >> 17);
      out[y * stride + 7 - k] = clamp8((e[k] + bias - o[k]) >> 17);
    }
  }
}

struct JComp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0;
  int bw = 0, bh = 0;      // blocks per row / column of the coefficient store (whole MCUs)
  int cw = 0, ch = 0;      // the component's own size in samples
  std::vector<int16_t> coef;
};

inline bool decode_jpeg(const uint8_t* d, size_t n, GrayImage& out) {
  auto fail = [&](const char* m) { out.err = m; return false; };
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG file (no SOI marker)");
  uint16_t qt[4][64];
  bool qt_ok[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  JComp comp[4];
  int ncomp = 0, W = 0, H = 0, restart = 0, hmax = 1, vmax = 1, mx = 0, my = 0;
  bool progressive = false, have_frame = false, saw_scan = false;
  int app14_transform = -1;
  size_t i = 2;
  while (i + 4 <= n) {
    if (d[i] != 0xFF) { i++; continue; }
    const int m = d[i + 1];
    if (m == 0xFF) { i++; continue; }
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7) || m == 0x00) { i += 2; continue; }
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
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // SOF0 / SOF1 / SOF2
      if (have_frame) return fail("more than one frame header");
      progressive = m == 0xC2;
      if (sl < 6 || s[0] != 8) return fail("only 8-bit JPEG samples are supported");
      H = (s[1] << 8) | s[2];
      W = (s[3] << 8) | s[4];
      ncomp = s[5];
      if ((ncomp != 1 && ncomp != 3) || sl < (size_t)6 + 3 * ncomp || W <= 0 || H <= 0)
        return fail("bad frame header (1 or 3 components expected)");
      for (int c = 0; c < ncomp; c++) {
        comp[c].id = s[6 + 3 * c];
        comp[c].h = s[7 + 3 * c] >> 4;
        comp[c].v = s[7 + 3 * c] & 15;
        comp[c].tq = s[8 + 3 * c] & 3;
        if (comp[c].h < 1 || comp[c].v < 1 || comp[c].h > 4 || comp[c].v > 4) return fail("bad sampling factors");
        hmax = comp[c].h > hmax ? comp[c].h : hmax;
        vmax = comp[c].v > vmax ? comp[c].v : vmax;
      }
      mx = (W + 8 * hmax - 1) / (8 * hmax);
      my = (H + 8 * vmax - 1) / (8 * vmax);
      for (int c = 0; c < ncomp; c++) {
        comp[c].bw = mx * comp[c].h;
        comp[c].bh = my * comp[c].v;
        comp[c].cw = (W * comp[c].h + hmax - 1) / hmax;
        comp[c].ch = (H * comp[c].v + vmax - 1) / vmax;
        comp[c].coef.assign((size_t)comp[c].bw * comp[c].bh * 64, 0);
      }
      have_frame = true;
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return fail("lossless / hierarchical / arithmetic-coded JPEG is not supported");
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
    } else if (m == 0xEE) {  // Adobe APP14: colour transform flag
      if (sl >= 12 && memcmp(s, "Adobe", 5) == 0) app14_transform = s[11];
    } else if (m == 0xDA) {  // SOS: one scan
      if (!have_frame) return fail("SOS before SOF");
      const int ns = s[0];
      if (ns < 1 || ns > ncomp || sl < (size_t)1 + 2 * ns + 3) return fail("bad SOS");
      int sc[4];
      for (int k = 0; k < ns; k++) {
        int c = 0;
        while (c < ncomp && comp[c].id != s[1 + 2 * k]) c++;
        if (c == ncomp) return fail("bad SOS component");
        comp[c].td = s[2 + 2 * k] >> 4;
        comp[c].ta = s[2 + 2 * k] & 15;
        if (comp[c].td > 3 || comp[c].ta > 3) return fail("bad SOS table index");
        comp[c].pred = 0;
        sc[k] = c;
      }
      int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns];
      const int Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
      if (!progressive) { Ss = 0; Se = 63; }
      if (progressive && (Ss > 63 || Se > 63 || Ss > Se || Ah > 13 || Al > 13 || (Ss > 0 && ns != 1) || (Ss == 0 && Se != 0)))
        return fail("bad progressive scan parameters");
      for (int k = 0; k < ns; k++) {
        const JComp& C = comp[sc[k]];
        if ((Ss == 0 && !(progressive && Ah) && !dc[C.td].present) || (Se > 0 && !ac[C.ta].present))
          return fail("scan refers to a missing Huffman table");
      }
      Bits br{d + i + 2 + len, d + n};
      int eobrun = 0;
      // one 8x8 block of component C at block coordinates (bx, by)
      auto block = [&](JComp& C, int bx, int by) -> bool {
        int16_t* q = &C.coef[((size_t)by * C.bw + bx) * 64];
        if (!progressive) {
          const int t = decode_symbol(br, dc[C.td]);
          if (t < 0 || t > 15) return false;
          C.pred += extend(br.bits(t), t);
          q[0] = (int16_t)C.pred;
          for (int k = 1; k < 64;) {
            const int rs = decode_symbol(br, ac[C.ta]);
            if (rs < 0) return false;
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
              if (r == 15) { k += 16; continue; }
              break;
            }
            k += r;
            if (k > 63) return false;
            q[ZIGZAG[k]] = (int16_t)extend(br.bits(sz), sz);
            k++;
          }
          return true;
        }
        if (Ss == 0) {  // DC scan
          if (Ah == 0) {
            const int t = decode_symbol(br, dc[C.td]);
            if (t < 0 || t > 15) return false;
            C.pred += extend(br.bits(t), t);
            q[0] = (int16_t)(C.pred * (1 << Al));
          } else if (br.bit()) {
            q[0] = (int16_t)(q[0] + (1 << Al));
          }
          return true;
        }
        if (Ah == 0) {  // AC first pass of a band
          if (eobrun) { eobrun--; return true; }
          for (int k = Ss; k <= Se;) {
            const int rs = decode_symbol(br, ac[C.ta]);
            if (rs < 0) return false;
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
              if (r < 15) {
                eobrun = (1 << r) - 1;
                if (r) eobrun += br.bits(r);
                break;
              }
              k += 16;
            } else {
              k += r;
              if (k > 63) return false;
              q[ZIGZAG[k]] = (int16_t)(extend(br.bits(sz), sz) * (1 << Al));
              k++;
            }
          }
          return true;
        }
        // AC refinement: one more bit of every coefficient of the band that is already non-zero, new +-1 values
        const int bitv = 1 << Al;
        auto refine = [&](int16_t& v) {
          if (br.bit() && (v & bitv) == 0) v = (int16_t)(v > 0 ? v + bitv : v - bitv);
        };
        if (eobrun) {
          eobrun--;
          for (int k = Ss; k <= Se; k++)
            if (q[ZIGZAG[k]]) refine(q[ZIGZAG[k]]);
          return true;
        }
        int k = Ss;
        do {
          const int rs = decode_symbol(br, ac[C.ta]);
          if (rs < 0) return false;
          int r = rs >> 4, sz = rs & 15, nv = 0;
          if (sz == 0) {
            if (r < 15) {
              eobrun = (1 << r) - 1;
              if (r) eobrun += br.bits(r);
              r = 64;  // run to the end of the band, refining what is there
            }
          } else {
            if (sz != 1) return false;
            nv = br.bit() ? bitv : -bitv;
          }
          while (k <= Se) {
            int16_t& v = q[ZIGZAG[k++]];
            if (v) refine(v);
            else {
              if (r == 0) { v = (int16_t)nv; break; }
              r--;
            }
          }
        } while (k <= Se);
        return true;
      };
      int count = 0;
      auto maybe_restart = [&]() {
        if (restart && count && count % restart == 0) {
          br.restart();
          for (int c = 0; c < ncomp; c++) comp[c].pred = 0;
          eobrun = 0;
        }
        count++;
      };
      if (ns == 1) {  // non-interleaved: the component's own blocks, row by row
        JComp& C = comp[sc[0]];
        const int nbx = (C.cw + 7) / 8, nby = (C.ch + 7) / 8;
        for (int by = 0; by < nby; by++)
          for (int bx = 0; bx < nbx; bx++) {
            maybe_restart();
            if (!block(C, bx, by)) return fail("corrupt JPEG data");
          }
      } else {
        for (int yb = 0; yb < my; yb++)
          for (int xb = 0; xb < mx; xb++) {
            maybe_restart();
            for (int k = 0; k < ns; k++) {
              JComp& C = comp[sc[k]];
              for (int by = 0; by < C.v; by++)
                for (int bx = 0; bx < C.h; bx++)
                  if (!block(C, xb * C.h + bx, yb * C.v + by)) return fail("corrupt JPEG data");
            }
          }
      }
      saw_scan = true;
      // continue behind the entropy-coded data: the next marker that is not a restart / stuffed byte
      size_t j = i + 2 + len;
      while (j + 1 < n && !(d[j] == 0xFF && d[j + 1] != 0x00 && !(d[j + 1] >= 0xD0 && d[j + 1] <= 0xD7) && d[j + 1] != 0xFF)) j++;
      i = j;
      continue;
    }
    i += 2 + len;
  }
  if (!saw_scan) return fail("no image data (SOS marker) found");
  // which planes: Y alone for a YCbCr (or grey) file; all three for an RGB-tagged file
  const bool rgb = ncomp == 3 && (app14_transform == 0 || (app14_transform < 0 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B'));
  const int nplanes = rgb ? 3 : 1;
  std::vector<uint8_t> plane[3];
  for (int c = 0; c < nplanes; c++) {
    JComp& C = comp[c];
    if (!qt_ok[C.tq]) return fail("frame refers to a missing quantisation table");
    const int pw = C.bw * 8;
    plane[c].resize((size_t)pw * C.bh * 8);
    int16_t blk[64];
    for (int by = 0; by < C.bh; by++)
      for (int bx = 0; bx < C.bw; bx++) {
        const int16_t* q = &C.coef[((size_t)by * C.bw + bx) * 64];
        for (int k = 0; k < 64; k++) blk[k] = (int16_t)(q[k] * qt[C.tq][k]);
        idct8x8(blk, &plane[c][(size_t)by * 8 * pw + (size_t)bx * 8], pw);
      }
  }
  out.w = W;
  out.h = H;
  out.channels = ncomp;
  out.pix.resize((size_t)W * H);
  // a plane that is subsampled relative to the image (rare for Y / R,G,B) is expanded by replication
  auto at = [&](int c, int x, int y) {
    const JComp& C = comp[c];
    return plane[c][(size_t)(y * C.v / vmax) * (C.bw * 8) + (size_t)(x * C.h / hmax)];
  };
  if (!rgb && comp[0].h == hmax && comp[0].v == vmax) {  // the usual case: the plane is at image resolution
    for (int y = 0; y < H; y++) memcpy(&out.pix[(size_t)y * W], &plane[0][(size_t)y * comp[0].bw * 8], (size_t)W);
    return true;
  }
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      out.pix[(size_t)y * W + x] = rgb ? luma8(at(0, x, y), at(1, x, y), at(2, x, y)) : at(0, x, y);
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// zlib inflate (RFC 1950 / 1951)
// ---------------------------------------------------------------------------------------------------------
struct Inflate {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int cnt = 0;
  bool bad = false;
  int bits(int n) {
    while (cnt < n) {
      const unsigned b = p < end ? *p++ : (bad = true, 0u);
      acc |= b << cnt;
      cnt += 8;
    }
    const int v = (int)(acc & ((1u << n) - 1));
    acc >>= n;
    cnt -= n;
    return v;
  }
  struct Table {
    uint16_t count[16], sym[288];
    void build(const uint8_t* len, int n) {
      memset(count, 0, sizeof count);
      for (int i = 0; i < n; i++) count[len[i]]++;
      count[0] = 0;
      uint16_t off[16];
      off[1] = 0;
      for (int i = 1; i < 15; i++) off[i + 1] = off[i] + count[i];
      for (int i = 0; i < n; i++)
        if (len[i]) sym[off[len[i]]++] = (uint16_t)i;
    }
  };
  int decode(const Table& t) {  // canonical code, one bit at a time
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
      code |= bits(1);
      const int c = t.count[len];
      if (code - c < first) return t.sym[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    bad = true;
    return -1;
  }
  bool run(std::vector<uint8_t>& out) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (end - p < 2) return false;
    const int cmf = p[0], flg = p[1];
    if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) return false;
    p += 2;
    int last;
    do {
      last = bits(1);
      const int type = bits(2);
      if (type == 0) {
        acc = 0;
        cnt = 0;
        if (end - p < 4) return false;
        const int len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
        p += 4;
        if ((len ^ 0xffff) != nlen || end - p < len) return false;
        out.insert(out.end(), p, p + len);
        p += len;
      } else if (type == 1 || type == 2) {
        Table tl, td;
        uint8_t lens[320];
        if (type == 1) {
          for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
          tl.build(lens, 288);
          for (int i = 0; i < 30; i++) lens[i] = 5;
          td.build(lens, 30);
        } else {
          static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
          const int hlit = bits(5) + 257, hdist = bits(5) + 1, hclen = bits(4) + 4;
          uint8_t cl[19] = {0};
          for (int i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)bits(3);
          Table tc;
          tc.build(cl, 19);
          int k = 0;
          while (k < hlit + hdist && !bad) {
            const int s = decode(tc);
            if (s < 0) return false;
            if (s < 16) lens[k++] = (uint8_t)s;
            else {
              int rep, v = 0;
              if (s == 16) {
                if (k == 0) return false;
                v = lens[k - 1];
                rep = 3 + bits(2);
              } else if (s == 17) rep = 3 + bits(3);
              else rep = 11 + bits(7);
              if (k + rep > hlit + hdist) return false;
              while (rep--) lens[k++] = (uint8_t)v;
            }
          }
          tl.build(lens, hlit);
          td.build(lens + hlit, hdist);
        }
        while (!bad) {
          const int s = decode(tl);
          if (s < 0) return false;
          if (s < 256) out.push_back((uint8_t)s);
          else if (s == 256) break;
          else {
            if (s > 285) return false;
            const int len = lbase[s - 257] + bits(lext[s - 257]);
            const int ds = decode(td);
            if (ds < 0 || ds > 29) return false;
            const size_t dist = dbase[ds] + (size_t)bits(dext[ds]);
            if (dist > out.size()) return false;
            const size_t from = out.size() - dist;
            for (int q = 0; q < len; q++) out.push_back(out[from + q]);
          }
        }
      } else return false;
    } while (!last && !bad);
    return !bad;
  }
};

// ---------------------------------------------------------------------------------------------------------
// PNG
// ---------------------------------------------------------------------------------------------------------
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

inline bool decode_png(const uint8_t* d, size_t n, GrayImage& out) {
  auto fail = [&](const char* m) { out.err = m; return false; };
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (n < 8 || memcmp(d, sig, 8) != 0) return fail("not a PNG file");
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> z, pal;
  bool have_hdr = false, have_trns_pal = false;
  size_t i = 8;
  while (i + 12 <= n) {
    const uint32_t len = be32(d + i);
    const uint8_t* t = d + i + 4;
    const uint8_t* body = d + i + 8;
    if ((size_t)len > n - i - 12) return fail("truncated PNG chunk");
    if (!memcmp(t, "IHDR", 4)) {
      if (len != 13) return fail("bad IHDR");
      W = be32(body);
      H = be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
      if (body[10] || body[11] || interlace > 1 || W == 0 || H == 0 || W > (1u << 24) || H > (1u << 24)) return fail("bad IHDR");
      if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) return fail("bad PNG bit depth");
      if (!(ctype == 0 || ctype == 2 || ctype == 3 || ctype == 4 || ctype == 6)) return fail("bad PNG colour type");
      if ((ctype == 3 && depth == 16) || ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8)) return fail("bad PNG colour type / depth");
      have_hdr = true;
    } else if (!memcmp(t, "PLTE", 4)) {
      pal.assign(body, body + len);
    } else if (!memcmp(t, "tRNS", 4)) {
      have_trns_pal = ctype == 3;
    } else if (!memcmp(t, "IDAT", 4)) {
      z.insert(z.end(), body, body + len);
    } else if (!memcmp(t, "IEND", 4)) {
      break;
    }
    i += 12 + (size_t)len;
  }
  if (!have_hdr || z.empty()) return fail("PNG without header or image data");
  const int nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
  if (ctype == 3 && pal.size() < 3) return fail("palette PNG without PLTE");
  std::vector<uint8_t> raw;
  raw.reserve((size_t)H * ((size_t)W * nch * depth / 8 + 2));
  Inflate inf{z.data(), z.data() + z.size()};
  if (!inf.run(raw)) return fail("corrupt PNG data (inflate)");
  out.w = (int)W;
  out.h = (int)H;
  out.channels = ctype == 3 ? (have_trns_pal ? 4 : 3) : nch;
  out.pix.assign((size_t)W * H, 0);
  const int bpp_bits = nch * depth, fb = (bpp_bits + 7) / 8;  // filter unit in bytes
  static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
  size_t pos = 0;
  std::vector<uint8_t> cur, prev;
  for (int pass = 0; pass < (interlace ? 7 : 1); pass++) {
    const uint32_t pw = interlace ? (W - xs[pass] + dx[pass] - 1) / dx[pass] : W;
    const uint32_t ph = interlace ? (H - ys[pass] + dy[pass] - 1) / dy[pass] : H;
    if (interlace && (W <= (uint32_t)xs[pass] || H <= (uint32_t)ys[pass])) continue;
    if (pw == 0 || ph == 0) continue;
    const size_t rb = ((size_t)pw * bpp_bits + 7) / 8;
    cur.assign(rb, 0);
    prev.assign(rb, 0);
    for (uint32_t y = 0; y < ph; y++) {
      if (pos + 1 + rb > raw.size()) return fail("PNG data too short");
      const int ft = raw[pos++];
      const uint8_t* src = &raw[pos];
      pos += rb;
      for (size_t k = 0; k < rb; k++) {
        const int a = k >= (size_t)fb ? cur[k - fb] : 0, b = prev[k], c = k >= (size_t)fb ? prev[k - fb] : 0;
        int pr = 0;
        switch (ft) {
          case 0: pr = 0; break;
          case 1: pr = a; break;
          case 2: pr = b; break;
          case 3: pr = (a + b) >> 1; break;
          case 4: {
            const int pp = a + b - c, pa = pp > a ? pp - a : a - pp, pb = pp > b ? pp - b : b - pp, pc = pp > c ? pp - c : c - pp;
            pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            break;
          }
          default: return fail("bad PNG filter type");
        }
        cur[k] = (uint8_t)(src[k] + pr);
      }
      // samples of this row -> one grey value per pixel
      const uint32_t oy = interlace ? ys[pass] + y * dy[pass] : y;
      for (uint32_t x = 0; x < pw; x++) {
        const uint32_t ox = interlace ? xs[pass] + x * dx[pass] : x;
        uint8_t g;
        if (depth == 16) {
          const uint8_t* q = &cur[(size_t)x * nch * 2];
          auto s16 = [&](int c) { return (q[2 * c] << 8) | q[2 * c + 1]; };
          const int v = nch >= 3 ? luma16(s16(0), s16(1), s16(2)) : s16(0);
          g = (uint8_t)(v >> 8);
        } else if (depth == 8) {
          const uint8_t* q = &cur[(size_t)x * nch];
          if (ctype == 3) {
            const size_t e = (size_t)q[0] * 3;
            g = e + 2 < pal.size() ? luma8(pal[e], pal[e + 1], pal[e + 2]) : 0;
          } else g = nch >= 3 ? luma8(q[0], q[1], q[2]) : q[0];
        } else {
          const int per = 8 / depth, sh = (per - 1 - (int)(x % per)) * depth;
          const int v = (cur[x / per] >> sh) & ((1 << depth) - 1);
          if (ctype == 3) {
            const size_t e = (size_t)v * 3;
            g = e + 2 < pal.size() ? luma8(pal[e], pal[e + 1], pal[e + 2]) : 0;
          } else g = (uint8_t)(v * (depth == 1 ? 255 : depth == 2 ? 85 : 17));
        }
        out.pix[(size_t)oy * W + ox] = g;
      }
      prev.swap(cur);
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// BMP, PNM
// ---------------------------------------------------------------------------------------------------------
inline uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

inline bool decode_bmp(const uint8_t* d, size_t n, GrayImage& out) {
  auto fail = [&](const char* m) { out.err = m; return false; };
  if (n < 30 || d[0] != 'B' || d[1] != 'M') return fail("not a BMP file");
  const uint32_t off = le32(d + 10), hsz = le32(d + 14);
  if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return fail("unknown BMP header");
  if (14 + (size_t)hsz > n) return fail("truncated BMP header");
  int W, H, bpp, compress = 0;
  uint32_t mr = 0, mg = 0, mb = 0;
  if (hsz == 12) {
    W = le16(d + 18);
    H = le16(d + 20);
    bpp = le16(d + 24);
  } else {
    W = (int32_t)le32(d + 18);
    H = (int32_t)le32(d + 22);
    bpp = le16(d + 28);
    compress = (int)le32(d + 30);
    if (compress == 1 || compress == 2) return fail("run-length-encoded BMP is not supported");
    if (compress == 3 && 14 + 40 + 12 <= n) {
      mr = le32(d + 54);
      mg = le32(d + 58);
      mb = le32(d + 62);
    }
  }
  const bool flip = H > 0;
  if (H < 0) H = -H;
  if (W <= 0 || H <= 0) return fail("bad BMP size");
  if (bpp == 16) return fail("16-bit BMP is not supported");
  if (bpp != 1 && bpp != 4 && bpp != 8 && bpp != 24 && bpp != 32) return fail("bad BMP bit depth");
  if (bpp == 32 && compress == 3 && !(mr == 0xff0000u && mg == 0xff00u && mb == 0xffu)) return fail("BMP bit masks other than 8-8-8 are not supported");
  const int pal_entry = hsz == 12 ? 3 : 4;
  int npal = 0;
  const uint8_t* pal = d + 14 + hsz + (hsz == 40 && compress == 3 ? 12 : 0);
  if (bpp <= 8) {
    npal = (int)((off - (size_t)(pal - d)) / pal_entry);
    if (npal <= 0 || npal > 256 || (size_t)(pal - d) + (size_t)npal * pal_entry > n) return fail("bad BMP palette");
  }
  const size_t rb = (((size_t)W * bpp + 31) / 32) * 4;
  if ((size_t)off + rb * H > n) return fail("truncated BMP data");
  out.w = W;
  out.h = H;
  out.channels = bpp == 32 ? 4 : 3;
  out.pix.resize((size_t)W * H);
  for (int y = 0; y < H; y++) {
    const uint8_t* row = d + off + rb * (size_t)(flip ? H - 1 - y : y);
    for (int x = 0; x < W; x++) {
      int r, g, b;
      if (bpp <= 8) {
        int v;
        if (bpp == 8) v = row[x];
        else if (bpp == 4) v = (row[x >> 1] >> ((x & 1) ? 0 : 4)) & 15;
        else v = (row[x >> 3] >> (7 - (x & 7))) & 1;
        if (v >= npal) v = 0;
        b = pal[v * pal_entry];
        g = pal[v * pal_entry + 1];
        r = pal[v * pal_entry + 2];
      } else {
        const uint8_t* q = row + (size_t)x * (bpp / 8);
        b = q[0];
        g = q[1];
        r = q[2];
      }
      out.pix[(size_t)y * W + x] = luma8(r, g, b);
    }
  }
  return true;
}

inline bool decode_pnm(const uint8_t* d, size_t n, GrayImage& out) {
  auto fail = [&](const char* m) { out.err = m; return false; };
  if (n < 3 || d[0] != 'P' || (d[1] != '5' && d[1] != '6')) return fail("not a binary PGM / PPM file");
  const int nch = d[1] == '5' ? 1 : 3;
  size_t i = 2;
  auto next_int = [&](int& v) -> bool {
    while (i < n) {
      if (d[i] == '#') while (i < n && d[i] != '\n' && d[i] != '\r') i++;
      else if (d[i] == ' ' || d[i] == '\t' || d[i] == '\n' || d[i] == '\r' || d[i] == '\f' || d[i] == '\v') i++;
      else break;
    }
    if (i >= n || d[i] < '0' || d[i] > '9') return false;
    long t = 0;
    while (i < n && d[i] >= '0' && d[i] <= '9' && t < (1 << 28)) t = t * 10 + (d[i++] - '0');
    v = (int)t;
    return true;
  };
  int W, H, maxv;
  if (!next_int(W) || !next_int(H) || !next_int(maxv) || W <= 0 || H <= 0) return fail("bad PNM header");
  if (maxv > 255) return fail("16-bit PNM is not supported");
  i++;  // the single whitespace byte behind maxval
  if (i + (size_t)W * H * nch > n) return fail("truncated PNM data");
  out.w = W;
  out.h = H;
  out.channels = nch;
  out.pix.resize((size_t)W * H);
  for (size_t k = 0; k < (size_t)W * H; k++) {
    const uint8_t* q = d + i + k * nch;
    out.pix[k] = nch == 1 ? q[0] : luma8(q[0], q[1], q[2]);
  }
  return true;
}

}  // namespace img_detail

inline bool decode_image_gray(const uint8_t* d, size_t n, GrayImage& out) {
  using namespace img_detail;
  if (n >= 2 && d[0] == 0xFF && d[1] == 0xD8) return decode_jpeg(d, n, out);
  if (n >= 8 && d[0] == 137 && d[1] == 'P' && d[2] == 'N' && d[3] == 'G') return decode_png(d, n, out);
  if (n >= 2 && d[0] == 'B' && d[1] == 'M') return decode_bmp(d, n, out);
  if (n >= 2 && d[0] == 'P' && (d[1] == '5' || d[1] == '6')) return decode_pnm(d, n, out);
  out.err = "unrecognised image format (JPEG, PNG, BMP and binary PGM/PPM are read)";
  return false;
}

inline bool load_image_gray(const std::string& path, GrayImage& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { out.err = "cannot open " + path; return false; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t r;
  while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
  fclose(f);
  return decode_image_gray(buf.data(), buf.size(), out);
}

// the name the host programs have used so far
inline bool load_jpeg_gray(const std::string& path, GrayImage& out) { return load_image_gray(path, out); }

}  // namespace kbhost
