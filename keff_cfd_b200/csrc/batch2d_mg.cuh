// batch2d_mg.cuh -- 2D batch fast path with an ON-CHIP multigrid preconditioner: one image per SM.
//
// Same outer scheme as batch2d_onchip.cuh (FP32 conjugate gradients on chip, FP64 entering through
// reliable updates: x += y flushed in FP64, true residual b - A x recomputed in FP64 with the FP64
// coefficients, every stop decision on FP64 quantities), but the preconditioner is one geometric
// multigrid V-cycle instead of D^-1 + a 64-unknown coarse correction: 12 iterations per 128x128
// image instead of 153.  The linear system is still exactly the reference's
// (DiscretizeMatrixCD2D, Keff2D/keff2D.h:482-582); the cycle only changes how fast CG gets there.
// Replaces the reference batch loop Keff2D/keff2D.cpp:341-471 (SerialGS per image) for images
// whose sides are multiples of 16 and <= 128; other shapes keep k_batch2d_onchip / k_batch2d_pcg.
// The same arithmetic in numpy: tests/helpers/onchip_mg_proto.py (checked on the CPU by tests/test_onchip_mg_design.py).
//
// Hierarchy (all FP32, all in shared memory): level 0 = the image (nx x ny), level l = 2x2
// aggregates of level l-1, down to level 4 (<= 8x8 = 64 cells) which is solved with a dense inverse
// (Gauss-Jordan in shared memory once per image).  Transfers are piecewise constant; a coarse face
// conductance is HALF the sum of the two fine face conductances crossing it (the re-discretised
// operator on a uniform medium -- same rule as mg.cuh), wall conductances likewise.
// Cycle on level 0 (one weighted-Jacobi sweep per side, weight w, u = D^-1 b):
//   x1 = w u;   r1 = b - A x1 = (1-w) D u + w sum_f c_f u_nb;   e = coarse(P^T r1);   x2 = x1 + P e;
//   out = x2 + w D^-1 (b - A x2) = (1-w) x2 + w u + w D^-1 sum_f c_f x2_nb
// x1 and x2 are never stored: a pass reads u (and the coarser correction) of the neighbours and forms them on the
// fly.  Levels 1..3 make TWO sweeps per side (weights w1, w2 down, w2, w1 up) around one stored iterate X (passes
// A-D above mb_sweep).  Same weights mirrored on the way up -> the cycle is symmetric positive definite, as CG needs.
//
// Level 0: a thread owns a 4(x) x 8(y) patch (512 threads, lane -> x, warp -> y).  u = D^-1 a lives in a guarded
// shared-memory image (written by the CG stencil / the reliable update, read row by row with LDS.128 by the
// restriction and post-smoothing passes; W/E neighbours come from warp shuffles, the rows above/below the patch are
// the neighbour warps' rows of the image).  Keeping u in registers instead (first light) cost 584 B of spills per
// thread.  a_P is the sum of the face conductances that every pass selects from the code bytes anyway, so no 1/a_P
// array exists.  w = M A p stays in registers from the post-smoothing pass to the vector update.  p (a guarded image:
// the CG stencil reads the neighbours' rows with ld.global.cg), y (solution increment) and z = M r are touched once
// or twice per iteration and stream through a per-CTA slice of global memory that stays L2-resident (200 KB per SM).
// Per iteration: 3 passes over the image (CG stencil, restriction, post-smoothing), 12 small passes over the
// coarse levels, one dense 64x64 mat-vec, 18 barriers, one block reduction.
#pragma once
#include "batch2d_onchip.cuh"

namespace kb {

constexpr int MB_THREADS = 512, MB_PITCH = 128, MB_ROWS = 130;
constexpr int MB_NA = 64;  // max cells of the dense level
__host__ __device__ constexpr int mb_pitch(int L) { return (128 >> L) + 4; }
__host__ __device__ constexpr int mb_rows(int L) { return (128 >> L) + 2; }
__host__ __device__ constexpr int mb_n(int L) { return mb_pitch(L) * mb_rows(L); }
__host__ __device__ constexpr int mb_off(int L) { return L <= 1 ? 0 : mb_off(L - 1) + 5 * mb_n(L - 1); }
// floats: u image | X1 | X2 | X3 | levels 1..3 (U, E, CX, CY, DI each) | dense-level correction (guarded) |
//         dense right-hand side | dense inverse | scratch
constexpr int MB_F_P = 0, MB_F_XCH = MB_F_P + MB_ROWS * MB_PITCH, MB_F_X2 = MB_F_XCH + mb_n(1), MB_F_X3 = MB_F_X2 + mb_n(2);
constexpr int MB_F_LV = MB_F_X3 + mb_n(3);
constexpr int MB_F_E4 = MB_F_LV + mb_off(4), MB_F_B4 = MB_F_E4 + mb_n(4), MB_F_AI = MB_F_B4 + MB_NA;
constexpr int MB_F_SCR = MB_F_AI + MB_NA * MB_NA, MB_F_END = MB_F_SCR + 256;
constexpr size_t MB_SMEM = (size_t)MB_F_END * 4 + 2 * 64 * 8 + 4 * 8 + 16 * 8 + 64;
// per-CTA global scratch (doubles): y (8192: 8 x 512 float4) | z (8192) | p image (8320: 130 x 128 floats) | x (n)
constexpr int MB_PG_DOUBLES = MB_ROWS * MB_PITCH / 2;
__host__ __device__ constexpr size_t mb_ws_stride(size_t n) { return n + 2 * 8192 + MB_PG_DOUBLES; }

struct Row5 {
  float cw, c01, c12, c23, ce;  // x faces of a 4-cell row: W of cell 0, 0|1, 1|2, 2|3, E of cell 3
  float s0, s1, s2, s3;         // S faces
};

// Face conductances of one row of my patch from its code word (one byte per cell).  left/right: the
// patch touches the x wall, whose face is a conductance c_wall towards the value 0 (the code bit of a
// boundary face is clear by construction).  nosouth: the row is the image's last row (adiabatic).
__device__ __forceinline__ Row5 mb_row(unsigned c, const CoefF& F, bool left, bool right, bool nosouth) {
  const bool p0 = c & (C_PH << 0), p1 = c & (C_PH << 8), p2 = c & (C_PH << 16), p3 = c & (C_PH << 24);
  const float a0 = p0 ? F.sx[1] : F.sx[0], a1 = p1 ? F.sx[1] : F.sx[0], a2 = p2 ? F.sx[1] : F.sx[0];
  const float a3 = p3 ? F.sx[1] : F.sx[0];
  Row5 r;
  r.cw = (c & (C_W << 0)) ? F.dx : (left ? (p0 ? F.wall[1] : F.wall[0]) : a0);
  r.c01 = (c & (C_E << 0)) ? F.dx : a0;
  r.c12 = (c & (C_E << 8)) ? F.dx : a1;
  r.c23 = (c & (C_E << 16)) ? F.dx : a2;
  r.ce = (c & (C_E << 24)) ? F.dx : (right ? (p3 ? F.wall[1] : F.wall[0]) : a3);
  r.s0 = nosouth ? 0.f : ((c & (C_S << 0)) ? F.dy : (p0 ? F.sy[1] : F.sy[0]));
  r.s1 = nosouth ? 0.f : ((c & (C_S << 8)) ? F.dy : (p1 ? F.sy[1] : F.sy[0]));
  r.s2 = nosouth ? 0.f : ((c & (C_S << 16)) ? F.dy : (p2 ? F.sy[1] : F.sy[0]));
  r.s3 = nosouth ? 0.f : ((c & (C_S << 24)) ? F.dy : (p3 ? F.sy[1] : F.sy[0]));
  return r;
}
// The face conductances are functions of the code words alone; left to itself the compiler computes all of
// them once per image and keeps ~100 values per thread alive across every pass (= spilled).  Passing the
// code words through an empty asm at the start of a pass makes it recompute them where they are used.
__device__ __forceinline__ void mb_fresh(unsigned (&cw)[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) asm volatile("" : "+r"(cw[j]));
}
// N faces of the patch's first row (the S faces of the row above, seen from this side)
__device__ __forceinline__ void mb_north(unsigned c, const CoefF& F, bool top, float (&n)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
    n[i] = top ? 0.f : ((c & (C_N << (8 * i))) ? F.dy : ((c & (C_PH << (8 * i))) ? F.sy[1] : F.sy[0]));
}

// the same in FP64 with the FP64 coefficients (true residual of the reliable updates)
struct Row5d {
  double cw, c01, c12, c23, ce, s0, s1, s2, s3;
};
__device__ __forceinline__ Row5d mb_row_d(unsigned c, const Coef& C, bool left, bool right, bool nosouth) {
  const bool p0 = c & (C_PH << 0), p1 = c & (C_PH << 8), p2 = c & (C_PH << 16), p3 = c & (C_PH << 24);
  const double sx0 = C.same[0][0], sx1 = C.same[0][1], sy0 = C.same[1][0], sy1 = C.same[1][1];
  const double a0 = p0 ? sx1 : sx0, a1 = p1 ? sx1 : sx0, a2 = p2 ? sx1 : sx0, a3 = p3 ? sx1 : sx0;
  Row5d r;
  r.cw = (c & (C_W << 0)) ? C.diff[0] : (left ? (p0 ? C.wall[1] : C.wall[0]) : a0);
  r.c01 = (c & (C_E << 0)) ? C.diff[0] : a0;
  r.c12 = (c & (C_E << 8)) ? C.diff[0] : a1;
  r.c23 = (c & (C_E << 16)) ? C.diff[0] : a2;
  r.ce = (c & (C_E << 24)) ? C.diff[0] : (right ? (p3 ? C.wall[1] : C.wall[0]) : a3);
  r.s0 = nosouth ? 0.0 : ((c & (C_S << 0)) ? C.diff[1] : (p0 ? sy1 : sy0));
  r.s1 = nosouth ? 0.0 : ((c & (C_S << 8)) ? C.diff[1] : (p1 ? sy1 : sy0));
  r.s2 = nosouth ? 0.0 : ((c & (C_S << 16)) ? C.diff[1] : (p2 ? sy1 : sy0));
  r.s3 = nosouth ? 0.0 : ((c & (C_S << 24)) ? C.diff[1] : (p3 ? sy1 : sy0));
  return r;
}
__device__ __forceinline__ void mb_north_d(unsigned c, const Coef& C, bool top, double (&n)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
    n[i] = top ? 0.0 : ((c & (C_N << (8 * i))) ? C.diff[1] : ((c & (C_PH << (8 * i))) ? C.same[1][1] : C.same[1][0]));
}

// idx / d for 0 <= idx <= 4096, 2 <= d <= 64 without a division: the high word of idx * (floor(2^32 / d) + 1)
// (exact on that range; checked exhaustively).  mb_rcp(d) is computed once per launch; 0 stands for d = 1.
__device__ __forceinline__ unsigned mb_rcp(int d) {
  return d <= 1 ? 0u : 0xffffffffu / (unsigned)d + ((0xffffffffu % (unsigned)d) + 1u == (unsigned)d ? 2u : 1u);
}
__device__ __forceinline__ int mb_div(int idx, unsigned rcp) { return rcp ? (int)__umulhi((unsigned)idx, rcp) : idx; }

// ---- coarse levels (shared-memory arrays with a zero guard ring: cell (X,Y) at (Y+1)*pitch + 2 + X) ----
template <int L>
struct MbLv {
  static constexpr int pitch = mb_pitch(L), n = mb_n(L);
  float *U, *E, *CX, *CY, *DI;
  __device__ __forceinline__ MbLv(float* lv) {
    float* b = lv + mb_off(L);
    U = b;
    E = b + n;
    CX = b + 2 * n;
    CY = b + 3 * n;
    DI = b + 4 * n;
  }
};

// 1/a_P of level L (E holds the wall conductances during setup)
template <int L>
__device__ __forceinline__ void mb_dinv(float* lv, int nxl, int nyl, int tid) {
  MbLv<L> V(lv);
  constexpr int pt = MbLv<L>::pitch;
  for (int idx = tid; idx < nxl * nyl; idx += MB_THREADS) {
    const int Y = idx / nxl, X = idx - Y * nxl, o = (Y + 1) * pt + 2 + X;
    V.DI[o] = 1.f / (((V.CX[o] + V.CX[o - 1]) + (V.CY[o] + V.CY[o - pt])) + V.E[o]);
  }
}

// coefficients of level L+1 from level L (L = 3: into the scratch arrays of the dense level)
template <int L>
__device__ __forceinline__ void mb_coarsen(float* lv, float* scr, int nxl, int nyl, int tid) {
  MbLv<L> V(lv);
  constexpr int pt = MbLv<L>::pitch;
  const int nxc = nxl >> 1, nyc = nyl >> 1;
  for (int idx = tid; idx < nxc * nyc; idx += MB_THREADS) {
    const int Yc = idx / nxc, Xc = idx - Yc * nxc, oa = (2 * Yc + 1) * pt + 2 + 2 * Xc;
    const float cx = 0.5f * (V.CX[oa + 1] + V.CX[oa + pt + 1]);
    const float cy = 0.5f * (V.CY[oa + pt] + V.CY[oa + pt + 1]);
    const float wl = 0.5f * ((V.E[oa] + V.E[oa + 1]) + (V.E[oa + pt] + V.E[oa + pt + 1]));
    if constexpr (L < 3) {
      MbLv<L + 1> W(lv);
      const int oc = (Yc + 1) * MbLv<L + 1>::pitch + 2 + Xc;
      W.CX[oc] = cx;
      W.CY[oc] = cy;
      W.E[oc] = wl;
    } else {
      scr[idx] = cx;
      scr[64 + idx] = cy;
      scr[128 + idx] = wl;
    }
  }
}

// Coarse levels run nu = 2 sweeps on either side of their coarse correction (weights w1, w2 before, w2, w1 after),
// which needs one stored iterate X per level:
//   A: X = x2 = (w1 + w2 - w1 w2) u + w1 w2 D^-1 sum c u_nb            (x1 = w1 u is implicit)
//   B: restrict  D (u - x2) + sum c x2_nb  -> u of the next level
//   C: E = x4 = (1 - w2) x3 + w2 u + w2 D^-1 sum c x3_nb,  x3 = x2 + P e_next
//   D: X = x5 = (1 - w1) x4 + w1 u + w1 D^-1 sum c x4_nb                (the level's correction e = x5)
// A, C, D are one routine: out_P = a x_P + b u_P + g D^-1_P sum_f c_f x_nb with x = in (+ Ec of the aggregate).
// Guard cells of `in` / Ec only ever meet a zero conductance, so they need to be finite, not zero.
template <int L, bool ADD>
__device__ __forceinline__ void mb_sweep(float* lv, const float* in, const float* Ec, float* out, int nxl, int nyl, float a,
                                         float b, float g, int tid) {
  MbLv<L> V(lv);
  constexpr int pt = MbLv<L>::pitch;
  constexpr int pc = mb_pitch(L + 1);
  for (int idx = tid; idx < nxl * nyl; idx += MB_THREADS) {
    const int Y = idx / nxl, X = idx - Y * nxl, o = (Y + 1) * pt + 2 + X;
    float xP = in[o], xW = in[o - 1], xE = in[o + 1], xN = in[o - pt], xS = in[o + pt];
    if (ADD) {  // aggregate of a cell (arithmetic shifts: -1 -> the guard ring)
      const int rP = ((Y >> 1) + 1) * pc + 2, rN = (((Y - 1) >> 1) + 1) * pc + 2, rS = (((Y + 1) >> 1) + 1) * pc + 2;
      const int cP = X >> 1, cW = (X - 1) >> 1, cE = (X + 1) >> 1;
      xP += Ec[rP + cP];
      xW += Ec[rP + cW];
      xE += Ec[rP + cE];
      xN += Ec[rN + cP];
      xS += Ec[rS + cP];
    }
    const float sm = (V.CX[o] * xE + V.CX[o - 1] * xW) + (V.CY[o] * xS + V.CY[o - pt] * xN);
    out[o] = a * xP + b * V.U[o] + g * V.DI[o] * sm;
  }
}

template <int L>
__device__ __forceinline__ void mb_restrict(float* lv, const float* in, float* b4, int nxl, int nyl, int tid) {
  MbLv<L> V(lv);
  constexpr int pt = MbLv<L>::pitch;
  const int nxc = nxl >> 1, nyc = nyl >> 1;
  for (int idx = tid; idx < nxc * nyc; idx += MB_THREADS) {
    const int Yc = idx / nxc, Xc = idx - Yc * nxc;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        const int o = (2 * Yc + dy + 1) * pt + 2 + 2 * Xc + dx;
        const float sm = (V.CX[o] * in[o + 1] + V.CX[o - 1] * in[o - 1]) + (V.CY[o] * in[o + pt] + V.CY[o - pt] * in[o - pt]);
        acc += __fdividef(V.U[o] - in[o], V.DI[o]) + sm;
      }
    if constexpr (L < 3) {
      MbLv<L + 1> W(lv);
      const int oc = (Yc + 1) * MbLv<L + 1>::pitch + 2 + Xc;
      W.U[oc] = acc * W.DI[oc];
    } else {
      b4[idx] = acc;
    }
  }
}

// ---- level 1 with the thread layout of level 0: a thread owns the 2(x) x 4(y) level-1 cells under its patch ----
// (the generic passes spend most of their instructions on index arithmetic and scalar loads; level 1 holds 3/4 of
// all coarse cells).  o = array offset of my first cell; rows -1 and 4 / columns -1 and 2 are my neighbours' cells
// or the guard ring.
struct MbBlk1 {
  float2 I[6];         // input field, rows -1 .. 4
  float iw[4], ie[4];  // its columns -1 and 2, rows 0 .. 3
  float2 u[4];         // u of my cells
  float2 cx[4], di[4], cy[5];  // cy rows -1 .. 3
  float cxw[4];
};
template <bool IN_IS_U>
__device__ __forceinline__ void mb_load1(float* lv, const float* in, int o, MbBlk1& B) {
  MbLv<1> V(lv);
  constexpr int pt = MbLv<1>::pitch;
#pragma unroll
  for (int r = 0; r < 6; r++) B.I[r] = *reinterpret_cast<const float2*>(in + o + (r - 1) * pt);
#pragma unroll
  for (int r = 0; r < 5; r++) B.cy[r] = *reinterpret_cast<const float2*>(V.CY + o + (r - 1) * pt);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    B.iw[r] = in[o + r * pt - 1];
    B.ie[r] = in[o + r * pt + 2];
    B.u[r] = IN_IS_U ? B.I[r + 1] : *reinterpret_cast<const float2*>(V.U + o + r * pt);
    B.cx[r] = *reinterpret_cast<const float2*>(V.CX + o + r * pt);
    B.cxw[r] = V.CX[o + r * pt - 1];
    B.di[r] = *reinterpret_cast<const float2*>(V.DI + o + r * pt);
  }
}
// passes A, C, D on my block; Ec = correction of level 2 (ADD)
template <bool ADD, bool IN_IS_U>
__device__ __forceinline__ void mb_sweep1(float* lv, const float* in, const float* Ec, float* out, int o, int lane, int wrp,
                                          float a, float b, float g) {
  MbBlk1 B;
  mb_load1<IN_IS_U>(lv, in, o, B);
  constexpr int pt = MbLv<1>::pitch, p2 = MbLv<2>::pitch;
  float ew0 = 0.f, ew1 = 0.f, ee0 = 0.f, ee1 = 0.f;
  if (ADD) {
    const int c2 = (2 * wrp + 1) * p2 + 2 + lane;  // level-2 cell under my rows 0, 1
    const float e0 = Ec[c2], e1 = Ec[c2 + p2], en = Ec[c2 - p2], es = Ec[c2 + 2 * p2];
    ew0 = Ec[c2 - 1]; ew1 = Ec[c2 + p2 - 1]; ee0 = Ec[c2 + 1]; ee1 = Ec[c2 + p2 + 1];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const float e = r == 0 ? en : (r <= 2 ? e0 : (r <= 4 ? e1 : es));
      B.I[r].x += e;
      B.I[r].y += e;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float xw = B.iw[r] + (r < 2 ? ew0 : ew1), xe = B.ie[r] + (r < 2 ? ee0 : ee1);
    const float2 xc = B.I[r + 1], xn = B.I[r], xs = B.I[r + 2];
    const float sa = (B.cx[r].x * xc.y + B.cxw[r] * xw) + (B.cy[r + 1].x * xs.x + B.cy[r].x * xn.x);
    const float sb = (B.cx[r].y * xe + B.cx[r].x * xc.x) + (B.cy[r + 1].y * xs.y + B.cy[r].y * xn.y);
    *reinterpret_cast<float2*>(out + o + r * pt) =
        make_float2(a * xc.x + b * B.u[r].x + g * B.di[r].x * sa, a * xc.y + b * B.u[r].y + g * B.di[r].y * sb);
  }
}
// pass B on my block: the two level-2 cells (column lane, rows 2 wrp + {0, 1}) under it
__device__ __forceinline__ void mb_restrict1(float* lv, const float* in, int o, int lane, int wrp) {
  MbBlk1 B;
  mb_load1<false>(lv, in, o, B);
  MbLv<2> W(lv);
  float acc[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float2 xc = B.I[r + 1], xn = B.I[r], xs = B.I[r + 2];
    const float sa = (B.cx[r].x * xc.y + B.cxw[r] * B.iw[r]) + (B.cy[r + 1].x * xs.x + B.cy[r].x * xn.x);
    const float sb = (B.cx[r].y * B.ie[r] + B.cx[r].x * xc.x) + (B.cy[r + 1].y * xs.y + B.cy[r].y * xn.y);
    acc[r >> 1] += (__fdividef(B.u[r].x - xc.x, B.di[r].x) + __fdividef(B.u[r].y - xc.y, B.di[r].y)) + (sa + sb);
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int oc = (2 * wrp + k + 1) * MbLv<2>::pitch + 2 + lane;
    W.U[oc] = acc[k] * W.DI[oc];
  }
}

__global__ void __launch_bounds__(MB_THREADS, 1)
    k_batch2d_mg(const uint8_t* __restrict__ masks, int n_img, int nx, int ny, const Coef C, const CoefF F, double thr2_in,
                 double flux_tol, long long max_iter, int const_init, float om, float wc1, float wc2, double delta2,
                 double* __restrict__ xws,
                 double* __restrict__ T_out, BatchOut* __restrict__ res, int* counter, const unsigned int* ready) {
  extern __shared__ __align__(16) unsigned char mb_smem[];
  float* SM = reinterpret_cast<float*>(mb_smem);
  float* P = SM + MB_F_P;
  float* X1 = SM + MB_F_XCH;  // iterates / final corrections of levels 1..3
  float* X2 = SM + MB_F_X2;
  float* X3 = SM + MB_F_X3;
  float* LV = SM + MB_F_LV;
  float* E4 = SM + MB_F_E4;
  float* B4 = SM + MB_F_B4;
  float* AI = SM + MB_F_AI;
  float* scr = SM + MB_F_SCR;
  double* red = reinterpret_cast<double*>(SM + MB_F_END);
  double* bc = red + 2 * 64;
  double* SC = bc + 4;  // block-uniform FP64 scalars that are only looked at around reliable updates (every thread
                        // writes the same value, and only plain assignments -- never read-modify-write): parked here so that they do not occupy registers in the passes
  int* s_img = reinterpret_cast<int*>(SC + 16);
  unsigned* s_rcp = reinterpret_cast<unsigned*>(s_img + 4);  // reciprocal for r / nax (mb_div) in the dense-level product

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int x0 = lane * 4, y0 = wrp * 8;
  const bool act = x0 < nx && y0 < ny;  // shapes are multiples of the patch: all 32 cells in or out
  const int n = nx * ny;
  const int nx1 = nx >> 1, ny1 = ny >> 1, nx2 = nx >> 2, ny2 = ny >> 2, nx3 = nx >> 3, ny3 = ny >> 3;
  const int nax = nx >> 4, na = nax * (ny >> 4);
  const bool left = x0 == 0, right = x0 + 4 == nx, top = y0 == 0, bot = y0 + 8 == ny;
  const float om1 = 1.f - om;

  for (int i = tid; i < MB_F_END; i += MB_THREADS) SM[i] = 0.f;  // guard rings stay zero for the whole launch
  if (tid == 0) s_rcp[0] = mb_rcp(nax);
  float* Uown = P + (y0 + 1) * MB_PITCH + x0;  // my patch's first row inside the guarded image of u
  // y and z slices: thread-major float4 rows at compile-time offsets from one base
  float4* Yg = reinterpret_cast<float4*>(xws + (size_t)blockIdx.x * mb_ws_stride(n)) + tid;
  constexpr int ZG = 8 * MB_THREADS;  // z relative to y (in float4)
  // p: guarded image [130][128] in the slice (the CG stencil reads the neighbours' rows from it); guard rows stay 0
  float* Pg = reinterpret_cast<float*>(Yg - tid) + 16 * MB_THREADS * 4;
  for (int i = tid; i < MB_ROWS * MB_PITCH; i += MB_THREADS) Pg[i] = 0.f;
  float* Pown = Pg + (y0 + 1) * MB_PITCH + x0;
  // level-1 cells of my patch: aggregate rows (y0>>1) + 0..3, aggregate columns (x0>>1) + 0..1
  MbLv<1> L1(LV);
  MbLv<2> L2(LV);
  MbLv<3> L3(LV);
  const float cA = wc1 + wc2 - wc1 * wc2, cG = wc1 * wc2;  // first stored iterate of a coarse level from u
  constexpr int p1 = MbLv<1>::pitch;
  const int o1 = ((y0 >> 1) + 1) * p1 + 2 + (x0 >> 1);

  while (true) {
    if (tid == 0) {
      int i = atomicAdd(counter, 1);
      // host batches stream in while the kernel runs: `ready` counts the images whose structure has landed in HBM
      // (written by the copy stream behind each chunk); images are taken in copy order
      if (ready && i < n_img && !wait_counter(ready, (unsigned)i + 1u)) i = n_img;
      *s_img = i;
    }
    __syncthreads();
    const int img = *s_img;
    __syncthreads();
    if (img >= n_img) break;
    const size_t moff = (size_t)img * n;
#define MB_M (masks + moff)
    double* x64 = T_out ? T_out + (size_t)img * n : reinterpret_cast<double*>(Yg - tid) + 2 * 8192 + MB_PG_DOUBLES;

    // ---- per-image setup: code words (registers), x = initial field, y = 0 ----
    unsigned cw[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      unsigned cwj = 0;
      if (act) {
        const int yy = y0 + j;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int xx = x0 + i, g = yy * nx + xx;
          const uint8_t* m = MB_M;
          const unsigned me = m[g] == 1;
          unsigned c = me;
          if (xx > 0 && (unsigned)(m[g - 1] == 1) != me) c |= C_W;
          if (xx < nx - 1 && (unsigned)(m[g + 1] == 1) != me) c |= C_E;
          if (yy < ny - 1 && (unsigned)(m[g + nx] == 1) != me) c |= C_S;
          if (yy > 0 && (unsigned)(m[g - nx] == 1) != me) c |= C_N;
          cwj |= c << (8 * i);
          x64[g] = const_init ? C.tl : C.tl + (C.tr - C.tl) * ((xx + 0.5) / nx);
        }
      }
      cw[j] = cwj;
      Yg[j * MB_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- hierarchy: level 1 from the code words, levels 2, 3 and the dense level from the level above ----
    if (act) {
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const Row5 ra = mb_row(cw[2 * jj], F, left, right, false);
        const Row5 rb = mb_row(cw[2 * jj + 1], F, left, right, jj == 3 && bot);
        const unsigned ca = cw[2 * jj], cb = cw[2 * jj + 1];
        const float wl0 = left ? 0.5f * (((ca & C_PH) ? F.wall[1] : F.wall[0]) + ((cb & C_PH) ? F.wall[1] : F.wall[0])) : 0.f;
        const float wl1 = right ? 0.5f * (((ca & (C_PH << 24)) ? F.wall[1] : F.wall[0]) + ((cb & (C_PH << 24)) ? F.wall[1] : F.wall[0])) : 0.f;
        const int o = o1 + jj * p1;
        *reinterpret_cast<float2*>(L1.CX + o) = make_float2(0.5f * (ra.c12 + rb.c12), right ? 0.f : 0.5f * (ra.ce + rb.ce));
        *reinterpret_cast<float2*>(L1.CY + o) = make_float2(0.5f * (rb.s0 + rb.s1), 0.5f * (rb.s2 + rb.s3));
        *reinterpret_cast<float2*>(L1.E + o) = make_float2(wl0, wl1);
      }
    }
    __syncthreads();
    mb_dinv<1>(LV, nx1, ny1, tid);
    mb_coarsen<1>(LV, scr, nx1, ny1, tid);
    __syncthreads();
    mb_dinv<2>(LV, nx2, ny2, tid);
    mb_coarsen<2>(LV, scr, nx2, ny2, tid);
    __syncthreads();
    mb_dinv<3>(LV, nx3, ny3, tid);
    mb_coarsen<3>(LV, scr, nx3, ny3, tid);
    for (int i = tid; i < MB_NA * MB_NA; i += MB_THREADS) AI[i] = 0.f;
    __syncthreads();
    if (tid < na) {  // dense level: 5-point matrix on the nax x nay grid
      const int by = tid / nax, ax = tid - by * nax;
      const float ce = scr[tid], cs = scr[64 + tid];
      float dg = ce + cs + scr[128 + tid];
      if (ax > 0) dg += scr[tid - 1];
      if (by > 0) dg += scr[64 + tid - nax];
      AI[tid * MB_NA + tid] = dg;
      if (ax + 1 < nax) AI[tid * MB_NA + tid + 1] = AI[(tid + 1) * MB_NA + tid] = -ce;
      if (tid + nax < na) AI[tid * MB_NA + tid + nax] = AI[(tid + nax) * MB_NA + tid] = -cs;
    } else if (tid < MB_NA) {
      AI[tid * MB_NA + tid] = 1.f;  // identity padding: every 8x8 pivot block below is regular
    }
    __syncthreads();
    // In-place inverse by BLOCK Gauss-Jordan without pivoting (SPD M-matrix): eight steps with 8x8 pivot blocks
    // instead of 64 scalar pivots.  Step kb, with P = (pivot block)^-1, C = the pivot columns, R = the pivot rows:
    //   rows of the pivot block:   A_kj <- P R_j  (j != kb),   A_kk <- P
    //   every other row i:         A_ij <- A_ij - C_i (P R_j)  (j != kb),   A_ik <- -C_i P
    // i.e. with T' = P R whose pivot columns are replaced by P:  A_i: <- (A_i: with its pivot columns zeroed) - C_i T'.
    // One warp inverts the pivot block (8 Gauss-Jordan steps, two entries per lane), all threads form T' (one entry
    // each) and snapshot C, then every thread updates its 8 entries with 64 FMAs: three barriers per block step, 24
    // per image instead of 64, and a quarter of the instructions of the scalar form.  Rows whose pivot columns are
    // all zero (the band has not reached them) are skipped.  Scratch: the level-1 iterate array (not live yet).
    {
      float* Tq = X1;          // T': 8 x 64
      float* Cq = X1 + 512;    // pivot columns: 64 x 8
      float* Pq = X1 + 1024;   // pivot block / its inverse: 8 x 8
      const int gi = tid >> 3, cg = tid & 7, gj = cg * 8;
      float4* arow = reinterpret_cast<float4*>(AI + gi * MB_NA + gj);
      const int nb = (na + 7) >> 3;
      for (int kb = 0; kb < nb; kb++) {
        if (wrp == 0) {
          const int r0 = lane >> 3, r1 = r0 + 4, c = lane & 7;
          float m0 = AI[(8 * kb + r0) * MB_NA + 8 * kb + c], m1 = AI[(8 * kb + r1) * MB_NA + 8 * kb + c];
          Pq[lane] = m0;
          Pq[lane + 32] = m1;
#pragma unroll
          for (int p = 0; p < 8; p++) {
            __syncwarp();
            const float piv = 1.f / Pq[p * 8 + p], rowc = Pq[p * 8 + c], col0 = Pq[r0 * 8 + p], col1 = Pq[r1 * 8 + p];
            __syncwarp();
            const float rj = c == p ? piv : rowc * piv;
            m0 = r0 == p ? rj : (c == p ? 0.f : m0) - col0 * rj;
            m1 = r1 == p ? rj : (c == p ? 0.f : m1) - col1 * rj;
            Pq[lane] = m0;
            Pq[lane + 32] = m1;
          }
        }
        __syncthreads();
        {
          const int r = tid >> 6, j = tid & 63;
          float t;
          if ((j >> 3) == kb) {
            t = Pq[r * 8 + (j & 7)];
          } else {
            t = 0.f;
#pragma unroll
            for (int m = 0; m < 8; m++) t += Pq[r * 8 + m] * AI[(8 * kb + m) * MB_NA + j];
          }
          Tq[r * 64 + j] = t;
          Cq[tid] = AI[gi * MB_NA + 8 * kb + cg];
        }
        __syncthreads();
        {
          const float4 c0 = *reinterpret_cast<const float4*>(Cq + gi * 8), c1 = *reinterpret_cast<const float4*>(Cq + gi * 8 + 4);
          const float cm[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          const bool prow = (gi >> 3) == kb;
          bool any = prow;
#pragma unroll
          for (int m = 0; m < 8; m++) any = any || cm[m] != 0.f;
          if (any) {
            float av[8];
            if (prow) {
              const float4 t0 = *reinterpret_cast<const float4*>(Tq + (gi & 7) * 64 + gj);
              const float4 t1 = *reinterpret_cast<const float4*>(Tq + (gi & 7) * 64 + gj + 4);
              av[0] = t0.x; av[1] = t0.y; av[2] = t0.z; av[3] = t0.w; av[4] = t1.x; av[5] = t1.y; av[6] = t1.z; av[7] = t1.w;
            } else {
              const float4 a0 = arow[0], a1 = arow[1];
              const bool pcol = cg == kb;
              av[0] = pcol ? 0.f : a0.x; av[1] = pcol ? 0.f : a0.y; av[2] = pcol ? 0.f : a0.z; av[3] = pcol ? 0.f : a0.w;
              av[4] = pcol ? 0.f : a1.x; av[5] = pcol ? 0.f : a1.y; av[6] = pcol ? 0.f : a1.z; av[7] = pcol ? 0.f : a1.w;
#pragma unroll
              for (int m = 0; m < 8; m++) {
                const float4 t0 = *reinterpret_cast<const float4*>(Tq + m * 64 + gj);
                const float4 t1 = *reinterpret_cast<const float4*>(Tq + m * 64 + gj + 4);
                av[0] -= cm[m] * t0.x; av[1] -= cm[m] * t0.y; av[2] -= cm[m] * t0.z; av[3] -= cm[m] * t0.w;
                av[4] -= cm[m] * t1.x; av[5] -= cm[m] * t1.y; av[6] -= cm[m] * t1.z; av[7] -= cm[m] * t1.w;
              }
            }
            arow[0] = make_float4(av[0], av[1], av[2], av[3]);
            arow[1] = make_float4(av[4], av[5], av[6], av[7]);
          }
        }
        __syncthreads();
      }
    }

    // ---- the V-cycle: u (= D^-1 a, in the shared-memory image U0) -> w (registers);
    //      swq = sum w.a, szq = sum z.a over my patch ----
    float w[8][4];
    auto vcycle = [&](bool with_z, float& swq, float& szq) {
      __syncthreads();  // every row of u is in the image
      auto urow = [&](int j, float (&v)[4]) {  // row j of my patch (-1 and 8: the neighbours' rows or the zero guard rows)
        const float4 t = *reinterpret_cast<const float4*>(Uown + j * MB_PITCH);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      };
      // ---- level 0, restriction: r1 = (1-w) d u + w sum_f c_f u_nb, summed over 2x2 aggregates ----
      {
        mb_fresh(cw);
        float nn[4], un[4], uc[4], us[4];
        mb_north(cw[0], F, top, nn);
        urow(-1, un);
        urow(0, uc);
        float a01 = 0.f, a23 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          urow(j + 1, us);
          const Row5 r = mb_row(cw[j], F, left, right, j == 7 && bot);
          float uw = __shfl_up_sync(0xffffffffu, uc[3], 1), ue = __shfl_down_sync(0xffffffffu, uc[0], 1);
          if (left) uw = 0.f;
          if (right || lane == 31) ue = 0.f;
          const float d0 = (r.cw + r.c01) + (nn[0] + r.s0), d1 = (r.c01 + r.c12) + (nn[1] + r.s1);
          const float d2 = (r.c12 + r.c23) + (nn[2] + r.s2), d3 = (r.c23 + r.ce) + (nn[3] + r.s3);
          const float s0 = (r.cw * uw + r.c01 * uc[1]) + (nn[0] * un[0] + r.s0 * us[0]);
          const float s1 = (r.c01 * uc[0] + r.c12 * uc[2]) + (nn[1] * un[1] + r.s1 * us[1]);
          const float s2 = (r.c12 * uc[1] + r.c23 * uc[3]) + (nn[2] * un[2] + r.s2 * us[2]);
          const float s3 = (r.c23 * uc[2] + r.ce * ue) + (nn[3] * un[3] + r.s3 * us[3]);
          a01 += (om1 * d0 * uc[0] + om * s0) + (om1 * d1 * uc[1] + om * s1);
          a23 += (om1 * d2 * uc[2] + om * s2) + (om1 * d3 * uc[3] + om * s3);
          if (j & 1) {
            if (act) {
              const int o = o1 + (j >> 1) * p1;
              const float2 di = *reinterpret_cast<const float2*>(L1.DI + o);
              *reinterpret_cast<float2*>(L1.U + o) = make_float2(a01 * di.x, a23 * di.y);
            }
            a01 = a23 = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            un[i] = uc[i];
            uc[i] = us[i];
          }
          nn[0] = r.s0; nn[1] = r.s1; nn[2] = r.s2; nn[3] = r.s3;
        }
      }
      __syncthreads();
      // ---- levels 1..3 down: two sweeps (the first implicit), restriction ----
      if (act) mb_sweep1<false, true>(LV, L1.U, nullptr, X1, o1, lane, wrp, cA, 0.f, cG);
      __syncthreads();
      if (act) mb_restrict1(LV, X1, o1, lane, wrp);
      __syncthreads();
      mb_sweep<2, false>(LV, L2.U, nullptr, X2, nx2, ny2, cA, 0.f, cG, tid);
      __syncthreads();
      mb_restrict<2>(LV, X2, B4, nx2, ny2, tid);
      __syncthreads();
      mb_sweep<3, false>(LV, L3.U, nullptr, X3, nx3, ny3, cA, 0.f, cG, tid);
      __syncthreads();
      mb_restrict<3>(LV, X3, B4, nx3, ny3, tid);
      __syncthreads();
      {  // dense level: e4 = inverse * b4; 8 threads per row, 8 columns each (columns >= na of the inverse and entries
         // >= na of b4 are zero: no bounds inside the product)
        const int r = tid >> 3, part = tid & 7;
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float4 av = *reinterpret_cast<const float4*>(AI + r * MB_NA + part * 8 + 4 * h);
          const float4 bv = *reinterpret_cast<const float4*>(B4 + part * 8 + 4 * h);
          sum += (av.x * bv.x + av.y * bv.y) + (av.z * bv.z + av.w * bv.w);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (part == 0 && r < na) {
          const int by = mb_div(r, s_rcp[0]), ax = r - by * nax;
          E4[(by + 1) * mb_pitch(4) + 2 + ax] = sum;
        }
      }
      __syncthreads();
      // ---- levels 3..1 up: correction added on the fly, two sweeps; a level's correction ends up in its X ----
      mb_sweep<3, true>(LV, X3, E4, L3.E, nx3, ny3, 1.f - wc2, wc2, wc2, tid);
      __syncthreads();
      mb_sweep<3, false>(LV, L3.E, nullptr, X3, nx3, ny3, 1.f - wc1, wc1, wc1, tid);
      __syncthreads();
      mb_sweep<2, true>(LV, X2, X3, L2.E, nx2, ny2, 1.f - wc2, wc2, wc2, tid);
      __syncthreads();
      mb_sweep<2, false>(LV, L2.E, nullptr, X2, nx2, ny2, 1.f - wc1, wc1, wc1, tid);
      __syncthreads();
      if (act) mb_sweep1<true, false>(LV, X1, X2, L1.E, o1, lane, wrp, 1.f - wc2, wc2, wc2);
      __syncthreads();
      if (act) mb_sweep1<false, false>(LV, L1.E, nullptr, X1, o1, lane, wrp, 1.f - wc1, wc1, wc1);
      __syncthreads();
      // ---- level 0, post-smoothing: w = (1-w) x2 + w u + w D^-1 sum_f c_f x2_nb,  x2 = w u + P e1 ----
      {
        mb_fresh(cw);
        float nn[4];
        mb_north(cw[0], F, top, nn);
        // corrections of my aggregate rows (own two columns, the column to the left and to the right), of the
        // aggregate row above the patch and below it
        float2 eN = make_float2(0.f, 0.f), eS = eN;
        if (act) {
          eN = *reinterpret_cast<const float2*>(X1 + o1 - p1);
          eS = *reinterpret_cast<const float2*>(X1 + o1 + 4 * p1);
        }
        float xp[4], xc[4], xn[4];  // x2 of the row above, this row, the row below
        float2 ec = act ? *reinterpret_cast<const float2*>(X1 + o1) : make_float2(0.f, 0.f);
        float uc[4], ub[4];  // u of this row and of the row below
        urow(-1, ub);
        urow(0, uc);
        xp[0] = om * ub[0] + eN.x; xp[1] = om * ub[1] + eN.x; xp[2] = om * ub[2] + eN.y; xp[3] = om * ub[3] + eN.y;
        xc[0] = om * uc[0] + ec.x; xc[1] = om * uc[1] + ec.x; xc[2] = om * uc[2] + ec.y; xc[3] = om * uc[3] + ec.y;
        float sw = 0.f, sz = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const Row5 r = mb_row(cw[j], F, left, right, j == 7 && bot);
          const int o = o1 + (j >> 1) * p1;
          float ew = 0.f, ee = 0.f;
          if (act) {
            ew = left ? 0.f : X1[o - 1];  // the wall is a face towards the value 0 (X1's guard ring is not zero)
            ee = right ? 0.f : X1[o + 2];
          }
          // x2 of the row below: row j+1 of my patch (its aggregate row changes at even j+1) or the halo row
          float2 en = ec;
          if (j == 7) en = eS;
          else if (j & 1) en = act ? *reinterpret_cast<const float2*>(X1 + o + p1) : make_float2(0.f, 0.f);
          urow(j + 1, ub);
          xn[0] = om * ub[0] + en.x; xn[1] = om * ub[1] + en.x; xn[2] = om * ub[2] + en.y; xn[3] = om * ub[3] + en.y;
          float uw = __shfl_up_sync(0xffffffffu, uc[3], 1), ue = __shfl_down_sync(0xffffffffu, uc[0], 1);
          if (left) uw = 0.f;
          if (right || lane == 31) ue = 0.f;
          const float xw = om * uw + ew, xe = om * ue + ee;
          const float d0 = (r.cw + r.c01) + (nn[0] + r.s0), d1 = (r.c01 + r.c12) + (nn[1] + r.s1);
          const float d2 = (r.c12 + r.c23) + (nn[2] + r.s2), d3 = (r.c23 + r.ce) + (nn[3] + r.s3);
          const float s0 = (r.cw * xw + r.c01 * xc[1]) + (nn[0] * xp[0] + r.s0 * xn[0]);
          const float s1 = (r.c01 * xc[0] + r.c12 * xc[2]) + (nn[1] * xp[1] + r.s1 * xn[1]);
          const float s2 = (r.c12 * xc[1] + r.c23 * xc[3]) + (nn[2] * xp[2] + r.s2 * xn[2]);
          const float s3 = (r.c23 * xc[2] + r.ce * xe) + (nn[3] * xp[3] + r.s3 * xn[3]);
          float4 wv;
          wv.x = om1 * xc[0] + om * uc[0] + om * __fdividef(s0, d0);
          wv.y = om1 * xc[1] + om * uc[1] + om * __fdividef(s1, d1);
          wv.z = om1 * xc[2] + om * uc[2] + om * __fdividef(s2, d2);
          wv.w = om1 * xc[3] + om * uc[3] + om * __fdividef(s3, d3);
          if (!act) wv = make_float4(0.f, 0.f, 0.f, 0.f);
          w[j][0] = wv.x; w[j][1] = wv.y; w[j][2] = wv.z; w[j][3] = wv.w;
          const float q0 = d0 * uc[0], q1 = d1 * uc[1], q2 = d2 * uc[2], q3 = d3 * uc[3];  // a = D u
          sw += (wv.x * q0 + wv.y * q1) + (wv.z * q2 + wv.w * q3);
          if (with_z) {
            const float4 zv = Yg[ZG + j * MB_THREADS];
            sz += (zv.x * q0 + zv.y * q1) + (zv.z * q2 + zv.w * q3);
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            xp[i] = xc[i];
            xc[i] = xn[i];
            uc[i] = ub[i];
          }
          ec = en;
          nn[0] = r.s0; nn[1] = r.s1; nn[2] = r.s2; nn[3] = r.s3;
        }
        swq = act ? sw : 0.f;
        szq = act ? sz : 0.f;
      }
    };

    double rz = 0, thr2 = thr2_in;  // thr2 is updated multiplicatively: it stays a per-thread register
#define S_rz_flush SC[0]
#define S_rr_true SC[1]
#define S_bb SC[2]
#define S_r1 SC[3]
#define S_ratio SC[4]
#define S_ql SC[6]
#define S_qr SC[7]
#define S_rz_stop SC[8]
    S_rz_flush = 0; S_rr_true = 0; S_bb = 0; S_r1 = 0; S_ratio = 1;  S_ql = 0; S_qr = 0;
    long long it = 0;
    int status = 1;
    bool first = true, restart = false, fresh_p = true;  // fresh_p: p has not been set for this image yet

    while (true) {
      // ================= reliable update: x += y (FP64), y = 0, z = M (b - A x), residual in FP64 =================
      if (!first) {
        if (act) {
          // all loads of four rows before their stores (x and y may alias as far as the compiler knows, which would
          // chain eight L2 round trips)
#pragma unroll
          for (int h = 0; h < 8; h += 4) {
            float4 yv[4];
            double2 xa[4], xb[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              yv[j] = Yg[(h + j) * MB_THREADS];
              const double2* xp = reinterpret_cast<const double2*>(x64 + (y0 + h + j) * nx + x0);
              xa[j] = xp[0];
              xb[j] = xp[1];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              double2* xp = reinterpret_cast<double2*>(x64 + (y0 + h + j) * nx + x0);
              xp[0] = make_double2(xa[j].x + yv[j].x, xa[j].y + yv[j].y);
              xp[1] = make_double2(xb[j].x + yv[j].z, xb[j].y + yv[j].w);
              Yg[(h + j) * MB_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        __syncthreads();
      }
      double s4[4] = {0, 0, 0, 0};
      if (act) {
        // TRUE residual in FP64, flux form r_P = sum_f c_f (x_nb - x_P) with the wall a face towards TL / TR; rows of x
        // roll through registers (each value is loaded once), the FP64 face conductances are selected per row
        mb_fresh(cw);
        float nn[4];
        double nd[4];
        mb_north(cw[0], F, top, nn);
        mb_north_d(cw[0], C, top, nd);
        auto load_row = [&](int yy, double (&v)[4]) {
          const double2* xp = reinterpret_cast<const double2*>(x64 + yy * nx + x0);
          const double2 a = xp[0], b = xp[1];
          v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        };
        double xn[4] = {0, 0, 0, 0}, xc[4], xs[4];
        if (!top) load_row(y0 - 1, xn);
        load_row(y0, xc);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int yy = y0 + j;
          if (j == 7 && bot) xs[0] = xs[1] = xs[2] = xs[3] = 0.0;
          else load_row(yy + 1, xs);
          const double xw = left ? C.tl : x64[yy * nx + x0 - 1], xe = right ? C.tr : x64[yy * nx + x0 + 4];
          const unsigned cwj = cw[j];
          const Row5 r = mb_row(cwj, F, left, right, j == 7 && bot);
          const Row5d rd = mb_row_d(cwj, C, left, right, j == 7 && bot);
          const float dd[4] = {(r.cw + r.c01) + (nn[0] + r.s0), (r.c01 + r.c12) + (nn[1] + r.s1),
                               (r.c12 + r.c23) + (nn[2] + r.s2), (r.c23 + r.ce) + (nn[3] + r.s3)};
          const double f01 = rd.c01 * (xc[1] - xc[0]), f12 = rd.c12 * (xc[2] - xc[1]), f23 = rd.c23 * (xc[3] - xc[2]);
          const double rv[4] = {(rd.cw * (xw - xc[0]) + f01) + (nd[0] * (xn[0] - xc[0]) + rd.s0 * (xs[0] - xc[0])),
                                (f12 - f01) + (nd[1] * (xn[1] - xc[1]) + rd.s1 * (xs[1] - xc[1])),
                                (f23 - f12) + (nd[2] * (xn[2] - xc[2]) + rd.s2 * (xs[2] - xc[2])),
                                (rd.ce * (xe - xc[3]) - f23) + (nd[3] * (xn[3] - xc[3]) + rd.s3 * (xs[3] - xc[3]))};
          const double b0 = left ? rd.cw * C.tl : 0.0, b3 = right ? rd.ce * C.tr : 0.0;  // right-hand side entries
          *reinterpret_cast<float4*>(Uown + j * MB_PITCH) =
              make_float4(__fdividef((float)rv[0], dd[0]), __fdividef((float)rv[1], dd[1]), __fdividef((float)rv[2], dd[2]),
                          __fdividef((float)rv[3], dd[3]));
#pragma unroll
          for (int i = 0; i < 4; i++) {
            s4[1] += rv[i] * rv[i];
            s4[3] += fabs(rv[i]);
            xn[i] = xc[i];
            xc[i] = xs[i];
          }
          s4[2] += b0 * b0 + b3 * b3;
          nn[0] = r.s0; nn[1] = r.s1; nn[2] = r.s2; nn[3] = r.s3;
          nd[0] = rd.s0; nd[1] = rd.s1; nd[2] = rd.s2; nd[3] = rd.s3;
        }
      }
      block_allsum<4>(s4, red, bc);
      S_rr_true = s4[1];
      S_r1 = s4[3];
      if (first) S_bb = s4[2];
      // ---- stop decisions use only FP64 quantities of the flushed field (and come before the cycle: the last
      //      reliable update of an image needs no z) ----
      const bool conv = !(S_rr_true > thr2 * S_bb);
      if (conv || it >= max_iter) {
        double f2[2] = {0, 0};
        const uint8_t* m = MB_M;
        for (int row = tid; row < ny; row += MB_THREADS) {
          const int l = row * nx, e = l + nx - 1;
          f2[0] += ((m[l] == 1) ? C.wall[1] : C.wall[0]) * (x64[l] - C.tl);
          f2[1] += ((m[e] == 1) ? C.wall[1] : C.wall[0]) * (C.tr - x64[e]);
        }
        block_allsum<2>(f2, red, bc);
        S_ql = f2[0];
        S_qr = f2[1];
        const double keff = (S_ql + S_qr) / 2 / (C.tr - C.tl);
        if (conv && (fabs(S_ql - S_qr) <= flux_tol * fabs(keff) || thr2 < 1e-28)) {
          status = 0;
          break;
        }
        if (it >= max_iter) break;
        thr2 *= 1e-2;
        if (!(S_rr_true > thr2 * S_bb)) {  // still satisfied: re-evaluate with the tighter target
          first = false;
          continue;
        }
      }
      {  // z = M r straight into its slice; r.z with r rounded to FP32
        float swq, szq;
        vcycle(false, swq, szq);
#pragma unroll
        for (int j = 0; j < 8; j++) Yg[ZG + j * MB_THREADS] = make_float4(w[j][0], w[j][1], w[j][2], w[j][3]);
        double s1[1] = {(double)swq};
        block_allsum<1>(s1, red, bc);
        rz = s1[0];
      }
      S_rz_flush = rz;
      S_ratio = rz > 0 ? S_rr_true / rz : 1.0;
      if (first || restart || fresh_p) {  // p = z (start, or restart after a window that did not reach its target)
        if (act)
#pragma unroll
          for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(Pown + j * MB_PITCH) = make_float4(w[j][0], w[j][1], w[j][2], w[j][3]);
        first = false;
        restart = false;
        fresh_p = false;
        __syncthreads();
      }

      // ================================ FP32 CG iterations ================================
      S_rz_stop = fmax(delta2 * S_rz_flush, thr2 * S_bb / S_ratio);
      bool broke = false;
      while (true) {
        float spq = 0.f, szq = 0.f, swq = 0.f;
        {  // q = A p from the guarded image of p; u = q / a_P
          mb_fresh(cw);
          float nn[4];
          mb_north(cw[0], F, top, nn);
          // all ten rows of p first (L2 round trips in flight together); rows -1 and 8 are the neighbours' or zero guards
          float4 prow[10];
#pragma unroll
          for (int j = 0; j < 10; j++) prow[j] = __ldcg(reinterpret_cast<const float4*>(Pown + (j - 1) * MB_PITCH));
          float4 hn = prow[0];
          float4 pc = prow[1];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 pn = prow[j + 2];
            float pw = __shfl_up_sync(0xffffffffu, pc.w, 1), pe = __shfl_down_sync(0xffffffffu, pc.x, 1);
            if (left) pw = 0.f;
            if (right || lane == 31) pe = 0.f;
            const Row5 r = mb_row(cw[j], F, left, right, j == 7 && bot);
            const float f01 = r.c01 * (pc.x - pc.y), f12 = r.c12 * (pc.y - pc.z), f23 = r.c23 * (pc.z - pc.w);
            const float q0 = (r.cw * (pc.x - pw) + f01) + (nn[0] * (pc.x - hn.x) + r.s0 * (pc.x - pn.x));
            const float q1 = (f12 - f01) + (nn[1] * (pc.y - hn.y) + r.s1 * (pc.y - pn.y));
            const float q2 = (f23 - f12) + (nn[2] * (pc.z - hn.z) + r.s2 * (pc.z - pn.z));
            const float q3 = (r.ce * (pc.w - pe) - f23) + (nn[3] * (pc.w - hn.w) + r.s3 * (pc.w - pn.w));
            const float d0 = (r.cw + r.c01) + (nn[0] + r.s0), d1 = (r.c01 + r.c12) + (nn[1] + r.s1);
            const float d2 = (r.c12 + r.c23) + (nn[2] + r.s2), d3 = (r.c23 + r.ce) + (nn[3] + r.s3);
            if (act)
              *reinterpret_cast<float4*>(Uown + j * MB_PITCH) =
                  make_float4(__fdividef(q0, d0), __fdividef(q1, d1), __fdividef(q2, d2), __fdividef(q3, d3));
            spq += (pc.x * q0 + pc.y * q1) + (pc.z * q2 + pc.w * q3);
            nn[0] = r.s0; nn[1] = r.s1; nn[2] = r.s2; nn[3] = r.s3;
            hn = pc;
            pc = pn;
          }
          if (!act) spq = 0.f;
        }
        vcycle(true, swq, szq);  // w = M q (registers), sums w.q and z.q
        // one barrier per reduction: FP32 inside a warp, FP64 across the 16 warps, every warp sums the 16
        // partials itself (fixed order -> deterministic); `red` is double-buffered by parity
        double s3[3];
        {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            spq += __shfl_xor_sync(0xffffffffu, spq, o);
            szq += __shfl_xor_sync(0xffffffffu, szq, o);
            swq += __shfl_xor_sync(0xffffffffu, swq, o);
          }
          double* rb = red + (it & 1) * 64;
          if (lane == 0) {
            rb[wrp * 4 + 0] = (double)spq;
            rb[wrp * 4 + 1] = (double)szq;
            rb[wrp * 4 + 2] = (double)swq;
          }
          __syncthreads();
          const int l16 = lane & 15;
          s3[0] = rb[l16 * 4 + 0];
          s3[1] = rb[l16 * 4 + 1];
          s3[2] = rb[l16 * 4 + 2];
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            s3[0] += __shfl_xor_sync(0xffffffffu, s3[0], o);
            s3[1] += __shfl_xor_sync(0xffffffffu, s3[1], o);
            s3[2] += __shfl_xor_sync(0xffffffffu, s3[2], o);
          }
        }
        if (!(s3[0] > 0.0)) {  // breakdown (cannot happen for an SPD operator short of NaN)
          status = 2;
          broke = true;
          break;
        }
        const double alpha = rz / s3[0];
        double rzn = rz - 2.0 * alpha * s3[1] + alpha * alpha * s3[2];
        const bool bad = !(rzn > 0.0);  // recurrence lost positivity: force a reliable update
        if (bad) rzn = rz * 1e-2;
        const float al = (float)alpha, be = (float)(rzn / rz);
        rz = rzn;
        it++;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float4 pv = __ldcg(reinterpret_cast<const float4*>(Pown + j * MB_PITCH));
          float4 yv = Yg[j * MB_THREADS];
          float4 zv = Yg[ZG + j * MB_THREADS];
          zv.x -= al * w[j][0]; zv.y -= al * w[j][1]; zv.z -= al * w[j][2]; zv.w -= al * w[j][3];
          yv.x += al * pv.x; yv.y += al * pv.y; yv.z += al * pv.z; yv.w += al * pv.w;
          pv.x = zv.x + be * pv.x; pv.y = zv.y + be * pv.y; pv.z = zv.z + be * pv.z; pv.w = zv.w + be * pv.w;
          Yg[j * MB_THREADS] = yv;
          Yg[ZG + j * MB_THREADS] = zv;
          if (act) *reinterpret_cast<float4*>(Pown + j * MB_PITCH) = pv;
        }
        __syncthreads();
        if (rz < S_rz_stop || it >= max_iter) break;
        if (bad || (it & 127) == 0) {  // recurrence lost positivity or stalled: reliable update + CG restart
          restart = true;
          break;
        }
      }
      if (broke) break;
    }

    if (tid == 0) {
      BatchOut o;
      o.ql = S_ql;
      o.qr = S_qr;
      o.rr = S_rr_true;
      o.bb = S_bb;
      o.r1 = S_r1;
      o.it = it;
      o.status = status;
      o.pad = 0;
      res[img] = o;
    }
    __syncthreads();
  }
}

#undef MB_M
#undef S_rz_flush
#undef S_rr_true
#undef S_bb
#undef S_r1
#undef S_ratio
#undef S_ql
#undef S_qr
#undef S_rz_stop

}  // namespace kb
