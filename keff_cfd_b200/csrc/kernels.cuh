// kernels.cuh -- device kernels of the grid solver (2D images are nz == 1 volumes).
//
// Data layout in HBM: every field is double[nz(+2 halo planes)][ny][nx], x fastest -- the
// reference's index z*ny*nx + y*nx + x (Keff3D/keff3D.h:691).  The structure is one "code" byte per
// cell (common.cuh).  Nothing else is stored: no coefficient matrix, no right-hand side.
#pragma once
#include "common.cuh"

namespace kb {

// ------------------------------------------------------------------------------------------------
// structure -> code bytes.  Replaces the phase->k maps (Keff2D/keff2D.cpp:108-121,
// Keff3D/keff3D.cpp:105-120) and the harmonic-mean selection inside DiscretizeMatrixCD2D/3D.
// m points at local plane 0; planes -1 and nz are readable when lo/hi are set (slab neighbours).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_make_code(const uint8_t* __restrict__ m, uint8_t* __restrict__ code, Dom D, int lo,
                                                   int hi) {
  const long plane = (long)D.nx * D.ny;
  const int rows = D.ny * D.nz;  // rows (y,z) round-robin over the grid: no per-cell division
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = row % D.ny, z = row / D.ny;
    const bool hasN = y > 0, hasS = y < D.ny - 1, hasF = z > 0 || lo, hasB = z < D.nz - 1 || hi;
    const long base = (long)row * D.nx;
    for (int x = threadIdx.x; x < D.nx; x += blockDim.x) {
      const long i = base + x;
      const unsigned me = m[i] == 1;
      unsigned c = me;
      if (x > 0 && (unsigned)(m[i - 1] == 1) != me) c |= C_W;
      if (x < D.nx - 1 && (unsigned)(m[i + 1] == 1) != me) c |= C_E;
      if (hasS && (unsigned)(m[i + D.nx] == 1) != me) c |= C_S;
      if (hasN && (unsigned)(m[i - D.nx] == 1) != me) c |= C_N;
      if (hasF && (unsigned)(m[i - plane] == 1) != me) c |= C_F;
      if (hasB && (unsigned)(m[i + plane] == 1) != me) c |= C_B;
      code[i] = (uint8_t)c;
    }
  }
}

// Initial field: linear ramp between the wall temperatures evaluated at the cell centres, or the
// constant TL the reference actually starts from (SolInitLinear, Keff2D/keff2D.h:584-611, whose
// integer division collapses the ramp).
__global__ void __launch_bounds__(256) k_init_field(double* __restrict__ T, Dom D, double tl, double tr, int constant) {
  const int rows = D.ny * D.nz;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    double* Tr = T + (long)row * D.nx;
    for (int x = threadIdx.x; x < D.nx; x += blockDim.x) Tr[x] = constant ? tl : tl + (tr - tl) * ((x + 0.5) / D.nx);
  }
}

enum { MODE_APPLY = 0, MODE_INIT = 1, MODE_RESID = 2, MODE_INIT_R = 3 };
// MODE_APPLY: o0 = A*in,              partials {in.A in, -, -}
// MODE_INIT : o0 = r = b - A*in, o1 = p = r/a_P, partials {r.z, r.r, b.b}
// MODE_RESID: no output,              partials {-, r.r, sum|r|}
// MODE_INIT_R: o0 = r = b - A*in,       partials {-, r.r, b.b}   (the multigrid path builds its own z)

template <int MODE>
__device__ __forceinline__ void cell_epilogue(const Coef& C, unsigned c, bool x0, bool x1, const Faces& F,
                                              double pc, double q, double& out0, double& out1, double* acc) {
  if (MODE == MODE_APPLY) {
    out0 = q;
    acc[0] += pc * q;
  } else {
    const double b = rhs_of(C, c, x0, x1);
    const double r = b - q;
    if (MODE == MODE_INIT) {
      const double z = r / F.diag;
      out0 = r;
      out1 = z;
      acc[0] += r * z;
      acc[1] += r * r;
      acc[2] += b * b;
    } else if (MODE == MODE_INIT_R) {
      out0 = r;
      acc[1] += r * r;
      acc[2] += b * b;
    } else {
      acc[1] += r * r;
      acc[2] += fabs(r);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Plain stencil kernel (any nx): rows are distributed round-robin over a fixed grid, neighbours are
// read through L1/L2.  Fallback for odd nx and cross-check for the TMA kernel.
// in/o0/o1 point at local plane 0; `in` has readable halo planes.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_stencil_plain(const double* __restrict__ in, const uint8_t* __restrict__ code,
                                                       double* __restrict__ o0, double* __restrict__ o1, const Coef C,
                                                       const Dom D, const Scal* __restrict__ S,
                                                       double* __restrict__ partials) {
  __shared__ double red[32 * 3];
  if (S->done) return;
  const long plane = (long)D.nx * D.ny;
  const long rows = (long)D.ny * D.nz;
  double acc[3] = {0, 0, 0};
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % D.ny), z = (int)(row / D.ny);
    const bool y0 = y == 0, y1 = y == D.ny - 1, z0 = D.z0 + z == 0, z1 = D.z0 + z == D.nzg - 1;
    const long base = row * D.nx;
    for (int x = threadIdx.x; x < D.nx; x += blockDim.x) {
      const long i = base + x;
      const bool x0 = x == 0, x1 = x == D.nx - 1;
      const unsigned c = code[i];
      const Faces F = faces_of(C, c, x0, x1, y0, y1, z0, z1);
      const double pc = in[i];
      double q = F.diag * pc;
      if (!x0) q -= F.w * in[i - 1];
      if (!x1) q -= F.e * in[i + 1];
      if (!y0) q -= F.n * in[i - D.nx];
      if (!y1) q -= F.s * in[i + D.nx];
      if (!z0) q -= F.f * in[i - plane];
      if (!z1) q -= F.b * in[i + plane];
      double a = 0, b2 = 0;
      cell_epilogue<MODE>(C, c, x0, x1, F, pc, q, a, b2, acc);
      if (MODE != MODE_RESID) o0[i] = a;
      if (MODE == MODE_INIT) o1[i] = b2;
    }
  }
  block_sum<3>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0];
    partials[blockIdx.x * 3 + 1] = acc[1];
    partials[blockIdx.x * 3 + 2] = acc[2];
  }
}

// ------------------------------------------------------------------------------------------------
// TMA stencil kernel (nx even).  Work item = (64 x 16) xy tile x one z chunk.  A CTA marches its item
// in z: planes (2-cell x halo, 1-row y halo, zero-filled outside the domain by the TMA unit) stream
// into an NST-deep shared-memory ring via cp.async.bulk.tensor + mbarrier.  Each thread owns a 2x2
// block of cells and keeps the current and next plane's values in registers, so a field value is
// fetched from L2/HBM once per tile and read from shared memory 3 times.
//
// Flux form: q_P = sum_faces c_f (p_P - p_f).  Each face conductance and flux is evaluated once: the
// four faces inside my 2x2 block feed both of their cells, and the flux towards the next plane is
// carried to the next step.  Domain boundaries cost nothing in the loop: the fixed-temperature wall
// is a face of conductance c_wall whose neighbour value is the TMA's zero fill, an adiabatic face is
// a face of conductance 0 -- both folded into per-thread "same phase" constants of the four outward
// faces (the code bits of boundary faces are 0 by construction).
// Items are dealt round-robin to a fixed grid (deterministic partial sums).
// ------------------------------------------------------------------------------------------------
constexpr int TX = 64, TY = 16, BOXW = TX + 4, BOXH = TY + 2, NST = 4, NTHR = 256;
constexpr int STAGE_BYTES = BOXW * BOXH * 8;
constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

// staged plane of the TMA stencil kernel per input type.  FP64 input (x, and p of the diagonally preconditioned path):
// box of TX + 4 doubles starting 2 cells left of the tile.  FP32 input (p of the multigrid path, which is built from
// the FP32 z anyway): box of TX + 8 floats starting 4 cells left, so that box rows stay multiples of 16 bytes.
template <typename T>
struct StencilTile;
template <>
struct StencilTile<double> {
  static constexpr int boxw = BOXW, hx = 2, bytes = STAGE_BYTES, stride = STAGE_STRIDE;
};
template <>
struct StencilTile<float> {
  static constexpr int boxw = TX + 8, hx = 4, bytes = (TX + 8) * BOXH * 4, stride = (bytes + 127) / 128 * 128;
};
__device__ __forceinline__ double2 tile_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 tile_ld2(const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return make_double2((double)v.x, (double)v.y);
}

struct Plan {
  int ntx, nty, nzc, zlen, nitems;
  int bfirst;  // z-slabs: deal the two chunks that touch the slab boundaries first (their planes travel to the
               // neighbours while the interior chunks are still being worked on)
};
__device__ __forceinline__ void plan_item(const Plan& P, int item, int& tx, int& ty, int& zc) {
  tx = item % P.ntx;
  const int t1 = item / P.ntx;
  ty = t1 % P.nty;
  const int zi = t1 / P.nty;
  zc = (!P.bfirst || P.nzc < 2) ? zi : (zi == 0 ? 0 : (zi == 1 ? P.nzc - 1 : zi - 1));
}

__device__ __forceinline__ double pick(bool c, double a, double b) { return c ? a : b; }

template <int MODE, typename TIN = double>
__global__ void __launch_bounds__(NTHR) k_stencil_tma(const __grid_constant__ CUtensorMap tm,
                                                      const uint8_t* __restrict__ code, double* __restrict__ o0,
                                                      double* __restrict__ o1, const Coef C, const Dom D, const Plan P,
                                                      const Scal* __restrict__ S, double* __restrict__ partials) {
  using TT = StencilTile<TIN>;
  __shared__ __align__(128) unsigned char tiles[NST * TT::stride];
  __shared__ __align__(8) uint64_t full[NST];
  __shared__ double red[32 * 3];
  if (S->done) return;
  const int tid = threadIdx.x, lx = (tid & 31) * 2, ly = (tid >> 5) * 2;
  if (tid == 0) {
    for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  uint32_t seq = 0;  // planes streamed so far by this CTA (identical in every thread)
  double acc[3] = {0, 0, 0};
  const long plane = (long)D.nx * D.ny;
  const int coff = (ly + 1) * TT::boxw + lx + TT::hx;  // cell a of my 2x2 block inside a staged plane
  // interior "same phase" conductances [phase]
  const double sx0 = C.same[0][0], sx1 = C.same[0][1], sy0 = C.same[1][0], sy1 = C.same[1][1];
  const double sz0 = C.same[2][0], sz1 = C.same[2][1];
  const double dfx = C.diff[0], dfy = C.diff[1], dfz = C.diff[2];

  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    int tx, ty, zc;
    plan_item(P, item, tx, ty, zc);
    const int xo = tx * TX, yo = ty * TY, zb = zc * P.zlen;
    const int len = min(D.nz, zb + P.zlen) - zb, nplanes = len + 2;
    const int gx = xo + lx, gy = yo + ly;
    const bool colv = gx < D.nx;                       // nx even: both columns valid or none
    const bool row0 = colv && gy < D.ny, row1 = colv && gy + 1 < D.ny;
    // outward faces of my block: W (cells a,c), E (b,d), N (a,b), S (c,d); inner y face (a|c, b|d)
    const bool ax0 = gx == 0, bx1 = gx + 1 == D.nx - 1, yt = gy == 0, yb = gy + 1 == D.ny - 1;
    const double w0 = ax0 ? C.wall[0] : sx0, w1 = ax0 ? C.wall[1] : sx1;
    const double e0 = bx1 ? C.wall[0] : sx0, e1 = bx1 ? C.wall[1] : sx1;
    const double n0 = yt ? 0.0 : sy0, n1 = yt ? 0.0 : sy1;
    const double s0 = (yb || !row1) ? 0.0 : sy0, s1 = (yb || !row1) ? 0.0 : sy1;
    const double m0 = row1 ? sy0 : 0.0, m1 = row1 ? sy1 : 0.0;

    __syncthreads();  // ring is idle (previous item fully consumed; barriers initialised)
    if (tid == 0) {
      const int npre = min(NST, nplanes);
      for (int k = 0; k < npre; k++) {
        const uint32_t st = (seq + k) % NST;
        mbar_expect_tx(&full[st], TT::bytes);
        // tensor z index = local plane + 1 (field arrays carry one halo plane below plane 0)
        tma_load_3d(tiles + st * TT::stride, &tm, &full[st], xo - TT::hx, yo - 1, zb + k);
      }
    }
    long ci = ((long)zb * D.ny + gy) * D.nx + gx;  // global index of cell a in plane zb
    unsigned cn0 = row0 ? *reinterpret_cast<const unsigned short*>(code + ci) : 0u;
    unsigned cn1 = row1 ? *reinterpret_cast<const unsigned short*>(code + ci + D.nx) : 0u;
    double2 pt, pb, nt, nb;               // current / next plane, top row (a,b) and bottom row (c,d)
    double ga = 0.0, gb = 0.0, gc = 0.0, gd = 0.0;  // flux into my cells from the plane below
    {
      const uint32_t q0 = seq % NST, q1 = (seq + 1) % NST;
      mbar_wait(&full[q0], (seq / NST) & 1);
      const TIN* lo = reinterpret_cast<const TIN*>(tiles + q0 * TT::stride) + coff;
      const double2 mt = tile_ld2(lo), mb = tile_ld2(lo + TT::boxw);
      mbar_wait(&full[q1], ((seq + 1) / NST) & 1);
      const TIN* c0 = reinterpret_cast<const TIN*>(tiles + q1 * TT::stride) + coff;
      pt = tile_ld2(c0);
      pb = tile_ld2(c0 + TT::boxw);
      if (D.z0 + zb != 0) {  // face between plane zb-1 and zb
        const unsigned ca = cn0 & 0xffu, cb = cn0 >> 8, cc = cn1 & 0xffu, cd = cn1 >> 8;
        ga = pick(ca & C_F, dfz, pick(ca & C_PH, sz1, sz0)) * (mt.x - pt.x);
        gb = pick(cb & C_F, dfz, pick(cb & C_PH, sz1, sz0)) * (mt.y - pt.y);
        gc = pick(cc & C_F, dfz, pick(cc & C_PH, sz1, sz0)) * (mb.x - pb.x);
        gd = pick(cd & C_F, dfz, pick(cd & C_PH, sz1, sz0)) * (mb.y - pb.y);
      }
    }

    for (int k = 1; k <= len; k++) {
      const int lz = zb + k - 1;
      __syncthreads();  // every thread is done with plane k-1's stage
      if (tid == 0 && k - 1 + NST < nplanes) {
        const uint32_t st = (seq + k - 1) % NST;  // == stage of plane k-1+NST
        mbar_expect_tx(&full[st], TT::bytes);
        tma_load_3d(tiles + st * TT::stride, &tm, &full[st], xo - TT::hx, yo - 1, zb + k - 1 + NST);
      }
      const uint32_t sc = (seq + k) % NST, sn = (seq + k + 1) % NST;
      mbar_wait(&full[sn], ((seq + k + 1) / NST) & 1);
      const TIN* cur = reinterpret_cast<const TIN*>(tiles + sc * TT::stride) + coff;
      const TIN* nxt = reinterpret_cast<const TIN*>(tiles + sn * TT::stride) + coff;
      nt = tile_ld2(nxt);
      nb = tile_ld2(nxt + TT::boxw);
      const unsigned k0 = cn0, k1 = cn1;
      if (k < len) {
        if (row0) cn0 = *reinterpret_cast<const unsigned short*>(code + ci + plane);
        if (row1) cn1 = *reinterpret_cast<const unsigned short*>(code + ci + plane + D.nx);
      }
      const double2 pn = tile_ld2(cur - TT::boxw);
      const double2 ps = tile_ld2(cur + 2 * TT::boxw);
      const double pwt = cur[-1], pet = cur[2], pwb = cur[TT::boxw - 1], peb = cur[TT::boxw + 2];
      const unsigned ca = k0 & 0xffu, cb = k0 >> 8, cc = k1 & 0xffu, cd = k1 >> 8;
      const bool fa = ca & C_PH, fb = cb & C_PH, fc = cc & C_PH, fd = cd & C_PH;
      // start from the flux that entered from the plane below
      double qa = -ga, qb = -gb, qc = -gc, qd = -gd;
      double g;
      // x faces: W wall/neighbour, inner, E  (top row, bottom row)
      qa += pick(ca & C_W, dfx, pick(fa, w1, w0)) * (pt.x - pwt);
      g = pick(ca & C_E, dfx, pick(fa, sx1, sx0)) * (pt.x - pt.y);
      qa += g;
      qb -= g;
      qb += pick(cb & C_E, dfx, pick(fb, e1, e0)) * (pt.y - pet);
      qc += pick(cc & C_W, dfx, pick(fc, w1, w0)) * (pb.x - pwb);
      g = pick(cc & C_E, dfx, pick(fc, sx1, sx0)) * (pb.x - pb.y);
      qc += g;
      qd -= g;
      qd += pick(cd & C_E, dfx, pick(fd, e1, e0)) * (pb.y - peb);
      // y faces: N, inner, S  (left column, right column)
      qa += pick(ca & C_N, dfy, pick(fa, n1, n0)) * (pt.x - pn.x);
      g = pick(ca & C_S, dfy, pick(fa, m1, m0)) * (pt.x - pb.x);
      qa += g;
      qc -= g;
      qc += pick(cc & C_S, dfy, pick(fc, s1, s0)) * (pb.x - ps.x);
      qb += pick(cb & C_N, dfy, pick(fb, n1, n0)) * (pt.y - pn.y);
      g = pick(cb & C_S, dfy, pick(fb, m1, m0)) * (pt.y - pb.y);
      qb += g;
      qd -= g;
      qd += pick(cd & C_S, dfy, pick(fd, s1, s0)) * (pb.y - ps.y);
      // z faces towards the next plane (absent at the global top)
      if (D.z0 + lz != D.nzg - 1) {
        ga = pick(ca & C_B, dfz, pick(fa, sz1, sz0)) * (pt.x - nt.x);
        gb = pick(cb & C_B, dfz, pick(fb, sz1, sz0)) * (pt.y - nt.y);
        gc = pick(cc & C_B, dfz, pick(fc, sz1, sz0)) * (pb.x - nb.x);
        gd = pick(cd & C_B, dfz, pick(fd, sz1, sz0)) * (pb.y - nb.y);
        qa += ga;
        qb += gb;
        qc += gc;
        qd += gd;
      }
      if (MODE == MODE_APPLY) {
        if (row0) {
          *reinterpret_cast<double2*>(o0 + ci) = make_double2(qa, qb);
          acc[0] += pt.x * qa + pt.y * qb;
        }
        if (row1) {
          *reinterpret_cast<double2*>(o0 + ci + D.nx) = make_double2(qc, qd);
          acc[0] += pb.x * qc + pb.y * qd;
        }
      } else {
        const bool z0 = D.z0 + lz == 0, z1 = D.z0 + lz == D.nzg - 1;
        Faces Fa, Fb;
        double2 u, v;
        if (row0) {
          if (MODE == MODE_INIT) {
            Fa.diag = diag_of(C, ca, ax0, false, yt, gy == D.ny - 1, z0, z1);
            Fb.diag = diag_of(C, cb, false, bx1, yt, gy == D.ny - 1, z0, z1);
          }
          cell_epilogue<MODE>(C, ca, ax0, false, Fa, pt.x, qa, u.x, v.x, acc);
          cell_epilogue<MODE>(C, cb, false, bx1, Fb, pt.y, qb, u.y, v.y, acc);
          if (MODE != MODE_RESID) *reinterpret_cast<double2*>(o0 + ci) = u;
          if (MODE == MODE_INIT) *reinterpret_cast<double2*>(o1 + ci) = v;
        }
        if (row1) {
          if (MODE == MODE_INIT) {
            Fa.diag = diag_of(C, cc, ax0, false, false, yb, z0, z1);
            Fb.diag = diag_of(C, cd, false, bx1, false, yb, z0, z1);
          }
          cell_epilogue<MODE>(C, cc, ax0, false, Fa, pb.x, qc, u.x, v.x, acc);
          cell_epilogue<MODE>(C, cd, false, bx1, Fb, pb.y, qd, u.y, v.y, acc);
          if (MODE != MODE_RESID) *reinterpret_cast<double2*>(o0 + ci + D.nx) = u;
          if (MODE == MODE_INIT) *reinterpret_cast<double2*>(o1 + ci + D.nx) = v;
        }
      }
      pt = nt;
      pb = nb;
      ci += plane;
    }
    seq += nplanes;
  }
  block_sum<3>(acc, red);
  if (tid == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0];
    partials[blockIdx.x * 3 + 1] = acc[1];
    partials[blockIdx.x * 3 + 2] = acc[2];
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised TMA stencil kernel (nx a multiple of 16): the same tiling and flux arithmetic as
// k_stencil_tma, but the plane ring is run by a dedicated producer warp and per-stage full/empty
// mbarriers instead of one __syncthreads per plane, and the code bytes of a plane arrive by TMA in
// the same stage as the field box (a second tensor map over the u8 code array; its row stride must be
// a multiple of 16 bytes, hence nx % 16 == 0).  Consumer warps (8: two tile rows each) only ever
// wait for "plane k+1 has landed"; a warp that is ahead keeps going until the ring is exhausted, and
// the ring keeps running across work items.  Stage layout: field box (BOXW x BOXH doubles, padded to
// 128 B) followed by the TX x TY code bytes.
// ------------------------------------------------------------------------------------------------
#ifndef WS_OCC
#define WS_OCC 2
#endif
constexpr int WS_NST = WS_OCC == 2 ? 9 : 6, WS_THR = NTHR + 32, WS_CODE_BYTES = TX * TY;
constexpr int WS_STAGE = STAGE_STRIDE + WS_CODE_BYTES;
constexpr int WS_SMEM = WS_NST * WS_STAGE + 128;

template <int MODE>
__global__ void __launch_bounds__(WS_THR, WS_OCC) k_stencil_ws(const __grid_constant__ CUtensorMap tm,
                                                          const __grid_constant__ CUtensorMap tmc,
                                                          double* __restrict__ o0, double* __restrict__ o1,
                                                          const Coef C, const Dom D, const Plan P,
                                                          const Scal* __restrict__ S, double* __restrict__ partials) {
  extern __shared__ unsigned char ws_raw[];
  __shared__ __align__(8) uint64_t full[WS_NST], empty[WS_NST];
  __shared__ double red[32 * 3];
  if (S->done) return;
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < WS_NST; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NTHR / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  double acc[3] = {0, 0, 0};

  if (tid >= NTHR) {
    // ---- producer warp: one lane streams every plane of every item of this CTA through the ring ----
    if (tid == NTHR) {
      uint32_t seq = 0;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        int tx, ty, zc;
    plan_item(P, item, tx, ty, zc);
        const int xo = tx * TX, yo = ty * TY, zb = zc * P.zlen;
        const int nplanes = min(D.nz, zb + P.zlen) - zb + 2;
        for (int k = 0; k < nplanes; k++, seq++) {
          const uint32_t st = seq % WS_NST, use = seq / WS_NST;
          if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
          unsigned char* dst = tiles + st * WS_STAGE;
          mbar_expect_tx(&full[st], STAGE_BYTES + WS_CODE_BYTES);
          // field tensor z index = local plane + 1 (one halo plane below plane 0); the code array has none
          tma_load_3d(dst, &tm, &full[st], xo - 2, yo - 1, zb + k);
          tma_load_3d(dst + STAGE_STRIDE, &tmc, &full[st], xo, yo, zb + k - 1);
        }
      }
    }
  } else {
    // ---- consumer warps ----
    const int lane = tid & 31, lx = lane * 2, ly = (tid >> 5) * 2;
    uint32_t seq = 0;  // planes consumed so far by this CTA (identical in every consumer thread)
    const long plane = (long)D.nx * D.ny;
    const int coff = (ly + 1) * BOXW + lx + 2;  // cell a of my 2x2 block inside a staged plane
    const int koff = STAGE_STRIDE + ly * TX + lx;  // its code byte
    const double sx0 = C.same[0][0], sx1 = C.same[0][1], sy0 = C.same[1][0], sy1 = C.same[1][1];
    const double sz0 = C.same[2][0], sz1 = C.same[2][1];
    const double dfx = C.diff[0], dfy = C.diff[1], dfz = C.diff[2];

    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      int tx, ty, zc;
    plan_item(P, item, tx, ty, zc);
      const int xo = tx * TX, yo = ty * TY, zb = zc * P.zlen;
      const int len = min(D.nz, zb + P.zlen) - zb, nplanes = len + 2;
      const int gx = xo + lx, gy = yo + ly;
      const bool colv = gx < D.nx;
      const bool row0 = colv && gy < D.ny, row1 = colv && gy + 1 < D.ny;
      const bool ax0 = gx == 0, bx1 = gx + 1 == D.nx - 1, yt = gy == 0, yb = gy + 1 == D.ny - 1;
      const double w0 = ax0 ? C.wall[0] : sx0, w1 = ax0 ? C.wall[1] : sx1;
      const double e0 = bx1 ? C.wall[0] : sx0, e1 = bx1 ? C.wall[1] : sx1;
      const double n0 = yt ? 0.0 : sy0, n1 = yt ? 0.0 : sy1;
      const double s0 = (yb || !row1) ? 0.0 : sy0, s1 = (yb || !row1) ? 0.0 : sy1;
      const double m0 = row1 ? sy0 : 0.0, m1 = row1 ? sy1 : 0.0;

      long ci = ((long)zb * D.ny + gy) * D.nx + gx;  // global index of cell a in plane zb
      double2 pt, pb, nt, nb;
      double ga = 0.0, gb = 0.0, gc = 0.0, gd = 0.0;  // flux into my cells from the plane below
      {
        const uint32_t q0 = seq % WS_NST, q1 = (seq + 1) % WS_NST;
        mbar_wait(&full[q0], (seq / WS_NST) & 1);
        const double* lo = reinterpret_cast<const double*>(tiles + q0 * WS_STAGE) + coff;
        const double2 mt = *reinterpret_cast<const double2*>(lo), mb = *reinterpret_cast<const double2*>(lo + BOXW);
        mbar_wait(&full[q1], ((seq + 1) / WS_NST) & 1);
        const unsigned char* s1p = tiles + q1 * WS_STAGE;
        const double* c0 = reinterpret_cast<const double*>(s1p) + coff;
        pt = *reinterpret_cast<const double2*>(c0);
        pb = *reinterpret_cast<const double2*>(c0 + BOXW);
        if (D.z0 + zb != 0) {  // face between plane zb-1 and zb
          const unsigned cn0 = *reinterpret_cast<const unsigned short*>(s1p + koff);
          const unsigned cn1 = *reinterpret_cast<const unsigned short*>(s1p + koff + TX);
          const unsigned ca = cn0 & 0xffu, cb = cn0 >> 8, cc = cn1 & 0xffu, cd = cn1 >> 8;
          ga = pick(ca & C_F, dfz, pick(ca & C_PH, sz1, sz0)) * (mt.x - pt.x);
          gb = pick(cb & C_F, dfz, pick(cb & C_PH, sz1, sz0)) * (mt.y - pt.y);
          gc = pick(cc & C_F, dfz, pick(cc & C_PH, sz1, sz0)) * (mb.x - pb.x);
          gd = pick(cd & C_F, dfz, pick(cd & C_PH, sz1, sz0)) * (mb.y - pb.y);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[q0]);  // the plane below the chunk is not read again
      }

      for (int k = 1; k <= len; k++) {
        const int lz = zb + k - 1;
        const uint32_t sc = (seq + k) % WS_NST, sn = (seq + k + 1) % WS_NST;
        mbar_wait(&full[sn], ((seq + k + 1) / WS_NST) & 1);
        const unsigned char* scp = tiles + sc * WS_STAGE;
        const double* cur = reinterpret_cast<const double*>(scp) + coff;
        const double* nxt = reinterpret_cast<const double*>(tiles + sn * WS_STAGE) + coff;
        nt = *reinterpret_cast<const double2*>(nxt);
        nb = *reinterpret_cast<const double2*>(nxt + BOXW);
        const unsigned k0 = *reinterpret_cast<const unsigned short*>(scp + koff);
        const unsigned k1 = *reinterpret_cast<const unsigned short*>(scp + koff + TX);
        const double2 pn = *reinterpret_cast<const double2*>(cur - BOXW);
        const double2 ps = *reinterpret_cast<const double2*>(cur + 2 * BOXW);
        const double pwt = cur[-1], pet = cur[2], pwb = cur[BOXW - 1], peb = cur[BOXW + 2];
        const unsigned ca = k0 & 0xffu, cb = k0 >> 8, cc = k1 & 0xffu, cd = k1 >> 8;
        const bool fa = ca & C_PH, fb = cb & C_PH, fc = cc & C_PH, fd = cd & C_PH;
        double qa = -ga, qb = -gb, qc = -gc, qd = -gd;
        double g;
        qa += pick(ca & C_W, dfx, pick(fa, w1, w0)) * (pt.x - pwt);
        g = pick(ca & C_E, dfx, pick(fa, sx1, sx0)) * (pt.x - pt.y);
        qa += g;
        qb -= g;
        qb += pick(cb & C_E, dfx, pick(fb, e1, e0)) * (pt.y - pet);
        qc += pick(cc & C_W, dfx, pick(fc, w1, w0)) * (pb.x - pwb);
        g = pick(cc & C_E, dfx, pick(fc, sx1, sx0)) * (pb.x - pb.y);
        qc += g;
        qd -= g;
        qd += pick(cd & C_E, dfx, pick(fd, e1, e0)) * (pb.y - peb);
        qa += pick(ca & C_N, dfy, pick(fa, n1, n0)) * (pt.x - pn.x);
        g = pick(ca & C_S, dfy, pick(fa, m1, m0)) * (pt.x - pb.x);
        qa += g;
        qc -= g;
        qc += pick(cc & C_S, dfy, pick(fc, s1, s0)) * (pb.x - ps.x);
        qb += pick(cb & C_N, dfy, pick(fb, n1, n0)) * (pt.y - pn.y);
        g = pick(cb & C_S, dfy, pick(fb, m1, m0)) * (pt.y - pb.y);
        qb += g;
        qd -= g;
        qd += pick(cd & C_S, dfy, pick(fd, s1, s0)) * (pb.y - ps.y);
        if (D.z0 + lz != D.nzg - 1) {
          ga = pick(ca & C_B, dfz, pick(fa, sz1, sz0)) * (pt.x - nt.x);
          gb = pick(cb & C_B, dfz, pick(fb, sz1, sz0)) * (pt.y - nt.y);
          gc = pick(cc & C_B, dfz, pick(fc, sz1, sz0)) * (pb.x - nb.x);
          gd = pick(cd & C_B, dfz, pick(fd, sz1, sz0)) * (pb.y - nb.y);
          qa += ga;
          qb += gb;
          qc += gc;
          qd += gd;
        }
        // every shared-memory value of plane k is in registers now: hand its stage back to the producer
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&empty[sc]);
          if (k == len) mbar_arrive(&empty[sn]);  // the plane above the chunk is not read again either
        }
        if (MODE == MODE_APPLY) {
          if (row0) {
            *reinterpret_cast<double2*>(o0 + ci) = make_double2(qa, qb);
            acc[0] += pt.x * qa + pt.y * qb;
          }
          if (row1) {
            *reinterpret_cast<double2*>(o0 + ci + D.nx) = make_double2(qc, qd);
            acc[0] += pb.x * qc + pb.y * qd;
          }
        } else {
          const bool z0 = D.z0 + lz == 0, z1 = D.z0 + lz == D.nzg - 1;
          Faces Fa, Fb;
          double2 u, v;
          if (row0) {
            if (MODE == MODE_INIT) {
              Fa.diag = diag_of(C, ca, ax0, false, yt, gy == D.ny - 1, z0, z1);
              Fb.diag = diag_of(C, cb, false, bx1, yt, gy == D.ny - 1, z0, z1);
            }
            cell_epilogue<MODE>(C, ca, ax0, false, Fa, pt.x, qa, u.x, v.x, acc);
            cell_epilogue<MODE>(C, cb, false, bx1, Fb, pt.y, qb, u.y, v.y, acc);
            if (MODE != MODE_RESID) *reinterpret_cast<double2*>(o0 + ci) = u;
            if (MODE == MODE_INIT) *reinterpret_cast<double2*>(o1 + ci) = v;
          }
          if (row1) {
            if (MODE == MODE_INIT) {
              Fa.diag = diag_of(C, cc, ax0, false, false, yb, z0, z1);
              Fb.diag = diag_of(C, cd, false, bx1, false, yb, z0, z1);
            }
            cell_epilogue<MODE>(C, cc, ax0, false, Fa, pb.x, qc, u.x, v.x, acc);
            cell_epilogue<MODE>(C, cd, false, bx1, Fb, pb.y, qd, u.y, v.y, acc);
            if (MODE != MODE_RESID) *reinterpret_cast<double2*>(o0 + ci + D.nx) = u;
            if (MODE == MODE_INIT) *reinterpret_cast<double2*>(o1 + ci + D.nx) = v;
          }
        }
        pt = nt;
        pb = nb;
        ci += plane;
      }
      seq += nplanes;
    }
  }
  block_sum<3>(acc, red);
  if (tid == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0];
    partials[blockIdx.x * 3 + 1] = acc[1];
    partials[blockIdx.x * 3 + 2] = acc[2];
  }
}

// ------------------------------------------------------------------------------------------------
// CG vector kernels.  Rows round-robin over a fixed grid; a_P is recomputed from the code byte.
//   k_cg_update : alpha = r.z / p.q;  x += alpha p;  r -= alpha q;  partials {r.(r/a_P), r.r}
//   k_cg_dir    : beta = r.z_new / r.z_old;  p = r/a_P + beta p
// ------------------------------------------------------------------------------------------------
template <int V>
struct VecD;
template <>
struct VecD<1> {
  double v[1];
  __device__ __forceinline__ void load(const double* p) { v[0] = *p; }
  __device__ __forceinline__ void store(double* p) const { *p = v[0]; }
};
template <>
struct VecD<2> {
  double v[2];
  __device__ __forceinline__ void load(const double* p) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x;
    v[1] = t.y;
  }
  __device__ __forceinline__ void store(double* p) const { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
};

// V cells per thread per step (V = 2 needs nx even: 16-byte aligned pairs)
template <int V>
__global__ void __launch_bounds__(256) k_cg_update(double* __restrict__ x, double* __restrict__ r,
                                                   const double* __restrict__ p, const double* __restrict__ q,
                                                   const uint8_t* __restrict__ code, const Coef C, const Dom D,
                                                   const Scal* __restrict__ S, int par,
                                                   double* __restrict__ partials) {
  __shared__ double red[32 * 2];
  if (S->done) return;
  const double alpha = S->rz[par] / S->pq;
  const long rows = (long)D.ny * D.nz;
  double acc[2] = {0, 0};
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % D.ny), z = (int)(row / D.ny);
    const bool y0 = y == 0, y1 = y == D.ny - 1, z0 = D.z0 + z == 0, z1 = D.z0 + z == D.nzg - 1;
    const long base = row * D.nx;
    for (int xx = threadIdx.x * V; xx < D.nx; xx += blockDim.x * V) {
      const long i = base + xx;
      VecD<V> vx, vr, vp, vq;
      vx.load(x + i);
      vr.load(r + i);
      vp.load(p + i);
      vq.load(q + i);
      unsigned cc = V == 2 ? *reinterpret_cast<const unsigned short*>(code + i) : code[i];
#pragma unroll
      for (int j = 0; j < V; j++) {
        const double d = diag_of(C, (cc >> (8 * j)) & 0xffu, xx + j == 0, xx + j == D.nx - 1, y0, y1, z0, z1);
        vx.v[j] += alpha * vp.v[j];
        const double ri = vr.v[j] - alpha * vq.v[j];
        vr.v[j] = ri;
        acc[0] += ri * (ri / d);
        acc[1] += ri * ri;
      }
      vx.store(x + i);
      vr.store(r + i);
    }
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2 + 0] = acc[0];
    partials[blockIdx.x * 2 + 1] = acc[1];
  }
}

template <int V>
__global__ void __launch_bounds__(256) k_cg_dir(double* __restrict__ p, const double* __restrict__ r,
                                                const uint8_t* __restrict__ code, const Coef C, const Dom D,
                                                const Scal* __restrict__ S, int par) {
  if (S->done) return;
  const double beta = S->rz[par ^ 1] / S->rz[par];
  const long rows = (long)D.ny * D.nz;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % D.ny), z = (int)(row / D.ny);
    const bool y0 = y == 0, y1 = y == D.ny - 1, z0 = D.z0 + z == 0, z1 = D.z0 + z == D.nzg - 1;
    const long base = row * D.nx;
    for (int xx = threadIdx.x * V; xx < D.nx; xx += blockDim.x * V) {
      const long i = base + xx;
      VecD<V> vr, vp;
      vr.load(r + i);
      vp.load(p + i);
      unsigned cc = V == 2 ? *reinterpret_cast<const unsigned short*>(code + i) : code[i];
#pragma unroll
      for (int j = 0; j < V; j++) {
        const double d = diag_of(C, (cc >> (8 * j)) & 0xffu, xx + j == 0, xx + j == D.nx - 1, y0, y1, z0, z1);
        vp.v[j] = vr.v[j] / d + beta * vp.v[j];
      }
      vp.store(p + i);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One-block reduction + all-reduce + scalar step.  Every dot product of the iteration ends here: the
// per-CTA partials of the producing kernel are summed in a fixed order (bit-reproducible), the sum is
// completed across the slab ranks THROUGH PEER MEMORY -- every rank stores its contribution into every
// other rank's mailbox (NVLink stores), waits for the others' to land in its own, and adds them in rank
// order, so all ranks hold bit-identical scalars and take identical branches -- and the scalar
// recurrence of the step (iteration count, stop flag) is applied, all in one launch.
//   phase 0: everything (single GPU, or peer memory)
//   phase 1: the local sum only, into S->tmp (an NCCL all-reduce of S->tmp follows on the stream)
//   phase 2: the scalar step from S->tmp
// ------------------------------------------------------------------------------------------------
enum { RS_PQ = 0, RS_RR_MG = 1, RS_RZ_MG = 2, RS_INIT = 3, RS_ITER = 4, RS_RESID = 5, RS_FLUX = 6 };
// what          partials (stride: slots)      step
// RS_PQ         3: 0                          S->pq
// RS_RR_MG      2: 1                          S->rr, ++S->it, stop flag              (multigrid path)
// RS_RZ_MG      3: 0                          S->rz[arg]
// RS_INIT       3: 0,1,2 = r.z, r.r, b.b      S->rz[0], S->rr, S->bb, it = 0, stop flag
// RS_ITER       2: 0,1   = r.z, r.r           S->rz[arg ^ 1], S->rr, ++S->it, stop flag   (diagonal path)
// RS_RESID      3: 1,2   = r.r, sum|r|        S->tmp[1], S->tmp[2]
// RS_FLUX       2: 0,1   = QL, QR             S->ql, S->qr        (runs once the stop flag is up, or arg = 1)

// complete v[0..3] (valid in thread 0 on entry, in every thread on return) over the ranks of the link
__device__ __forceinline__ void ar_sum(double (&v)[4], const ArLink& ar, Scal* S) {
  __shared__ unsigned long long s_seq;
  __shared__ double s_in[KB_MAX_RANKS][4];
  __shared__ double s_out[4];
  const int t = threadIdx.x;
  if (t == 0) {
    s_seq = ++(*ar.seq);
#pragma unroll
    for (int k = 0; k < 4; k++) s_out[k] = v[k];
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  if (t < ar.world) {
    MailSlot* dst = &ar.mail[t]->ar[seq & 1][ar.rank];
#pragma unroll
    for (int k = 0; k < 4; k++) reinterpret_cast<volatile double*>(dst->v)[k] = s_out[k];
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    const MailSlot* src = &ar.mail[ar.rank]->ar[seq & 1][t];
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while (ld_acquire_sys(&src->seq) != seq) {
      if (globaltimer_ns() - t0 > KB_WAIT_NS) {
        ok = false;
        break;
      }
    }
    if (!ok) S->fault = 1;
#pragma unroll
    for (int k = 0; k < 4; k++) s_in[t][k] = ok ? reinterpret_cast<const volatile double*>(src->v)[k] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double a = 0.0;
    for (int r = 0; r < ar.world; r++) a += s_in[r][k];
    v[k] = a;
  }
}

template <int WHAT>
__global__ void __launch_bounds__(512) k_reduce_step(const double* __restrict__ partials, int n, Scal* S, int arg, int phase,
                                                     const ArLink ar) {
  __shared__ double red[32 * 3];
  const bool iter_step = WHAT == RS_PQ || WHAT == RS_RR_MG || WHAT == RS_RZ_MG || WHAT == RS_ITER;
  if (iter_step && S->done) return;
  if (WHAT == RS_FLUX && !S->done && !arg) return;
  constexpr int stride = (WHAT == RS_RR_MG || WHAT == RS_ITER || WHAT == RS_FLUX) ? 2 : 3;
  constexpr int s0 = (WHAT == RS_RR_MG || WHAT == RS_RESID) ? 1 : 0;                     // first slot used
  constexpr int nv = WHAT == RS_INIT ? 3 : (WHAT == RS_ITER || WHAT == RS_RESID || WHAT == RS_FLUX) ? 2 : 1;
  double v[4] = {0, 0, 0, 0};
  if (phase != 2) {
    double acc[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
      for (int k = 0; k < nv; k++) acc[k] += partials[i * stride + s0 + k];
    }
    block_sum<3>(acc, red);
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = acc[k];
    if (phase == 1) {
      if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < nv; k++) S->tmp[k] = v[k];
      return;
    }
    if (ar.world > 1) ar_sum(v, ar, S);
  } else {
#pragma unroll
    for (int k = 0; k < nv; k++) v[k] = S->tmp[k];
  }
  if (threadIdx.x != 0) return;
  if (WHAT == RS_PQ) S->pq = v[0];
  if (WHAT == RS_RZ_MG) S->rz[arg] = v[0];
  if (WHAT == RS_RR_MG) {
    S->rr = v[0];
    S->it += 1;
    if (!(v[0] > S->thr2 * S->bb)) S->done = 1;
  }
  if (WHAT == RS_INIT) {
    S->rz[0] = v[0];
    S->rr = v[1];
    S->bb = v[2];
    S->it = 0;
    S->done = !(v[1] > S->thr2 * v[2]);
  }
  if (WHAT == RS_ITER) {
    S->rz[arg ^ 1] = v[0];
    S->rr = v[1];
    S->it += 1;
    if (!(v[1] > S->thr2 * S->bb)) S->done = 1;
  }
  if (WHAT == RS_RESID) {
    S->tmp[1] = v[0];
    S->tmp[2] = v[1];
  }
  if (WHAT == RS_FLUX) {
    S->ql = v[0];
    S->qr = v[1];
  }
  if (S->fault) S->done = 1;  // a rank went silent: let every queued kernel drain, the host reports it
}

// ------------------------------------------------------------------------------------------------
// boundary-flux integration: QL = sum c_wall (T(x=0) - TL), QR = sum c_wall (TR - T(x=nx-1))
// (Keff2D/keff2D.cpp:151-170, Keff3D/keff3D.cpp:145-163).  One thread per (z,y) row.
// ------------------------------------------------------------------------------------------------
// S != nullptr: in-loop use -- the sums are only needed once the stop flag is up (the flux balance is part of the
// stop rule), so the kernel returns at once while the iteration is still running.
__global__ void __launch_bounds__(256) k_flux(const double* __restrict__ T, const uint8_t* __restrict__ code,
                                              const Coef C, const Dom D, double* __restrict__ qL,
                                              double* __restrict__ qR, double* __restrict__ partials,
                                              const Scal* __restrict__ S) {
  __shared__ double red[32 * 2];
  if (S && !S->done) return;
  const long rows = (long)D.ny * D.nz;
  double acc[2] = {0, 0};
  for (long row = blockIdx.x * (long)blockDim.x + threadIdx.x; row < rows; row += (long)gridDim.x * blockDim.x) {
    const long l = row * D.nx, e = l + D.nx - 1;
    const double a = ((code[l] & C_PH) ? C.wall[1] : C.wall[0]) * (T[l] - C.tl);
    const double b = ((code[e] & C_PH) ? C.wall[1] : C.wall[0]) * (C.tr - T[e]);
    if (qL) qL[row] = a;
    if (qR) qR[row] = b;
    acc[0] += a;
    acc[1] += b;
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2 + 0] = acc[0];
    partials[blockIdx.x * 2 + 1] = acc[1];
  }
}

}  // namespace kb
