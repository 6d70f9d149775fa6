// mg.cuh -- geometric multigrid V-cycle used as the CG preconditioner of the grid solver.
//
// No reference counterpart: the reference relaxes the assembled system with lexicographic
// Gauss-Seidel (ParallelGS Keff2D/keff2D.h:758-850, ParallelGS3D Keff3D/keff3D.h:492-595), O(N^2)
// sweeps.  The linear system solved here is still exactly the reference's (DiscretizeMatrixCD2D/3D);
// the V-cycle only changes how fast CG gets to its solution (iteration counts independent of N).
//
// Hierarchy: level 0 is the voxel grid (operator evaluated from the code bytes).  Level l+1 aggregates
// 2 cells per coarsened direction of level l (a lone last cell for odd sizes).  Transfer operators are
// piecewise constant: restriction = sum over the aggregate, prolongation = copy to the aggregate's
// cells.  A coarse face conductance is the sum of the fine face conductances crossing that coarse
// face, halved along coarsened directions (the Galerkin product P^T A P of piecewise-constant
// transfers overestimates the coarse-grid stiffness by 2; halving restores the re-discretised operator
// on a uniform medium).  Coarse levels are stored: FP32 east/south/back face conductances, wall
// conductance, 1/a_P, and the vectors of the cycle.  The whole cycle is FP32.
//
// Smoother: nu weighted-Jacobi sweeps with weights w_1..w_nu before the coarse correction and the
// same weights in reverse order after it, so the cycle is a symmetric positive definite operator (what
// CG needs); the weights are Chebyshev points for the upper part of the spectrum of D^-1 A.
// Every level keeps its right-hand side in the scaled form u = D^-1 b.  One Jacobi sweep for A x = b is
//     x' = (1-w) x + w u + w D^-1 sum_f c_f x_nb                                   ("sweep" pass)
// and the residual of an iterate is  b - A x = D (u - x) + sum_f c_f x_nb          ("restrict" pass),
// so every pass is one 7-point stencil over one field; the first iterate x_1 = w_1 u is never stored
// (passes take a scale factor for their input), and the prolongated correction P e is never
// materialised either (the first sweep after the coarse solve adds e of the right aggregate to each
// value it reads).  The coarsest level (<= 64 cells) is solved by a dense inverse.
#pragma once
#include "kernels.cuh"

namespace kb {

constexpr int MG_MAX_LEVELS = 24;
constexpr int MG_DENSE_MAX = 64;
constexpr long MG_REPLICATE_CELLS = 1L << 18;  // z-slabs: levels of at most this many cells are replicated

// one stored (coarse) level; every array has one spare plane below plane 0 and one above plane nz-1
struct MgLevel {
  int nx, ny, nz;      // extent of this level
  int fx, fy, fz;      // 1 if the direction is coarsened to build the NEXT level
  float *cx, *cy, *cz; // conductance of the face towards x+1 / y+1 / z+1 (0 at the domain end)
  float* wsum;         // fixed-temperature wall conductance of the cell (non-zero in the two wall columns)
  float* dinv;         // 1 / a_P
  float *u, *e;        // scaled right-hand side D^-1 b (the dense level keeps b itself), correction
  float *t0, *t1;      // iterates of the smoothing sweeps (nu >= 2)
};

// coefficients of one pass:  in = sc * (stored value) [+ e of its aggregate];
//   sweep   : out = a in_P + b u_P + g D^-1_P sum_f c_f in_nb
//   restrict: res = D_P (a in_P + b u_P) + g sum_f c_f in_nb
struct MgPass {
  float sc, a, b, g;
};

struct MgLevelView {
  int nx, ny, nz;
  const float *cx, *cy, *cz, *dinv;
};

// ------------------------------------------------------------------------------------------------
// hierarchy construction
// ------------------------------------------------------------------------------------------------
// level 1 from the code bytes of the voxel grid (fine dims in D, global z end decided from D.nzg)
__global__ void __launch_bounds__(256) k_mg_coarsen_fine(const uint8_t* __restrict__ code, const Coef C, const Dom D, int fx,
                                                         int fy, int fz, MgLevel L, int bx) {
  const float sx = fx ? 0.5f : 1.0f, sy = fy ? 0.5f : 1.0f, sz = fz ? 0.5f : 1.0f;
  const int tx = threadIdx.x & (bx - 1), trow = threadIdx.x / bx, rpb = blockDim.x / bx;  // bx: power of two
  for (int Z = blockIdx.y; Z < L.nz; Z += gridDim.y)
  for (int Y = blockIdx.x * rpb + trow; Y < L.ny; Y += gridDim.x * rpb)
  for (int X = tx; X < L.nx; X += bx) {
    const long i = ((long)Z * L.ny + Y) * L.nx + X;
    double ax = 0, ay = 0, az = 0, aw = 0;
    for (int dz = 0; dz <= fz; dz++) {
      const int z = fz ? 2 * Z + dz : Z;
      if (z >= D.nz) break;
      for (int dy = 0; dy <= fy; dy++) {
        const int y = fy ? 2 * Y + dy : Y;
        if (y >= D.ny) break;
        for (int dx = 0; dx <= fx; dx++) {
          const int x = fx ? 2 * X + dx : X;
          if (x >= D.nx) break;
          const unsigned c = code[((long)z * D.ny + y) * D.nx + x];
          const int ph = c & C_PH;
          // a face crosses to the next aggregate when this is the aggregate's last cell in that direction
          if (x < D.nx - 1 && (!fx || dx == 1)) ax += (c & C_E) ? C.diff[0] : C.same[0][ph];
          if (y < D.ny - 1 && (!fy || dy == 1)) ay += (c & C_S) ? C.diff[1] : C.same[1][ph];
          if (D.z0 + z < D.nzg - 1 && (!fz || dz == 1)) az += (c & C_B) ? C.diff[2] : C.same[2][ph];
          if (x == 0) aw += C.wall[ph];
          if (x == D.nx - 1) aw += C.wall[ph];
        }
      }
    }
    L.cx[i] = sx * (float)ax;
    L.cy[i] = sy * (float)ay;
    L.cz[i] = sz * (float)az;
    L.wsum[i] = sx * (float)aw;
  }
}

// level l+1 from stored level l
__global__ void k_mg_coarsen_level(const MgLevel F, MgLevel L) {
  const long n1 = (long)L.nx * L.ny * L.nz;
  const int fx = F.fx, fy = F.fy, fz = F.fz;
  const float sx = fx ? 0.5f : 1.0f, sy = fy ? 0.5f : 1.0f, sz = fz ? 0.5f : 1.0f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n1; i += (long)gridDim.x * blockDim.x) {
    const int X = (int)(i % L.nx);
    const long t = i / L.nx;
    const int Y = (int)(t % L.ny), Z = (int)(t / L.ny);
    float ax = 0, ay = 0, az = 0, aw = 0;
    for (int dz = 0; dz <= fz; dz++) {
      const int z = fz ? 2 * Z + dz : Z;
      if (z >= F.nz) break;
      for (int dy = 0; dy <= fy; dy++) {
        const int y = fy ? 2 * Y + dy : Y;
        if (y >= F.ny) break;
        for (int dx = 0; dx <= fx; dx++) {
          const int x = fx ? 2 * X + dx : X;
          if (x >= F.nx) break;
          const long j = ((long)z * F.ny + y) * F.nx + x;
          if (!fx || dx == 1) ax += F.cx[j];  // zero at the domain end by construction
          if (!fy || dy == 1) ay += F.cy[j];
          if (!fz || dz == 1) az += F.cz[j];
          aw += F.wsum[j];
        }
      }
    }
    L.cx[i] = sx * ax;
    L.cy[i] = sy * ay;
    L.cz[i] = sz * az;
    L.wsum[i] = sx * aw;
  }
}

// 1/a_P of a stored level.  lo: plane -1 of cz holds the face towards the lower slab neighbour.
__global__ void k_mg_dinv(MgLevel L, int lo) {
  const long plane = (long)L.nx * L.ny, n = plane * L.nz;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % L.nx);
    const long t = i / L.nx;
    const int y = (int)(t % L.ny), z = (int)(t / L.ny);
    float d = L.cx[i] + L.cy[i] + L.cz[i] + L.wsum[i];
    if (x > 0) d += L.cx[i - 1];
    if (y > 0) d += L.cy[i - L.nx];
    if (z > 0 || lo) d += L.cz[i - plane];
    L.dinv[i] = 1.0f / d;
  }
}

// ------------------------------------------------------------------------------------------------
// stored-level smoothing passes (FP32, neighbours through L1/L2; these levels hold <= 1/8 of the cells)
// ------------------------------------------------------------------------------------------------
// lo/hi: the level's z-halo planes (b, dinv, e of the slab neighbours) are valid

// Thread layout of the stored-level passes: a block is bx (32..256, covering a row of the output
// when it fits) x 256/bx threads; blockIdx.x strides over y, blockIdx.y over z -- 32-bit index arithmetic,
// no division.  lo/hi: the z-halo planes of the input arrays hold the slab neighbours' values.
// restrict pass: un[aggregate] = D^-1_next sum over the aggregate of  D_P (a in_P + b u_P) + g sum_f c_f in_nb
// (dn == nullptr: the sum itself, for the dense level)
__global__ void __launch_bounds__(256) k_mg_restrict(const MgLevelView L, const float* __restrict__ in,
                                                     const float* __restrict__ u, const MgPass K, int fx, int fy, int fz,
                                                     int cnx, int cny, int cnz, float* __restrict__ un,
                                                     const float* __restrict__ dn, int lo, int hi, int bx,
                                                     const Scal* __restrict__ S) {
  if (S->done) return;
  const int plane = L.nx * L.ny;  // stored levels hold < 2^31 cells
  const int tx = threadIdx.x & (bx - 1), trow = threadIdx.x / bx, rpb = blockDim.x / bx;  // bx: power of two
  const float as = K.a * K.sc, gs = K.g * K.sc;
  for (int Z = blockIdx.y; Z < cnz; Z += gridDim.y)
  for (int Y = blockIdx.x * rpb + trow; Y < cny; Y += gridDim.x * rpb) {
    const int crow = Z * cny + Y;
    for (int X = tx; X < cnx; X += bx) {
      float acc = 0.f;
      for (int dz = 0; dz <= fz; dz++) {
        const int z = fz ? 2 * Z + dz : Z;
        if (z >= L.nz) break;
        for (int dy = 0; dy <= fy; dy++) {
          const int y = fy ? 2 * Y + dy : Y;
          if (y >= L.ny) break;
          for (int dx = 0; dx <= fx; dx++) {
            const int x = fx ? 2 * X + dx : X;
            if (x >= L.nx) break;
            const int i = (z * L.ny + y) * L.nx + x;
            // no boundary branches: the conductance of a face that leaves the domain is zero by construction (so is
            // cz[plane -1] without a lower neighbour, from the pool memset) and the value it multiplies is finite
            // (spare planes of every level array are zeroed at setup), so all 13 loads go out together
            const float q = ((L.cx[i - 1] * in[i - 1] + L.cx[i] * in[i + 1]) + (L.cy[i - L.nx] * in[i - L.nx] + L.cy[i] * in[i + L.nx])) +
                            (L.cz[i - plane] * in[i - plane] + L.cz[i] * in[i + plane]);
            acc += __fdividef(as * in[i] + K.b * u[i], L.dinv[i]) + gs * q;
          }
        }
      }
      const int ic = crow * cnx + X;
      un[ic] = dn ? acc * dn[ic] : acc;
    }
  }
}

// sweep pass: out = a in_P + b u_P + g D^-1_P sum_f c_f in_nb,  in = sc * stored [+ ec of the aggregate]
template <int ADD_E>
__global__ void __launch_bounds__(256) k_mg_sweep(const MgLevelView L, const float* __restrict__ in,
                                                  const float* __restrict__ u, const float* __restrict__ ec,
                                                  const MgPass K, int fx, int fy, int fz, int cnx, int cny,
                                                  float* __restrict__ out, int lo, int hi, int bx,
                                                  const Scal* __restrict__ S) {
  if (S->done) return;
  const int plane = L.nx * L.ny, cplane = cnx * cny;
  const int tx = threadIdx.x & (bx - 1), trow = threadIdx.x / bx, rpb = blockDim.x / bx;  // bx: power of two
  const float sc = K.sc;
  for (int z = blockIdx.y; z < L.nz; z += gridDim.y)
  for (int y = blockIdx.x * rpb + trow; y < L.ny; y += gridDim.x * rpb) {
    const int row = z * L.ny + y;
    const int Y = fy ? y >> 1 : y, Z = fz ? z >> 1 : z;
    const float *e0 = nullptr, *eN = nullptr, *eS = nullptr, *eF = nullptr, *eB = nullptr;
    if (ADD_E) {
      e0 = ec + Z * cplane + Y * cnx;                                 // my aggregate row
      eN = ec + Z * cplane + (fy ? (y - 1) >> 1 : y - 1) * cnx;       // of the cell north / south
      eS = ec + Z * cplane + (fy ? (y + 1) >> 1 : y + 1) * cnx;
      eF = ec + (fz ? (z - 1) >> 1 : z - 1) * cplane + Y * cnx;       // of the cell in front / behind
      eB = ec + (fz ? (z + 1) >> 1 : z + 1) * cplane + Y * cnx;
    }
    for (int x = tx; x < L.nx; x += bx) {
      const int i = row * L.nx + x;
      const int X = fx ? x >> 1 : x;
      // no boundary branches (see k_mg_restrict): faces that leave the domain carry a zero conductance and meet a
      // finite value (previous / next row, spare planes), so every load of a cell is issued at once
      if (ADD_E) {
        const float q = ((L.cx[i - 1] * (sc * in[i - 1] + e0[fx ? (x - 1) >> 1 : x - 1]) +
                          L.cx[i] * (sc * in[i + 1] + e0[fx ? (x + 1) >> 1 : x + 1])) +
                         (L.cy[i - L.nx] * (sc * in[i - L.nx] + eN[X]) + L.cy[i] * (sc * in[i + L.nx] + eS[X]))) +
                        (L.cz[i - plane] * (sc * in[i - plane] + eF[X]) + L.cz[i] * (sc * in[i + plane] + eB[X]));
        out[i] = K.a * (sc * in[i] + e0[X]) + K.b * u[i] + K.g * L.dinv[i] * q;
      } else {
        const float q = ((L.cx[i - 1] * in[i - 1] + L.cx[i] * in[i + 1]) + (L.cy[i - L.nx] * in[i - L.nx] + L.cy[i] * in[i + L.nx])) +
                        (L.cz[i - plane] * in[i - plane] + L.cz[i] * in[i + plane]);
        out[i] = K.a * sc * in[i] + K.b * u[i] + K.g * sc * L.dinv[i] * q;
      }
    }
  }
}

// x += P e: the correction of the next level added to a stored iterate in place (planes z_first .. z_last: the
// z-halo planes of the iterate take the correction of the neighbour's aggregates from the halo planes of e).
// The sweep that follows is then a plain one: adding P e on the fly costs it seven more loads per cell (measured
// at 256^3: 281 us against 118 us for the plain sweep; this pass moves 8.5 B per cell and takes 68 us).
__global__ void __launch_bounds__(256) k_mg_prolong_add(int nx, int ny, int z_first, int z_last, float* __restrict__ x,
                                                        const float* __restrict__ ec, int fx, int fy, int fz, int cnx,
                                                        int cny, int bx, const Scal* __restrict__ S) {
  if (S->done) return;
  const int plane = nx * ny, cplane = cnx * cny;
  const int tx = threadIdx.x & (bx - 1), trow = threadIdx.x / bx, rpb = blockDim.x / bx;
  // four cells = two aggregates per thread when rows allow it (x coarsened, nx a multiple of 4: rows of both arrays
  // are then 16 / 8 byte aligned -- level arrays start on 128-byte boundaries and their planes are multiples of 4)
  const bool vec = fx && (nx & 3) == 0 && (cnx & 1) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(ec) & 7)) == 0;
  for (int z = z_first + (int)blockIdx.y; z <= z_last; z += gridDim.y)
  for (int y = blockIdx.x * rpb + trow; y < ny; y += gridDim.x * rpb) {
    const float* er = ec + (fz ? z >> 1 : z) * cplane + (fy ? y >> 1 : y) * cnx;
    float* xr = x + z * plane + y * nx;
    if (vec) {
      float4* xr4 = reinterpret_cast<float4*>(xr);
      const float2* er2 = reinterpret_cast<const float2*>(er);
      for (int q = tx; q < (nx >> 2); q += bx) {
        float4 v = xr4[q];
        const float2 e = er2[q];
        v.x += e.x;
        v.y += e.x;
        v.z += e.y;
        v.w += e.y;
        xr4[q] = v;
      }
    } else {
      for (int xx = tx; xx < nx; xx += bx) xr[xx] += er[fx ? xx >> 1 : xx];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// coarsest level: dense inverse (n <= MG_DENSE_MAX), built and inverted in shared memory by one CTA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MG_DENSE_MAX) k_mg_dense_setup(const MgLevel L, double* __restrict__ inv) {
  __shared__ double A[MG_DENSE_MAX][MG_DENSE_MAX + 1];
  __shared__ double col[MG_DENSE_MAX];
  const int n = L.nx * L.ny * L.nz, i = threadIdx.x;
  const int plane = L.nx * L.ny;
  if (i < n) {
    for (int j = 0; j < n; j++) A[i][j] = 0.0;
    const int x = i % L.nx, y = (i / L.nx) % L.ny, z = i / plane;
    double d = L.wsum[i];
    if (x < L.nx - 1) { A[i][i + 1] = -(double)L.cx[i]; d += L.cx[i]; }
    if (x > 0) { A[i][i - 1] = -(double)L.cx[i - 1]; d += L.cx[i - 1]; }
    if (y < L.ny - 1) { A[i][i + L.nx] = -(double)L.cy[i]; d += L.cy[i]; }
    if (y > 0) { A[i][i - L.nx] = -(double)L.cy[i - L.nx]; d += L.cy[i - L.nx]; }
    if (z < L.nz - 1) { A[i][i + plane] = -(double)L.cz[i]; d += L.cz[i]; }
    if (z > 0) { A[i][i - plane] = -(double)L.cz[i - plane]; d += L.cz[i - plane]; }
    A[i][i] = d;
  }
  __syncthreads();
  // in-place Gauss-Jordan inversion without pivoting (symmetric positive definite M-matrix);
  // thread i owns row i
  for (int k = 0; k < n; k++) {
    if (i < n) col[i] = A[i][k];
    __syncthreads();
    const double pinv = 1.0 / col[k];
    if (i == k) {
      for (int j = 0; j < n; j++) A[k][j] *= pinv;
      A[k][k] = pinv;
    }
    __syncthreads();
    if (i < n && i != k) {
      const double f = col[i];
      for (int j = 0; j < n; j++)
        if (j != k) A[i][j] -= f * A[k][j];
      A[i][k] = -f * pinv;
    }
    __syncthreads();
  }
  if (i < n)
    for (int j = 0; j < n; j++) inv[i * n + j] = A[i][j];
}

__global__ void __launch_bounds__(MG_DENSE_MAX) k_mg_dense_apply(const double* __restrict__ inv, const float* __restrict__ b,
                                                                 float* __restrict__ e, int n,
                                                                 const Scal* __restrict__ S) {
  __shared__ double sb[MG_DENSE_MAX];
  if (S->done) return;
  const int i = threadIdx.x;
  if (i < n) sb[i] = b[i];
  __syncthreads();
  if (i < n) {
    double a = 0;
    for (int j = 0; j < n; j++) a += inv[i * n + j] * sb[j];
    e[i] = (float)a;
  }
}

// ------------------------------------------------------------------------------------------------
// voxel-grid (level 0) passes.  The CG update kernel leaves
//     u = r / d     (FP32, pitched rows, one halo plane below and above;  d = a_P rounded to FP32)
// next to the FP64 residual; the cycle is the exact V-cycle of the operator whose diagonal is d
// (r_P is recovered as d u_P where needed) -- symmetric, as CG needs.  The passes are 7-point stencils
// with the same tiling as k_stencil_tma: a CTA marches a 64x16 xy tile in z, planes of the input field
// (u or an iterate, both pitched FP32 with halo planes) arrive by TMA in a shared-memory ring,
// zero-filled outside the domain, so a face towards the outside multiplies a zero and needs no
// boundary logic.  1/d and d come from a table indexed by position class and code byte.
//   OUT 0: restrict pass -> u of level 1         OUT 1: sweep pass -> pitched iterate
//   OUT 2: last sweep -> z (FP32, unpitched) and the partial r.z
//   ADD_E: the input is  sc * stored + e1 of the cell's aggregate.  P e1 is never materialised: the
//          2x2 cell block of a thread is one aggregate in xy, so the thread adds its own aggregate's
//          correction and those of the four in-plane neighbour aggregates (and of the aggregate
//          below/above) to the staged values on the fly.
//   IN_IS_U: the staged field is u itself (otherwise my cells' u is read from global memory).
// ------------------------------------------------------------------------------------------------
// the staged box starts 4 cells left of the tile: a TMA box must start on a 16-byte boundary of the row
constexpr int MG_NST = 6, MG_BOXW = TX + 8, MG_HX = 4;
constexpr int MG_STAGE_BYTES = MG_BOXW * BOXH * 4;
constexpr int MG_STAGE_STRIDE = (MG_STAGE_BYTES + 127) / 128 * 128;

// face conductances of the voxel grid rounded to FP32 (what the FP32 cycle works with)
struct MgCoefF {
  float same[3][2], diff[3];
};

struct MgFine {
  const float* e1;     // level-1 correction (ADD_E)
  float* u1;           // level-1 scaled right-hand side (OUT 0)
  const float* dinv1;  // level-1 1/a_P (nullptr: level 1 is the dense level, store the plain sum)
  int nx1, ny1, nz1;
  int fy, fz;          // y / z coarsened (x always is: nx even >= 4)
  int lo, hi;          // z-slabs: halo planes of the fields / e1 hold the lower / upper neighbour's values
  int pitch;           // floats per row of the pitched fields
};

// position class of a cell along one axis: 0 interior, 1 first, 2 last
__host__ __device__ __forceinline__ int mg_cls(int i, int n) { return i == 0 ? 1 : (i == n - 1 ? 2 : 0); }

// a_P of a cell of position class cls = xcls + 3 ycls + 9 zcls with code byte c, rounded to FP32
__device__ __forceinline__ float mg_diag32(const Coef& C, int cls, unsigned c, bool flat) {
  const int cx = cls % 3, cy = (cls / 3) % 3, cz = cls / 9;
  return (float)diag_of(C, c, cx == 1, cx == 2, cy == 1, cy == 2, flat || cz == 1, flat || cz == 2);
}

__device__ __forceinline__ float pickf(bool c, float a, float b) { return c ? a : b; }

// the two per-solve tables over (position class, code byte): [0] 1/d, [1] d
constexpr int MG_TAB = 27 * 128;
__global__ void k_mg_tables(float* __restrict__ t, const Coef C, int flat) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= MG_TAB) return;
  const float d = mg_diag32(C, e >> 7, (unsigned)(e & 127), flat != 0);
  t[e] = 1.f / d;
  t[MG_TAB + e] = d;
}

template <int OUT, int ADD_E, int IN_IS_U>
__global__ void __launch_bounds__(NTHR, 3) k_mg_fine(const __grid_constant__ CUtensorMap tm,
                                                     const uint8_t* __restrict__ code, const float* __restrict__ gtab,
                                                     const float* __restrict__ uown, float* __restrict__ fout,
                                                     float* __restrict__ zout, const MgFine M, const MgPass K,
                                                     const MgCoefF C, const Dom D, const Plan P,
                                                     const Scal* __restrict__ S, double* __restrict__ partials,
                                                     const HaloPush H) {
  __shared__ __align__(128) unsigned char tiles[MG_NST * MG_STAGE_STRIDE];
  __shared__ __align__(8) uint64_t full[MG_NST];
  __shared__ float tab[27 * 128];  // OUT 0: d    otherwise: 1/d
  __shared__ double red[32 * 3];
  const bool push = OUT == 1 && (H.lo_dst || H.hi_dst);
  if (S->done) {
    if (push) {  // the neighbours count arrivals whether or not the iteration is still running
      unsigned int nlo = 0, nhi = 0;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        int tx, ty, zc;
        plan_item(P, item, tx, ty, zc);
        nlo += zc == 0;
        nhi += zc == P.nzc - 1;
      }
      halo_signal(H, nlo, nhi);
    }
    return;
  }
  const int tid = threadIdx.x, lx = (tid & 31) * 2, ly = (tid >> 5) * 2;
  if (tid == 0) {
    for (int s = 0; s < MG_NST; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  for (int e = tid; e < MG_TAB; e += NTHR) tab[e] = gtab[(OUT == 0 ? MG_TAB : 0) + e];
  uint32_t rs = 0, rp = 0;  // ring slot and mbarrier parity of the next plane this CTA consumes (identical in every thread)
  double acc[3] = {0, 0, 0};
  const long plane = (long)D.nx * D.ny, pplane = (long)M.pitch * D.ny;
  const int coff = (ly + 1) * MG_BOXW + lx + MG_HX;
  const float sc = K.sc, as = K.a * K.sc, gs = K.g * K.sc;
  const float sx0 = C.same[0][0], sx1 = C.same[0][1], sy0 = C.same[1][0];
  const float sy1 = C.same[1][1], sz0 = C.same[2][0], sz1 = C.same[2][1];
  const float dfx = C.diff[0], dfy = C.diff[1], dfz = C.diff[2];

  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    int tx, ty, zc;
    plan_item(P, item, tx, ty, zc);
    const int xo = tx * TX, yo = ty * TY, zb = zc * P.zlen;
    const int len = min(D.nz, zb + P.zlen) - zb, nplanes = len + 2;
    const int gx = xo + lx, gy = yo + ly;
    const bool colv = gx < D.nx;
    const bool row0 = colv && gy < D.ny, row1 = colv && gy + 1 < D.ny;
    // xy position classes of my four cells a,b (row gy) and c,d (row gy+1), pre-shifted table offsets
    const int ya = 3 * mg_cls(gy, D.ny), yc = 3 * mg_cls(gy + 1, D.ny);
    const int xa = mg_cls(gx, D.nx), xb = mg_cls(gx + 1, D.nx);
    const int ta = (ya + xa) << 7, tb = (ya + xb) << 7, tc = (yc + xa) << 7, td = (yc + xb) << 7;
    // aggregate of my block and its in-plane neighbours
    const int X = gx >> 1, Y = M.fy ? gy >> 1 : gy;
    const bool hW = ADD_E && row0 && X > 0, hE = ADD_E && row0 && X + 1 < M.nx1;
    const bool hN = ADD_E && row0 && Y > 0, hS = ADD_E && row0 && Y + 1 < M.ny1;
    const long eoff = (long)Y * M.nx1 + X;
    const long eplane = (long)M.nx1 * M.ny1;

    __syncthreads();  // ring idle (previous item consumed; barriers and table initialised)
    if (tid == 0) {
      const int npre = min(MG_NST, nplanes);
      uint32_t st = rs;
      for (int k = 0; k < npre; k++) {
        mbar_expect_tx(&full[st], MG_STAGE_BYTES);
        tma_load_3d(tiles + st * MG_STAGE_STRIDE, &tm, &full[st], xo - MG_HX, yo - 1, zb + k);
        st = st + 1 == MG_NST ? 0 : st + 1;
      }
    }
    long ci = ((long)zb * D.ny + gy) * D.nx + gx;    // global index of cell a in plane zb
    long cu = ((long)zb * D.ny + gy) * M.pitch + gx;  // the same cell in a pitched field
    unsigned cn0 = row0 ? *reinterpret_cast<const unsigned short*>(code + ci) : 0u;
    unsigned cn1 = row1 ? *reinterpret_cast<const unsigned short*>(code + ci + D.nx) : 0u;
    float2 un0 = make_float2(0.f, 0.f), un1 = make_float2(0.f, 0.f);  // u of my cells (when the staged field is an iterate)
    if (!IN_IS_U) {
      if (row0) un0 = *reinterpret_cast<const float2*>(uown + cu);
      if (row1) un1 = *reinterpret_cast<const float2*>(uown + cu + M.pitch);
    }
    // corrections of my aggregate column: own, W,E,N,S neighbours; of the aggregate plane below; prefetched next
    float E0 = 0.f, EW = 0.f, EE = 0.f, EN = 0.f, ES = 0.f, Eprev = 0.f;
    float nE0 = 0.f, nEW = 0.f, nEE = 0.f, nEN = 0.f, nES = 0.f;
    auto load_E = [&](int Z, float& a0, float& aw, float& ae, float& an, float& as_) {
      a0 = aw = ae = an = as_ = 0.f;
      if ((Z < 0 && !(M.lo && Z == -1)) || (Z >= M.nz1 && !(M.hi && Z == M.nz1))) return;
      const float* ep = M.e1 + (long)Z * eplane + eoff;
      if (row0) a0 = ep[0];
      if (hW) aw = ep[-1];
      if (hE) ae = ep[1];
      if (hN) an = ep[-M.nx1];
      if (hS) as_ = ep[M.nx1];
    };
    if (ADD_E) {
      const int Z = M.fz ? zb >> 1 : zb;
      float d1, d2, d3, d4;
      load_E(Z - 1, Eprev, d1, d2, d3, d4);
      load_E(Z, nE0, nEW, nEE, nEN, nES);
    }
    float2 pt, pb, lt, lb;  // current plane (top row a,b / bottom row c,d) and the plane below
    float fa, fb, fc, fd;   // conductances of the faces towards the plane below
    // slots of the plane below the chunk (sprev), of plane k (scur, parity pcur) -- advanced once per plane
    uint32_t sprev = rs, scur = rs + 1 == MG_NST ? 0 : rs + 1, pcur = rs + 1 == MG_NST ? rp ^ 1 : rp;
    {
      mbar_wait(&full[sprev], rp);
      const float* lo = reinterpret_cast<const float*>(tiles + sprev * MG_STAGE_STRIDE) + coff;
      lt = *reinterpret_cast<const float2*>(lo);
      lb = *reinterpret_cast<const float2*>(lo + MG_BOXW);
      mbar_wait(&full[scur], pcur);
      const float* c0 = reinterpret_cast<const float*>(tiles + scur * MG_STAGE_STRIDE) + coff;
      pt = *reinterpret_cast<const float2*>(c0);
      pb = *reinterpret_cast<const float2*>(c0 + MG_BOXW);
      const unsigned ca = cn0 & 0xffu, cb = cn0 >> 8, cc = cn1 & 0xffu, cd = cn1 >> 8;
      fa = pickf(ca & C_F, dfz, pickf(ca & C_PH, sz1, sz0));
      fb = pickf(cb & C_F, dfz, pickf(cb & C_PH, sz1, sz0));
      fc = pickf(cc & C_F, dfz, pickf(cc & C_PH, sz1, sz0));
      fd = pickf(cd & C_F, dfz, pickf(cd & C_PH, sz1, sz0));
    }
    float sacc = 0.f;

    for (int k = 1; k <= len; k++) {
      const int lz = zb + k - 1;
      __syncthreads();  // every thread is done with plane k-1's stage
      if (tid == 0 && k - 1 + MG_NST < nplanes) {  // the stage of plane k-1 takes plane k-1+NST
        mbar_expect_tx(&full[sprev], MG_STAGE_BYTES);
        tma_load_3d(tiles + sprev * MG_STAGE_STRIDE, &tm, &full[sprev], xo - MG_HX, yo - 1, zb + k - 1 + MG_NST);
      }
      const uint32_t sn = scur + 1 == MG_NST ? 0 : scur + 1, pn_ = scur + 1 == MG_NST ? pcur ^ 1 : pcur;
      mbar_wait(&full[sn], pn_);
      const float* cur = reinterpret_cast<const float*>(tiles + scur * MG_STAGE_STRIDE) + coff;
      const float* nxt = reinterpret_cast<const float*>(tiles + sn * MG_STAGE_STRIDE) + coff;
      sprev = scur;
      scur = sn;
      pcur = pn_;
      const float2 nt0 = *reinterpret_cast<const float2*>(nxt), nb0 = *reinterpret_cast<const float2*>(nxt + MG_BOXW);
      float2 ut = nt0, ub = nb0;
      const unsigned k0 = cn0, k1 = cn1;
      const float2 u0 = IN_IS_U ? pt : un0, u1 = IN_IS_U ? pb : un1;  // u of my cells
      if (k < len) {
        if (row0) cn0 = *reinterpret_cast<const unsigned short*>(code + ci + plane);
        if (row1) cn1 = *reinterpret_cast<const unsigned short*>(code + ci + plane + D.nx);
        if (!IN_IS_U) {
          if (row0) un0 = *reinterpret_cast<const float2*>(uown + cu + pplane);
          if (row1) un1 = *reinterpret_cast<const float2*>(uown + cu + pplane + M.pitch);
        }
      }
      float2 pn = *reinterpret_cast<const float2*>(cur - MG_BOXW);
      float2 ps = *reinterpret_cast<const float2*>(cur + 2 * MG_BOXW);
      float pwt = cur[-1], pet = cur[2], pwb = cur[MG_BOXW - 1], peb = cur[MG_BOXW + 2];
      const float2 ct = pt, cb2 = pb;  // stored values of my own cells
      float2 bt = lt, bb = lb;         // plane below
      if (ADD_E) {
        // corrections: a new aggregate plane starts at every plane (z not coarsened) or at even planes
        const bool first_of_pair = !M.fz || !(lz & 1);
        if (first_of_pair) {
          if (k > 1) Eprev = E0;
          E0 = nE0; EW = nEW; EE = nEE; EN = nEN; ES = nES;
          load_E((M.fz ? lz >> 1 : lz) + 1, nE0, nEW, nEE, nEN, nES);
        }
        const bool has_lo = D.z0 + lz > 0, has_hi = D.z0 + lz < D.nzg - 1;
        const float Eb = !has_lo ? 0.f : ((M.fz && (lz & 1)) ? E0 : Eprev);
        const float Ea = !has_hi ? 0.f : ((M.fz && !(lz & 1) && lz + 1 < D.nz) ? E0 : nE0);
        const float e0c = row1 ? E0 : 0.f;
        // in = sc * stored + e;  cells outside the domain (stored 0, e 0) stay 0
        pt.x = fmaf(sc, pt.x, E0); pt.y = fmaf(sc, pt.y, E0); pb.x = fmaf(sc, pb.x, e0c); pb.y = fmaf(sc, pb.y, e0c);
        pwt = fmaf(sc, pwt, EW); pwb = fmaf(sc, pwb, EW); pet = fmaf(sc, pet, EE); peb = fmaf(sc, peb, EE);
        pn.x = fmaf(sc, pn.x, EN); pn.y = fmaf(sc, pn.y, EN); ps.x = fmaf(sc, ps.x, ES); ps.y = fmaf(sc, ps.y, ES);
        bt.x = fmaf(sc, bt.x, Eb); bt.y = fmaf(sc, bt.y, Eb); bb.x = fmaf(sc, bb.x, Eb); bb.y = fmaf(sc, bb.y, Eb);
        ut.x = fmaf(sc, ut.x, Ea); ut.y = fmaf(sc, ut.y, Ea); ub.x = fmaf(sc, ub.x, Ea); ub.y = fmaf(sc, ub.y, Ea);
      }
      const unsigned ca = k0 & 0xffu, cb = k0 >> 8, cc = k1 & 0xffu, cd = k1 >> 8;
      const bool ha = ca & C_PH, hb = cb & C_PH, hc = cc & C_PH, hd = cd & C_PH;
      float g;
      // plane below, then the plane above (its conductances are kept for the next step)
      float qa = fa * bt.x, qb = fb * bt.y, qc = fc * bb.x, qd = fd * bb.y;
      fa = pickf(ca & C_B, dfz, pickf(ha, sz1, sz0));
      fb = pickf(cb & C_B, dfz, pickf(hb, sz1, sz0));
      fc = pickf(cc & C_B, dfz, pickf(hc, sz1, sz0));
      fd = pickf(cd & C_B, dfz, pickf(hd, sz1, sz0));
      qa += fa * ut.x;
      qb += fb * ut.y;
      qc += fc * ub.x;
      qd += fd * ub.y;
      // x faces
      qa += pickf(ca & C_W, dfx, pickf(ha, sx1, sx0)) * pwt;
      g = pickf(ca & C_E, dfx, pickf(ha, sx1, sx0));
      qa += g * pt.y;
      qb += g * pt.x;
      qb += pickf(cb & C_E, dfx, pickf(hb, sx1, sx0)) * pet;
      qc += pickf(cc & C_W, dfx, pickf(hc, sx1, sx0)) * pwb;
      g = pickf(cc & C_E, dfx, pickf(hc, sx1, sx0));
      qc += g * pb.y;
      qd += g * pb.x;
      qd += pickf(cd & C_E, dfx, pickf(hd, sx1, sx0)) * peb;
      // y faces
      qa += pickf(ca & C_N, dfy, pickf(ha, sy1, sy0)) * pn.x;
      g = pickf(ca & C_S, dfy, pickf(ha, sy1, sy0));
      qa += g * pb.x;
      qc += g * pt.x;
      qc += pickf(cc & C_S, dfy, pickf(hc, sy1, sy0)) * ps.x;
      qb += pickf(cb & C_N, dfy, pickf(hb, sy1, sy0)) * pn.y;
      g = pickf(cb & C_S, dfy, pickf(hb, sy1, sy0));
      qb += g * pb.y;
      qd += g * pt.y;
      qd += pickf(cd & C_S, dfy, pickf(hd, sy1, sy0)) * ps.y;

      const float* tz = tab + 9 * 128 * mg_cls(D.z0 + lz, D.nzg);
      const float wa = tz[ta | (ca & 127u)], wb = tz[tb | (cb & 127u)], wc = tz[tc | (cc & 127u)], wd = tz[td | (cd & 127u)];
      // local part a in_P + b u_P and the neighbour sum scaled by g (times sc when the inputs were not rescaled above)
      const float ka = ADD_E ? K.a : as, kg = ADD_E ? K.g : gs;
      const float la = ka * pt.x + K.b * u0.x, lb2 = ka * pt.y + K.b * u0.y;
      const float lc = ka * pb.x + K.b * u1.x, ld = ka * pb.y + K.b * u1.y;
      if (OUT == 0) {
        // res = d (a in + b u) + g sum
        float s = 0.f;
        if (row0) s += (wa * la + kg * qa) + (wb * lb2 + kg * qb);
        if (row1) s += (wc * lc + kg * qc) + (wd * ld + kg * qd);
        const bool last = !M.fz || (lz & 1) || lz == D.nz - 1;
        sacc = (M.fz && (lz & 1)) ? sacc + s : s;
        if (row0 && last) {
          const long ic = (long)(M.fz ? lz >> 1 : lz) * eplane + eoff;
          M.u1[ic] = M.dinv1 ? sacc * M.dinv1[ic] : sacc;
        }
      } else {
        // out = a in + b u + g (1/d) sum
        const float za = la + kg * wa * qa, zb2 = lb2 + kg * wb * qb, zc2 = lc + kg * wc * qc, zd = ld + kg * wd * qd;
        if (OUT == 1) {
          if (row0) *reinterpret_cast<float2*>(fout + cu) = make_float2(za, zb2);
          if (row1) *reinterpret_cast<float2*>(fout + cu + M.pitch) = make_float2(zc2, zd);
          if (push && (lz == 0 || lz == D.nz - 1)) {
            // a slab-boundary plane: the same values go straight into the neighbour's halo plane (NVLink stores)
            const long ro = (long)gy * M.pitch + gx;
            if (lz == 0 && H.lo_dst) {
              float* d = reinterpret_cast<float*>(H.lo_dst) + ro;
              if (row0) *reinterpret_cast<float2*>(d) = make_float2(za, zb2);
              if (row1) *reinterpret_cast<float2*>(d + M.pitch) = make_float2(zc2, zd);
            }
            if (lz == D.nz - 1 && H.hi_dst) {
              float* d = reinterpret_cast<float*>(H.hi_dst) + ro;
              if (row0) *reinterpret_cast<float2*>(d) = make_float2(za, zb2);
              if (row1) *reinterpret_cast<float2*>(d + M.pitch) = make_float2(zc2, zd);
            }
          }
        } else {
          // r_P = d u_P
          if (row0) {
            *reinterpret_cast<float2*>(zout + ci) = make_float2(za, zb2);
            acc[0] += (double)(__fdividef(u0.x, wa) * za) + (double)(__fdividef(u0.y, wb) * zb2);
          }
          if (row1) {
            *reinterpret_cast<float2*>(zout + ci + D.nx) = make_float2(zc2, zd);
            acc[0] += (double)(__fdividef(u1.x, wc) * zc2) + (double)(__fdividef(u1.y, wd) * zd);
          }
        }
      }
      lt = ct;
      lb = cb2;
      pt = nt0;
      pb = nb0;
      ci += plane;
      cu += pplane;
    }
    rs = scur + 1 == MG_NST ? 0 : scur + 1;  // scur: the plane above the chunk, the item's last
    rp = scur + 1 == MG_NST ? pcur ^ 1 : pcur;
    if (push && (zb == 0 || zb + len == D.nz)) halo_signal(H, zb == 0, zb + len == D.nz);
  }
  if (OUT == 2) {
    block_sum<3>(acc, red);
    if (tid == 0) {
      partials[blockIdx.x * 3 + 0] = acc[0];
      partials[blockIdx.x * 3 + 1] = 0.0;
      partials[blockIdx.x * 3 + 2] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CG vector kernels of the multigrid-preconditioned iteration.
//   k_cg_update_mg<1>: r -= alpha q;  u = r / d (FP32);  partial r.r                       (29 B per cell)
//   k_cg_update_mg<0>: u = r / d alone (after the initial residual, and for the test hook)
//   k_cg_dir_mg      : x += alpha p;  p = z + beta p  (z is FP32; first = 1: p = z only)     (36 B per cell; 28 with p in FP32)
// x += alpha p rides with the direction update because that kernel reads p anyway (p is then read
// once per iteration instead of twice).  It must happen exactly once per completed residual update
// even though every kernel of the blindly enqueued iterations after convergence returns early: the
// host passes the iteration's index `kit` and the update is applied iff S->it == kit + 1, i.e. iff
// this iteration's k_step_mg_rr counted it.
// u rows are `pitch` floats apart (multiple of 4: a TMA row stride must be a multiple of 16 bytes);
// u points at plane 0 of a field with one halo plane below.
// ------------------------------------------------------------------------------------------------
template <int UPDATE>
__global__ void __launch_bounds__(256) k_cg_update_mg(double* __restrict__ r, const double* __restrict__ q,
                                                      const uint8_t* __restrict__ code, const float* __restrict__ gtab,
                                                      float* __restrict__ u, int pitch, const Dom D,
                                                      const Scal* __restrict__ S, int par,
                                                      double* __restrict__ partials, const HaloPush H) {
  __shared__ double red[32 * 2];
  __shared__ float tab[27 * 128];  // 1 / d
  const bool push = H.lo_dst || H.hi_dst;
  if (S->done) {
    if (push) halo_signal(H, 1, 1);
    return;
  }
  for (int e = threadIdx.x; e < MG_TAB; e += blockDim.x) tab[e] = gtab[e];
  __syncthreads();
  const double alpha = UPDATE ? S->rz[par] / S->pq : 0.0;
  const long rows = (long)D.ny * D.nz;
  double acc[2] = {0, 0};
  // one row (y, z); `remote`: the same row of a neighbour's halo plane (slab-boundary planes only)
  auto do_row = [&](long row, float* remote) {
    const int y = (int)(row % D.ny), z = (int)(row / D.ny);
    const float* trow = tab + ((3 * mg_cls(y, D.ny) + 9 * mg_cls(D.z0 + z, D.nzg)) << 7);
    const long base = row * D.nx;
    float* urow = u + ((long)z * D.ny + y) * pitch;
    for (int xx = threadIdx.x * 2; xx < D.nx; xx += blockDim.x * 2) {
      const long i = base + xx;
      double2 vr = *reinterpret_cast<const double2*>(r + i);
      if (UPDATE) {
        const double2 vq = *reinterpret_cast<const double2*>(q + i);
        vr.x -= alpha * vq.x;
        vr.y -= alpha * vq.y;
        acc[1] += vr.x * vr.x + vr.y * vr.y;
        *reinterpret_cast<double2*>(r + i) = vr;
      }
      const unsigned cc = *reinterpret_cast<const unsigned short*>(code + i);
      const float t0 = trow[(mg_cls(xx, D.nx) << 7) | (cc & 127u)], t1 = trow[(mg_cls(xx + 1, D.nx) << 7) | ((cc >> 8) & 127u)];
      const float2 uv = make_float2((float)vr.x * t0, (float)vr.y * t1);
      *reinterpret_cast<float2*>(urow + xx) = uv;
      if (remote) *reinterpret_cast<float2*>(remote + xx) = uv;
    }
  };
  long first = 0, last = rows;  // rows of the interior planes
  if (push && D.nz >= 2) {
    // the two slab-boundary planes first: their u goes into the neighbours' halo planes while the rest is worked on
    first = D.ny;
    last = rows - D.ny;
    for (long y = blockIdx.x; y < D.ny; y += gridDim.x) {
      do_row(y, H.lo_dst ? reinterpret_cast<float*>(H.lo_dst) + y * pitch : nullptr);
      do_row(last + y, H.hi_dst ? reinterpret_cast<float*>(H.hi_dst) + y * pitch : nullptr);
    }
    halo_signal(H, 1, 1);
  }
  for (long row = first + blockIdx.x; row < last; row += gridDim.x) do_row(row, nullptr);
  if (UPDATE) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      partials[blockIdx.x * 2 + 0] = 0.0;
      partials[blockIdx.x * 2 + 1] = acc[1];
    }
  }
}

template <typename TP>
struct Pair2;
template <>
struct Pair2<double> { using type = double2; };
template <>
struct Pair2<float> { using type = float2; };

// TP: storage type of p.  double: as the other CG vectors.  float: p = z + beta p is built from the FP32 z, and x, r
// are both advanced with the SAME stored p (x += alpha p here, r -= alpha A p from the stencil of that p), so r stays
// the residual of x to FP64 rounding -- rounding p only perturbs the search direction (by 6e-8, next to the FP32
// preconditioner) and saves 12 of the iteration's 131 bytes per cell.
template <typename TP>
__global__ void __launch_bounds__(256) k_cg_dir_mg(double* __restrict__ x, TP* __restrict__ p,
                                                   const float* __restrict__ z, long n2, long n,
                                                   const Scal* __restrict__ S, int par, long long kit, int first,
                                                   const HaloPush H, long plane2) {
  using P2 = typename Pair2<TP>::type;
  const bool upd_x = !first && S->it == kit + 1, upd_p = !S->done;
  const bool push = (H.lo_dst || H.hi_dst) && n2 >= 2 * plane2;
  if (!upd_x && !upd_p) {
    if (push) halo_signal(H, 1, 1);
    return;
  }
  const double alpha = upd_x ? S->rz[par] / S->pq : 0.0;
  const double beta = first ? 0.0 : S->rz[par ^ 1] / S->rz[par];
  auto body = [&](long i, P2* remote) {
    P2 vp;  // first: p = z, and what the buffer held before (another solve's FP64 values, say) must not be looked at
    if (first) vp.x = vp.y = (TP)0;
    else vp = reinterpret_cast<P2*>(p)[i];
    if (upd_x) {
      double2 vx = reinterpret_cast<double2*>(x)[i];
      vx.x += alpha * (double)vp.x;
      vx.y += alpha * (double)vp.y;
      reinterpret_cast<double2*>(x)[i] = vx;
    }
    if (upd_p) {
      const float2 vz = reinterpret_cast<const float2*>(z)[i];
      vp.x = (TP)((double)vz.x + beta * (double)vp.x);
      vp.y = (TP)((double)vz.y + beta * (double)vp.y);
      reinterpret_cast<P2*>(p)[i] = vp;
      if (remote) *remote = vp;
    }
  };
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  long lo = 0, hi = n2;  // the interior planes
  if (push) {
    // the two slab-boundary planes first: the new p goes into the neighbours' halo planes (NVLink stores) while
    // the interior is still being updated
    lo = plane2;
    hi = n2 - plane2;
    for (long i = g; i < plane2; i += stride) {
      body(i, H.lo_dst ? reinterpret_cast<P2*>(H.lo_dst) + i : nullptr);
      body(hi + i, H.hi_dst ? reinterpret_cast<P2*>(H.hi_dst) + i : nullptr);
    }
    halo_signal(H, 1, 1);
  }
  for (long i = lo + g; i < hi; i += stride) body(i, nullptr);
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    if (upd_x) x[n - 1] += alpha * (double)p[n - 1];
    if (upd_p) p[n - 1] = (TP)((double)z[n - 1] + (first ? 0.0 : beta * (double)p[n - 1]));
  }
}

}  // namespace kb
