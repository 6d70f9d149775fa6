// mg.cuh -- geometric multigrid V-cycle used as the CG preconditioner of the grid solver.
//
// No reference counterpart: the reference relaxes the assembled system with lexicographic
// Gauss-Seidel (ParallelGS Keff2D/keff2D.h:758-850, ParallelGS3D Keff3D/keff3D.h:492-595), O(N^2)
// sweeps.  The linear system solved here is still exactly the reference's (DiscretizeMatrixCD2D/3D);
// the V-cycle only changes how fast CG gets to its solution (iteration counts independent of N).
//
// Hierarchy: level 0 is the voxel grid (operator evaluated from the code bytes, FP64 vectors).
// Level l+1 aggregates 2 cells per direction of level l (a lone last cell for odd sizes; directions
// of extent 1 are left alone).  Transfer operators are piecewise constant: restriction = sum over the
// aggregate, prolongation = copy to the aggregate's cells.  A coarse face conductance is the sum of
// the fine face conductances crossing that coarse face, halved along coarsened directions (the
// Galerkin product P^T A P of piecewise-constant transfers overestimates the coarse-grid stiffness by
// 2; halving restores the re-discretised operator on a uniform medium).  Coarse levels are stored:
// FP32 east/south/back face conductances, wall conductance, 1/a_P, right-hand side and correction.
// Smoother: one weighted-Jacobi sweep before and after the coarse correction (same omega -> the cycle
// is a symmetric positive definite operator, as CG requires).  With u = omega D^-1 b,
//    pre :  residual  b - A u               = (1-omega) b_P + sum_f c_f u_nb        -> restricted
//    post:  v = u + P e;   out = v + omega D^-1 (b - A v) = (1-omega) v_P + omega D^-1_P (b_P + sum_f c_f v_nb)
// so each level costs two stencil passes and never stores u, v or the residual.
// The coarsest level (<= 64 cells) is solved exactly by a dense inverse kept in global memory.
#pragma once
#include "kernels.cuh"

namespace kb {

constexpr int MG_MAX_LEVELS = 24;
constexpr int MG_DENSE_MAX = 64;

// one stored (coarse) level; every array has one spare plane below plane 0 and one above plane nz-1
struct MgLevel {
  int nx, ny, nz;      // extent of this level
  int fx, fy, fz;      // 1 if the direction is coarsened to build the NEXT level
  float *cx, *cy, *cz; // conductance of the face towards x+1 / y+1 / z+1 (0 at the domain end)
  float* wsum;         // fixed-temperature wall conductance of the cell (non-zero in the two wall columns)
  float* dinv;         // 1 / a_P
  float *b, *e;        // right-hand side, correction
};

struct MgLevelView {
  int nx, ny, nz;
  const float *cx, *cy, *cz, *dinv;
};

// ------------------------------------------------------------------------------------------------
// hierarchy construction
// ------------------------------------------------------------------------------------------------
// level 1 from the code bytes of the voxel grid (fine dims in D, global z end decided from D.nzg)
__global__ void k_mg_coarsen_fine(const uint8_t* __restrict__ code, const Coef C, const Dom D, int fx, int fy, int fz,
                                  MgLevel L) {
  const long n1 = (long)L.nx * L.ny * L.nz;
  const float sx = fx ? 0.5f : 1.0f, sy = fy ? 0.5f : 1.0f, sz = fz ? 0.5f : 1.0f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n1; i += (long)gridDim.x * blockDim.x) {
    const int X = (int)(i % L.nx);
    const long t = i / L.nx;
    const int Y = (int)(t % L.ny), Z = (int)(t / L.ny);
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
__device__ __forceinline__ float mg_u(const MgLevelView& L, const float* __restrict__ b, long i, float om) {
  return om * b[i] * L.dinv[i];
}

// restricted residual after the pre-smoothing sweep: bc = P^T ((1-omega) b + sum_f c_f u_nb)
__global__ void __launch_bounds__(256) k_mg_pre(const MgLevelView L, const float* __restrict__ b, int fx, int fy, int fz,
                                                int cnx, int cny, int cnz, float* __restrict__ bc, float om, int lo,
                                                int hi, const Scal* __restrict__ S) {
  if (S->done) return;
  const long plane = (long)L.nx * L.ny;
  const long nc = (long)cnx * cny * cnz;
  for (long ic = blockIdx.x * (long)blockDim.x + threadIdx.x; ic < nc; ic += (long)gridDim.x * blockDim.x) {
    const int X = (int)(ic % cnx);
    const long t = ic / cnx;
    const int Y = (int)(t % cny), Z = (int)(t / cny);
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
          const long i = ((long)z * L.ny + y) * L.nx + x;
          float s = (1.f - om) * b[i];
          if (x > 0) s += L.cx[i - 1] * mg_u(L, b, i - 1, om);
          if (x < L.nx - 1) s += L.cx[i] * mg_u(L, b, i + 1, om);
          if (y > 0) s += L.cy[i - L.nx] * mg_u(L, b, i - L.nx, om);
          if (y < L.ny - 1) s += L.cy[i] * mg_u(L, b, i + L.nx, om);
          if (z > 0 || lo) s += L.cz[i - plane] * mg_u(L, b, i - plane, om);
          if (z < L.nz - 1 || hi) s += L.cz[i] * mg_u(L, b, i + plane, om);
          acc += s;
        }
      }
    }
    bc[ic] = acc;
  }
}

__device__ __forceinline__ float mg_v(const MgLevelView& L, const float* __restrict__ b, const float* __restrict__ ec,
                                      int x, int y, int z, long i, int fx, int fy, int fz, int cnx, int cny,
                                      float om) {
  const int X = fx ? x >> 1 : x, Y = fy ? y >> 1 : y, Z = fz ? z >> 1 : z;  // z = -1 -> Z = -1 (halo plane)
  return om * b[i] * L.dinv[i] + ec[((long)Z * cny + Y) * cnx + X];
}

// e = (1-omega) v_P + omega D^-1_P (b_P + sum_f c_f v_nb),  v = omega D^-1 b + P ec
__global__ void __launch_bounds__(256) k_mg_post(const MgLevelView L, const float* __restrict__ b,
                                                 const float* __restrict__ ec, int fx, int fy, int fz, int cnx, int cny,
                                                 float* __restrict__ e, float om, int lo, int hi,
                                                 const Scal* __restrict__ S) {
  if (S->done) return;
  const long plane = (long)L.nx * L.ny, n = plane * L.nz;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % L.nx);
    const long t = i / L.nx;
    const int y = (int)(t % L.ny), z = (int)(t / L.ny);
    float s = b[i];
    if (x > 0) s += L.cx[i - 1] * mg_v(L, b, ec, x - 1, y, z, i - 1, fx, fy, fz, cnx, cny, om);
    if (x < L.nx - 1) s += L.cx[i] * mg_v(L, b, ec, x + 1, y, z, i + 1, fx, fy, fz, cnx, cny, om);
    if (y > 0) s += L.cy[i - L.nx] * mg_v(L, b, ec, x, y - 1, z, i - L.nx, fx, fy, fz, cnx, cny, om);
    if (y < L.ny - 1) s += L.cy[i] * mg_v(L, b, ec, x, y + 1, z, i + L.nx, fx, fy, fz, cnx, cny, om);
    if (z > 0 || lo) s += L.cz[i - plane] * mg_v(L, b, ec, x, y, z - 1, i - plane, fx, fy, fz, cnx, cny, om);
    if (z < L.nz - 1 || hi) s += L.cz[i] * mg_v(L, b, ec, x, y, z + 1, i + plane, fx, fy, fz, cnx, cny, om);
    const float vP = mg_v(L, b, ec, x, y, z, i, fx, fy, fz, cnx, cny, om);
    e[i] = (1.f - om) * vP + om * L.dinv[i] * s;
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
// voxel-grid (level 0) passes.  Same tiling as k_stencil_tma: a CTA marches a 64x16 xy tile in z, raw
// planes of r arrive by TMA in a ring; every arrived plane is turned once into u = omega r / a_P
// (POST: + the prolongated coarse correction) over the whole staged box -- halo included, zero outside
// the domain -- and kept in a second ring of 4 planes; the stencil then needs no boundary logic at
// all (a face towards the outside multiplies a zero).  1/a_P comes from a 128-entry table indexed by
// the code byte (interior cells) or is evaluated exactly (cells on the domain surface).
//   PRE : b1[aggregate] = sum over its cells of (1-omega) r_P + sum_f c_f u_nb        (FP32 out)
//   POST: z_P = (1-omega) v_P + omega/a_P (r_P + sum_f c_f v_nb),  partial r.z        (FP64 out)
// ------------------------------------------------------------------------------------------------
constexpr int MG_NST = 4, MG_NU = 4, MG_UPLANE = BOXW * BOXH;
constexpr int MG_SMEM = MG_NST * STAGE_STRIDE + MG_NU * MG_UPLANE * 8;

struct MgFine {
  const float* e1;  // level-1 correction (POST)
  float* b1;        // level-1 right-hand side (PRE)
  int nx1, ny1;
  int fy, fz;       // y / z coarsened (x always is: nx even >= 2)
  double omega;
};

// position class of a cell along one axis: 0 interior, 1 first, 2 last
__device__ __forceinline__ int mg_cls(int i, int n) { return i == 0 ? 1 : (i == n - 1 ? 2 : 0); }

template <int POST>
__global__ void __launch_bounds__(NTHR, 2) k_mg_fine(const __grid_constant__ CUtensorMap tm,
                                                     const double* __restrict__ r, const uint8_t* __restrict__ code,
                                                     double* __restrict__ zout, const MgFine M, const Coef C,
                                                     const Dom D, const Plan P, const Scal* __restrict__ S,
                                                     double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t full[MG_NST];
  __shared__ float tab[27 * 128];  // omega / a_P by [z class][y class][x class][code & 127]
  __shared__ double red[32 * 3];
  if (S->done) return;
  unsigned char* raw = dyn;
  double* U = reinterpret_cast<double*>(dyn + MG_NST * STAGE_STRIDE);
  const int tid = threadIdx.x, lx = (tid & 31) * 2, ly = (tid >> 5) * 2;
  if (tid == 0) {
    for (int s = 0; s < MG_NST; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  const bool flat = D.nzg == 1;  // 2D grid: no z faces anywhere
  for (int e = tid; e < 27 * 128; e += NTHR) {
    const int cls = e >> 7, cx = cls % 3, cy = (cls / 3) % 3, cz = cls / 9;
    tab[e] = (float)(M.omega / diag_of(C, (unsigned)(e & 127), cx == 1, cx == 2, cy == 1, cy == 2, flat || cz == 1,
                                       flat || cz == 2));
  }
  uint32_t seq = 0;
  double acc[3] = {0, 0, 0};
  const long plane = (long)D.nx * D.ny;
  const int coff = (ly + 1) * BOXW + lx + 2;
  const double om1 = 1.0 - M.omega;
  const double sx0 = C.same[0][0], sx1 = C.same[0][1], sy0 = C.same[1][0], sy1 = C.same[1][1];
  const double sz0 = C.same[2][0], sz1 = C.same[2][1];
  const double dfx = C.diff[0], dfy = C.diff[1], dfz = C.diff[2];

  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const int tx = item % P.ntx, t1 = item / P.ntx, ty = t1 % P.nty, zc = t1 / P.nty;
    const int xo = tx * TX, yo = ty * TY, zb = zc * P.zlen;
    const int len = min(D.nz, zb + P.zlen) - zb, nplanes = len + 2;
    const int gx = xo + lx, gy = yo + ly;
    const bool colv = gx < D.nx;
    const bool row0 = colv && gy < D.ny, row1 = colv && gy + 1 < D.ny;
    // xy classes of my four cells a,b,c,d (4 bits each)
    unsigned ocls = (unsigned)(3 * mg_cls(gy, D.ny) + mg_cls(gx, D.nx)) | (unsigned)(3 * mg_cls(gy, D.ny) + mg_cls(gx + 1, D.nx)) << 4 |
                    (unsigned)(3 * mg_cls(gy + 1, D.ny) + mg_cls(gx, D.nx)) << 8 |
                    (unsigned)(3 * mg_cls(gy + 1, D.ny) + mg_cls(gx + 1, D.nx)) << 12;
    asm volatile("" : "+r"(ocls));

    __syncthreads();  // rings idle, barriers + table initialised
    int issued = 0;
    if (tid == 0) {
      for (; issued < min(MG_NST, nplanes); issued++) {
        const uint32_t st = (seq + issued) % MG_NST;
        mbar_expect_tx(&full[st], STAGE_BYTES);
        tma_load_3d(raw + st * STAGE_STRIDE, &tm, &full[st], xo - 2, yo - 1, zb + issued);
      }
    }
    // Staging assignment: thread t owns row t/14 of the 68x18 box and the cells tc, tc+14, .. tc+56 of
    // it (tc = t%14), so every global address is one base register plus an immediate.
    const int tr = tid / 14, tc = tid - tr * 14;
    const int sy = yo - 1 + tr, sx = xo - 2 + tc;
    const bool rowok = tid < 252 && sy >= 0 && sy < D.ny;
    unsigned vmask = 0, scls = 0;  // valid bits; 4-bit xy class (3*ycls + xcls) of each of my 5 cells
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const int x = sx + 14 * i;
      if (rowok && tc + 14 * i < BOXW && x >= 0 && x < D.nx) vmask |= 1u << i;
      scls |= (unsigned)(3 * mg_cls(sy, D.ny) + mg_cls(x, D.nx)) << (4 * i);
    }
    // keep them packed (the compiler would otherwise re-derive every bit from coordinates per use)
    asm volatile("" : "+r"(vmask), "+r"(scls));
    const long cbase = (long)sy * D.nx + sx;
    const long ebase = POST ? (long)(M.fy ? sy >> 1 : sy) * M.nx1 + (sx >> 1) : 0;
    const int ubase = tr * BOXW + tc;
    unsigned pc[5];  // code bytes of the plane staged next, fetched one step ahead
    float pe[5];     // (POST) level-1 corrections of the same cells
    auto prefetch = [&](int j) {
      const int lz = zb - 1 + j;
      if (lz < 0 || lz >= D.nz) return;
      const uint8_t* cp = code + (long)lz * plane + cbase;
      const float* ep = POST ? M.e1 + (long)(M.fz ? lz >> 1 : lz) * M.ny1 * M.nx1 + ebase : nullptr;
#pragma unroll
      for (int i = 0; i < 5; i++)
        if (vmask & (1u << i)) {
          pc[i] = cp[14 * i];
          if (POST) pe[i] = ep[7 * i];
        }
    };
    // plane j of the item (local plane zb-1+j)  ->  U ring slot j & 3
    auto transform = [&](int j) {
      const uint32_t sq = seq + j, st = sq % MG_NST;
      mbar_wait(&full[st], (sq / MG_NST) & 1);
      const double* src = reinterpret_cast<const double*>(raw + st * STAGE_STRIDE) + ubase;
      double* dst = U + (j & (MG_NU - 1)) * MG_UPLANE + ubase;
      const int lz = zb - 1 + j;
      const bool zin = lz >= 0 && lz < D.nz;
      const float* tz = tab + 9 * 128 * mg_cls(D.z0 + lz, D.nzg);
      unsigned live = zin ? vmask : 0u;
      asm volatile("" : "+r"(live));
      double u[5];
#pragma unroll
      for (int i = 0; i < 5; i++) {
        u[i] = 0.0;
        if (live & (1u << i)) {
          u[i] = src[14 * i] * (double)tz[(((scls >> (4 * i)) & 15u) << 7) | (pc[i] & 127u)];
          if (POST) u[i] += (double)pe[i];
        }
      }
      if (j + 1 < nplanes) prefetch(j + 1);
      if (tid < 252) {
#pragma unroll
        for (int i = 0; i < 4; i++) dst[14 * i] = u[i];
        if (tc < BOXW - 56) dst[56] = u[4];
      }
    };
    prefetch(0);
    transform(0);
    transform(1);
    if (nplanes > 2) transform(2);

    long ci = ((long)zb * D.ny + gy) * D.nx + gx;  // global index of cell a in plane zb
    unsigned cn0 = row0 ? *reinterpret_cast<const unsigned short*>(code + ci) : 0u;
    unsigned cn1 = row1 ? *reinterpret_cast<const unsigned short*>(code + ci + D.nx) : 0u;
    double2 rn0 = row0 ? *reinterpret_cast<const double2*>(r + ci) : make_double2(0, 0);
    double2 rn1 = row1 ? *reinterpret_cast<const double2*>(r + ci + D.nx) : make_double2(0, 0);
    // conductances of the faces towards the plane below (zero-valued neighbour when there is none)
    double fa, fb, fc, fd;
    {
      const unsigned ca = cn0 & 0xffu, cb = cn0 >> 8, cc = cn1 & 0xffu, cd = cn1 >> 8;
      fa = pick(ca & C_F, dfz, pick(ca & C_PH, sz1, sz0));
      fb = pick(cb & C_F, dfz, pick(cb & C_PH, sz1, sz0));
      fc = pick(cc & C_F, dfz, pick(cc & C_PH, sz1, sz0));
      fd = pick(cd & C_F, dfz, pick(cd & C_PH, sz1, sz0));
    }
    float sacc = 0.f;

    for (int k = 1; k <= len; k++) {
      const int lz = zb + k - 1;
      __syncthreads();  // plane k+1 of the U ring is visible; raw stages of planes <= k+1 are consumed
      if (tid == 0) {
        for (; issued < nplanes && issued <= k + 1 + MG_NST; issued++) {
          // stage of plane `issued` was last used by plane issued-MG_NST <= k+1: consumed
          const uint32_t st = (seq + issued) % MG_NST;
          mbar_expect_tx(&full[st], STAGE_BYTES);
          tma_load_3d(raw + st * STAGE_STRIDE, &tm, &full[st], xo - 2, yo - 1, zb + issued);
        }
      }
      const unsigned k0 = cn0, k1 = cn1;
      const double2 r0 = rn0, r1 = rn1;
      if (k < len) {  // prefetch my codes and residuals of the next plane
        if (row0) {
          cn0 = *reinterpret_cast<const unsigned short*>(code + ci + plane);
          rn0 = *reinterpret_cast<const double2*>(r + ci + plane);
        }
        if (row1) {
          cn1 = *reinterpret_cast<const unsigned short*>(code + ci + plane + D.nx);
          rn1 = *reinterpret_cast<const double2*>(r + ci + plane + D.nx);
        }
      }
      if (k + 2 < nplanes) transform(k + 2);

      const double* cur = U + (k & (MG_NU - 1)) * MG_UPLANE + coff;
      const double* blw = U + ((k - 1) & (MG_NU - 1)) * MG_UPLANE + coff;
      const double* abv = U + ((k + 1) & (MG_NU - 1)) * MG_UPLANE + coff;
      const double2 pt = *reinterpret_cast<const double2*>(cur), pb = *reinterpret_cast<const double2*>(cur + BOXW);
      const double2 pn = *reinterpret_cast<const double2*>(cur - BOXW);
      const double2 ps = *reinterpret_cast<const double2*>(cur + 2 * BOXW);
      const double pwt = cur[-1], pet = cur[2], pwb = cur[BOXW - 1], peb = cur[BOXW + 2];
      const double2 lt = *reinterpret_cast<const double2*>(blw), lb = *reinterpret_cast<const double2*>(blw + BOXW);
      const double2 ut = *reinterpret_cast<const double2*>(abv), ub = *reinterpret_cast<const double2*>(abv + BOXW);
      const unsigned ca = k0 & 0xffu, cb = k0 >> 8, cc = k1 & 0xffu, cd = k1 >> 8;
      const bool ha = ca & C_PH, hb = cb & C_PH, hc = cc & C_PH, hd = cd & C_PH;
      double g;
      // from the plane below, then towards the plane above (kept for the next step)
      double qa = fa * lt.x, qb = fb * lt.y, qc = fc * lb.x, qd = fd * lb.y;
      fa = pick(ca & C_B, dfz, pick(ha, sz1, sz0));
      fb = pick(cb & C_B, dfz, pick(hb, sz1, sz0));
      fc = pick(cc & C_B, dfz, pick(hc, sz1, sz0));
      fd = pick(cd & C_B, dfz, pick(hd, sz1, sz0));
      qa += fa * ut.x;
      qb += fb * ut.y;
      qc += fc * ub.x;
      qd += fd * ub.y;
      // x faces
      qa += pick(ca & C_W, dfx, pick(ha, sx1, sx0)) * pwt;
      g = pick(ca & C_E, dfx, pick(ha, sx1, sx0));
      qa += g * pt.y;
      qb += g * pt.x;
      qb += pick(cb & C_E, dfx, pick(hb, sx1, sx0)) * pet;
      qc += pick(cc & C_W, dfx, pick(hc, sx1, sx0)) * pwb;
      g = pick(cc & C_E, dfx, pick(hc, sx1, sx0));
      qc += g * pb.y;
      qd += g * pb.x;
      qd += pick(cd & C_E, dfx, pick(hd, sx1, sx0)) * peb;
      // y faces
      qa += pick(ca & C_N, dfy, pick(ha, sy1, sy0)) * pn.x;
      g = pick(ca & C_S, dfy, pick(ha, sy1, sy0));
      qa += g * pb.x;
      qc += g * pt.x;
      qc += pick(cc & C_S, dfy, pick(hc, sy1, sy0)) * ps.x;
      qb += pick(cb & C_N, dfy, pick(hb, sy1, sy0)) * pn.y;
      g = pick(cb & C_S, dfy, pick(hb, sy1, sy0));
      qb += g * pb.y;
      qd += g * pt.y;
      qd += pick(cd & C_S, dfy, pick(hd, sy1, sy0)) * ps.y;

      if (!POST) {
        float s = 0.f;
        if (row0) s += (float)((om1 * r0.x + qa) + (om1 * r0.y + qb));
        if (row1) s += (float)((om1 * r1.x + qc) + (om1 * r1.y + qd));
        if (M.fz) {
          sacc = (lz & 1) ? sacc + s : s;
          if (row0 && ((lz & 1) || lz == D.nz - 1))
            M.b1[((long)(lz >> 1) * M.ny1 + (M.fy ? gy >> 1 : gy)) * M.nx1 + (gx >> 1)] = sacc;
        } else if (row0) {
          M.b1[((long)lz * M.ny1 + (M.fy ? gy >> 1 : gy)) * M.nx1 + (gx >> 1)] = s;
        }
      } else {
        const float* tz = tab + 9 * 128 * mg_cls(D.z0 + lz, D.nzg);
        const double da = tz[(ocls & 0xf) << 7 | (ca & 127u)], db = tz[(ocls >> 4 & 0xf) << 7 | (cb & 127u)];
        const double dc = tz[(ocls >> 8 & 0xf) << 7 | (cc & 127u)], dd = tz[(ocls >> 12 & 0xf) << 7 | (cd & 127u)];
        if (row0) {
          const double za = om1 * pt.x + da * (r0.x + qa), zb2 = om1 * pt.y + db * (r0.y + qb);
          *reinterpret_cast<double2*>(zout + ci) = make_double2(za, zb2);
          acc[0] += r0.x * za + r0.y * zb2;
        }
        if (row1) {
          const double zc2 = om1 * pb.x + dc * (r1.x + qc), zd = om1 * pb.y + dd * (r1.y + qd);
          *reinterpret_cast<double2*>(zout + ci + D.nx) = make_double2(zc2, zd);
          acc[0] += r1.x * zc2 + r1.y * zd;
        }
      }
      ci += plane;
    }
    seq += nplanes;
  }
  if (POST) {
    block_sum<3>(acc, red);
    if (tid == 0) {
      partials[blockIdx.x * 3 + 0] = acc[0];
      partials[blockIdx.x * 3 + 1] = 0.0;
      partials[blockIdx.x * 3 + 2] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CG vector kernels of the multigrid-preconditioned iteration (no a_P needed -> no code bytes read)
//   k_cg_update_mg: x += alpha p;  r -= alpha q;  partial r.r          (48 B per cell)
//   k_cg_dir_mg   : p = z + beta p   (first = 1: p = z)                 (24 B per cell)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cg_update_mg(double* __restrict__ x, double* __restrict__ r,
                                                      const double* __restrict__ p, const double* __restrict__ q, long n2,
                                                      long n, const Scal* __restrict__ S, int par,
                                                      double* __restrict__ partials) {
  __shared__ double red[32 * 2];
  if (S->done) return;
  const double alpha = S->rz[par] / S->pq;
  double acc[2] = {0, 0};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
    double2 vx = reinterpret_cast<double2*>(x)[i], vr = reinterpret_cast<double2*>(r)[i];
    const double2 vp = reinterpret_cast<const double2*>(p)[i], vq = reinterpret_cast<const double2*>(q)[i];
    vx.x += alpha * vp.x;
    vx.y += alpha * vp.y;
    vr.x -= alpha * vq.x;
    vr.y -= alpha * vq.y;
    acc[1] += vr.x * vr.x + vr.y * vr.y;
    reinterpret_cast<double2*>(x)[i] = vx;
    reinterpret_cast<double2*>(r)[i] = vr;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long i = n - 1;
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc[1] += ri * ri;
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2 + 0] = 0.0;
    partials[blockIdx.x * 2 + 1] = acc[1];
  }
}

__global__ void __launch_bounds__(256) k_cg_dir_mg(double* __restrict__ p, const double* __restrict__ z, long n2, long n,
                                                   const Scal* __restrict__ S, int par, int first) {
  if (S->done) return;
  const double beta = first ? 0.0 : S->rz[par ^ 1] / S->rz[par];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
    const double2 vz = reinterpret_cast<const double2*>(z)[i];
    double2 vp = reinterpret_cast<double2*>(p)[i];
    vp.x = vz.x + beta * vp.x;
    vp.y = vz.y + beta * vp.y;
    reinterpret_cast<double2*>(p)[i] = vp;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = z[n - 1] + beta * p[n - 1];
}

// scalar steps of the multigrid-preconditioned iteration
__global__ void k_step_mg_rr(Scal* S) {
  if (S->done) return;
  S->rr = S->tmp[1];
  S->it += 1;
  if (!(S->tmp[1] > S->thr2 * S->bb)) S->done = 1;
}
__global__ void k_step_mg_rz(Scal* S, int slot) {
  if (S->done) return;
  S->rz[slot] = S->tmp[0];
}

}  // namespace kb
