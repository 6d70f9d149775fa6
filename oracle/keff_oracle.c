/* keff_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker).
 *
 * A plain-C, single-threaded, FP64 restatement of the reference Keff-CFD CPU algorithm for the hot
 * path (conduction solve on a binary 2D/3D voxel image + boundary-flux integration).  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED.  The reference ships no golden outputs, so this file is pinned
 * against outputs of the unmodified reference itself, compiled here from /root/reference into
 * oracle/_ref (see oracle/Makefile, oracle/ref_wrap{2d,3d}.cpp) -- tests/test_oracle_vs_reference.py
 * requires bit-identical A, b, sweep counts and Keff -- and against the committed fixtures in
 * tests/golden/ produced by oracle/make_golden.py from the same unmodified reference.
 *
 * Every function cites the reference lines it restates.  The arithmetic (operation order and
 * association) follows the reference so that trajectories are reproduced bit for bit; the code
 * structure is this repo's own (one dimension-generic assembler and one sweep routine instead of the
 * reference's per-program copies).
 *
 * Layout everywhere: cell index = (z*ny + y)*nx + x, x fastest (keff3D.h:691); 2D is nz == 1.
 * Stencil slot order in A (keff2D.h:503-509, keff3D.h:668-677):
 *   0:P  1:W(x-1)  2:E(x+1)  3:S(y+1)  4:N(y-1)  5:F(z-1)  6:B(z+1)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- a3: WeightedHarmonicMean, keff2D.h:334-347 / keff3D.h:316-329 ------------------------------ */
static double hmean(double w1, double w2, double x1, double x2) { return (w1 + w2) / (w1 / x1 + w2 / x2); }

/* ---- a1: phase -> conductivity, 2D (keff2D.cpp:108-121): pixel < 150 is fluid; the mesh
 *      amplification is an integer-division replicate of the source pixel. ----------------------- */
void orc_kmap2d(const uint8_t* img, int W, int H, int ampX, int ampY, double ks, double kf, double* K) {
  int nx = W * ampX, ny = H * ampY;
  for (int r = 0; r < ny; r++)
    for (int c = 0; c < nx; c++) K[(size_t)r * nx + c] = img[(r / ampY) * W + c / ampX] < 150 ? kf : ks;
}

/* ---- a2: phase -> conductivity, 3D (keff3D.cpp:105-120): voxel == 1 is solid; amp is fixed to 1. */
void orc_kmap3d(const uint8_t* solid, size_t n, double ks, double kf, double* K) {
  for (size_t i = 0; i < n; i++) K[i] = solid[i] == 1 ? ks : kf;
}

/* Cell-centre coordinate as the reference builds it (keff2D.cpp:87-98). */
static double centre(int i, int n) { return (double)i * (1.0 / n) + 0.5 * (1.0 / n); }

/* ---- a4: DiscretizeMatrixCD2D, keff2D.h:482-582.  A[n*5], b[n].
 * The reference measures x and y spacings as differences of cell centres (1 ulp from 1/n) and
 * uses them both in the harmonic-mean weights and in the face conductance; the cross-direction
 * length is the exact dx or dy. ----------------------------------------------------------------- */
void orc_assemble2d(const double* K, int ny, int nx, double TL, double TR, double* A, double* b) {
  double dx = (double)1.0 / nx, dy = (double)1.0 / ny;
  double* xc = (double*)malloc(sizeof(double) * nx);
  double* yc = (double*)malloc(sizeof(double) * ny);
  for (int i = 0; i < nx; i++) xc[i] = centre(i, nx);
  for (int i = 0; i < ny; i++) yc[i] = centre(i, ny);
  for (int r = 0; r < ny; r++)
    for (int c = 0; c < nx; c++) {
      size_t p = (size_t)r * nx + c;
      double* a = A + p * 5;
      double gw = 0, ge = 0, gs = 0, gn = 0;
      a[0] = a[1] = a[2] = a[3] = a[4] = 0;
      b[p] = 0;
      /* x direction: interior faces use the harmonic mean; the two x walls are fixed-temperature
       * faces half a cell away (keff2D.h:529-544). */
      if (c > 0) {
        double d = xc[c] - xc[c - 1];
        gw = hmean(d / 2, d / 2, K[p], K[p - 1]) * dy / d;
        a[1] = -gw;
      } else {
        gw = K[p] * dy / xc[c];
        b[p] += TL * K[p] * dy / xc[c];
      }
      if (c < nx - 1) {
        double d = xc[c + 1] - xc[c];
        ge = hmean(d / 2, d / 2, K[p], K[p + 1]) * dy / d;
        a[2] = -ge;
      } else {
        double d = 1 - xc[c];
        ge = K[p] * dy / d;
        b[p] += TR * K[p] * dy / d;
      }
      a[0] += (ge + gw);
      /* y direction: row 0 has no N, last row no S (adiabatic), keff2D.h:555-577. */
      if (r > 0) {
        double d = yc[r] - yc[r - 1];
        gn = hmean(d / 2, d / 2, K[p], K[p - nx]) * dx / d;
        a[4] = -gn;
      }
      if (r < ny - 1) {
        double d = yc[r + 1] - yc[r];
        gs = hmean(d / 2, d / 2, K[p + nx], K[p]) * dx / d;
        a[3] = -gs;
      }
      if (r > 0 && r < ny - 1) a[0] += (gn + gs);
      else if (r > 0) a[0] += gn;
      else a[0] += (gs);
    }
  free(xc);
  free(yc);
}

/* ---- a5: DiscretizeMatrixCD3D, keff3D.h:666-763.  A[n*7], b[n].  Exact dx,dy,dz throughout. ------ */
void orc_assemble3d(const double* K, int nx, int ny, int nz, double TL, double TR, double* A, double* b) {
  double dx = (double)1.0 / nx, dy = (double)1.0 / ny, dz = (double)1.0 / nz;
  size_t sy = (size_t)nx, sz = (size_t)nx * ny;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        size_t p = (size_t)z * sz + (size_t)y * sy + x;
        double* a = A + p * 7;
        for (int t = 0; t < 7; t++) a[t] = 0;
        b[p] = 0;
        /* z faces (keff3D.h:698-714): F = slice z-1 (slot 5), B = slice z+1 (slot 6) */
        double gf = 0, gb = 0;
        if (z > 0) { gf = hmean(dz / 2, dz / 2, K[p], K[p - sz]) * dy * dx / dz; a[5] = -gf; }
        if (z < nz - 1) { gb = hmean(dz / 2, dz / 2, K[p], K[p + sz]) * dy * dx / dz; a[6] = -gb; }
        if (z > 0 && z < nz - 1) a[0] += (gf + gb);
        else if (z > 0) a[0] += gf;
        else if (z < nz - 1) a[0] += gb;
        /* y faces (keff3D.h:717-733): S = row y+1 (slot 3), N = row y-1 (slot 4) */
        double gs = 0, gn = 0;
        if (y < ny - 1) { gs = hmean(dy / 2, dy / 2, K[p], K[p + sy]) * dx * dz / dy; a[3] = -gs; }
        if (y > 0) { gn = hmean(dy / 2, dy / 2, K[p], K[p - sy]) * dx * dz / dy; a[4] = -gn; }
        if (y > 0 && y < ny - 1) a[0] += (gn + gs);
        else if (y > 0) a[0] += gn;
        else if (y < ny - 1) a[0] += gs;
        /* x faces (keff3D.h:736-757): fixed-temperature walls at x = 0 and x = nx-1 */
        double ge, gw;
        if (x == 0) {
          ge = hmean(dx / 2, dx / 2, K[p], K[p + 1]) * dy * dz / dx;
          gw = K[p] * dy * dz / (dx / 2);
          a[2] = -ge;
          b[p] += TL * K[p] * dy * dz / (dx / 2);
        } else if (x == nx - 1) {
          ge = K[p] * dy * dz / (dx / 2);
          gw = hmean(dx / 2, dx / 2, K[p], K[p - 1]) * dy * dz / dx;
          a[1] = -gw;
          b[p] += TR * K[p] * dy * dz / (dx / 2);
        } else {
          ge = hmean(dx / 2, dx / 2, K[p], K[p + 1]) * dy * dz / dx;
          gw = hmean(dx / 2, dx / 2, K[p], K[p - 1]) * dy * dz / dx;
          a[1] = -gw;
          a[2] = -ge;
        }
        a[0] += (ge + gw);
      }
}

/* ---- a6: SolInitLinear as it actually behaves (keff2D.h:584-611, keff3D.h:598-627): the integer
 * division 1/numCols is 0, so the "ramp" is the constant TL when TR > TL (2D is called with the
 * arguments swapped, keff2D.cpp:135, and lands on TL as well: the tR<tL branch gives "+ tR" of the
 * swapped pair); when TR < TL the constant is TR; untouched when equal. ------------------------- */
void orc_init_T(double* T, size_t n, double TL, double TR) {
  double v = TL;
  if (TR < TL) v = TR;
  if (TR == TL) return;
  for (size_t i = 0; i < n; i++) T[i] = v;
}

/* ---- a10: final flux + Keff.  2D keff2D.cpp:151-170, 3D keff3D.cpp:145-163.  out = {keff,Q1,Q2}.
 * 2D: left face is half the first centre away (xc[0]), right face 1 - xc[nx-1].
 * 3D: dx here is xCenters[0] = half a cell; dy, dz are doubled first centres. ------------------- */
void orc_flux2d(const double* K, const double* T, int ny, int nx, double TL, double TR, double* qL,
                double* qR, double* out) {
  double dy = centre(0, ny) * 2, xl = centre(0, nx), xr = 1 - centre(nx - 1, nx), Q1 = 0, Q2 = 0;
  for (int r = 0; r < ny; r++) {
    size_t l = (size_t)r * nx, e = (size_t)(r + 1) * nx - 1;
    double a = K[l] * dy * (T[l] - TL) / xl, c = K[e] * dy * (TR - T[e]) / xr;
    if (qL) qL[r] = a;
    if (qR) qR[r] = c;
    Q1 += a;
    Q2 += c;
  }
  out[0] = (Q2 + Q1) / 2 / (TR - TL);
  out[1] = Q1;
  out[2] = Q2;
}

void orc_flux3d(const double* K, const double* T, int nx, int ny, int nz, double TL, double TR, double* qL,
                double* qR, double* out) {
  double dxh = (double)0 * (1.0) + 0.5 * (1.0 / nx), dy = centre(0, ny) * 2, dz = centre(0, nz) * 2;
  double Q1 = 0, Q2 = 0;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++) {
      size_t l = ((size_t)z * ny + y) * nx, e = l + nx - 1, f = (size_t)z * ny + y;
      double a = K[l] * dy * dz * (T[l] - TL) / dxh, c = K[e] * dy * dz * (TR - T[e]) / dxh;
      if (qL) qL[f] = a;
      if (qR) qR[f] = c;
      Q1 += a;
      Q2 += c;
    }
  out[0] = (Q1 + Q2) / 2 / (TR - TL);
  out[1] = Q1;
  out[2] = Q2;
}

/* ---- a7/a8/a9: the reference's stationary iteration and stopping rule.
 * SerialGS keff2D.h:853-940 (== ParallelGS :758-850 on one thread) and ParallelGS3D keff3D.h:492-595:
 * lexicographic in-place Gauss-Seidel, sigma accumulated over slots 1.. in order skipping exact zeros,
 * T = 1/aP * (b - sigma).  Every `cadence` sweeps (1000, dropping to 100 once the relative Keff change
 * is < 1e-3; the "< 1e-4 -> 10" branch is unreachable) the boundary flux is summed and the relative
 * change of Keff against the previous check decides convergence; keffOld starts at 1.
 * 2D quirks kept: K is indexed with stride ny in the check (keff2D.h:830-831) and the in-loop Keff is
 * divided by (TR-TL)/ny.  Returns the sweep count. ---------------------------------------------- */
long orc_gs(const double* A, const double* b, double* T, const double* K, int nd, int nx, int ny, int nz,
            double TL, double TR, double tol, long maxIter) {
  const size_t n = (size_t)nx * ny * nz;
  const long off[7] = {0, -1, 1, nx, -nx, -(long)nx * ny, (long)nx * ny};
  long it = 0, cadence = 1000;
  double change = 1, kOld = 1;
  while (change > tol && it < maxIter) {
    for (size_t i = 0; i < n; i++) {
      const double* a = A + i * nd;
      double sigma = 0;
      for (int j = 1; j < nd; j++)
        if (a[j] != 0) sigma += a[j] * T[(long)i + off[j]];
      T[i] = 1 / a[0] * (b[i] - sigma);
    }
    it++;
    if (it % cadence == 0) {
      double Q1 = 0, Q2 = 0, kNew;
      if (nd == 5) {
        double dy = centre(0, ny) * 2, xl = centre(0, nx), xr = 1 - centre(nx - 1, nx);
        for (int r = 0; r < ny; r++) {
          Q1 += K[(size_t)r * ny] * dy * (T[(size_t)r * nx] - TL) / xl;
          Q2 += K[(size_t)(r + 1) * ny - 1] * dy * (TR - T[(size_t)(r + 1) * nx - 1]) / xr;
        }
        kNew = (Q1 + Q2) / 2 / ((TR - TL) / ny);
      } else {
        double o[3];
        orc_flux3d(K, T, nx, ny, nz, TL, TR, 0, 0, o);
        kNew = o[0];
      }
      change = fabs((kNew - kOld) / kOld);
      kOld = kNew;
      if (change < 0.001) cadence = 100;
    }
  }
  return it;
}

/* ---- a11: Residual, keff2D.h:384-424: sum over cells of |qW - qE + qN - qS| (energy imbalance),
 * harmonic means recomputed from K; note the reference uses dy/dx for the y fluxes too. ---------- */
double orc_residual2d(const double* K, const double* T, int ny, int nx, double TL, double TR) {
  double dx = 1.0 / nx, dy = 1.0 / ny, R = 0;
  for (int r = 0; r < ny; r++)
    for (int c = 0; c < nx; c++) {
      size_t p = (size_t)r * nx + c;
      double qW, qE, qN = 0, qS = 0;
      qW = c == 0 ? dy / (dx / 2) * K[p] * (T[p] - TL)
                  : dy / (dx)*hmean(dx / 2, dx / 2, K[p], K[p - 1]) * (T[p] - T[p - 1]);
      qE = c == nx - 1 ? dy / (dx / 2) * K[p] * (TR - T[p])
                       : dy / (dx)*hmean(dx / 2, dx / 2, K[p], K[p + 1]) * (T[p + 1] - T[p]);
      if (r < ny - 1) qS = dy / dx * hmean(dx / 2, dx / 2, K[p + nx], K[p]) * (T[p + nx] - T[p]);
      if (r > 0) qN = dy / dx * hmean(dx / 2, dx / 2, K[p - nx], K[p]) * (T[p] - T[p - nx]);
      R += fabs(qW - qE + qN - qS);
    }
  return R;
}

/* ---- whole-path drivers (what main() does between reading the structure and writing the row) ---- */
/* 2D: keff2D.cpp:102-170.  out = {keff, Q1, Q2, iters, residual}. */
int orc_solve2d(const uint8_t* img, int W, int H, int ampX, int ampY, double ks, double kf, double TL,
                double TR, double tol, long maxIter, double* T_out, double* qL, double* qR, double* out) {
  int nx = W * ampX, ny = H * ampY;
  size_t n = (size_t)nx * ny;
  double *K = malloc(n * 8), *A = malloc(n * 40), *b = malloc(n * 8), *T = calloc(n, 8);
  orc_kmap2d(img, W, H, ampX, ampY, ks, kf, K);
  orc_assemble2d(K, ny, nx, TL, TR, A, b);
  orc_init_T(T, n, TL, TR);
  out[3] = (double)orc_gs(A, b, T, K, 5, nx, ny, 1, TL, TR, tol, maxIter);
  orc_flux2d(K, T, ny, nx, TL, TR, qL, qR, out);
  out[4] = orc_residual2d(K, T, ny, nx, TL, TR);
  if (T_out) memcpy(T_out, T, n * 8);
  free(K); free(A); free(b); free(T);
  return 0;
}

/* 3D: keff3D.cpp:99-163.  out = {keff, Q1, Q2, iters}. */
int orc_solve3d(const uint8_t* solid, int nx, int ny, int nz, double ks, double kf, double TL, double TR,
                double tol, long maxIter, double* T_out, double* out) {
  size_t n = (size_t)nx * ny * nz;
  double *K = malloc(n * 8), *A = malloc(n * 56), *b = malloc(n * 8), *T = calloc(n, 8);
  orc_kmap3d(solid, n, ks, kf, K);
  orc_assemble3d(K, nx, ny, nz, TL, TR, A, b);
  orc_init_T(T, n, TL, TR);
  out[3] = (double)orc_gs(A, b, T, K, 7, nx, ny, nz, TL, TR, tol, maxIter);
  orc_flux3d(K, T, nx, ny, nz, TL, TR, 0, 0, out);
  if (T_out) memcpy(T_out, T, n * 8);
  free(K); free(A); free(b); free(T);
  return 0;
}

/* ---- the discrete solution of the reference's own system --------------------------------------
 * y = A x with the assembled reference matrix (the operator the iteration above relaxes). */
void orc_apply(const double* A, int nd, int nx, int ny, int nz, const double* x, double* y) {
  const size_t n = (size_t)nx * ny * nz;
  const long off[7] = {0, -1, 1, nx, -nx, -(long)nx * ny, (long)nx * ny};
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    const double* a = A + i * nd;
    double s = a[0] * x[i];
    for (int j = 1; j < nd; j++)
      if (a[j] != 0) s += a[j] * x[(long)i + off[j]];
    y[i] = s;
  }
}

/* Diagonally preconditioned conjugate gradients on the assembled reference system A T = b, run to
 * ||b - A T|| / ||b|| <= eps: the limit the reference's Gauss-Seidel converges to when its
 * `Convergence` is tightened (checked against the tightened reference in tests/golden).  T holds the
 * initial guess on entry.  Returns iterations; *relres gets the final TRUE relative residual. */
long orc_pcg(const double* A, const double* b, double* T, int nd, int nx, int ny, int nz, double eps,
             long maxIter, double* relres) {
  const size_t n = (size_t)nx * ny * nz;
  double *r = malloc(n * 8), *p = malloc(n * 8), *q = malloc(n * 8);
  double bb = 0, rz = 0, rr = 0;
  long it = 0;
  orc_apply(A, nd, nx, ny, nz, T, q);
  for (size_t i = 0; i < n; i++) {
    r[i] = b[i] - q[i];
    p[i] = r[i] / A[i * nd];
    bb += b[i] * b[i];
    rz += r[i] * p[i];
    rr += r[i] * r[i];
  }
  while (it < maxIter && sqrt(rr) > eps * sqrt(bb)) {
    double pq = 0, rzn = 0;
    orc_apply(A, nd, nx, ny, nz, p, q);
    for (size_t i = 0; i < n; i++) pq += p[i] * q[i];
    double alpha = rz / pq;
    rr = 0;
    for (size_t i = 0; i < n; i++) {
      T[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rzn += r[i] * r[i] / A[i * nd];
      rr += r[i] * r[i];
    }
    double beta = rzn / rz;
    rz = rzn;
    for (size_t i = 0; i < n; i++) p[i] = r[i] / A[i * nd] + beta * p[i];
    it++;
  }
  orc_apply(A, nd, nx, ny, nz, T, q);
  rr = 0;
  for (size_t i = 0; i < n; i++) rr += (b[i] - q[i]) * (b[i] - q[i]);
  if (relres) *relres = sqrt(rr / bb);
  free(r); free(p); free(q);
  return it;
}

/* Discrete solution for a 3D (or nz = 1 with the 3D formulas) / 2D structure: assemble as the
 * reference does, solve by orc_pcg, integrate the flux as the reference does.
 * dim = 2 uses the 2D assembler (img semantics, pixel<150 fluid), dim = 3 the 3D one (voxel==1 solid).
 * out = {keff, Q1, Q2, pcg iterations, true relative residual}. */
int orc_discrete(int dim, const uint8_t* s, int nx, int ny, int nz, double ks, double kf, double TL,
                 double TR, double eps, long maxIter, double* T_out, double* out) {
  size_t n = (size_t)nx * ny * nz;
  int nd = dim == 2 ? 5 : 7;
  double *K = malloc(n * 8), *A = malloc(n * 8 * nd), *b = malloc(n * 8), *T = calloc(n, 8);
  if (dim == 2) {
    orc_kmap2d(s, nx, ny, 1, 1, ks, kf, K);
    orc_assemble2d(K, ny, nx, TL, TR, A, b);
  } else {
    orc_kmap3d(s, n, ks, kf, K);
    orc_assemble3d(K, nx, ny, nz, TL, TR, A, b);
  }
  out[3] = (double)orc_pcg(A, b, T, nd, nx, ny, nz, eps, maxIter, &out[4]);
  if (dim == 2) orc_flux2d(K, T, ny, nx, TL, TR, 0, 0, out);
  else orc_flux3d(K, T, nx, ny, nz, TL, TR, 0, 0, out);
  if (T_out) memcpy(T_out, T, n * 8);
  free(K); free(A); free(b); free(T);
  return 0;
}

/* Bounded CPU timing sample for bench.py: n_sweeps of the reference-style Gauss-Seidel on a 3D
 * structure (no convergence test fires: tol < 0). Returns sweeps done. */
long orc_gs3d_sweeps(const uint8_t* solid, int nx, int ny, int nz, double ks, double kf, double TL,
                     double TR, long nSweeps) {
  double out[4];
  orc_solve3d(solid, nx, ny, nz, ks, kf, TL, TR, -1.0, nSweeps, 0, out);
  return (long)out[3];
}
