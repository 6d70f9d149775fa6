// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Shim that compiles the UNMODIFIED reference 3D CPU solver (header #included
// from /root/reference/Keff3D where it lies) into oracle/_ref/libkeffref3d.so
// with in-memory entry points.
//
// Reference functions called (all from Keff3D/keff3D.h):
//   DiscretizeMatrixCD3D :666   SolInitLinear :598   ParallelGS3D :492
// Glue restating main(): k-map keff3D.cpp:105-120, centres keff3D.cpp:85-95
// (including the xCenters quirk at :86), final flux/Keff keff3D.cpp:145-163.
#include "keff3D.h"

namespace {

struct Case3D {
  int nx, ny, nz;
  size_t n;
  double *xc, *yc, *zc, *K, *A, *b, *T, *qL, *qR;
  Case3D(int nx_, int ny_, int nz_) : nx(nx_), ny(ny_), nz(nz_), n((size_t)nx_ * ny_ * nz_) {
    xc = (double*)malloc(sizeof(double) * nx);
    yc = (double*)malloc(sizeof(double) * ny);
    zc = (double*)malloc(sizeof(double) * nz);
    K = (double*)malloc(sizeof(double) * n);
    A = (double*)malloc(sizeof(double) * n * 7);
    b = (double*)malloc(sizeof(double) * n);
    T = (double*)calloc(n, sizeof(double));
    qL = (double*)calloc((size_t)ny * nz, sizeof(double));
    qR = (double*)calloc((size_t)ny * nz, sizeof(double));
    for (int i = 0; i < nx; i++) xc[i] = (double)i * (1.0) + 0.5 * (1.0 / nx);
    for (int i = 0; i < ny; i++) yc[i] = (double)i * (1.0 / ny) + 0.5 * (1.0 / ny);
    for (int i = 0; i < nz; i++) zc[i] = (double)i * (1.0 / nz) + 0.5 * (1.0 / nz);
  }
  ~Case3D() { free(xc); free(yc); free(zc); free(K); free(A); free(b); free(T); free(qL); free(qR); }
  void kmap(const unsigned char* s, double ks, double kf) {
    for (size_t i = 0; i < n; i++) K[i] = s[i] == 1 ? ks : kf;
  }
};

}  // namespace

extern "C" {

// out[0..3] = keff, Q1, Q2, iters.  Thread count: the reference never sets it for 3D
// (keff3D.cpp has no omp_set_num_threads) -> controlled by nThreads here via omp_set_num_threads,
// 1 = the deterministic configuration.
int ref3d_solve(const unsigned char* solid, int nx, int ny, int nz, double ks, double kf, double TL,
                double TR, double tol, long maxIter, int nThreads, double* T_out, double* out) {
  Case3D c(nx, ny, nz);
  c.kmap(solid, ks, kf);
  DiscretizeMatrixCD3D(c.K, ny, nx, nz, c.A, c.b, TL, TR, c.xc, c.yc, c.zc);
  SolInitLinear(c.T, TL, TR, nx, ny, nz);
  omp_set_num_threads(nThreads);
  options o;
  int iters = ParallelGS3D(c.A, c.b, c.T, c.qL, c.qR, c.K, maxIter, tol, nx, ny, nz, TL, TR, c.xc, c.yc,
                           c.zc, nThreads, &o);
  double Q1 = 0, Q2 = 0, dx = c.xc[0], dy = c.yc[0] * 2, dz = c.zc[0] * 2;
  for (int j = 0; j < nz; j++)
    for (int k = 0; k < ny; k++) {
      size_t l = (size_t)j * ny * nx + (size_t)k * nx, r = (size_t)j * ny * nx + (size_t)(k + 1) * nx - 1;
      c.qL[(size_t)j * ny + k] = c.K[l] * dy * dz * (c.T[l] - TL) / dx;
      c.qR[(size_t)j * ny + k] = c.K[r] * dy * dz * (TR - c.T[r]) / dx;
      Q1 += c.qL[(size_t)j * ny + k];
      Q2 += c.qR[(size_t)j * ny + k];
    }
  out[0] = (Q1 + Q2) / 2 / (TR - TL);
  out[1] = Q1; out[2] = Q2; out[3] = iters;
  if (T_out) memcpy(T_out, c.T, c.n * sizeof(double));
  return 0;
}

int ref3d_discretize(const unsigned char* solid, int nx, int ny, int nz, double ks, double kf, double TL,
                     double TR, double* A, double* b) {
  Case3D c(nx, ny, nz);
  c.kmap(solid, ks, kf);
  DiscretizeMatrixCD3D(c.K, ny, nx, nz, c.A, c.b, TL, TR, c.xc, c.yc, c.zc);
  memcpy(A, c.A, c.n * 7 * sizeof(double));
  memcpy(b, c.b, c.n * sizeof(double));
  return 0;
}

// Fixed number of reference GS sweeps (throughput only; tol < 0 never satisfies the stop test).
int ref3d_sweeps(const unsigned char* solid, int nx, int ny, int nz, double ks, double kf, double TL,
                 double TR, long nSweeps, int nThreads) {
  double out[4];
  ref3d_solve(solid, nx, ny, nz, ks, kf, TL, TR, -1.0, nSweeps, nThreads, 0, out);
  return (int)out[3];
}

}  // extern "C"
