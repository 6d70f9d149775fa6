// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Shim that compiles the UNMODIFIED reference 2D CPU solver (the header is
// #included from /root/reference/Keff2D where it lies, nothing is copied) into
// oracle/_ref/libkeffref2d.so and exposes in-memory entry points, so that the
// real reference functions can be driven on arbitrary masks without JPEG files.
//
// Reference functions called (all from Keff2D/keff2D.h):
//   DiscretizeMatrixCD2D :482   SolInitLinear :584   SerialGS :853
//   ParallelGS :758             Residual :384        calcPorosity :306
// The glue between them restates what main() does inline:
//   k-map loop      keff2D.cpp:108-121      centres  keff2D.cpp:87-98
//   final flux/Keff keff2D.cpp:151-170      batch loop keff2D.cpp:341-471
#include "keff2D.h"

namespace {

struct Case2D {
  int nx, ny;
  double *xc, *yc, *K, *A, *b, *T, *qL, *qR;
  Case2D(int nx_, int ny_) : nx(nx_), ny(ny_) {
    size_t n = (size_t)nx * ny;
    xc = (double*)malloc(sizeof(double) * nx);
    yc = (double*)malloc(sizeof(double) * ny);
    K = (double*)malloc(sizeof(double) * n);
    A = (double*)malloc(sizeof(double) * n * 5);
    b = (double*)malloc(sizeof(double) * n);
    T = (double*)calloc(n, sizeof(double));
    qL = (double*)calloc(ny, sizeof(double));
    qR = (double*)calloc(ny, sizeof(double));
    for (int i = 0; i < nx; i++) xc[i] = (double)i * (1.0 / nx) + 0.5 * (1.0 / nx);
    for (int i = 0; i < ny; i++) yc[i] = (double)i * (1.0 / ny) + 0.5 * (1.0 / ny);
  }
  ~Case2D() { free(xc); free(yc); free(K); free(A); free(b); free(T); free(qL); free(qR); }
  void kmap(const unsigned char* img, int W, int ampX, int ampY, double ks, double kf) {
    for (int i = 0; i < ny; i++)
      for (int j = 0; j < nx; j++)
        K[(size_t)i * nx + j] = img[(i / ampY) * W + (j / ampX)] < 150 ? kf : ks;
  }
};

}  // namespace

extern "C" {

// out[0..5] = keff, Q1, Q2, iters, residual(Residual()), porosity
// solver: 0 = SerialGS (batch-mode solver), 1 = ParallelGS (single-image solver)
int ref2d_solve(const unsigned char* img, int W, int H, int ampX, int ampY, double ks, double kf,
                double TL, double TR, double tol, long maxIter, int solver, int nThreads,
                double* T_out, double* qL_out, double* qR_out, double* out) {
  Case2D c(W * ampX, H * ampY);
  c.kmap(img, W, ampX, ampY, ks, kf);
  DiscretizeMatrixCD2D(c.K, c.ny, c.nx, c.A, c.b, TL, TR, c.xc, c.yc);
  SolInitLinear(c.T, TR, TL, c.nx, c.ny);  // argument order as at keff2D.cpp:135
  int iters;
  if (solver == 1) {
    omp_set_num_threads(nThreads);
    iters = ParallelGS(c.A, c.b, c.T, c.qL, c.qR, c.K, maxIter, tol, c.nx, c.ny, TL, TR, c.xc, c.yc, nThreads);
  } else {
    iters = SerialGS(c.A, c.b, c.T, c.qL, c.qR, c.K, maxIter, tol, c.nx, c.ny, TL, TR, c.xc, c.yc);
  }
  double dy = c.yc[0] * 2, Q1 = 0, Q2 = 0;
  for (int j = 0; j < c.ny; j++) {
    c.qL[j] = c.K[(size_t)j * c.nx] * dy * (c.T[(size_t)j * c.nx] - TL) / (c.xc[0]);
    c.qR[j] = c.K[(size_t)(j + 1) * c.nx - 1] * dy * (TR - c.T[(size_t)(j + 1) * c.nx - 1]) / (1 - c.xc[c.nx - 1]);
  }
  for (int j = 0; j < c.ny; j++) { Q1 += c.qL[j]; Q2 += c.qR[j]; }
  options o;
  o.TempLeft = TL; o.TempRight = TR;
  out[0] = (Q2 + Q1) / 2 / (TR - TL);
  out[1] = Q1; out[2] = Q2; out[3] = iters;
  out[4] = Residual(c.ny, c.nx, &o, c.T, c.K);
  out[5] = calcPorosity((unsigned char*)img, W, H);
  size_t n = (size_t)c.nx * c.ny;
  if (T_out) memcpy(T_out, c.T, n * sizeof(double));
  if (qL_out) memcpy(qL_out, c.qL, c.ny * sizeof(double));
  if (qR_out) memcpy(qR_out, c.qR, c.ny * sizeof(double));
  return 0;
}

// Expose the assembled system for operator-level parity: A[n*5], b[n], K[n].
int ref2d_discretize(const unsigned char* img, int W, int H, int ampX, int ampY, double ks, double kf,
                     double TL, double TR, double* A, double* b, double* K) {
  Case2D c(W * ampX, H * ampY);
  c.kmap(img, W, ampX, ampY, ks, kf);
  DiscretizeMatrixCD2D(c.K, c.ny, c.nx, c.A, c.b, TL, TR, c.xc, c.yc);
  size_t n = (size_t)c.nx * c.ny;
  memcpy(A, c.A, n * 5 * sizeof(double));
  memcpy(b, c.b, n * sizeof(double));
  if (K) memcpy(K, c.K, n * sizeof(double));
  return 0;
}

// Reference batch mode (keff2D.cpp:341): OpenMP over images, SerialGS per image.
// imgs = n_img consecutive W*H u8 images; results[n_img*4] = keff, Q1, Q2, iters.
int ref2d_batch(const unsigned char* imgs, int n_img, int W, int H, double ks, double kf, double TL,
                double TR, double tol, long maxIter, int nThreads, double* results) {
  omp_set_num_threads(nThreads);
#pragma omp parallel for schedule(dynamic, 1)
  for (int im = 0; im < n_img; im++) {
    double out[6];
    ref2d_solve(imgs + (size_t)im * W * H, W, H, 1, 1, ks, kf, TL, TR, tol, maxIter, 0, 1, 0, 0, 0, out);
    for (int k = 0; k < 4; k++) results[(size_t)im * 4 + k] = out[k];
  }
  return 0;
}

// Time a fixed number of reference GS sweeps (throughput only): returns sweeps done.
int ref2d_sweeps(const unsigned char* img, int W, int H, double ks, double kf, double TL, double TR,
                 long nSweeps) {
  Case2D c(W, H);
  c.kmap(img, W, 1, 1, ks, kf);
  DiscretizeMatrixCD2D(c.K, c.ny, c.nx, c.A, c.b, TL, TR, c.xc, c.yc);
  SolInitLinear(c.T, TR, TL, c.nx, c.ny);
  return SerialGS(c.A, c.b, c.T, c.qL, c.qR, c.K, nSweeps, -1.0, c.nx, c.ny, TL, TR, c.xc, c.yc);
}

}  // extern "C"
