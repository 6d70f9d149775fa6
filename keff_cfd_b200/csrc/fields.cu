// fields.cu -- field-level helpers either side of the solve: the x-direction heat-flux field behind the
// Q-maps, and the interpolation of a converged field onto a finer mesh (warm start of the mesh-convergence mode).
//
// Replaces printQMAP's flux evaluation (Keff2D/keff2D.h:427-479: harmonic-mean face flux between x-neighbours,
// the wall fluxes at both ends; the reference never finished the 3D one, Keff3D/keff3D.cpp:179-181) and
// SolExpandSimple (Keff2D/keff2D.h:614-642, Keff3D: same with a z loop), which replicates the coarse field
// without interpolating.
#include "context.cuh"

namespace kb {

// Q[z][y][j], j = 0..nx: flux through the face between cell j-1 and cell j (j = 0 / nx: the two walls), positive
// in +x.  One thread per face; T and the structure are read through L1 (each value twice).
__global__ void __launch_bounds__(256) k_flux_field(const double* __restrict__ T, const uint8_t* __restrict__ solid, int nx,
                                                    long rows, double cs, double cf, double cd, double ws, double wf,
                                                    double tl, double tr, double* __restrict__ Q) {
  const long faces = rows * (nx + 1);
  for (long f = blockIdx.x * (long)blockDim.x + threadIdx.x; f < faces; f += (long)gridDim.x * blockDim.x) {
    const long row = f / (nx + 1);
    const int j = (int)(f - row * (nx + 1));
    const long base = row * nx;
    double q;
    if (j == 0) q = (solid[base] == 1 ? ws : wf) * (T[base] - tl);
    else if (j == nx) q = (solid[base + nx - 1] == 1 ? ws : wf) * (tr - T[base + nx - 1]);
    else {
      const bool a = solid[base + j - 1] == 1, b = solid[base + j] == 1;
      q = (a != b ? cd : (a ? cs : cf)) * (T[base + j] - T[base + j - 1]);
    }
    Q[f] = q;
  }
}

// Cell-centred (tri)linear interpolation of a field on an sx*sy*sz mesh of the unit cube onto an nx*ny*nz mesh.
// Along x the two walls are interpolation nodes (T = tl at x = 0, tr at x = 1: the fixed-temperature faces of
// Keff2D/keff2D.h:529-544); along y and z the field is extended with zero gradient (adiabatic faces).
__device__ __forceinline__ void interp_axis(int i, int n, int s, int& i0, int& i1, double& w) {
  const double t = (i + 0.5) * s / n - 0.5;  // position in units of source cells, relative to source centre 0
  const double fl = floor(t);
  i0 = (int)fl;
  w = t - fl;
  i1 = i0 + 1;
  if (i0 < 0) { i0 = 0; i1 = 0; w = 0; }
  if (i1 > s - 1) { i1 = s - 1; if (i0 > s - 1) i0 = s - 1; }
}
__global__ void __launch_bounds__(256) k_interpolate(const double* __restrict__ src, int sx, int sy, int sz, double* __restrict__ dst,
                                                     int nx, int ny, int nz, double tl, double tr) {
  const long n = (long)nx * ny * nz;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx);
    const long t = i / nx;
    const int y = (int)(t % ny), z = (int)(t / ny);
    int y0, y1, z0, z1;
    double wy, wz;
    interp_axis(y, ny, sy, y0, y1, wy);
    interp_axis(z, nz, sz, z0, z1, wz);
    // x: nodes are the wall (x = 0), the source centres, the wall (x = 1)
    const double tx = (x + 0.5) * sx / nx - 0.5;
    const double fl = floor(tx);
    const int x0 = (int)fl;
    const double wx = tx - fl;
    auto row = [&](int yy, int zz) -> double {
      const double* r = src + ((long)zz * sy + yy) * sx;
      if (x0 < 0) return tl + (r[0] - tl) * (2.0 * (tx + 0.5));               // between the left wall and centre 0
      if (x0 >= sx - 1) return r[sx - 1] + (tr - r[sx - 1]) * (2.0 * (tx - (sx - 1)));  // last centre .. right wall
      return r[x0] + (r[x0 + 1] - r[x0]) * wx;
    };
    const double a = row(y0, z0) * (1 - wy) + row(y1, z0) * wy;
    const double b = row(y0, z1) * (1 - wy) + row(y1, z1) * wy;
    dst[i] = a * (1 - wz) + b * wz;
  }
}

}  // namespace kb

using namespace kb;

extern "C" {

int keffb200_flux_field(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* T, double* Qx) {
  Ctx& c = ctx();
  if (!c.ready) return fail(KEFFB200_ESTATE, "keffb200_init has not been called");
  if (!solid || !T || !Qx || nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "bad argument");
  const keffb200_params P = norm_params(p);
  const Coef C = make_coef(nx, ny, nz, P);
  const long rows = (long)ny * nz;
  const size_t n = (size_t)rows * nx, nq = (size_t)rows * (nx + 1);
  double *dT = nullptr, *dQ = nullptr;
  uint8_t* dS = nullptr;
  int rc = 0;
  if (cudaMalloc(&dT, n * 8) != cudaSuccess || cudaMalloc(&dQ, nq * 8) != cudaSuccess || cudaMalloc(&dS, n) != cudaSuccess)
    rc = fail(KEFFB200_ENOMEM, "cannot allocate the flux field (%zu cells)", n);
  if (rc == 0) {
    cudaMemcpyAsync(dT, T, n * 8, cudaMemcpyHostToDevice, c.st);
    cudaMemcpyAsync(dS, solid, n, cudaMemcpyHostToDevice, c.st);
    const int grid = (int)std::min<long>((long)c.sms * 8, (long)((nq + 255) / 256));
    k_flux_field<<<grid, 256, 0, c.st>>>(dT, dS, nx, rows, C.same[0][1], C.same[0][0], C.diff[0], C.wall[1], C.wall[0], P.tl, P.tr, dQ);
    KB_LAUNCHED();
    cudaMemcpyAsync(Qx, dQ, nq * 8, cudaMemcpyDeviceToHost, c.st);
    const cudaError_t e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) rc = fail(KEFFB200_ECUDA, "flux field: %s", cudaGetErrorString(e));
  }
  cudaFree(dT);
  cudaFree(dQ);
  cudaFree(dS);
  return rc;
}

int keffb200_interpolate(const double* T_src, int sx, int sy, int sz, double* T_dst, int nx, int ny, int nz, double tl, double tr) {
  Ctx& c = ctx();
  if (!c.ready) return fail(KEFFB200_ESTATE, "keffb200_init has not been called");
  if (!T_src || !T_dst || sx < 1 || sy < 1 || sz < 1 || nx < 1 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "bad argument");
  const size_t ns = (size_t)sx * sy * sz, n = (size_t)nx * ny * nz;
  double *ds = nullptr, *dd = nullptr;
  int rc = 0;
  if (cudaMalloc(&ds, ns * 8) != cudaSuccess || cudaMalloc(&dd, n * 8) != cudaSuccess)
    rc = fail(KEFFB200_ENOMEM, "cannot allocate the interpolation buffers (%zu cells)", n);
  if (rc == 0) {
    cudaMemcpyAsync(ds, T_src, ns * 8, cudaMemcpyHostToDevice, c.st);
    const int grid = (int)std::min<long>((long)c.sms * 8, (long)((n + 255) / 256));
    k_interpolate<<<grid, 256, 0, c.st>>>(ds, sx, sy, sz, dd, nx, ny, nz, tl, tr);
    KB_LAUNCHED();
    cudaMemcpyAsync(T_dst, dd, n * 8, cudaMemcpyDeviceToHost, c.st);
    const cudaError_t e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) rc = fail(KEFFB200_ECUDA, "interpolate: %s", cudaGetErrorString(e));
  }
  cudaFree(ds);
  cudaFree(dd);
  return rc;
}

// Page-locked host memory for structures / fields a caller hands to the host-buffer entry points: copies from it
// run at full PCIe rate and overlap with the solves (pageable memory works too, slower).
void* keffb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void keffb200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
