// batch2d.cu -- batches of independent 2D images: one persistent CTA solves one image at a time.
//
// Replaces the reference batch loops (Keff2D/keff2D.cpp:341-471: OpenMP over images, each image
// assembled by DiscretizeMatrixCD2D and relaxed by SerialGS; Keff2DGPU/keff2D.cuh:1007-1168:
// BatchSim, a serial loop with cudaMalloc..cudaDeviceReset per image).  Here the whole solve of an
// image -- code bytes, initial field, diagonally preconditioned CG, convergence control, boundary
// flux, true residual -- runs inside one CTA with block-level reductions and no host round trip;
// CTAs pull images from an atomic counter, so the grid stays full whatever the per-image iteration
// counts are, and per-image results do not depend on the schedule.
#include "context.cuh"
#include "batch2d_onchip.cuh"
#include "batch2d_mg.cuh"

namespace kb {

Coef make_coef(int nx, int ny, int nzg, const keffb200_params& P);

// v1: CG vectors live in a per-CTA slice of global memory that stays L2-resident
// (148 CTAs x 4 vectors x 128 KiB = 74 MiB < 126 MB L2 for 128x128 images).
__global__ void __launch_bounds__(1024, 1)
    k_batch2d_pcg(const uint8_t* __restrict__ masks, int n_img, int nx, int ny, const Coef C, double thr2_in,
                  double flux_tol, long long max_iter, int const_init, double* __restrict__ ws,
                  uint8_t* __restrict__ codes, double* __restrict__ T_out, BatchOut* __restrict__ res, int* counter) {
  __shared__ double red[32 * 3];
  __shared__ double bc[4];
  __shared__ int s_img;
  const int n = nx * ny;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* x = ws + (size_t)blockIdx.x * 4 * n;
  double* r = x + n;
  double* p = r + n;
  double* q = p + n;
  uint8_t* code = codes + (size_t)blockIdx.x * n;

  while (true) {
    if (threadIdx.x == 0) s_img = atomicAdd(counter, 1);
    __syncthreads();
    const int img = s_img;
    __syncthreads();
    if (img >= n_img) break;
    const uint8_t* m = masks + (size_t)img * n;

    // structure -> code bytes, initial field
    for (int row = wrp; row < ny; row += nw)
      for (int xx = lane; xx < nx; xx += 32) {
        const int i = row * nx + xx;
        const unsigned me = m[i] == 1;
        unsigned c = me;
        if (xx > 0 && (unsigned)(m[i - 1] == 1) != me) c |= C_W;
        if (xx < nx - 1 && (unsigned)(m[i + 1] == 1) != me) c |= C_E;
        if (row < ny - 1 && (unsigned)(m[i + nx] == 1) != me) c |= C_S;
        if (row > 0 && (unsigned)(m[i - nx] == 1) != me) c |= C_N;
        code[i] = (uint8_t)c;
        x[i] = const_init ? C.tl : C.tl + (C.tr - C.tl) * ((xx + 0.5) / nx);
      }
    __syncthreads();

    // r = b - A x, p = z = r / a_P
    double a3[3] = {0, 0, 0};
    for (int row = wrp; row < ny; row += nw) {
      const bool y0 = row == 0, y1 = row == ny - 1;
      for (int xx = lane; xx < nx; xx += 32) {
        const int i = row * nx + xx;
        const bool x0 = xx == 0, x1 = xx == nx - 1;
        const unsigned c = code[i];
        const Faces F = faces_of(C, c, x0, x1, y0, y1, true, true);
        double qv = F.diag * x[i];
        if (!x0) qv -= F.w * x[i - 1];
        if (!x1) qv -= F.e * x[i + 1];
        if (!y0) qv -= F.n * x[i - nx];
        if (!y1) qv -= F.s * x[i + nx];
        const double b = rhs_of(C, c, x0, x1), rv = b - qv, z = rv / F.diag;
        r[i] = rv;
        p[i] = z;
        a3[0] += rv * z;
        a3[1] += rv * rv;
        a3[2] += b * b;
      }
    }
    block_allsum<3>(a3, red, bc);
    double rz = a3[0], rr = a3[1];
    const double bb = a3[2];
    double thr2 = thr2_in;
    long long it = 0;
    int status = 1;
    double ql = 0, qr = 0;

    while (true) {
      bool conv = !(rr > thr2 * bb);
      if (conv || it >= max_iter) {
        // boundary flux (Keff2D/keff2D.cpp:151-170); also the flux-balance part of the stop test
        double f2[2] = {0, 0};
        for (int row = threadIdx.x; row < ny; row += blockDim.x) {
          const int l = row * nx, e = l + nx - 1;
          f2[0] += ((code[l] & C_PH) ? C.wall[1] : C.wall[0]) * (x[l] - C.tl);
          f2[1] += ((code[e] & C_PH) ? C.wall[1] : C.wall[0]) * (C.tr - x[e]);
        }
        block_allsum<2>(f2, red, bc);
        ql = f2[0];
        qr = f2[1];
        const double keff = (ql + qr) / 2 / (C.tr - C.tl);
        if (conv && (fabs(ql - qr) <= flux_tol * fabs(keff) || thr2 < 1e-28)) {
          status = 0;
          break;
        }
        if (it >= max_iter) break;
        thr2 *= 1e-2;  // residual target met but wall fluxes still disagree: keep iterating
        continue;
      }
      // q = A p, p.q
      double a1[1] = {0};
      for (int row = wrp; row < ny; row += nw) {
        const bool y0 = row == 0, y1 = row == ny - 1;
        for (int xx = lane; xx < nx; xx += 32) {
          const int i = row * nx + xx;
          const bool x0 = xx == 0, x1 = xx == nx - 1;
          const Faces F = faces_of(C, code[i], x0, x1, y0, y1, true, true);
          const double pc = p[i];
          double qv = F.diag * pc;
          if (!x0) qv -= F.w * p[i - 1];
          if (!x1) qv -= F.e * p[i + 1];
          if (!y0) qv -= F.n * p[i - nx];
          if (!y1) qv -= F.s * p[i + nx];
          q[i] = qv;
          a1[0] += pc * qv;
        }
      }
      block_allsum<1>(a1, red, bc);
      const double alpha = rz / a1[0];
      // x += alpha p, r -= alpha q, r.z, r.r
      double a2[2] = {0, 0};
      for (int row = wrp; row < ny; row += nw) {
        const bool y0 = row == 0, y1 = row == ny - 1;
        for (int xx = lane; xx < nx; xx += 32) {
          const int i = row * nx + xx;
          const double d = diag_of(C, code[i], xx == 0, xx == nx - 1, y0, y1, true, true);
          x[i] += alpha * p[i];
          const double rv = r[i] - alpha * q[i];
          r[i] = rv;
          a2[0] += rv * (rv / d);
          a2[1] += rv * rv;
        }
      }
      block_allsum<2>(a2, red, bc);
      const double beta = a2[0] / rz;
      rz = a2[0];
      rr = a2[1];
      it++;
      // p = z + beta p
      for (int row = wrp; row < ny; row += nw) {
        const bool y0 = row == 0, y1 = row == ny - 1;
        for (int xx = lane; xx < nx; xx += 32) {
          const int i = row * nx + xx;
          const double d = diag_of(C, code[i], xx == 0, xx == nx - 1, y0, y1, true, true);
          p[i] = r[i] / d + beta * p[i];
        }
      }
      __syncthreads();
    }

    // true residual of the returned field + optional temperature map
    double t2[2] = {0, 0};
    for (int row = wrp; row < ny; row += nw) {
      const bool y0 = row == 0, y1 = row == ny - 1;
      for (int xx = lane; xx < nx; xx += 32) {
        const int i = row * nx + xx;
        const bool x0 = xx == 0, x1 = xx == nx - 1;
        const unsigned c = code[i];
        const Faces F = faces_of(C, c, x0, x1, y0, y1, true, true);
        double qv = F.diag * x[i];
        if (!x0) qv -= F.w * x[i - 1];
        if (!x1) qv -= F.e * x[i + 1];
        if (!y0) qv -= F.n * x[i - nx];
        if (!y1) qv -= F.s * x[i + nx];
        const double rv = rhs_of(C, c, x0, x1) - qv;
        t2[0] += rv * rv;
        t2[1] += fabs(rv);
        if (T_out) T_out[(size_t)img * n + i] = x[i];
      }
    }
    block_allsum<2>(t2, red, bc);
    if (threadIdx.x == 0) {
      BatchOut o;
      o.ql = ql;
      o.qr = qr;
      o.rr = t2[0];
      o.bb = bb;
      o.r1 = t2[1];
      o.it = it;
      o.status = status;
      o.pad = 0;
      res[img] = o;
    }
  }
}

// Enqueue the solve of images [first, first+count) of a device-resident batch on the compute stream
// (asynchronous).  Result records go to the device array slot of each image; counters[slot] is the
// work counter of this launch.
static bool batch2d_onchip_shape(int nx, int ny, const keffb200_params& P) {
  return !(P.flags & KEFFB200_F_BATCH_GLOBAL) && nx <= 128 && nx % 4 == 0 && ny <= 128 && ny % 8 == 0;
}
// on-chip multigrid kernel: sides multiples of 16 (four 2x coarsenings down to the dense level)
static bool batch2d_uses_mg(int nx, int ny, const keffb200_params& P) {
  static const bool no_mgk = getenv("KEFFB200_BATCH_NO_MG") != nullptr;
  return batch2d_onchip_shape(nx, ny, P) && !no_mgk && !(P.flags & KEFFB200_F_BATCH_ADDITIVE) && nx % 16 == 0 && ny % 16 == 0 &&
         nx >= 16 && ny >= 16;
}

// ready (nullable, multigrid kernel only): device counter of the images already resident (see batch2d_solve_host)
static int batch2d_enqueue(const uint8_t* d_solid, int first, int count, int nx, int ny, const keffb200_params& P,
                           double* d_T_out, double* ws, uint8_t* codes, BatchOut* res, int* counters, int slot,
                           const unsigned int* ready = nullptr) {
  Ctx& c = ctx();
  const size_t n = (size_t)nx * ny;
  const int grid = std::min(count, c.sms);
  const Coef C = make_coef(nx, ny, 1, P);
  const bool onchip = batch2d_onchip_shape(nx, ny, P);
  const uint8_t* m = d_solid + (size_t)first * n;
  double* T = d_T_out ? d_T_out + (size_t)first * n : nullptr;
  const bool mgk = batch2d_uses_mg(nx, ny, P);
  if (ready && !mgk) return fail(KEFFB200_ESTATE, "streamed batches need the on-chip multigrid kernel");
  if (onchip) {
    CoefF F;
    for (int ph = 0; ph < 2; ph++) {
      F.sx[ph] = (float)C.same[0][ph];
      F.sy[ph] = (float)C.same[1][ph];
      F.wall[ph] = (float)C.wall[ph];
    }
    F.dx = (float)C.diff[0];
    F.dy = (float)C.diff[1];
    if (!c.attr_batch) {  // function attributes are per device: the flag lives in the context
      KB_CUDA(cudaFuncSetAttribute(k_batch2d_onchip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OC_SMEM));
      KB_CUDA(cudaFuncSetAttribute(k_batch2d_mg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM));
      c.attr_batch = true;
    }
    if (mgk) {
      static const float om = getenv("KEFFB200_BATCH_OMEGA") ? (float)atof(getenv("KEFFB200_BATCH_OMEGA")) : 0.8f;
      // smoothing weights of the coarse levels (two sweeps per side); "w1,w2" overrides
      static float wc1 = 0.7f, wc2 = 1.4f;
      static bool wc_read = false;
      if (!wc_read) {
        if (const char* e = getenv("KEFFB200_BATCH_WC")) sscanf(e, "%f,%f", &wc1, &wc2);
        wc_read = true;
      }
      static const double delta2 = getenv("KEFFB200_BATCH_DELTA") ? atof(getenv("KEFFB200_BATCH_DELTA")) : 1e-4;
      k_batch2d_mg<<<grid, MB_THREADS, MB_SMEM, c.st>>>(m, count, nx, ny, C, F, P.rel_residual * P.rel_residual, P.flux_balance,
                                                        (long long)P.max_iter, (P.flags & KEFFB200_F_ZERO_INIT) ? 1 : 0, om, wc1, wc2, delta2,
                                                        ws, T, res + first, counters + slot, ready);
    } else
    k_batch2d_onchip<<<grid, OC_THREADS, OC_SMEM, c.st>>>(m, count, nx, ny, C, F, P.rel_residual * P.rel_residual,
                                                          P.flux_balance, (long long)P.max_iter,
                                                          (P.flags & KEFFB200_F_ZERO_INIT) ? 1 : 0, ws, T, res + first,
                                                          counters + slot);
  } else {
    // diagonally preconditioned CG leaves its error in the smooth modes: at ks:kf = 300 a relative residual of 1e-8
    // was seen 5e-5 away from the direct solution in Keff, so this kernel aims 100x lower than asked
    k_batch2d_pcg<<<grid, 1024, 0, c.st>>>(m, count, nx, ny, C, 1e-4 * P.rel_residual * P.rel_residual, P.flux_balance,
                                           (long long)P.max_iter, (P.flags & KEFFB200_F_ZERO_INIT) ? 1 : 0, ws, codes, T,
                                           res + first, counters + slot);
  }
  KB_LAUNCHED();
  KB_CUDA(cudaGetLastError());
  return 0;
}

struct BatchWs {
  double* ws;
  BatchOut* res;
  int* counters;
  uint8_t* codes;
};

static int batch2d_workspace(int n_img, int nx, int ny, BatchWs& w) {
  Ctx& c = ctx();
  const size_t n = (size_t)nx * ny;
  const int grid = std::min(n_img, c.sms);
  // per CTA: 4 FP64 vectors + code bytes (k_batch2d_pcg); k_batch2d_mg needs x + two 64 KB FP32 slices
  const size_t per_cta = std::max<size_t>(4 * n, mb_ws_stride(n)) * 8;
  const size_t need = (size_t)grid * (per_cta + n) + (size_t)n_img * sizeof(BatchOut) + 256;
  if (need > c.cap_batch) {
    if (c.batch_ws) cudaFree(c.batch_ws);
    c.batch_ws = nullptr;
    c.cap_batch = 0;
    KB_CUDA(cudaMalloc(&c.batch_ws, need));
    c.cap_batch = need;
  }
  char* base = (char*)c.batch_ws;
  w.ws = (double*)base;
  w.res = (BatchOut*)(base + (size_t)grid * per_cta);
  w.counters = (int*)(w.res + n_img);
  w.codes = (uint8_t*)(w.counters + 16);
  return 0;
}

static int batch2d_collect(const BatchWs& w, int n_img, const keffb200_params& P, keffb200_result* out) {
  Ctx& c = ctx();
  KB_CUDA(cudaEventRecord(c.ev1, c.st));
  std::vector<BatchOut> h(n_img);
  KB_CUDA(cudaMemcpyAsync(h.data(), w.res, (size_t)n_img * sizeof(BatchOut), cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  float ms = 0;
  KB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  for (int i = 0; i < n_img; i++)
    if (h[i].status == -1 && h[i].pad == -1)
      return fail(KEFFB200_ECUDA, "image %d was never solved: its structure did not reach the GPU (host copy failed or stalled)", i);
  for (int i = 0; i < n_img; i++) {
    keffb200_result& o = out[i];
    o.q_left = h[i].ql;
    o.q_right = h[i].qr;
    o.keff = (h[i].ql + h[i].qr) / 2 / (P.tr - P.tl);
    o.rel_residual = h[i].bb > 0 ? sqrt(h[i].rr / h[i].bb) : 0.0;
    o.energy_residual = h[i].r1;
    o.iters = (long)h[i].it;
    o.status = h[i].status;
    o.gpu_ms = ms;  // whole-batch device time
  }
  return 0;
}

// An on-chip mixed-precision solve that reports a breakdown (status 2: the FP32 recurrence left the positive
// cone; seen with a reliable-update window of 1e-5 at conductivity ratios ~ 1e4) is re-solved by the all-FP64
// kernel, so a caller never receives a field that did not converge without being told (status stays 1 or 2 only if
// the FP64 solve fails as well).
static int batch2d_rescue(const uint8_t* d_solid, int n_img, int nx, int ny, const keffb200_params& P, double* d_T_out,
                          const BatchWs& w, keffb200_result* out) {
  Ctx& c = ctx();
  std::vector<int> bad;
  for (int i = 0; i < n_img; i++)
    if (out[i].status == 2 || !std::isfinite(out[i].keff)) bad.push_back(i);
  if (bad.empty() || (P.flags & KEFFB200_F_BATCH_GLOBAL)) return 0;
  keffb200_params Q = P;
  Q.flags |= KEFFB200_F_BATCH_GLOBAL;
  for (int i : bad) {
    KB_CUDA(cudaMemsetAsync(w.counters, 0, 16 * sizeof(int), c.st));
    KB_TRY(batch2d_enqueue(d_solid, i, 1, nx, ny, Q, d_T_out, w.ws, w.codes, w.res, w.counters, 0));
  }
  std::vector<BatchOut> h(n_img);
  KB_CUDA(cudaMemcpyAsync(h.data(), w.res, (size_t)n_img * sizeof(BatchOut), cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  for (int i : bad) {
    keffb200_result& o = out[i];
    o.q_left = h[i].ql;
    o.q_right = h[i].qr;
    o.keff = (h[i].ql + h[i].qr) / 2 / (P.tr - P.tl);
    o.rel_residual = h[i].bb > 0 ? sqrt(h[i].rr / h[i].bb) : 0.0;
    o.energy_residual = h[i].r1;
    o.iters += (long)h[i].it;
    o.status = h[i].status;
  }
  return 0;
}

static int batch2d_check(int nx, int ny, const keffb200_params& P) {
  if (nx < 2 || ny < 2) return fail(KEFFB200_EINVAL, "image %dx%d too small", nx, ny);
  if (P.tr == P.tl) return fail(KEFFB200_EINVAL, "TR == TL: Keff undefined");
  if (!(P.ks > 0) || !(P.kf > 0)) return fail(KEFFB200_EINVAL, "conductivities must be positive");
  return 0;
}

int batch2d_solve_dev(const uint8_t* d_solid, int n_img, int nx, int ny, const keffb200_params& P, double* d_T_out,
                      keffb200_result* out) {
  Ctx& c = ctx();
  KB_TRY(batch2d_check(nx, ny, P));
  BatchWs w;
  KB_TRY(batch2d_workspace(n_img, nx, ny, w));
  KB_CUDA(cudaEventRecord(c.ev0, c.st));
  KB_CUDA(cudaMemsetAsync(w.counters, 0, 16 * sizeof(int), c.st));
  KB_TRY(batch2d_enqueue(d_solid, 0, n_img, nx, ny, P, d_T_out, w.ws, w.codes, w.res, w.counters, 0));
  KB_TRY(batch2d_collect(w, n_img, P, out));
  return batch2d_rescue(d_solid, n_img, nx, ny, P, d_T_out, w, out);
}

// Host structures.  On-chip multigrid kernel: ONE launch over the whole batch is enqueued first and the structures
// stream in behind it on a second stream, chunk by chunk, each chunk followed by a 4-byte copy that publishes how many
// images are resident; a CTA that draws an image which has not arrived yet waits on that counter.  Only the first
// chunk (1/32 of the batch) is exposed, there is one kernel tail instead of one per piece, and a caller with pageable
// memory -- whose cudaMemcpyAsync blocks the host chunk by chunk -- still overlaps its copy with the solve.
// Other kernels: three pieces (1/16, 3/16, 3/4), each piece's launch waits only for its own copy.
// d_stage: device staging buffer for the whole batch.
int batch2d_solve_host(const uint8_t* h_solid, uint8_t* d_stage, int n_img, int nx, int ny, const keffb200_params& P,
                       double* d_T_out, keffb200_result* out) {
  Ctx& c = ctx();
  KB_TRY(batch2d_check(nx, ny, P));
  BatchWs w;
  KB_TRY(batch2d_workspace(n_img, nx, ny, w));
  const size_t n = (size_t)nx * ny;
  KB_CUDA(cudaEventRecord(c.ev0, c.st));
  KB_CUDA(cudaMemsetAsync(w.counters, 0, 16 * sizeof(int), c.st));
  KB_CUDA(cudaEventRecord(c.evx, c.st));
  KB_CUDA(cudaStreamWaitEvent(c.st2, c.evx, 0));  // staging buffer free (previous call's kernels are done), counters zeroed
  // experiments: KEFFB200_BATCH_PIECES keeps the three-piece scheme; a launch that is synchronised at once
  // (KEFFB200_DEBUG_SYNC) must not wait for copies that are enqueued after it
  static const bool no_stream = getenv("KEFFB200_BATCH_PIECES") != nullptr || getenv("KEFFB200_DEBUG_SYNC") != nullptr;
  if (batch2d_uses_mg(nx, ny, P) && c.h_ready && !no_stream) {
    unsigned int* ready = reinterpret_cast<unsigned int*>(w.counters + 8);
    // a record that keeps this pattern was never written: its image did not arrive within the kernel's wait limit
    KB_CUDA(cudaMemsetAsync(w.res, 0xff, (size_t)n_img * sizeof(BatchOut), c.st));
    KB_TRY(batch2d_enqueue(d_stage, 0, n_img, nx, ny, P, d_T_out, w.ws, w.codes, w.res, w.counters, 0, ready));
    const int nchunk = std::max(1, std::min(KB_READY_SLOTS, n_img / std::max(1, c.sms)));
    for (int k = 0; k < nchunk; k++) {
      const int first = (int)((long long)n_img * k / nchunk), last = (int)((long long)n_img * (k + 1) / nchunk);
      if (last <= first) continue;
      KB_CUDA(cudaMemcpyAsync(d_stage + (size_t)first * n, h_solid + (size_t)first * n, (size_t)(last - first) * n,
                              cudaMemcpyHostToDevice, c.st2));
      c.h_ready[k] = (unsigned int)last;
      KB_CUDA(cudaMemcpyAsync(ready, &c.h_ready[k], sizeof(unsigned int), cudaMemcpyHostToDevice, c.st2));
    }
    KB_TRY(batch2d_collect(w, n_img, P, out));
    return batch2d_rescue(d_stage, n_img, nx, ny, P, d_T_out, w, out);
  }
  // pieces of 1/16, 3/16 and 3/4 of the batch: only the first copy (1/16) is exposed, and every later copy is
  // shorter than the compute it hides behind (a 128x128 image takes ~6x longer to solve than to copy)
  int cut[4] = {0, n_img / 16, n_img / 4, n_img};
  if (n_img < 8 * c.sms) cut[1] = cut[2] = 0;  // small batches: one piece
  for (int k = 0; k < 3; k++) {
    const int first = cut[k], count = cut[k + 1] - cut[k];
    if (count <= 0) continue;
    KB_CUDA(cudaMemcpyAsync(d_stage + (size_t)first * n, h_solid + (size_t)first * n, (size_t)count * n,
                            cudaMemcpyHostToDevice, c.st2));
    KB_CUDA(cudaEventRecord(c.evp[k], c.st2));
  }
  for (int k = 0; k < 3; k++) {
    const int first = cut[k], count = cut[k + 1] - cut[k];
    if (count <= 0) continue;
    KB_CUDA(cudaStreamWaitEvent(c.st, c.evp[k], 0));
    KB_TRY(batch2d_enqueue(d_stage, first, count, nx, ny, P, d_T_out, w.ws, w.codes, w.res, w.counters, k));
  }
  KB_TRY(batch2d_collect(w, n_img, P, out));
  return batch2d_rescue(d_stage, n_img, nx, ny, P, d_T_out, w, out);
}

}  // namespace kb
