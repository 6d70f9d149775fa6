// solver.cu -- host driver of the grid solver (diagonally preconditioned CG on the matrix-free
// finite-volume operator), the z-slab multi-process variant, and the extern "C" ABI.
//
// Replaces, per solve: DiscretizeMatrixCD2D/3D + SolInitLinear + ParallelGS/SerialGS/ParallelGS3D
// + the flux loops of main() (Keff2D/keff2D.cpp:131-170, Keff3D/keff3D.cpp:134-163) and
// JacobiGPU (Keff2DGPU/keff2D.cuh:726-849).  All convergence control is on the device; the host
// only reads a 150-byte scalar record every `check_every` iterations.
#include "context.cuh"
#include "kernels.cuh"
#include "mg.cuh"
#include "peer.cuh"

#include <thread>

namespace kb {

// batch2d.cu
int batch2d_solve_dev(const uint8_t* d_solid, int n_img, int nx, int ny, const keffb200_params& P, double* d_T_out,
                      keffb200_result* out);
int batch2d_solve_host(const uint8_t* h_solid, uint8_t* d_stage, int n_img, int nx, int ny, const keffb200_params& P,
                       double* d_T_out, keffb200_result* out);

keffb200_params norm_params(const keffb200_params* p) {
  keffb200_params P;
  if (p) P = *p; else keffb200_default_params(&P);
  if (P.check_every <= 0) P.check_every = 32;
  if (P.max_iter <= 0) P.max_iter = 500000;
  if (!(P.convergence > 0)) P.convergence = 1e-5;
  // The reference's only stop control is `Convergence` (relative change of Keff between two checks,
  // Keff2D/keff2D.h:825-847).  Here a check happens once the residual target is met and the field is frozen from
  // then on, so the key acts through the two targets it implies: at the shipped 1e-5 they are the defaults
  // (1e-8, 1e-6); a tighter Convergence tightens both in proportion, a looser one never loosens them.
  if (!(P.rel_residual > 0)) P.rel_residual = std::max(1e-13, std::min(1e-8, 1e-3 * P.convergence));
  if (!(P.flux_balance > 0)) P.flux_balance = std::max(1e-11, std::min(1e-6, 1e-1 * P.convergence));
  return P;
}

Coef make_coef(int nx, int ny, int nzg, const keffb200_params& P) {
  // face conductances in the reference's own operation order (Keff3D/keff3D.h:698-757; with
  // dz = 1 they reduce to the 2D expressions of Keff2D/keff2D.h:529-577)
  const double dx = 1.0 / nx, dy = 1.0 / ny, dz = nzg > 1 ? 1.0 / nzg : 1.0;
  const double k[2] = {P.kf, P.ks};
  auto hmean = [](double w, double a, double b) { return (w + w) / (w / a + w / b); };
  Coef C;
  for (int ph = 0; ph < 2; ph++) {
    C.same[0][ph] = k[ph] * dy * dz / dx;
    C.same[1][ph] = k[ph] * dx * dz / dy;
    C.same[2][ph] = k[ph] * dy * dx / dz;
    C.wall[ph] = k[ph] * dy * dz / (dx / 2);
  }
  C.diff[0] = hmean(dx / 2, P.ks, P.kf) * dy * dz / dx;
  C.diff[1] = hmean(dy / 2, P.ks, P.kf) * dx * dz / dy;
  C.diff[2] = hmean(dz / 2, P.ks, P.kf) * dy * dx / dz;
  C.tl = P.tl;
  C.tr = P.tr;
  return C;
}

static int mg_pitch(int nx) { return (nx + 3) / 4 * 4; }

static int ensure_ws(const Dom& D) {
  Ctx& c = ctx();
  const size_t plane = (size_t)D.nx * D.ny, need = plane * (D.nz + 2);
  if (need > c.cap_field) {
    link_release_window(c, W_X, c.x);
    link_release_window(c, W_P, c.p);
    c.x = c.p = nullptr;
    for (double** f : {&c.r, &c.q}) {
      if (*f) cudaFree(*f);
      *f = nullptr;
    }
    if (c.code) cudaFree(c.code);
    if (c.mask) cudaFree(c.mask);
    c.code = c.mask = nullptr;
    c.cap_field = 0;
    for (double** f : {&c.x, &c.r, &c.p, &c.q}) KB_CUDA(cudaMalloc(f, need * sizeof(double)));
    KB_CUDA(cudaMalloc(&c.code, need));
    KB_CUDA(cudaMalloc(&c.mask, need));
    c.cap_field = need;
  }
  if (plane * D.nz > c.cap_z) {
    if (c.z) cudaFree(c.z);
    c.z = nullptr;
    c.cap_z = 0;
    KB_CUDA(cudaMalloc(&c.z, plane * D.nz * sizeof(float)));
    c.cap_z = plane * D.nz;
  }
  const size_t need_u = (size_t)mg_pitch(D.nx) * D.ny * (D.nz + 2);
  if (need_u > c.cap_u) {
    link_release_window(c, W_U, c.u);
    c.u = nullptr;
    c.cap_u = 0;
    KB_CUDA(cudaMalloc(&c.u, need_u * sizeof(float)));
    c.cap_u = need_u;
  }
  const size_t rows = (size_t)D.ny * D.nz;
  if (rows > c.cap_aux) {
    if (c.qL) cudaFree(c.qL);
    if (c.qR) cudaFree(c.qR);
    c.qL = c.qR = nullptr;
    c.cap_aux = 0;
    KB_CUDA(cudaMalloc(&c.qL, rows * sizeof(double)));
    KB_CUDA(cudaMalloc(&c.qR, rows * sizeof(double)));
    c.cap_aux = rows;
  }
  return 0;
}

// ---- phase trace (KEFFB200_TRACE=k: CUDA events around the phases of iteration k of the grid solver, printed to
// stderr at the end of the solve; 0 / unset = off).  Diagnostic only: an event per phase, nothing else changes.
struct Trace {
  int want = -1;        // iteration to trace (-1 = off)
  bool live = false;    // marks are being recorded right now
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> name;
};
static Trace& trace() {
  static thread_local Trace t;
  static const int want = getenv("KEFFB200_TRACE") ? atoi(getenv("KEFFB200_TRACE")) : -1;
  t.want = want;
  return t;
}
static void trace_mark(const char* what) {
  Trace& t = trace();
  if (!t.live) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx().st);
  t.ev.push_back(e);
  t.name.push_back(what);
}
static void trace_report(int rank) {
  Trace& t = trace();
  if (t.ev.size() < 2) return;
  cudaEventSynchronize(t.ev.back());
  std::string line = "[keffb200 trace] rank " + std::to_string(rank) + " iteration " + std::to_string(t.want) + ":";
  float total = 0;
  for (size_t i = 1; i < t.ev.size(); i++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t.ev[i - 1], t.ev[i]);
    total += ms;
    char b[96];
    snprintf(b, sizeof b, " %s %.0f", t.name[i], ms * 1e3);
    line += b;
  }
  char b[64];
  snprintf(b, sizeof b, " | total %.0f us", total * 1e3);
  line += b;
  line += "\n";
  if (write(2, line.data(), line.size()) < 0) {  // one write: lines of several ranks must not interleave
  }
  for (cudaEvent_t e : t.ev) cudaEventDestroy(e);
  t.ev.clear();
  t.name.clear();
}

// ---- NCCL helpers (no-ops when world == 1) -----------------------------------------------------
static int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  Nccl& n = ctx().nccl;
  return fail(KEFFB200_ENCCL, "%s: %s", what, n.GetErrorString ? n.GetErrorString(rc) : "nccl error");
}

static int allreduce_tmp(double* dptr, int count) {
  Nccl& n = ctx().nccl;
  if (n.world <= 1 || !n.comm || !ctx().dist_active || ctx().link.p2p) return 0;
  return nccl_check(n.AllReduce(dptr, dptr, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, n.comm, ctx().st), "ncclAllReduce");
}

static ArLink ar_link() {
  Ctx& c = ctx();
  ArLink a;
  memset(&a, 0, sizeof a);
  a.world = 1;
  if (c.dist_active && c.link.p2p) {
    a.world = c.link.world;
    a.rank = c.link.rank;
    a.seq = c.link.ar_seq;
    for (int r = 0; r < c.link.world; r++) a.mail[r] = c.link.mail[r];
  }
  return a;
}

// sum of the per-CTA partials of the kernel just launched + all-reduce over the slab ranks + scalar step:
// one launch (single GPU, peer memory), or sum / ncclAllReduce / step when the ranks only share NCCL
template <int WHAT>
static int reduce_step(int n, int arg) {
  Ctx& c = ctx();
  constexpr int nv = WHAT == RS_INIT ? 3 : (WHAT == RS_ITER || WHAT == RS_RESID || WHAT == RS_FLUX) ? 2 : 1;
  const bool via_nccl = c.dist_active && c.nccl.world > 1 && c.nccl.comm && !c.link.p2p;
  k_reduce_step<WHAT><<<1, 512, 0, c.st>>>(c.partials, n, c.S, arg, via_nccl ? 1 : 0, ar_link());
  KB_LAUNCHED();
  if (via_nccl) {
    KB_TRY(allreduce_tmp(c.S->tmp, nv));
    k_reduce_step<WHAT><<<1, 512, 0, c.st>>>(c.partials, n, c.S, arg, 2, ar_link());
    KB_LAUNCHED();
  }
  return 0;
}

// exchange the boundary planes of a halo'ed fine-grid field (x, or p: `esz` bytes per value -- p is FP32 on the
// multigrid path) with the z neighbours (f points at the halo plane)
static int exchange_halo(void* f_halo_v, const Dom& D, cudaStream_t st, size_t esz = 8) {
  Ctx& c = ctx();
  Nccl& n = c.nccl;
  if (!c.dist_active) return 0;
  const size_t pb = (size_t)D.nx * D.ny * esz;  // bytes per plane
  char* f_halo = (char*)f_halo_v;
  if (c.link.p2p) {
    const bool is_x = f_halo_v == (void*)c.x;
    return link_exchange(c, is_x ? HF_X : HF_P, is_x ? W_X : W_P, f_halo + pb, pb, D.nz, true);
  }
  if (n.world <= 1 || !n.comm) return 0;
  char* lo_halo = f_halo;
  char* first = f_halo + pb;
  char* last = f_halo + pb * D.nz;
  char* hi_halo = f_halo + pb * (D.nz + 1);
  KB_TRY(nccl_check(n.GroupStart(), "ncclGroupStart"));
  if (n.rank > 0) {
    KB_TRY(nccl_check(n.Send(first, pb, /*ncclInt8*/ 0, n.rank - 1, n.comm, st), "ncclSend"));
    KB_TRY(nccl_check(n.Recv(lo_halo, pb, 0, n.rank - 1, n.comm, st), "ncclRecv"));
  }
  if (n.rank < n.world - 1) {
    KB_TRY(nccl_check(n.Send(last, pb, 0, n.rank + 1, n.comm, st), "ncclSend"));
    KB_TRY(nccl_check(n.Recv(hi_halo, pb, 0, n.rank + 1, n.comm, st), "ncclRecv"));
  }
  KB_TRY(nccl_check(n.GroupEnd(), "ncclGroupEnd"));
  return 0;
}

// same for any plane-major array inside window `win` of the workspace: plane0 = local plane 0, halo planes at
// -1 and nz; `field` names its arrival counters
static int exchange_planes(int field, int win, void* plane0, size_t plane_bytes, int nz, cudaStream_t st) {
  Ctx& c = ctx();
  Nccl& n = c.nccl;
  if (!c.dist_active) return 0;
  if (c.link.p2p) return link_exchange(c, field, win, (char*)plane0, plane_bytes, nz);
  if (n.world <= 1 || !n.comm) return 0;
  char* b = (char*)plane0;
  KB_TRY(nccl_check(n.GroupStart(), "ncclGroupStart"));
  if (n.rank > 0) {
    KB_TRY(nccl_check(n.Send(b, plane_bytes, /*ncclInt8*/ 0, n.rank - 1, n.comm, st), "ncclSend"));
    KB_TRY(nccl_check(n.Recv(b - plane_bytes, plane_bytes, 0, n.rank - 1, n.comm, st), "ncclRecv"));
  }
  if (n.rank < n.world - 1) {
    KB_TRY(nccl_check(n.Send(b + (size_t)(nz - 1) * plane_bytes, plane_bytes, 0, n.rank + 1, n.comm, st), "ncclSend"));
    KB_TRY(nccl_check(n.Recv(b + (size_t)nz * plane_bytes, plane_bytes, 0, n.rank + 1, n.comm, st), "ncclRecv"));
  }
  KB_TRY(nccl_check(n.GroupEnd(), "ncclGroupEnd"));
  return 0;
}
static int level_field(int level, int which) { return HF_LEVEL0 + 6 * level + which; }

// complete a replicated array of the level pool: src = my slice inside my copy dst
static int allgather_f32(const float* src, float* dst, size_t count) {
  Ctx& c = ctx();
  if (c.link.p2p) return link_gather(c, (const char*)src, count * sizeof(float));
  Nccl& n = c.nccl;
  return nccl_check(n.AllGather(src, dst, count, /*ncclFloat32*/ 7, n.comm, c.st), "ncclAllGather");
}

// ---- halo planes pushed by the producing kernel itself -------------------------------------------------
// The producers of p, u and the two iterate fields of the voxel grid (20 of the 22 MB a rank sends per iteration at
// 1024^2 planes) store their slab-boundary planes into the neighbours' halo planes themselves, first thing, and the
// transfer rides under the rest of the kernel; the consumer side is a one-thread wait on the arrival counter.
// All these arrays are halo'ed buffers that start at their window's base: plane -1 at offset 0, plane nz at nz + 1.
struct Fused {
  bool on = false;
  int field = 0;
  HaloPush H{nullptr, nullptr, nullptr, nullptr};
};
static Fused fused_push(int field, int win, size_t plane_bytes, int nz) {
  Ctx& c = ctx();
  Fused f;
  f.field = field;
  static const bool off = getenv("KEFFB200_SLAB_NOFUSE") != nullptr;  // experiments: separate push kernels
  if (!c.dist_active || !c.link.p2p || nz < 2 || off) return f;
  Link& L = c.link;
  const bool lo = L.rank > 0, hi = L.rank < L.world - 1;
  if ((lo && !L.nb[0][win]) || (hi && !L.nb[1][win])) return f;
  f.on = true;
  f.H.lo_dst = lo ? L.nb[0][win] + (size_t)(nz + 1) * plane_bytes : nullptr;
  f.H.hi_dst = hi ? L.nb[1][win] : nullptr;
  f.H.lo_flag = lo ? &L.mail[L.rank - 1]->halo[field][1] : nullptr;
  f.H.hi_flag = hi ? &L.mail[L.rank + 1]->halo[field][0] : nullptr;
  return f;
}
// after the producer's launch: `units` arrivals per side are on their way into my halos as well
static int fused_wait(const Fused& f, unsigned units) {
  if (!f.on) return 0;
  Ctx& c = ctx();
  Link& L = c.link;
  Mail* mine = L.mail[L.rank];
  const bool lo = L.rank > 0, hi = L.rank < L.world - 1;
  if (lo) L.expect[f.field][0] += units;
  if (hi) L.expect[f.field][1] += units;
  k_wait_halo<<<1, 32, 0, c.st>>>(lo ? &mine->halo[f.field][0] : nullptr, L.expect[f.field][0], hi ? &mine->halo[f.field][1] : nullptr,
                                  L.expect[f.field][1], c.S);
  KB_LAUNCHED();
  return 0;
}

// ---- stencil launcher ---------------------------------------------------------------------------
struct Stencil {
  bool tma = false;
  bool ws = false;       // warp-specialised kernel (code bytes by TMA: nx % 16 == 0)
  CUtensorMap tm_code;
  Plan plan{};
  int grid = 0;
  Dom D{};
  Coef C{};
  CUtensorMap tm_x, tm_p;  // descriptors for the two fields that are ever stencil inputs
  CUtensorMap tm_p32;      // p stored in FP32 (multigrid path; same buffer as p)
  bool p32_ok = false;     // tm_p32 is valid (TMA path, nx % 4 == 0: FP32 rows must be multiples of 16 bytes)
};

static int make_tmap(CUtensorMap* tm, double* field_halo, const Dom& D) {
  Ctx& c = ctx();
  cuuint64_t dims[3] = {(cuuint64_t)D.nx, (cuuint64_t)D.ny, (cuuint64_t)D.nz + 2};
  cuuint64_t strides[2] = {(cuuint64_t)D.nx * 8, (cuuint64_t)D.nx * D.ny * 8};
  cuuint32_t box[3] = {BOXW, BOXH, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = c.encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, field_halo, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(KEFFB200_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// FP32 view of a halo'ed field buffer (p on the multigrid path): box of TX + 8 floats starting 4 cells left of the tile
static int make_tmap_f32(CUtensorMap* tm, float* field_halo, const Dom& D) {
  Ctx& c = ctx();
  cuuint64_t dims[3] = {(cuuint64_t)D.nx, (cuuint64_t)D.ny, (cuuint64_t)D.nz + 2};
  cuuint64_t strides[2] = {(cuuint64_t)D.nx * 4, (cuuint64_t)D.nx * D.ny * 4};
  cuuint32_t box[3] = {(cuuint32_t)StencilTile<float>::boxw, BOXH, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = c.encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, field_halo, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(KEFFB200_ECUDA, "cuTensorMapEncodeTiled (FP32 field) failed (%d)", (int)r);
  return 0;
}

static int make_tmap_code(CUtensorMap* tm, uint8_t* code, const Dom& D) {
  Ctx& c = ctx();
  cuuint64_t dims[3] = {(cuuint64_t)D.nx, (cuuint64_t)D.ny, (cuuint64_t)D.nz};
  cuuint64_t strides[2] = {(cuuint64_t)D.nx, (cuuint64_t)D.nx * D.ny};
  cuuint32_t box[3] = {TX, TY, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = c.encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, code, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(KEFFB200_ECUDA, "cuTensorMapEncodeTiled (code bytes) failed (%d)", (int)r);
  return 0;
}

template <int MODE>
static int ws_prepare() {  // function attributes are per device: applied on every setup (cheap)
  KB_CUDA(cudaFuncSetAttribute(k_stencil_ws<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
  return 0;
}

template <int MODE>
static int occupancy_tma() {
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stencil_tma<MODE>, NTHR, 0);
  return std::max(occ, 1);
}

static int stencil_setup(Stencil& s, const Dom& D, const Coef& C, int flags) {
  Ctx& c = ctx();
  s.D = D;
  s.C = C;
  s.tma = !(flags & KEFFB200_F_NO_TMA) && (D.nx % 2 == 0) && c.encode != nullptr;
  if (s.tma) {
    KB_TRY(make_tmap(&s.tm_x, c.x, D));
    KB_TRY(make_tmap(&s.tm_p, c.p, D));
    static const bool no_p32 = getenv("KEFFB200_P64") != nullptr;  // experiments: keep p in FP64 on the multigrid path
    s.p32_ok = D.nx % 4 == 0 && !no_p32;
    if (s.p32_ok) KB_TRY(make_tmap_f32(&s.tm_p32, reinterpret_cast<float*>(c.p), D));
    // experiment switch: the warp-specialised ring (producer warp + full/empty mbarriers, code bytes by TMA).
    // Measured at 512^3 on B200: 0.445 ms (3 CTAs/SM, 72 registers, 6 stages) / 0.433 ms (2 CTAs/SM, 9 stages)
    // against 0.421 ms for the __syncthreads ring, so it is off by default.  Also measured and rejected: a 6-stage
    // __syncthreads ring (no change: prefetch depth is not the limiter) and an even split of (tile, plane) runs
    // over the grid (0.608 ms: x-adjacent tiles stop streaming the same DRAM rows at the same time).
    static const bool want_ws = getenv("KEFFB200_STENCIL_WS") != nullptr;
    s.ws = want_ws && D.nx % 16 == 0;
    if (s.ws) s.p32_ok = false;
    int occ = occupancy_tma<MODE_APPLY>();
    if (s.ws) {
      KB_TRY(make_tmap_code(&s.tm_code, c.code, D));
      KB_TRY(ws_prepare<MODE_APPLY>());
      KB_TRY(ws_prepare<MODE_INIT>());
      KB_TRY(ws_prepare<MODE_RESID>());
      KB_TRY(ws_prepare<MODE_INIT_R>());
      occ = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stencil_ws<MODE_APPLY>, WS_THR, WS_SMEM);
      occ = std::max(occ, 1);
    }
    const int G = c.sms * occ;
    Plan pl;
    pl.bfirst = 0;
    pl.ntx = (D.nx + TX - 1) / TX;
    pl.nty = (D.ny + TY - 1) / TY;
    const long tiles = (long)pl.ntx * pl.nty;
    // choose the z chunking that minimises (items per CTA) x (planes streamed per item)
    long best_cost = -1;
    int best_nzc = 1;
    for (int nzc = 1; nzc <= std::max(1, D.nz / 4); nzc++) {
      const int zlen = (D.nz + nzc - 1) / nzc;
      const int real_nzc = (D.nz + zlen - 1) / zlen;
      const long items = tiles * real_nzc;
      const long per_cta = (items + G - 1) / G;
      const long cost = per_cta * (zlen + 2 + (s.ws ? 1 : 6));  // +6: per-item pipeline fill/drain (the ws ring runs on)
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_nzc = real_nzc;
      }
    }
    pl.zlen = (D.nz + best_nzc - 1) / best_nzc;
    pl.nzc = (D.nz + pl.zlen - 1) / pl.zlen;
    pl.nitems = (int)(tiles * pl.nzc);
    s.plan = pl;
    s.grid = (int)std::min<long>(pl.nitems, G);
  } else {
    const long rows = (long)D.ny * D.nz;
    s.grid = (int)std::min<long>(rows, (long)c.sms * 8);
  }
  if (s.grid > PARTIAL_SLOTS) s.grid = PARTIAL_SLOTS;
  return 0;
}

// in_halo: halo'ed field base; which: 0 -> x, 1 -> p, 2 -> p stored in FP32 (selects the tensor map)
template <int MODE>
static int stencil_launch(const Stencil& s, int which, double* o0, double* o1) {
  Ctx& c = ctx();
  const size_t plane = (size_t)s.D.nx * s.D.ny;
  double* in_halo = which == 0 ? c.x : c.p;
  if (which == 2) {
    if (MODE != MODE_APPLY || !s.p32_ok) return fail(KEFFB200_ESTATE, "FP32 stencil input outside the multigrid iteration");
    k_stencil_tma<MODE_APPLY, float><<<s.grid, NTHR, 0, c.st>>>(s.tm_p32, c.code, o0, o1, s.C, s.D, s.plan, c.S, c.partials);
  } else if (s.ws) {
    k_stencil_ws<MODE><<<s.grid, WS_THR, WS_SMEM, c.st>>>(which == 0 ? s.tm_x : s.tm_p, s.tm_code, o0, o1, s.C, s.D, s.plan, c.S,
                                                          c.partials);
  } else if (s.tma) {
    k_stencil_tma<MODE><<<s.grid, NTHR, 0, c.st>>>(which == 0 ? s.tm_x : s.tm_p, c.code, o0, o1, s.C, s.D, s.plan, c.S,
                                                    c.partials);
  } else {
    k_stencil_plain<MODE><<<s.grid, 256, 0, c.st>>>(in_halo + plane, c.code, o0, o1, s.C, s.D, c.S, c.partials);
  }
  KB_LAUNCHED();
  KB_CUDA(cudaGetLastError());
  return 0;
}


static int vec_threads(int nx, int V) {
  const int lanes = (nx + V - 1) / V;
  return std::min(256, std::max(32, (lanes + 31) / 32 * 32));
}

// ---- multigrid preconditioner: hierarchy, plan, V-cycle -----------------------------------------
struct Mg {
  bool on = false;
  int nlev = 0;                // stored levels 1..nlev; level nlev is the dense one
  MgLevel L[MG_MAX_LEVELS];    // L[0]: extent and coarsening flags of the voxel grid only
  int ndense = 0;              // cells of the dense level (global)
  int z0[MG_MAX_LEVELS] = {0}, nzg[MG_MAX_LEVELS] = {0};  // per level: first global plane of my slab, global planes
  bool dist = false, lo = false, hi = false;              // z-slab mode; a neighbour below / above
  int la = 1 << 20;            // z-slabs: levels >= la are replicated on every rank (global extent, no exchanges)
  int nzl_la = 0, z0_la = 0;   // my planes of level la (its u is written / its e is read through this slice)
  CUtensorMap tm_u, tm_x[2];   // u and the two iterate fields of the voxel grid
  int pitch = 0;               // floats per row of those fields
  Plan plan{};
  int grid = 0, gu = 0;        // CTAs of the stencil passes / of the row-wise vector kernels
  int nu = 2;                  // smoothing sweeps before and after the coarse correction
  float om[4] = {0.595f, 1.614f, 0.f, 0.f};  // their weights (Chebyshev points of [0.4, 1.9] for nu = 2)
  bool u_pushed = false;       // the kernel that wrote u has already delivered its boundary planes to the neighbours
  int nu0 = 2;                 // the same for the voxel grid alone (KEFFB200_MG_OMEGA0 overrides; default = nu, om)
  float om0[4] = {0.595f, 1.614f, 0.f, 0.f};
  Dom D{};
  Coef C{};
};

static bool mg_eligible(const Dom& D, int flags) {
  return !(flags & (KEFFB200_F_NO_MG | KEFFB200_F_NO_TMA)) && D.nx % 2 == 0 && D.nx >= 4 && D.ny >= 2 &&
         (long)D.nx * D.ny * D.nzg > MG_DENSE_MAX && ctx().encode != nullptr;
}

static int mg_grid(long n) { return (int)std::max<long>(1, std::min<long>((n + 255) / 256, (long)ctx().sms * 8)); }
// block shape of the stored-level passes: bx threads along x, 256/bx rows per block
static int mg_bx(int nx) {
  int bx = 32;
  while (bx < nx && bx < 256) bx *= 2;
  return bx;
}
static dim3 mg_row_grid(int nx, int ny, int nz) {
  const int rpb = 256 / mg_bx(nx);
  const int gx = (ny + rpb - 1) / rpb;                                       // blocks covering y once
  const int gz = std::max(1, std::min(nz, (ctx().sms * 8 + gx - 1) / gx));   // ... and enough z for ~8 CTAs per SM
  return dim3((unsigned)gx, (unsigned)std::min(gz, 65535), 1);
}

static bool mg_prolong_on_the_fly() {  // experiments: KEFFB200_MG_PROLONG_FUSED=1 keeps the on-the-fly prolongation
  static const bool v = getenv("KEFFB200_MG_PROLONG_FUSED") != nullptr;
  return v;
}
static MgLevelView mg_view(const MgLevel& L) { return MgLevelView{L.nx, L.ny, L.nz, L.cx, L.cy, L.cz, L.dinv}; }

static int mg_setup(Mg& mg, const Dom& D, const Coef& C) {
  Ctx& c = ctx();
  mg.D = D;
  mg.C = C;
  if (const char* e = getenv("KEFFB200_MG_OMEGA")) {  // "w1,w2,..": overrides the smoothing weights (experiments)
    int k = 0;
    for (const char* q = e; *q && k < 4; k++) {
      char* end = nullptr;
      mg.om[k] = strtof(q, &end);
      if (end == q) break;
      q = *end == ',' ? end + 1 : end;
    }
    if (k >= 1) mg.nu = k;
  }
  mg.nu0 = mg.nu;
  for (int k = 0; k < 4; k++) mg.om0[k] = mg.om[k];
  if (const char* e = getenv("KEFFB200_MG_OMEGA0")) {  // voxel grid: its own sweep count / weights (experiments)
    int k = 0;
    for (const char* q = e; *q && k < 4; k++) {
      char* end = nullptr;
      mg.om0[k] = strtof(q, &end);
      if (end == q) break;
      q = *end == ',' ? end + 1 : end;
    }
    if (k >= 1) mg.nu0 = k;
  }
  Nccl& nc = c.nccl;
  mg.dist = c.dist_active && (c.link.p2p ? c.link.world : nc.world) > 1;
  mg.lo = mg.dist && D.z0 > 0;
  mg.hi = mg.dist && D.z0 + D.nz < D.nzg;
  // z-slabs: an aggregate must not straddle two ranks, so z is coarsened only while every rank's first
  // plane and plane count are even; equal slabs are required (the dense level is all-gathered)
  int kz = 1 << 20;
  auto even_depth = [](int z0, int nz) {
    int k = 0;
    while (((z0 >> k) & 1) == 0 && ((nz >> k) & 1) == 0 && (nz >> k) > 1) k++;
    return k;
  };
  if (mg.dist && c.link.p2p) {
    // every rank's slab geometry arrived with the window tables (link_publish): no collective needed
    for (int r = 0; r < c.link.world; r++) {
      const WinTable& o = c.link.h_win[r];
      if (o.nz != D.nz) return 0;  // unequal slabs: stay on the diagonally preconditioned path
      kz = std::min(kz, even_depth(o.z0, o.nz));
    }
  } else if (mg.dist) {
    const int k = even_depth(D.z0, D.nz);
    c.hS->tmp[0] = k;
    c.hS->tmp[1] = D.nz;
    c.hS->tmp[2] = -D.nz;
    KB_CUDA(cudaMemcpyAsync(c.S, c.hS, sizeof(Scal), cudaMemcpyHostToDevice, c.st));
    KB_TRY(nccl_check(nc.AllReduce(c.S->tmp, c.S->tmp, 3, /*ncclFloat64*/ 8, /*ncclMin*/ 3, nc.comm, c.st), "ncclAllReduce"));
    KB_CUDA(cudaMemcpyAsync(c.hS, c.S, sizeof(Scal), cudaMemcpyDeviceToHost, c.st));
    KB_CUDA(cudaStreamSynchronize(c.st));
    if (c.hS->tmp[1] != -c.hS->tmp[2]) return 0;  // unequal slabs: stay on the diagonally preconditioned path
    kz = (int)c.hS->tmp[0];
  }
  // extents.  Semi-coarsening: on the unit cube a grid with unequal cell counts has unequal cell sizes, and
  // point smoothing only works along strongly coupled (small-h) directions, so a direction is coarsened
  // only while its cell size is within 1.5x of the smallest one.  (The voxel-grid passes always coarsen
  // x and y: they handle the 2x2 xy block of a thread as one aggregate.)
  int nx = D.nx, ny = D.ny, nz = D.nz, nzg = D.nzg, z0 = D.z0, l = 0, zdone = 0;
  double hx = 1.0 / nx, hy = 1.0 / ny, hz = 1.0 / nzg;
  size_t floats = 0;
  bool rep = false;  // this level is replicated
  mg.la = 1 << 20;
  while (true) {
    MgLevel& L = mg.L[l];
    // z-slabs: once a level is small its halo exchanges cost more than its arithmetic, so from there on every
    // rank holds the whole level (the level above all-gathers its restricted residual into it)
    if (mg.dist && !rep && l > 0 && (long)nx * ny * nzg <= MG_REPLICATE_CELLS) {
      rep = true;
      mg.la = l;
      mg.nzl_la = nz;
      mg.z0_la = z0;
      nz = nzg;
      z0 = 0;
    }
    L.nx = nx;
    L.ny = ny;
    L.nz = nz;
    const bool cx = nx > 1, cy = ny > 1, cz = (mg.dist && !rep) ? (zdone < kz && nz > 1) : nz > 1;
    double hmin = 1e300;
    if (cx) hmin = std::min(hmin, hx);
    if (cy) hmin = std::min(hmin, hy);
    if (cz) hmin = std::min(hmin, hz);
    if (l == 0) {
      L.fx = L.fy = 1;
      L.fz = cz && hz < 1.5 * std::min(hx, hy);
    } else {
      L.fx = cx && hx < 1.5 * hmin;
      L.fy = cy && hy < 1.5 * hmin;
      L.fz = cz && hz < 1.5 * hmin;
    }
    mg.z0[l] = z0;
    mg.nzg[l] = nzg;
    if (l > 0) floats += 9 * (((size_t)nx * ny * (nz + 2) + 31) / 32 * 32);
    if (l > 0 && (long)nx * ny * nzg <= MG_DENSE_MAX) break;
    if (!L.fx && !L.fy && !L.fz) return 0;  // cannot get down to a dense level (only with odd slabs)
    if (l + 1 >= MG_MAX_LEVELS) return fail(KEFFB200_EINVAL, "multigrid hierarchy too deep");
    if (L.fx) {
      nx = (nx + 1) / 2;
      hx *= 2;
    }
    if (L.fy) {
      ny = (ny + 1) / 2;
      hy *= 2;
    }
    if (L.fz) {
      nz = (nz + 1) / 2;
      nzg = (nzg + 1) / 2;
      z0 /= 2;
      hz *= 2;
      zdone++;
    }
    l++;
  }
  mg.nlev = l;
  mg.ndense = mg.L[l].nx * mg.L[l].ny * nzg;
  if (floats > c.cap_mg) {
    link_release_window(c, W_POOL, c.mg_pool);
    c.mg_pool = nullptr;
    c.cap_mg = 0;
    KB_CUDA(cudaMalloc(&c.mg_pool, floats * sizeof(float)));
    c.cap_mg = floats;
  }
  if (!c.mg_inv) KB_CUDA(cudaMalloc(&c.mg_inv, (size_t)MG_DENSE_MAX * MG_DENSE_MAX * sizeof(double)));
  if (!c.mg_tab) KB_CUDA(cudaMalloc(&c.mg_tab, (size_t)2 * MG_TAB * sizeof(float)));
  mg.pitch = mg_pitch(D.nx);
  {
    const size_t need = (size_t)mg.pitch * D.ny * (D.nz + 2);
    if (mg.nu0 >= 2 && need > c.cap_xf) {
      for (int k = 0; k < 2; k++) {
        link_release_window(c, k == 0 ? W_XF0 : W_XF1, c.xf[k]);
        c.xf[k] = nullptr;
      }
      c.cap_xf = 0;
      for (int k = 0; k < 2; k++) KB_CUDA(cudaMalloc(&c.xf[k], need * sizeof(float)));
      c.cap_xf = need;
    }
  }
  // Everything a neighbour may store into once the group has met must be zeroed BEFORE the meeting: a fast
  // neighbour pushes its first halo planes while this rank is still setting up.
  KB_CUDA(cudaMemsetAsync(c.mg_pool, 0, floats * sizeof(float), c.st));
  {
    const size_t uplane = (size_t)mg.pitch * D.ny, need = uplane * (D.nz + 2);
    KB_CUDA(cudaMemsetAsync(c.u, 0, uplane * sizeof(float), c.st));
    KB_CUDA(cudaMemsetAsync(c.u + uplane * (D.nz + 1), 0, uplane * sizeof(float), c.st));
    if (mg.nu0 >= 2)
      for (int k = 0; k < 2; k++) KB_CUDA(cudaMemsetAsync(c.xf[k], 0, need * sizeof(float), c.st));
  }
  // the level pool and the iterate fields may just have moved: the neighbours need their new addresses
  if (mg.dist && c.link.p2p) KB_TRY(link_publish(c, D, false));
  k_mg_tables<<<(MG_TAB + 255) / 256, 256, 0, c.st>>>(c.mg_tab, C, D.nzg == 1);
  KB_LAUNCHED();
  float* base = c.mg_pool;
  for (int k = 1; k <= mg.nlev; k++) {
    MgLevel& L = mg.L[k];
    const size_t plane = (size_t)L.nx * L.ny, stride = (plane * (L.nz + 2) + 31) / 32 * 32;
    float** arr[9] = {&L.cx, &L.cy, &L.cz, &L.wsum, &L.dinv, &L.u, &L.e, &L.t0, &L.t1};
    for (int a = 0; a < 9; a++) {
      *arr[a] = base + plane;  // plane 0 (one spare plane below)
      base += stride;
    }
  }
  // operators.  Level k is built from level k-1; where level k is the first replicated one, every rank builds
  // its own planes inside the global arrays and the planes of the others arrive by in-place all-gathers.
  for (int k = 1; k <= mg.nlev; k++) {
    MgLevel& L = mg.L[k];
    const bool rep_k = k >= mg.la, first_rep = k == mg.la;
    const size_t plane = (size_t)L.nx * L.ny, pb = plane * sizeof(float);
    MgLevel T = L;  // target of the coarsening kernel: my slice when this is the first replicated level
    if (first_rep) {
      T.nz = mg.nzl_la;
      const size_t off = (size_t)mg.z0_la * plane;
      T.cx += off;
      T.cy += off;
      T.cz += off;
      T.wsum += off;
    }
    const int gk = mg_grid((long)T.nx * T.ny * T.nz);
    if (k == 1)
      k_mg_coarsen_fine<<<mg_row_grid(T.nx, T.ny, T.nz), 256, 0, c.st>>>(c.code, C, D, mg.L[0].fx, mg.L[0].fy, mg.L[0].fz, T,
                                                                          mg_bx(T.nx));
    else k_mg_coarsen_level<<<gk, 256, 0, c.st>>>(mg.L[k - 1], T);
    KB_LAUNCHED();
    if (first_rep) {
      const size_t cnt = plane * mg.nzl_la;
      KB_TRY(allgather_f32(T.cx, L.cx, cnt));
      KB_TRY(allgather_f32(T.cy, L.cy, cnt));
      KB_TRY(allgather_f32(T.cz, L.cz, cnt));
      KB_TRY(allgather_f32(T.wsum, L.wsum, cnt));
    }
    if (!rep_k) KB_TRY(exchange_planes(level_field(k, HL_CZ), W_POOL, L.cz, pb, L.nz, c.st));  // cz[-1]: the face towards the lower neighbour
    k_mg_dinv<<<mg_grid((long)L.nx * L.ny * L.nz), 256, 0, c.st>>>(L, rep_k ? 0 : (int)mg.lo);
    KB_LAUNCHED();
    if (!rep_k) KB_TRY(exchange_planes(level_field(k, HL_DINV), W_POOL, L.dinv, pb, L.nz, c.st));
  }
  k_mg_dense_setup<<<1, MG_DENSE_MAX, 0, c.st>>>(mg.L[mg.nlev], c.mg_inv);
  KB_LAUNCHED();
  KB_CUDA(cudaGetLastError());
  // voxel-grid passes: tensor maps on u and the iterate fields, work plan (z chunks of even length so
  // aggregates stay whole)
  {
    const size_t uplane = (size_t)mg.pitch * D.ny;
    float* fields[3] = {c.u, mg.nu0 >= 2 ? c.xf[0] : nullptr, mg.nu0 >= 2 ? c.xf[1] : nullptr};
    CUtensorMap* maps[3] = {&mg.tm_u, &mg.tm_x[0], &mg.tm_x[1]};
    for (int k = 0; k < 3; k++) {
      if (!fields[k]) continue;
      // (halo planes and row padding were zeroed before the group met, above)
      cuuint64_t dims[3] = {(cuuint64_t)D.nx, (cuuint64_t)D.ny, (cuuint64_t)D.nz + 2};
      cuuint64_t strides[2] = {(cuuint64_t)mg.pitch * 4, (cuuint64_t)uplane * 4};
      cuuint32_t box[3] = {MG_BOXW, BOXH, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult rc = c.encode(maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fields[k], dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) return fail(KEFFB200_ECUDA, "cuTensorMapEncodeTiled (multigrid field) failed (%d)", (int)rc);
    }
  }
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mg_fine<2, 1, 1>, NTHR, 0);
  occ = std::max(occ, 1);
  {
    const long rows = (long)D.ny * D.nz;
    const int vt = vec_threads(D.nx, 2);
    mg.gu = (int)std::min<long>(std::min<long>(rows, (long)c.sms * (2048 / vt)), (long)PARTIAL_SLOTS);
  }
  const int G = c.sms * occ;
  Plan pl;
  pl.bfirst = mg.dist ? 1 : 0;
  pl.ntx = (D.nx + TX - 1) / TX;
  pl.nty = (D.ny + TY - 1) / TY;
  const long tiles = (long)pl.ntx * pl.nty;
  long best_cost = -1;
  int best_zlen = D.nz;
  for (int nzc = 1; nzc <= std::max(1, D.nz / 4); nzc++) {
    int zlen = (D.nz + nzc - 1) / nzc;
    if (zlen & 1) zlen++;
    const int real_nzc = (D.nz + zlen - 1) / zlen;
    const long items = tiles * real_nzc;
    const long per_cta = (items + G - 1) / G;
    const long cost = per_cta * (zlen + 2 + 8);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_zlen = zlen;
    }
  }
  pl.zlen = best_zlen;
  pl.nzc = (D.nz + pl.zlen - 1) / pl.zlen;
  pl.nitems = (int)(tiles * pl.nzc);
  mg.plan = pl;
  mg.grid = (int)std::min<long>(std::min<long>(pl.nitems, G), PARTIAL_SLOTS);
  mg.on = true;
  return 0;
}

// u = r / d from a residual that some other kernel produced
static int mg_make_u(Mg& mg, double* r) {
  Ctx& c = ctx();
  const int vt = vec_threads(mg.D.nx, 2);
  const Fused f = fused_push(HF_U, W_U, (size_t)mg.pitch * mg.D.ny * sizeof(float), mg.D.nz);
  k_cg_update_mg<0><<<mg.gu, vt, 0, c.st>>>(r, nullptr, c.code, c.mg_tab, c.u + (size_t)mg.pitch * mg.D.ny, mg.pitch, mg.D,
                                            c.S, 0, c.partials, f.H);
  KB_LAUNCHED();
  mg.u_pushed = f.on;
  return fused_wait(f, (unsigned)mg.gu);
}

// One pass over the voxel grid.  in: 0 = u, 1/2 = iterate field 0/1;  out: -1 restrict, -2 z, 0/1 iterate field.
static int mg_fine_pass(const Mg& mg, int in, int out, bool add_e, const MgPass& K, float* z) {
  Ctx& c = ctx();
  const MgLevel& L0 = mg.L[0];
  const MgLevel& L1 = mg.L[1];
  const size_t uplane = (size_t)mg.pitch * mg.D.ny;
  // level 1 replicated: my planes are a slice of its global arrays (the planes either side are the neighbours')
  const size_t o1 = mg.la == 1 ? (size_t)mg.z0_la * L1.nx * L1.ny : 0;
  MgFine mf{L1.e + o1, L1.u + o1, mg.nlev > 1 ? L1.dinv + o1 : nullptr, L1.nx, L1.ny, mg.la == 1 ? mg.nzl_la : L1.nz,
            L0.fy, L0.fz, mg.lo, mg.hi, mg.pitch};
  const CUtensorMap& tm = in == 0 ? mg.tm_u : mg.tm_x[in - 1];
  const float* uown = c.u + uplane;
  float* fout = out >= 0 ? c.xf[out] + uplane : nullptr;
  Fused f;
  if (out >= 0) f = fused_push(out == 0 ? HF_XF0 : HF_XF1, out == 0 ? W_XF0 : W_XF1, uplane * sizeof(float), mg.D.nz);
  MgCoefF cf;
  for (int d = 0; d < 3; d++) {
    cf.same[d][0] = (float)mg.C.same[d][0];
    cf.same[d][1] = (float)mg.C.same[d][1];
    cf.diff[d] = (float)mg.C.diff[d];
  }
#define KB_FINE(O, E, U) \
  k_mg_fine<O, E, U><<<mg.grid, NTHR, 0, c.st>>>(tm, c.code, c.mg_tab, uown, fout, z, mf, K, cf, mg.D, mg.plan, c.S, c.partials, f.H)
  const bool isu = in == 0;
  if (out == -1) {
    if (isu) KB_FINE(0, 0, 1); else KB_FINE(0, 0, 0);
  } else if (out == -2) {
    if (add_e) { if (isu) KB_FINE(2, 1, 1); else KB_FINE(2, 1, 0); }
    else KB_FINE(2, 0, 0);
  } else {
    if (add_e) KB_FINE(1, 1, 0);
    else if (isu) KB_FINE(1, 0, 1);
    else KB_FINE(1, 0, 0);
  }
#undef KB_FINE
  KB_LAUNCHED();
  if (out >= 0) {
    if (f.on) return fused_wait(f, (unsigned)(mg.plan.ntx * mg.plan.nty));  // one arrival per tile of a boundary chunk
    KB_TRY(exchange_planes(out == 0 ? HF_XF0 : HF_XF1, out == 0 ? W_XF0 : W_XF1, c.xf[out] + uplane, uplane * sizeof(float), mg.D.nz,
                           c.st));
  }
  return 0;
}

// z = M r, r given through u = r / d (ctx().u); partial r.z lands in ctx().partials, stride 3.
// Per level: nu-1 sweeps (the first iterate w_1 u is implicit), the restrict pass, ... coarse solve ...,
// the sweep that adds the correction, nu-1 more sweeps (weights in reverse order).
// z-slabs: every array that a pass reads across a slab boundary has its halo planes exchanged first.
static int mg_vcycle(const Mg& mg, float* z) {
  Ctx& c = ctx();
  const int lo = mg.lo, hi = mg.hi, nu = mg.nu, nu0 = mg.nu0;
  const float* w = mg.om;
  const float* w0 = mg.om0;
  if (!mg.u_pushed)
    KB_TRY(exchange_planes(HF_U, W_U, c.u + (size_t)mg.pitch * mg.D.ny, (size_t)mg.pitch * mg.D.ny * sizeof(float), mg.D.nz, c.st));
  // ---- voxel grid, down ----
  int cur = 0;  // field holding the current iterate (0 = u, scaled by w_1)
  float sc = w0[0];
  for (int k = 1; k < nu0; k++) {
    const int out = cur == 1 ? 1 : 0;  // iterate fields alternate: u -> x0 -> x1 -> x0 ..
    KB_TRY(mg_fine_pass(mg, cur, out, false, MgPass{sc, 1.f - w0[k], w0[k], w0[k]}, z));
    cur = out + 1;
    sc = 1.f;
  }
  trace_mark("fineA+wait");
  KB_TRY(mg_fine_pass(mg, cur, -1, false, MgPass{sc, -1.f, 1.f, 1.f}, z));
  trace_mark("fineB");
  const int fine_cur = cur;
  const float fine_sc = sc;
  // ---- stored levels, down ----
  // gather_u(l): level l is the first replicated one and its u was just restricted into my slice of it
  auto gather_u = [&](int l) -> int {
    if (l != mg.la) return 0;
    const MgLevel& L = mg.L[l];
    const size_t plane = (size_t)L.nx * L.ny;
    return allgather_f32(L.u + (size_t)mg.z0_la * plane, L.u, plane * mg.nzl_la);
  };
  KB_TRY(gather_u(1));
  const float* lcur[MG_MAX_LEVELS];
  float lsc[MG_MAX_LEVELS];
  for (int l = 1; l < mg.nlev; l++) {
    const MgLevel& L = mg.L[l];
    const MgLevel& N = mg.L[l + 1];
    const bool rep = l >= mg.la, nslice = l + 1 == mg.la;  // nslice: I fill only my planes of the next level
    const int llo = rep ? 0 : lo, lhi = rep ? 0 : hi;
    const size_t pb = (size_t)L.nx * L.ny * sizeof(float);
    const size_t on = nslice ? (size_t)mg.z0_la * N.nx * N.ny : 0;
    const int nnz = nslice ? mg.nzl_la : N.nz;
    const dim3 gl = mg_row_grid(L.nx, L.ny, L.nz), gn = mg_row_grid(N.nx, N.ny, nnz);
    if (!rep) KB_TRY(exchange_planes(level_field(l, HL_U), W_POOL, L.u, pb, L.nz, c.st));
    const float* in = L.u;
    float s1 = w[0];
    for (int k = 1; k < nu; k++) {
      float* out = in == L.t0 ? L.t1 : L.t0;
      k_mg_sweep<0><<<gl, 256, 0, c.st>>>(mg_view(L), in, L.u, nullptr, MgPass{s1, 1.f - w[k], w[k], w[k]}, L.fx, L.fy, L.fz,
                                          N.nx, N.ny, out, llo, lhi, mg_bx(L.nx), c.S);
      KB_LAUNCHED();
      if (!rep) KB_TRY(exchange_planes(level_field(l, out == L.t0 ? HL_T0 : HL_T1), W_POOL, out, pb, L.nz, c.st));
      in = out;
      s1 = 1.f;
    }
    k_mg_restrict<<<gn, 256, 0, c.st>>>(mg_view(L), in, L.u, MgPass{s1, -1.f, 1.f, 1.f}, L.fx, L.fy, L.fz, N.nx, N.ny, nnz,
                                        N.u + on, l + 1 < mg.nlev ? N.dinv + on : nullptr, llo, lhi, mg_bx(N.nx), c.S);
    KB_LAUNCHED();
    KB_TRY(gather_u(l + 1));
    lcur[l] = in;
    lsc[l] = s1;
  }
  trace_mark("levels-down");
  // ---- dense level (replicated; its u array holds the plain right-hand side) ----
  {
    const MgLevel& L = mg.L[mg.nlev];
    k_mg_dense_apply<<<1, MG_DENSE_MAX, 0, c.st>>>(c.mg_inv, L.u, L.e, mg.ndense, c.S);
    KB_LAUNCHED();
  }
  // ---- stored levels, up ----
  for (int l = mg.nlev - 1; l >= 1; l--) {
    const MgLevel& L = mg.L[l];
    const MgLevel& N = mg.L[l + 1];
    const bool rep = l >= mg.la, nslice = l + 1 == mg.la;
    const int llo = rep ? 0 : lo, lhi = rep ? 0 : hi;
    const size_t pb = (size_t)L.nx * L.ny * sizeof(float);
    const size_t on = nslice ? (size_t)mg.z0_la * N.nx * N.ny : 0;  // my planes of the replicated level below
    const dim3 gl = mg_row_grid(L.nx, L.ny, L.nz);
    const float* in = lcur[l];
    float s1 = lsc[l];
    for (int k = nu - 1; k >= 0; k--) {
      float* out = k == 0 ? L.e : (in == L.t0 ? L.t1 : L.t0);
      const MgPass K{s1, 1.f - w[k], w[k], w[k]};
      if (k == nu - 1 && in != L.u && s1 == 1.f && !mg_prolong_on_the_fly()) {
        // the iterate is a stored array of its own (nu >= 2): add P e to it in place (halo planes included), then sweep
        float* xin = in == L.t0 ? L.t0 : L.t1;
        const int zf = llo ? -1 : 0, zl = L.nz - 1 + (lhi ? 1 : 0);
        // (four cells per thread along x where rows allow it: the block shape follows nx / 4)
        const int pnx = (L.fx && L.nx % 4 == 0) ? L.nx / 4 : L.nx;
        dim3 gp = mg_row_grid(pnx, L.ny, zl - zf + 1);
        k_mg_prolong_add<<<gp, 256, 0, c.st>>>(L.nx, L.ny, zf, zl, xin, N.e + on, L.fx, L.fy, L.fz, N.nx, N.ny, mg_bx(pnx), c.S);
        KB_LAUNCHED();
        k_mg_sweep<0><<<gl, 256, 0, c.st>>>(mg_view(L), in, L.u, nullptr, K, L.fx, L.fy, L.fz, N.nx, N.ny, out, llo, lhi,
                                            mg_bx(L.nx), c.S);
      } else if (k == nu - 1)
        k_mg_sweep<1><<<gl, 256, 0, c.st>>>(mg_view(L), in, L.u, N.e + on, K, L.fx, L.fy, L.fz, N.nx, N.ny, out, llo, lhi,
                                            mg_bx(L.nx), c.S);
      else
        k_mg_sweep<0><<<gl, 256, 0, c.st>>>(mg_view(L), in, L.u, nullptr, K, L.fx, L.fy, L.fz, N.nx, N.ny, out, llo, lhi,
                                            mg_bx(L.nx), c.S);
      KB_LAUNCHED();
      if (!rep) KB_TRY(exchange_planes(level_field(l, out == L.e ? HL_E : (out == L.t0 ? HL_T0 : HL_T1)), W_POOL, out, pb, L.nz, c.st));
      in = out;
      s1 = 1.f;
    }
  }
  trace_mark("dense+levels-up");
  // ---- voxel grid, up ----
  cur = fine_cur;
  sc = fine_sc;
  for (int k = nu0 - 1; k >= 0; k--) {
    const int out = k == 0 ? -2 : (cur == 1 ? 1 : 0);
    KB_TRY(mg_fine_pass(mg, cur, out, k == nu0 - 1, MgPass{sc, 1.f - w0[k], w0[k], w0[k]}, z));
    trace_mark(k == 0 ? "fineD" : "fineC+wait");
    cur = out + 1;
    sc = 1.f;
  }
  KB_CUDA(cudaGetLastError());
  return 0;
}

// ---- the solve ----------------------------------------------------------------------------------
// d_mask_halo: (nz+2) planes (plane 0 and nz+1 are the slab neighbours' planes or don't-care);
// d_Tinit / d_Tout: nz planes, nullable.
static int solve_grid(const uint8_t* d_mask_halo, const Dom& D, const keffb200_params& P, const double* d_Tinit,
                      double* d_Tout, double* d_qL, double* d_qR, keffb200_result* out) {
  Ctx& c = ctx();
  if (D.nx < 2 || D.ny < 1 || D.nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", D.nx, D.ny, D.nz);
  if (P.tr == P.tl) return fail(KEFFB200_EINVAL, "TR == TL: Keff undefined");
  if (!(P.ks > 0) || !(P.kf > 0)) return fail(KEFFB200_EINVAL, "conductivities must be positive");
  KB_TRY(ensure_ws(D));
  const size_t plane = (size_t)D.nx * D.ny, n = plane * D.nz;
  // halo planes of the two stencil inputs start at zero (global ends keep them: zero coefficient)
  KB_CUDA(cudaMemsetAsync(c.x, 0, plane * 8, c.st));
  KB_CUDA(cudaMemsetAsync(c.x + plane * (D.nz + 1), 0, plane * 8, c.st));
  KB_CUDA(cudaMemsetAsync(c.p, 0, plane * 8, c.st));
  KB_CUDA(cudaMemsetAsync(c.p + plane * (D.nz + 1), 0, plane * 8, c.st));
  // (p is stored in FP32 on the multigrid path, decided below: its upper halo plane then sits at half the offset;
  //  in the FP64 layout that is interior, which the first direction update rewrites)
  KB_CUDA(cudaMemsetAsync(reinterpret_cast<float*>(c.p) + plane * (D.nz + 1), 0, plane * 4, c.st));
  // z-slabs over peer memory: meet the group (everybody has finished its previous solve), clear my arrival
  // counters, publish where my workspace lives and learn where the neighbours' does
  if (c.dist_active && c.link.p2p) KB_TRY(link_publish(c, D, true));
  const Coef C = make_coef(D.nx, D.ny, D.nzg, P);
  const bool lo = c.dist_active && D.z0 > 0, hi = c.dist_active && D.z0 + D.nz < D.nzg;
  double *x = c.x + plane, *r = c.r + plane, *p = c.p + plane, *q = c.q + plane;
  const int V = D.nx % 2 == 0 ? 2 : 1;
  const int vt = vec_threads(D.nx, V);
  const long rows = (long)D.ny * D.nz;
  const int gv = (int)std::min<long>(rows, (long)c.sms * (2048 / vt));
  const int gf = (int)std::min<long>((rows + 255) / 256, (long)c.sms * 4);

  KB_CUDA(cudaEventRecord(c.ev0, c.st));
  k_make_code<<<c.sms * 8, 256, 0, c.st>>>(d_mask_halo + plane, c.code, D, lo, hi);
  KB_LAUNCHED();
  if (d_Tinit) {
    KB_CUDA(cudaMemcpyAsync(x, d_Tinit, n * 8, cudaMemcpyDeviceToDevice, c.st));
  } else {
    k_init_field<<<c.sms * 8, 256, 0, c.st>>>(x, D, P.tl, P.tr, (P.flags & KEFFB200_F_ZERO_INIT) ? 1 : 0);
    KB_LAUNCHED();
  }
  Stencil s;
  KB_TRY(stencil_setup(s, D, C, P.flags));
  Mg mg;
  if (mg_eligible(D, P.flags)) KB_TRY(mg_setup(mg, D, C));
  const long n2 = (long)(n / 2);
  const int gm = mg_grid(n2);
  // p in FP32 on the multigrid path (x and r advance with the same stored p, so r stays the residual of x)
  const bool p32 = mg.on && s.p32_ok;
  const size_t pesz = p32 ? 4 : 8;
  float* pf = reinterpret_cast<float*>(c.p) + plane;
  // iterations between two convergence records.  The multigrid path takes ~10 iterations of milliseconds: it
  // records every iteration; the diagonally preconditioned path takes thousands of short ones.
  const long batch_len = mg.on ? 1 : P.check_every;

  Scal h0;
  memset(&h0, 0, sizeof h0);
  // Diagonally preconditioned CG leaves its error in the smooth modes: at ks:kf = 300 a relative residual of 1e-8 was
  // seen 5e-5 away from a direct solve in Keff (the multigrid path: 1e-7), so the fallback aims 100x lower than asked.
  double thr = mg.on ? P.rel_residual : 0.01 * P.rel_residual;
  h0.thr2 = thr * thr;
  *c.hS = h0;
  KB_CUDA(cudaMemcpyAsync(c.S, c.hS, sizeof(Scal), cudaMemcpyHostToDevice, c.st));

  bool p_pushed = false;  // the kernel that wrote p has already delivered its boundary planes to the neighbours
  // (re)start: r = b - A x, p = z = M r
  auto enqueue_init = [&]() -> int {
    KB_TRY(exchange_halo(c.x, D, c.st));
    if (mg.on) KB_TRY(stencil_launch<MODE_INIT_R>(s, 0, r, nullptr));  // the cycle builds z itself
    else KB_TRY(stencil_launch<MODE_INIT>(s, 0, r, p));
    KB_TRY(reduce_step<RS_INIT>(s.grid, 0));
    if (mg.on) {  // p = z = M r, r.z
      KB_TRY(mg_make_u(mg, r));
      KB_TRY(mg_vcycle(mg, c.z));
      KB_TRY(reduce_step<RS_RZ_MG>(mg.grid, 0));
      const Fused f = fused_push(HF_P, W_P, plane * pesz, D.nz);
      if (p32) k_cg_dir_mg<float><<<gm, 256, 0, c.st>>>(x, pf, c.z, n2, (long)n, c.S, 0, -1, 1, f.H, (long)(plane / 2));
      else k_cg_dir_mg<double><<<gm, 256, 0, c.st>>>(x, p, c.z, n2, (long)n, c.S, 0, -1, 1, f.H, (long)(plane / 2));
      KB_LAUNCHED();
      p_pushed = f.on;
      KB_TRY(fused_wait(f, (unsigned)gm));
    }
    return 0;
  };
  // iteration k since the last (re)start; every kernel returns at once when the stop flag is up
  auto enqueue_iter = [&](long k) -> int {
    const int par = (int)(k & 1);
    trace().live = trace().want >= 0 && k == trace().want && trace().ev.empty();
    trace_mark("start");
    if (!p_pushed) KB_TRY(exchange_halo(c.p, D, c.st, pesz));
    KB_TRY(stencil_launch<MODE_APPLY>(s, p32 ? 2 : 1, q, nullptr));
    trace_mark("stencil");
    KB_TRY(reduce_step<RS_PQ>(s.grid, 0));
    trace_mark("allreduce-pq");
    if (mg.on) {
      const Fused fu = fused_push(HF_U, W_U, (size_t)mg.pitch * D.ny * sizeof(float), D.nz);
      k_cg_update_mg<1><<<mg.gu, vt, 0, c.st>>>(r, q, c.code, c.mg_tab, c.u + (size_t)mg.pitch * D.ny, mg.pitch, D, c.S, par,
                                                c.partials, fu.H);
      KB_LAUNCHED();
      mg.u_pushed = fu.on;
      trace_mark("update");
      KB_TRY(reduce_step<RS_RR_MG>(mg.gu, 0));
      trace_mark("allreduce-rr");
      KB_TRY(fused_wait(fu, (unsigned)mg.gu));
      trace_mark("wait-u");
      KB_TRY(mg_vcycle(mg, c.z));
      KB_TRY(reduce_step<RS_RZ_MG>(mg.grid, par ^ 1));
      trace_mark("allreduce-rz");
      const Fused fp = fused_push(HF_P, W_P, plane * pesz, D.nz);
      if (p32) k_cg_dir_mg<float><<<gm, 256, 0, c.st>>>(x, pf, c.z, n2, (long)n, c.S, par, (long long)k, 0, fp.H, (long)(plane / 2));
      else k_cg_dir_mg<double><<<gm, 256, 0, c.st>>>(x, p, c.z, n2, (long)n, c.S, par, (long long)k, 0, fp.H, (long)(plane / 2));
      KB_LAUNCHED();
      p_pushed = fp.on;
      trace_mark("dir");
      KB_TRY(fused_wait(fp, (unsigned)gm));
      trace_mark("wait-p");
      trace().live = false;
      return 0;
    }
    if (V == 2) k_cg_update<2><<<gv, vt, 0, c.st>>>(x, r, p, q, c.code, C, D, c.S, par, c.partials);
    else k_cg_update<1><<<gv, vt, 0, c.st>>>(x, r, p, q, c.code, C, D, c.S, par, c.partials);
    KB_LAUNCHED();
    KB_TRY(reduce_step<RS_ITER>(gv, par));
    if (V == 2) k_cg_dir<2><<<gv, vt, 0, c.st>>>(p, r, c.code, C, D, c.S, par);
    else k_cg_dir<1><<<gv, vt, 0, c.st>>>(p, r, c.code, C, D, c.S, par);
    KB_LAUNCHED();
    return 0;
  };
  // convergence record `slot`: wall fluxes (computed only once the stop flag is up -- they are part of the stop
  // rule, not of the iteration), then the scalars travel to pinned memory; the host reads them one record behind
  auto enqueue_record = [&](int slot, bool force_flux) -> int {
    k_flux<<<gf, 256, 0, c.st>>>(x, c.code, C, D, nullptr, nullptr, c.partials, force_flux ? nullptr : c.S);
    KB_LAUNCHED();
    KB_TRY(reduce_step<RS_FLUX>(gf, force_flux ? 1 : 0));
    KB_CUDA(cudaMemcpyAsync(&c.hring[slot], c.S, sizeof(Scal), cudaMemcpyDeviceToHost, c.st));
    KB_CUDA(cudaEventRecord(c.evb[slot], c.st));
    return 0;
  };

  long total_iters = 0;  // iterations of the (re)starts already closed
  double keff = NAN, ql = 0, qr = 0;
  int status = 1;
  bool finished = false;
  while (!finished) {
    KB_TRY(enqueue_init());
    long queued = 0;         // iterations enqueued since this (re)start
    int head = 0, tail = 0;  // records enqueued / read
    while (true) {
      // two records in flight: the device never waits for the host
      while (head - tail < 2 && (head == 0 || total_iters + queued < P.max_iter)) {
        const long len = std::min<long>(batch_len, P.max_iter - total_iters - queued);
        for (long i = 0; i < len; i++) KB_TRY(enqueue_iter(queued + i));
        queued += std::max<long>(len, 0);
        const bool last = total_iters + queued >= P.max_iter;
        KB_TRY(enqueue_record(head & 1, last));
        head++;
      }
      KB_CUDA(cudaEventSynchronize(c.evb[tail & 1]));
      const Scal rec = c.hring[tail & 1];
      tail++;
      if (rec.fault) {
        cudaStreamSynchronize(c.st);
        return fail(KEFFB200_ENCCL, "a z-slab neighbour never delivered (peer-memory wait timed out)");
      }
      const long its_now = total_iters + rec.it;
      if (rec.done) {
        KB_CUDA(cudaStreamSynchronize(c.st));  // whatever was queued behind this record has returned at once
        ql = rec.ql;
        qr = rec.qr;
        keff = (ql + qr) / 2 / (P.tr - P.tl);
        if (!std::isfinite(keff)) return fail(KEFFB200_ECUDA, "solver produced a non-finite Keff");
        const double imb = fabs(ql - qr) / fabs(keff);
        total_iters = its_now;
        if (imb <= P.flux_balance || thr < 1e-14) {
          status = 0;
          finished = true;
          break;
        }
        if (total_iters >= P.max_iter) {
          finished = true;
          break;
        }
        // residual target met but the two wall fluxes still disagree: tighten and restart CG from x
        thr *= 0.01;
        *c.hS = rec;
        c.hS->thr2 = thr * thr;
        c.hS->done = 0;
        KB_CUDA(cudaMemcpyAsync(c.S, c.hS, sizeof(Scal), cudaMemcpyHostToDevice, c.st));
        break;
      }
      if (tail == head && its_now >= P.max_iter) {  // iteration cap: the last record carries the fluxes
        ql = rec.ql;
        qr = rec.qr;
        keff = (ql + qr) / 2 / (P.tr - P.tl);
        if (!std::isfinite(keff)) return fail(KEFFB200_ECUDA, "solver produced a non-finite Keff");
        total_iters = its_now;
        finished = true;
        break;
      }
    }
  }

  // true residual of the returned field, per-row fluxes, outputs
  KB_CUDA(cudaMemsetAsync(&c.S->done, 0, sizeof(int), c.st));
  KB_TRY(exchange_halo(c.x, D, c.st));
  KB_TRY(stencil_launch<MODE_RESID>(s, 0, nullptr, nullptr));
  KB_TRY(reduce_step<RS_RESID>(s.grid, 0));
  KB_CUDA(cudaMemcpyAsync(c.hS, c.S, sizeof(Scal), cudaMemcpyDeviceToHost, c.st));
  if (d_qL || d_qR) {
    k_flux<<<gf, 256, 0, c.st>>>(x, c.code, C, D, d_qL, d_qR, c.partials, nullptr);
    KB_LAUNCHED();
  }
  if (d_Tout) KB_CUDA(cudaMemcpyAsync(d_Tout, x, n * 8, cudaMemcpyDeviceToDevice, c.st));
  KB_CUDA(cudaEventRecord(c.ev1, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  float ms = 0;
  KB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  trace().live = false;
  trace_report(c.dist_active ? (c.link.p2p ? c.link.rank : c.nccl.rank) : 0);
  if (out) {
    out->keff = keff;
    out->q_left = ql;
    out->q_right = qr;
    out->rel_residual = c.hS->bb > 0 ? sqrt(c.hS->tmp[1] / c.hS->bb) : 0.0;
    out->energy_residual = c.hS->tmp[2];
    out->iters = total_iters;
    out->status = status;
    out->gpu_ms = ms;
  }
  return 0;
}

// stage a host (or device) structure into the halo'ed mask buffer of the workspace
static int stage_mask(const uint8_t* src, bool src_is_device, const Dom& D) {
  Ctx& c = ctx();
  KB_TRY(ensure_ws(D));
  const size_t plane = (size_t)D.nx * D.ny;
  KB_CUDA(cudaMemsetAsync(c.mask, 0, plane, c.st));
  KB_CUDA(cudaMemsetAsync(c.mask + plane * (D.nz + 1), 0, plane, c.st));
  KB_CUDA(cudaMemcpyAsync(c.mask + plane, src, plane * D.nz,
                          src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.st));
  return 0;
}

// a device-buffer entry point was handed memory of another GPU: refuse (a kernel on this device would fault on
// it, or -- with peer access enabled -- silently do another rank's work here)
static int check_dev_ptr(const void* p, const char* what) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return fail(KEFFB200_EINVAL, "%s: not a CUDA pointer", what);
  }
  if (a.type == cudaMemoryTypeDevice && a.device != ctx().device)
    return fail(KEFFB200_EINVAL, "%s lives on device %d but the library is bound to device %d", what, a.device, ctx().device);
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    return fail(KEFFB200_EINVAL, "%s: expected device memory", what);
  return 0;
}

static int need_ready() {
  if (!ctx().ready) return fail(KEFFB200_ESTATE, "keffb200_init has not been called");
  return 0;
}

}  // namespace kb

using namespace kb;

// number of contexts in use: 1 after keffb200_init, n after keffb200_init_multi(n)
static int g_nctx = 0;

extern "C" {

const char* keffb200_version(void) { return "keffb200 0.1 (sm_100a)"; }
const char* keffb200_last_error(void) { return ctx().err.c_str(); }
long keffb200_launch_count(void) {
  long n = 0;
  for (int i = 0; i < std::max(g_nctx, 1); i++) n += ctx_table()[i].launches;
  return n;
}

void keffb200_default_params(keffb200_params* p) {
  memset(p, 0, sizeof *p);
  p->ks = 10;
  p->kf = 1;
  p->tl = 0;
  p->tr = 1;
  p->convergence = 1e-5;
  p->max_iter = 500000;
  p->rel_residual = 0;  /* 0 = derived from `convergence`: 1e-8 at the shipped 1e-5 */
  p->flux_balance = 0;  /* 0 = derived from `convergence`: 1e-6 at the shipped 1e-5 */
  p->check_every = 32;
  p->flags = 0;
}

static int init_context(Ctx& c, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(KEFFB200_ENOGPU, "no CUDA device visible: this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(KEFFB200_EINVAL, "device %d out of range (%d visible)", device, ndev);
  KB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(KEFFB200_ENOGPU, "device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major,
                prop.minor);
  c.device = device;
  c.sms = prop.multiProcessorCount;
  KB_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
  KB_CUDA(cudaStreamCreateWithFlags(&c.st2, cudaStreamNonBlocking));
  KB_CUDA(cudaEventCreate(&c.ev0));
  KB_CUDA(cudaEventCreate(&c.ev1));
  KB_CUDA(cudaEventCreate(&c.evt0));
  KB_CUDA(cudaEventCreate(&c.evt1));
  KB_CUDA(cudaEventCreateWithFlags(&c.evx, cudaEventDisableTiming));
  KB_CUDA(cudaEventCreateWithFlags(&c.evy, cudaEventDisableTiming));
  for (int k = 0; k < 3; k++) KB_CUDA(cudaEventCreateWithFlags(&c.evp[k], cudaEventDisableTiming));
  KB_CUDA(cudaMalloc(&c.partials, (size_t)PARTIAL_SLOTS * 3 * sizeof(double)));
  KB_CUDA(cudaMalloc(&c.S, sizeof(Scal)));
  KB_CUDA(cudaMemset(c.S, 0, sizeof(Scal)));
  KB_CUDA(cudaMallocHost(&c.hS, sizeof(Scal)));
  KB_CUDA(cudaMallocHost(&c.hring, 2 * sizeof(Scal)));
  KB_CUDA(cudaMallocHost(&c.h_ready, KB_READY_SLOTS * sizeof(unsigned int)));
  for (int k = 0; k < 2; k++) KB_CUDA(cudaEventCreateWithFlags(&c.evb[k], cudaEventDisableTiming));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
      qr == cudaDriverEntryPointSuccess)
    c.encode = (PFN_encodeTiled)fn;
  else
    return fail(KEFFB200_ECUDA, "driver does not export cuTensorMapEncodeTiled");
  c.launches = 0;
  c.ready = true;
  return 0;
}

static void destroy_context(Ctx& c) {
  if (!c.ready) return;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  link_free(c);
  for (double** f : {&c.x, &c.r, &c.p, &c.q, &c.qL, &c.qR, &c.partials}) {
    if (*f) cudaFree(*f);
    *f = nullptr;
  }
  if (c.code) cudaFree(c.code);
  if (c.mask) cudaFree(c.mask);
  if (c.z) cudaFree(c.z);
  if (c.u) cudaFree(c.u);
  c.u = nullptr;
  c.cap_u = 0;
  for (int k = 0; k < 2; k++) {
    if (c.xf[k]) cudaFree(c.xf[k]);
    c.xf[k] = nullptr;
  }
  c.cap_xf = 0;
  if (c.mg_pool) cudaFree(c.mg_pool);
  if (c.mg_inv) cudaFree(c.mg_inv);
  if (c.mg_tab) cudaFree(c.mg_tab);
  c.mg_tab = nullptr;
  c.z = nullptr;
  c.mg_pool = nullptr;
  c.mg_inv = nullptr;
  c.cap_z = c.cap_mg = 0;
  if (c.batch_ws) cudaFree(c.batch_ws);
  if (c.stage) cudaFree(c.stage);
  if (c.stageT) cudaFree(c.stageT);
  c.stage = nullptr;
  c.stageT = nullptr;
  c.cap_stage = c.cap_stageT = 0;
  for (int k = 0; k < 3; k++) cudaEventDestroy(c.evp[k]);
  if (c.S) cudaFree(c.S);
  if (c.hS) cudaFreeHost(c.hS);
  if (c.hring) cudaFreeHost(c.hring);
  c.hring = nullptr;
  if (c.h_ready) cudaFreeHost(c.h_ready);
  c.h_ready = nullptr;
  for (int k = 0; k < 2; k++) {
    if (c.evb[k]) cudaEventDestroy(c.evb[k]);
    c.evb[k] = nullptr;
  }
  c.attr_batch = c.attr_ws = false;
  c.code = c.mask = nullptr;
  c.batch_ws = nullptr;
  c.S = c.hS = nullptr;
  c.cap_field = c.cap_aux = c.cap_batch = 0;
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  cudaEventDestroy(c.evt0);
  cudaEventDestroy(c.evt1);
  cudaEventDestroy(c.evx);
  cudaEventDestroy(c.evy);
  cudaStreamDestroy(c.st);
  cudaStreamDestroy(c.st2);
  c.ready = false;
}

int keffb200_init(int device) {
  Ctx& c = ctx_table()[0];
  if (c.ready && g_nctx == 1 && c.device == device) return 0;
  if (g_nctx > 0) keffb200_shutdown();
  KB_TRY(init_context(c, device));
  g_nctx = 1;
  return 0;
}

int keffb200_init_multi(int n_gpus) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(KEFFB200_ENOGPU, "no CUDA device visible: this library has no CPU path");
  if (n_gpus <= 0) n_gpus = ndev;
  if (n_gpus > ndev || n_gpus > KB_MAX_GPUS || n_gpus > KB_MAX_RANKS)
    return fail(KEFFB200_EINVAL, "%d GPUs requested, %d visible (at most %d per process)", n_gpus, ndev, KB_MAX_RANKS);
  if (g_nctx == n_gpus && ctx_table()[0].ready && (n_gpus == 1 || ctx_table()[0].link.world == n_gpus)) return 0;
  if (g_nctx > 0) keffb200_shutdown();
  Ctx* T = ctx_table();
  for (int i = 0; i < n_gpus; i++) {
    const int rc = init_context(T[i], i);
    if (rc != 0) {
      if (i > 0) T[0].err = T[i].err.empty() ? T[0].err : T[i].err;
      return rc;
    }
  }
  g_nctx = n_gpus;
  if (n_gpus > 1) {
    // one slab group inside the process: every GPU maps every other GPU's memory (NVLink peer access)
    bool all = true;
    for (int i = 0; i < n_gpus && all; i++)
      for (int j = 0; j < n_gpus && all; j++) {
        int can = 1;
        if (i != j) cudaDeviceCanAccessPeer(&can, i, j);
        all = can != 0;
      }
    if (!all) return fail(KEFFB200_ECUDA, "GPUs 0..%d cannot all access each other's memory", n_gpus - 1);
    for (int i = 0; i < n_gpus; i++) {
      KB_CUDA(cudaSetDevice(i));
      for (int j = 0; j < n_gpus; j++) {
        if (i == j) continue;
        const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail(KEFFB200_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", i, j, cudaGetErrorString(e));
      }
      T[i].link.rank = i;
      T[i].link.world = n_gpus;
      KB_TRY(link_alloc_mail(T[i]));
    }
    for (int i = 0; i < n_gpus; i++) {
      for (int j = 0; j < n_gpus; j++) T[i].link.mail[j] = T[j].link.mail[j];
      T[i].link.p2p = true;
    }
  }
  KB_CUDA(cudaSetDevice(0));
  return 0;
}

int keffb200_device_count(void) { return g_nctx; }

void keffb200_shutdown(void) {
  Ctx* T = ctx_table();
  if (g_nctx == 0 && !T[0].ready) return;
  keffb200_dist_shutdown();
  for (int i = 0; i < KB_MAX_GPUS; i++) destroy_context(T[i]);
  g_nctx = 0;
}

}  // extern "C"

// run fn(i) on context i (i < n), one host thread per GPU; the first error wins and lands in context 0
template <class F>
static int fan_out(int n, F fn) {
  std::vector<std::thread> th;
  std::vector<int> rc(n, 0);
  Ctx* T = ctx_table();
  for (int i = 0; i < n; i++)
    th.emplace_back([&, i] {
      ctx_current() = &T[i];
      if (cudaSetDevice(T[i].device) != cudaSuccess) rc[i] = fail(KEFFB200_ECUDA, "cudaSetDevice(%d) failed", T[i].device);
      else rc[i] = fn(i);
      ctx_current() = nullptr;
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i] != 0) {
      if (i > 0) T[0].err = T[i].err;
      return rc[i];
    }
  return 0;
}

// contiguous shares of n units over `parts` owners, sizes differing by at most one
static void share(long n, int part, int parts, long* first, long* count) {
  const long base = n / parts, rem = n % parts;
  *count = base + (part < rem ? 1 : 0);
  *first = part * base + std::min<long>(part, rem);
}

static long g_slab_min_cells = 1L << 26;  // volumes below this are solved on one GPU even when more are bound
static unsigned long long g_epoch = 0;    // fan-outs so far: ranks of a sub-group agree on barrier numbers through it

static int solve3d_one(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params& P, const double* T_init,
                       double* T_out, double* qL, double* qR, keffb200_result* out) {
  Ctx& c = ctx();
  Dom D{nx, ny, nz, 0, nz};
  KB_TRY(stage_mask(solid, false, D));
  const size_t n = (size_t)nx * ny * nz;
  double* dT = nullptr;
  if (T_init) {
    // q is free until the first iteration: use it to stage the warm start
    dT = c.q + (size_t)nx * ny;
    KB_CUDA(cudaMemcpyAsync(dT, T_init, n * 8, cudaMemcpyHostToDevice, c.st));
  }
  KB_TRY(solve_grid(c.mask, D, P, dT, nullptr, qL ? c.qL : nullptr, qR ? c.qR : nullptr, out));
  if (T_out) KB_CUDA(cudaMemcpyAsync(T_out, c.x + (size_t)nx * ny, n * 8, cudaMemcpyDeviceToHost, c.st));
  const size_t rows = (size_t)ny * nz;
  if (qL) KB_CUDA(cudaMemcpyAsync(qL, c.qL, rows * 8, cudaMemcpyDeviceToHost, c.st));
  if (qR) KB_CUDA(cudaMemcpyAsync(qR, c.qR, rows * 8, cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  return 0;
}

// One volume in g z-slabs on the first g GPUs of this process (z is the slowest index, Keff3D/keff3D.h:691: a
// slab is a contiguous range of the caller's arrays).  Rank i stages its planes plus one structure plane either
// side, joins the collective solve and copies its part of the outputs back.
static int solve3d_slabs(int g, const uint8_t* solid, int nx, int ny, int nz, const keffb200_params& P, const double* T_init,
                         double* T_out, double* qL, double* qR, keffb200_result* out) {
  Ctx* T = ctx_table();
  const int world = T[0].link.world;
  g_epoch++;
  std::vector<keffb200_result> res(g);
  const int rc = fan_out(g, [&](int i) -> int {
    Ctx& c = ctx();
    long z0 = 0, nzl = 0;
    share(nz, i, g, &z0, &nzl);
    Dom D{nx, ny, (int)nzl, (int)z0, nz};
    const size_t plane = (size_t)nx * ny, n = plane * nzl;
    c.link.world = g;  // the group of this solve: the first g ranks
    c.link.bar_seq = g_epoch << 8;
    int r = ensure_ws(D);
    if (r == 0) {
      const long zlo = std::max<long>(z0 - 1, 0), zhi = std::min<long>(z0 + nzl + 1, nz);
      if (cudaMemsetAsync(c.mask, 0, plane, c.st) != cudaSuccess ||
          cudaMemsetAsync(c.mask + plane * (nzl + 1), 0, plane, c.st) != cudaSuccess ||
          cudaMemcpyAsync(c.mask + plane * (zlo - (z0 - 1)), solid + plane * zlo, plane * (zhi - zlo), cudaMemcpyHostToDevice,
                          c.st) != cudaSuccess)
        r = fail(KEFFB200_ECUDA, "staging slab %d failed: %s", i, cudaGetErrorString(cudaGetLastError()));
    }
    double* dT = nullptr;
    if (r == 0 && T_init) {
      dT = c.q + plane;
      if (cudaMemcpyAsync(dT, T_init + plane * z0, n * 8, cudaMemcpyHostToDevice, c.st) != cudaSuccess)
        r = fail(KEFFB200_ECUDA, "staging the warm start of slab %d failed", i);
    }
    // every rank must enter the collective solve, also one whose staging failed (its result is void anyway)
    c.dist_active = true;
    const int rs = solve_grid(c.mask, D, P, dT, nullptr, qL ? c.qL : nullptr, qR ? c.qR : nullptr, &res[i]);
    c.dist_active = false;
    c.link.world = world;
    if (r == 0) r = rs;
    if (r != 0) return r;
    if (T_out) KB_CUDA(cudaMemcpyAsync(T_out + plane * z0, c.x + plane, n * 8, cudaMemcpyDeviceToHost, c.st));
    const size_t rows = (size_t)ny * nzl;
    if (qL) KB_CUDA(cudaMemcpyAsync(qL + (size_t)ny * z0, c.qL, rows * 8, cudaMemcpyDeviceToHost, c.st));
    if (qR) KB_CUDA(cudaMemcpyAsync(qR + (size_t)ny * z0, c.qR, rows * 8, cudaMemcpyDeviceToHost, c.st));
    KB_CUDA(cudaStreamSynchronize(c.st));
    return 0;
  });
  if (rc != 0) return rc;
  *out = res[0];
  for (int i = 1; i < g; i++) out->gpu_ms = std::max(out->gpu_ms, res[i].gpu_ms);
  return 0;
}

// how many slabs a volume of nz planes gets on up to `avail` GPUs: the largest count that divides nz (equal slabs
// keep the multigrid path; with an even plane count per slab where possible)
static int slab_count(int nz, int avail) {
  int best = 1;
  for (int g = 2; g <= avail; g++)
    if (nz % g == 0 && nz / g >= 4) best = g;
  for (int g = best; g >= 2; g--)
    if (nz % g == 0 && (nz / g) % 2 == 0 && 2 * g > best) return g;
  return best;
}

extern "C" {

int keffb200_solve3d(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* T_init,
                     double* T_out, double* qL, double* qR, keffb200_result* out) {
  KB_TRY(need_ready());
  if (!solid || !out) return fail(KEFFB200_EINVAL, "null argument");
  const keffb200_params P = norm_params(p);
  if (nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", nx, ny, nz);
  if (g_nctx > 1 && ctx_current() == nullptr && ctx_table()[0].link.p2p && (long)nx * ny * nz >= g_slab_min_cells) {
    const int g = slab_count(nz, g_nctx);
    if (g > 1) return solve3d_slabs(g, solid, nx, ny, nz, P, T_init, T_out, qL, qR, out);
  }
  return solve3d_one(solid, nx, ny, nz, P, T_init, T_out, qL, qR, out);
}

void keffb200_set_slab_min_cells(long cells) { g_slab_min_cells = cells < 0 ? (1L << 26) : cells; }

int keffb200_solve2d(const uint8_t* solid, int nx, int ny, const keffb200_params* p, const double* T_init,
                     double* T_out, double* qL, double* qR, keffb200_result* out) {
  return keffb200_solve3d(solid, nx, ny, 1, p, T_init, T_out, qL, qR, out);
}

int keffb200_solve3d_dev(const uint8_t* d_solid, int nx, int ny, int nz, const keffb200_params* p,
                         const double* d_T_init, double* d_T_out, keffb200_result* out) {
  KB_TRY(need_ready());
  if (!d_solid || !out) return fail(KEFFB200_EINVAL, "null argument");
  KB_TRY(check_dev_ptr(d_solid, "d_solid"));
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz, 0, nz};
  if (nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", nx, ny, nz);
  KB_TRY(stage_mask(d_solid, true, D));
  return solve_grid(ctx().mask, D, P, d_T_init, d_T_out, nullptr, nullptr, out);
}

int keffb200_solve3d_slab_dev(const uint8_t* d_solid_halo, int nx, int ny, int nz_local, int z0, int nz_global,
                              const keffb200_params* p, double* d_T_out, keffb200_result* out) {
  KB_TRY(need_ready());
  if (!d_solid_halo || !out) return fail(KEFFB200_EINVAL, "null argument");
  KB_TRY(check_dev_ptr(d_solid_halo, "d_solid_halo"));
  if (z0 < 0 || nz_local < 1 || z0 + nz_local > nz_global) return fail(KEFFB200_EINVAL, "bad slab extent");
  Ctx& c = ctx();
  const bool group = c.nccl.world > 1 && (c.nccl.comm || c.link.p2p);
  // a partial slab without a communicator would be solved with adiabatic cut faces and report a wrong Keff
  if ((z0 > 0 || nz_local < nz_global) && !group)
    return fail(KEFFB200_ESTATE, "partial slab [%d, %d) of %d planes, but keffb200_dist_init has not been called", z0,
                z0 + nz_local, nz_global);
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz_local, z0, nz_global};
  KB_TRY(ensure_ws(D));
  c.dist_active = group;
  const int rc = solve_grid(d_solid_halo, D, P, nullptr, d_T_out, nullptr, nullptr, out);
  c.dist_active = false;
  return rc;
}

int keffb200_solve2d_batch_dev(const uint8_t* d_solid, int n_img, int nx, int ny, const keffb200_params* p,
                               double* d_T_out, keffb200_result* out) {
  KB_TRY(need_ready());
  if (!d_solid || !out || n_img < 1) return fail(KEFFB200_EINVAL, "bad argument");
  KB_TRY(check_dev_ptr(d_solid, "d_solid"));
  return batch2d_solve_dev(d_solid, n_img, nx, ny, norm_params(p), d_T_out, out);
}

static int solve2d_batch_one(const uint8_t* solid, int n_img, int nx, int ny, const keffb200_params& P, double* T_out,
                             keffb200_result* out) {
  Ctx& c = ctx();
  const size_t n = (size_t)nx * ny, tot = n * n_img;
  if (tot > c.cap_stage) {
    if (c.stage) cudaFree(c.stage);
    c.stage = nullptr;
    c.cap_stage = 0;
    KB_CUDA(cudaMalloc(&c.stage, tot));
    c.cap_stage = tot;
  }
  if (T_out && tot > c.cap_stageT) {
    if (c.stageT) cudaFree(c.stageT);
    c.stageT = nullptr;
    c.cap_stageT = 0;
    if (cudaMalloc(&c.stageT, tot * 8) != cudaSuccess) return fail(KEFFB200_ENOMEM, "cannot allocate batch temperature output");
    c.cap_stageT = tot;
  }
  KB_TRY(batch2d_solve_host(solid, c.stage, n_img, nx, ny, P, T_out ? c.stageT : nullptr, out));
  if (T_out) {
    KB_CUDA(cudaMemcpyAsync(T_out, c.stageT, tot * 8, cudaMemcpyDeviceToHost, c.st));
    KB_CUDA(cudaStreamSynchronize(c.st));
  }
  return 0;
}

// Images are independent (the reference loops over them, Keff2D/keff2D.cpp:341): with several GPUs bound every GPU
// takes a contiguous share of the batch on its own host thread -- no communication, results land in place.
int keffb200_solve2d_batch(const uint8_t* solid, int n_img, int nx, int ny, const keffb200_params* p, double* T_out,
                           keffb200_result* out) {
  KB_TRY(need_ready());
  if (!solid || !out || n_img < 1) return fail(KEFFB200_EINVAL, "bad argument");
  const keffb200_params P = norm_params(p);
  const size_t n = (size_t)nx * ny;
  const int g = (g_nctx > 1 && ctx_current() == nullptr) ? std::min<long>(g_nctx, std::max(1, n_img / 64)) : 1;
  if (g <= 1) return solve2d_batch_one(solid, n_img, nx, ny, P, T_out, out);
  return fan_out(g, [&](int i) -> int {
    long first = 0, count = 0;
    share(n_img, i, g, &first, &count);
    return solve2d_batch_one(solid + first * n, (int)count, nx, ny, P, T_out ? T_out + first * n : nullptr, out + first);
  });
}

// ---- pieces ------------------------------------------------------------------------------------
int keffb200_apply(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* xin,
                   double* y, double* b) {
  KB_TRY(need_ready());
  if (!solid || !xin || !y) return fail(KEFFB200_EINVAL, "null argument");
  Ctx& c = ctx();
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz, 0, nz};
  if (nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", nx, ny, nz);
  KB_TRY(stage_mask(solid, false, D));
  const size_t plane = (size_t)nx * ny, n = plane * nz;
  const Coef C = make_coef(nx, ny, nz, P);
  k_make_code<<<c.sms * 8, 256, 0, c.st>>>(c.mask + plane, c.code, D, 0, 0);
  KB_LAUNCHED();
  KB_CUDA(cudaMemsetAsync(c.p, 0, plane * (nz + 2) * 8, c.st));
  KB_CUDA(cudaMemsetAsync(c.x, 0, plane * (nz + 2) * 8, c.st));
  KB_CUDA(cudaMemcpyAsync(c.p + plane, xin, n * 8, cudaMemcpyHostToDevice, c.st));
  KB_CUDA(cudaMemsetAsync(c.S, 0, sizeof(Scal), c.st));
  Stencil s;
  KB_TRY(stencil_setup(s, D, C, P.flags));
  KB_TRY(stencil_launch<MODE_APPLY>(s, 1, c.q + plane, nullptr));
  KB_CUDA(cudaMemcpyAsync(y, c.q + plane, n * 8, cudaMemcpyDeviceToHost, c.st));
  if (b) {
    // b = r of the INIT mode evaluated at x = 0
    KB_TRY(stencil_launch<MODE_INIT>(s, 0, c.r + plane, c.q + plane));
    KB_CUDA(cudaMemcpyAsync(b, c.r + plane, n * 8, cudaMemcpyDeviceToHost, c.st));
  }
  KB_CUDA(cudaStreamSynchronize(c.st));
  return 0;
}

int keffb200_flux(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* T, double* qL,
                  double* qR, keffb200_result* out) {
  KB_TRY(need_ready());
  if (!solid || !T || !out) return fail(KEFFB200_EINVAL, "null argument");
  Ctx& c = ctx();
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz, 0, nz};
  if (nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", nx, ny, nz);
  KB_TRY(stage_mask(solid, false, D));
  const size_t plane = (size_t)nx * ny, n = plane * nz;
  const long rows = (long)ny * nz;
  const Coef C = make_coef(nx, ny, nz, P);
  k_make_code<<<c.sms * 8, 256, 0, c.st>>>(c.mask + plane, c.code, D, 0, 0);
  KB_LAUNCHED();
  KB_CUDA(cudaMemsetAsync(c.x, 0, plane * (nz + 2) * 8, c.st));
  KB_CUDA(cudaMemcpyAsync(c.x + plane, T, n * 8, cudaMemcpyHostToDevice, c.st));
  KB_CUDA(cudaMemsetAsync(c.S, 0, sizeof(Scal), c.st));
  const int gf = (int)std::min<long>((rows + 255) / 256, (long)c.sms * 4);
  k_flux<<<gf, 256, 0, c.st>>>(c.x + plane, c.code, C, D, c.qL, c.qR, c.partials, nullptr);
  KB_LAUNCHED();
  KB_TRY(reduce_step<RS_FLUX>(gf, 1));
  Stencil s;
  KB_TRY(stencil_setup(s, D, C, P.flags));
  KB_TRY(stencil_launch<MODE_RESID>(s, 0, nullptr, nullptr));
  KB_TRY(reduce_step<RS_RESID>(s.grid, 0));
  KB_CUDA(cudaMemcpyAsync(c.hS, c.S, sizeof(Scal), cudaMemcpyDeviceToHost, c.st));
  if (qL) KB_CUDA(cudaMemcpyAsync(qL, c.qL, rows * 8, cudaMemcpyDeviceToHost, c.st));
  if (qR) KB_CUDA(cudaMemcpyAsync(qR, c.qR, rows * 8, cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  memset(out, 0, sizeof *out);
  out->q_left = c.hS->ql;
  out->q_right = c.hS->qr;
  out->keff = (c.hS->ql + c.hS->qr) / 2 / (P.tr - P.tl);
  out->energy_residual = c.hS->tmp[2];
  out->rel_residual = sqrt(c.hS->tmp[1]);  // absolute 2-norm here: no ||b|| without an INIT pass
  return 0;
}

int keffb200_precond(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* rin,
                     double* zout) {
  KB_TRY(need_ready());
  if (!solid || !rin || !zout) return fail(KEFFB200_EINVAL, "null argument");
  Ctx& c = ctx();
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz, 0, nz};
  if (nx < 2 || ny < 1 || nz < 1) return fail(KEFFB200_EINVAL, "grid %dx%dx%d too small", nx, ny, nz);
  if (!mg_eligible(D, P.flags)) return fail(KEFFB200_EINVAL, "grid %dx%dx%d does not take the multigrid path", nx, ny, nz);
  KB_TRY(stage_mask(solid, false, D));
  const size_t plane = (size_t)nx * ny, n = plane * nz;
  const Coef C = make_coef(nx, ny, nz, P);
  k_make_code<<<c.sms * 8, 256, 0, c.st>>>(c.mask + plane, c.code, D, 0, 0);
  KB_LAUNCHED();
  KB_CUDA(cudaMemsetAsync(c.r, 0, plane * (nz + 2) * 8, c.st));
  KB_CUDA(cudaMemcpyAsync(c.r + plane, rin, n * 8, cudaMemcpyHostToDevice, c.st));
  KB_CUDA(cudaMemsetAsync(c.S, 0, sizeof(Scal), c.st));
  Mg mg;
  KB_TRY(mg_setup(mg, D, C));
  KB_TRY(mg_make_u(mg, c.r + plane));
  KB_TRY(mg_vcycle(mg, c.z));
  KB_CUDA(cudaGetLastError());
  // widen the FP32 result on the host side of the hook (q is free here)
  std::vector<float> zf(n);
  KB_CUDA(cudaMemcpyAsync(zf.data(), c.z, n * sizeof(float), cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  for (size_t i = 0; i < n; i++) zout[i] = zf[i];
  return 0;
}

int keffb200_bench_apply_dev(const uint8_t* d_solid, int nx, int ny, int nz, const keffb200_params* p, int reps,
                             float* ms_per_apply) {
  KB_TRY(need_ready());
  if (!d_solid || !ms_per_apply || reps < 1) return fail(KEFFB200_EINVAL, "bad argument");
  Ctx& c = ctx();
  const keffb200_params P = norm_params(p);
  Dom D{nx, ny, nz, 0, nz};
  KB_TRY(stage_mask(d_solid, true, D));
  const size_t plane = (size_t)nx * ny;
  const Coef C = make_coef(nx, ny, nz, P);
  k_make_code<<<c.sms * 8, 256, 0, c.st>>>(c.mask + plane, c.code, D, 0, 0);
  KB_LAUNCHED();
  KB_CUDA(cudaMemsetAsync(c.p, 0, plane * (nz + 2) * 8, c.st));
  k_init_field<<<c.sms * 8, 256, 0, c.st>>>(c.p + plane, D, P.tl, P.tr, 0);
  KB_LAUNCHED();
  KB_CUDA(cudaMemsetAsync(c.S, 0, sizeof(Scal), c.st));
  Stencil s;
  KB_TRY(stencil_setup(s, D, C, P.flags));
  int which = 1;
  if (P.flags & KEFFB200_F_APPLY_F32) {  // the form the multigrid iteration uses: p stored in FP32
    if (!s.p32_ok) return fail(KEFFB200_EINVAL, "KEFFB200_F_APPLY_F32 needs the TMA path and nx %% 4 == 0");
    KB_CUDA(cudaMemcpyAsync(c.r, c.p, plane * (nz + 2) * 8, cudaMemcpyDeviceToDevice, c.st));
    k_f64_to_f32<<<c.sms * 8, 256, 0, c.st>>>(c.r, reinterpret_cast<float*>(c.p), (long)(plane * (nz + 2)));
    KB_LAUNCHED();
    which = 2;
  }
  for (int i = 0; i < 3; i++) KB_TRY(stencil_launch<MODE_APPLY>(s, which, c.q + plane, nullptr));
  KB_CUDA(cudaEventRecord(c.ev0, c.st));
  for (int i = 0; i < reps; i++) KB_TRY(stencil_launch<MODE_APPLY>(s, which, c.q + plane, nullptr));
  KB_CUDA(cudaEventRecord(c.ev1, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  float ms = 0;
  KB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  *ms_per_apply = ms / reps;
  return 0;
}

// ---- region timer: CUDA events on the library's compute stream -------------------------------
int keffb200_timer_start(void) {
  KB_TRY(need_ready());
  for (int i = 0; i < g_nctx; i++) {
    Ctx& c = ctx_table()[i];
    if (g_nctx > 1) KB_CUDA(cudaSetDevice(c.device));
    KB_CUDA(cudaEventRecord(c.evt0, c.st));
  }
  if (g_nctx > 1) KB_CUDA(cudaSetDevice(ctx_table()[0].device));
  return 0;
}

int keffb200_timer_stop(float* ms) {
  KB_TRY(need_ready());
  if (!ms) return fail(KEFFB200_EINVAL, "null argument");
  *ms = 0.f;
  for (int i = 0; i < g_nctx; i++) {  // several GPUs: the longest of the regions
    Ctx& c = ctx_table()[i];
    if (g_nctx > 1) KB_CUDA(cudaSetDevice(c.device));
    KB_CUDA(cudaEventRecord(c.evt1, c.st));
    KB_CUDA(cudaEventSynchronize(c.evt1));
    float t = 0.f;
    KB_CUDA(cudaEventElapsedTime(&t, c.evt0, c.evt1));
    *ms = std::max(*ms, t);
  }
  if (g_nctx > 1) KB_CUDA(cudaSetDevice(ctx_table()[0].device));
  return 0;
}

// Order the library's compute stream behind the work already queued on `stream` (a cudaStream_t; NULL = the
// legacy default stream): call it before a *_dev entry point whose inputs another stream is still writing.
int keffb200_wait_stream(void* stream) {
  KB_TRY(need_ready());
  Ctx& c = ctx();
  KB_CUDA(cudaEventRecord(c.evy, (cudaStream_t)stream));
  KB_CUDA(cudaStreamWaitEvent(c.st, c.evy, 0));
  return 0;
}

// ---- synthetic structures (bench / test utility) -----------------------------------------------
__global__ void k_gen_bernoulli(uint8_t* out, size_t n, unsigned long long seed, unsigned long long thr,
                                unsigned long long first) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long z = ((first + i) ^ seed) + 0x9E3779B97F4A7C15ull;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = z < thr;
  }
}

int keffb200_gen_bernoulli_dev(uint8_t* d_out, size_t n, unsigned long long first_index, unsigned long long seed,
                               double p) {
  KB_TRY(need_ready());
  if (!d_out) return fail(KEFFB200_EINVAL, "null argument");
  const double s = p * 18446744073709551616.0;
  const unsigned long long thr = s >= 18446744073709551615.0 ? ~0ull : (unsigned long long)s;
  k_gen_bernoulli<<<ctx().sms * 8, 256, 0, ctx().st>>>(d_out, n, seed, thr, first_index);
  KB_LAUNCHED();
  KB_CUDA(cudaGetLastError());
  KB_CUDA(cudaStreamSynchronize(ctx().st));
  return 0;
}

// ---- z-slab communicator -----------------------------------------------------------------------
static int nccl_load() {
  Nccl& n = ctx().nccl;
  if (n.h) return 0;
  n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) return fail(KEFFB200_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define KB_SYM(field, name)                                                  \
  *(void**)(&n.field) = dlsym(n.h, name);                                    \
  if (!n.field) return fail(KEFFB200_ENCCL, "libnccl lacks symbol %s", name)
  KB_SYM(GetUniqueId, "ncclGetUniqueId");
  KB_SYM(CommInitRank, "ncclCommInitRank");
  KB_SYM(CommDestroy, "ncclCommDestroy");
  KB_SYM(AllReduce, "ncclAllReduce");
  KB_SYM(AllGather, "ncclAllGather");
  KB_SYM(Send, "ncclSend");
  KB_SYM(Recv, "ncclRecv");
  KB_SYM(GroupStart, "ncclGroupStart");
  KB_SYM(GroupEnd, "ncclGroupEnd");
  KB_SYM(GetErrorString, "ncclGetErrorString");
#undef KB_SYM
  return 0;
}

int keffb200_comm_id(void* id128) {
  KB_TRY(nccl_load());
  return nccl_check(ctx().nccl.GetUniqueId(id128), "ncclGetUniqueId");
}

// One process per GPU: map every rank's mailbox through CUDA IPC (the handles travel over the NCCL communicator
// that was just created -- its only job besides being the fallback).  Every rank must succeed, or all stay on NCCL.
static int link_bootstrap_ipc(Ctx& c, int rank, int world) {
  Nccl& n = c.nccl;
  Link& L = c.link;
  if (getenv("KEFFB200_SLAB_NCCL")) return 0;  // experiments: keep the NCCL data path
  if (world > KB_MAX_RANKS) return 0;
  L.rank = rank;
  L.world = world;
  KB_TRY(link_alloc_mail(c));
  struct Boot {
    unsigned char handle[64];
    long long pid;
    long long ok;
  };
  std::vector<Boot> hb(world);
  Boot mine;
  memset(&mine, 0, sizeof mine);
  cudaIpcMemHandle_t h;
  mine.ok = cudaIpcGetMemHandle(&h, L.mail[rank]) == cudaSuccess;
  if (mine.ok) memcpy(mine.handle, &h, sizeof h);
  else cudaGetLastError();
  mine.pid = (long long)getpid();
  Boot* db = nullptr;
  KB_CUDA(cudaMalloc(&db, sizeof(Boot) * world));
  KB_CUDA(cudaMemcpyAsync(db + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c.st));
  KB_TRY(nccl_check(n.AllGather(db + rank, db, sizeof(Boot), /*ncclInt8*/ 0, n.comm, c.st), "ncclAllGather"));
  KB_CUDA(cudaMemcpyAsync(hb.data(), db, sizeof(Boot) * world, cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  bool ok = true;
  for (int r = 0; r < world && ok; r++) {
    if (r == rank) continue;
    void* ptr = nullptr;
    if (!hb[r].ok || hb[r].pid == mine.pid || link_ipc_open(L.ipc_mail[r], hb[r].handle, 0, &ptr) != 0) ok = false;
    else L.mail[r] = (Mail*)ptr;
  }
  double* flag = (double*)db;
  const double mine_ok = ok ? 1.0 : 0.0;
  KB_CUDA(cudaMemcpyAsync(flag, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, c.st));
  KB_TRY(nccl_check(n.AllReduce(flag, flag, 1, /*ncclFloat64*/ 8, /*ncclMin*/ 3, n.comm, c.st), "ncclAllReduce"));
  double all_ok = 0;
  KB_CUDA(cudaMemcpyAsync(&all_ok, flag, sizeof(double), cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  cudaFree(db);
  L.p2p = all_ok > 0.5;
  L.ipc = L.p2p;
  if (!L.p2p && getenv("KEFFB200_VERBOSE"))
    fprintf(stderr, "keffb200: rank %d: peer-memory link unavailable (%s), z-slabs stay on NCCL\n", rank, c.err.c_str());
  return 0;
}

int keffb200_dist_init(int rank, int world, const void* id128) {
  KB_TRY(need_ready());
  Ctx& c = ctx();
  Nccl& n = c.nccl;
  if (world < 1 || rank < 0 || rank >= world) return fail(KEFFB200_EINVAL, "bad rank/world");
  if (g_nctx > 1) return fail(KEFFB200_ESTATE, "keffb200_dist_init joins one-GPU processes; this process drives %d GPUs", g_nctx);
  n.rank = rank;
  n.world = world;
  if (world == 1) return 0;
  KB_TRY(nccl_load());
  Id128 id;
  memcpy(id.b, id128, 128);
  KB_CUDA(cudaSetDevice(c.device));
  KB_TRY(nccl_check(n.CommInitRank(&n.comm, world, id, rank), "ncclCommInitRank"));
  return link_bootstrap_ipc(c, rank, world);
}

// 1 = z-slab halos and reductions travel through peer memory, 0 = through NCCL (or no group)
int keffb200_dist_peer_memory(void) { return ctx().link.p2p ? 1 : 0; }

void keffb200_dist_shutdown(void) {
  Ctx& c = ctx();
  Nccl& n = c.nccl;
  if (g_nctx <= 1 && c.ready) {
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    if (c.link.p2p && c.link.ipc) {
      // nobody frees a window before every rank has unmapped it: unmap the workspace windows, meet through the
      // mailboxes (a rank that shuts down alone waits two seconds, then goes on), unmap the mailboxes, free
      Link& L = c.link;
      for (int r = 0; r < KB_MAX_RANKS; r++)
        for (int w = 0; w < W_COUNT; w++) {
          if (L.ipcw[r][w].ptr) cudaIpcCloseMemHandle(L.ipcw[r][w].ptr);
          L.ipcw[r][w].ptr = nullptr;
        }
      if (link_barrier(c, 2000000000ull) == 0) cudaStreamSynchronize(c.st);
    }
    link_free(c);
  }
  if (n.comm && n.CommDestroy) n.CommDestroy(n.comm);
  n.comm = nullptr;
  n.rank = 0;
  n.world = 1;
}

}  // extern "C"
