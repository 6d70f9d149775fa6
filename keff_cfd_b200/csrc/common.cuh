// common.cuh -- shared device/host helpers of the B200-native Keff hot path.
//
// Discretisation being evaluated matrix-free (never assembled): the finite-volume 5-/7-point
// operator of Keff2D/keff2D.h:482-582 and Keff3D/keff3D.h:666-763 --
//   face conductance  c = k_face * A_face / d,  k_face = harmonic mean of the two cell conductivities
//   fixed-temperature x walls half a cell away:  c_wall = k_P * A_face / (dx/2)
//   a_P = sum of face conductances (+ c_wall),   b_P = TL*c_wall (x = 0) + TR*c_wall (x = nx-1)
// With two phases a face conductance takes 3 values per direction, so the whole operator is a
// function of a one-byte "code" per cell: bit0 = phase of the cell, bits1..6 = "neighbour across
// face {W,E,S,N,F,B} has the other phase".  Domain-boundary faces are decided from coordinates.
#pragma once
#include <cuda.h>          // CUtensorMap type only; the encoder entry point is fetched at run time
#include <cuda_runtime.h>
#include <stdint.h>

namespace kb {

enum : unsigned { C_PH = 1, C_W = 2, C_E = 4, C_S = 8, C_N = 16, C_F = 32, C_B = 64 };
// slot names follow the reference: W = x-1, E = x+1, S = y+1, N = y-1, F = z-1, B = z+1

struct Coef {
  double same[3][2];  // [direction x,y,z][phase]: g_dir * k_phase
  double diff[3];     // g_dir * 2 ks kf / (ks + kf)
  double wall[2];     // [phase]: k_phase * A_x / (dx/2)
  double tl, tr;
};

struct Dom {
  int nx, ny, nz;  // local extent (nz = planes held by this rank)
  int z0, nzg;     // first global plane of this slab, global plane count
};

// device-resident solver scalars (one per solve)
struct Scal {
  double rz[2];   // r.z, double-buffered by iteration parity
  double pq;      // p.Ap of the current iteration
  double rr, bb;  // ||r||^2, ||b||^2
  double tmp[4];  // reduction outputs before an NCCL all-reduce (peer-memory all-reduces never leave the kernel)
  double ql, qr;  // boundary flux sums
  double thr2;    // squared relative-residual threshold
  long long it;   // iterations completed
  int done;       // sticky convergence flag
  int fault;      // a peer-memory wait timed out (a neighbour never delivered): the solve is void
};

// ---- peer memory (z-slabs over NVLink): mailboxes, flags, system-scope accesses -------------------
// Every rank owns one Mail block in its own HBM; every other rank of the slab group has it mapped
// (cudaDeviceEnablePeerAccess inside one process, CUDA IPC between processes) and writes into it
// with ordinary stores.  Nothing is ever read remotely.
constexpr int KB_MAX_RANKS = 16;
constexpr int KB_HALO_FIELDS = 160;  // p, x, u, two iterates, 6 arrays per stored level
struct MailSlot {
  double v[4];
  unsigned long long seq;
  unsigned long long pad[3];
};
// windows: the buffers of a rank that its peers store into
enum { W_X = 0, W_P, W_U, W_XF0, W_XF1, W_POOL, W_COUNT };
struct WinTable {                        // what a rank publishes about its workspace before a slab solve
  unsigned long long base[W_COUNT];      // device addresses in the owner's address space
  unsigned char handle[W_COUNT][64];     // the same allocations as CUDA IPC handles (other processes map these)
  unsigned long long gen[W_COUNT];       // bumped whenever the owner re-allocates the window: peers re-map
  long long pid;                         // owner process: peers inside it use `base` directly
  int nz, z0;                            // the owner's slab: plane count, first global plane
};
struct Mail {
  MailSlot ar[2][KB_MAX_RANKS];          // all-reduce contributions: [sequence parity][source rank]
  unsigned int halo[KB_HALO_FIELDS][2];  // cumulative arrival counters: [field][0 = my lower halo, 1 = my upper halo]
  unsigned int gath[KB_MAX_RANKS];       // cumulative arrival counters of all-gathers: [source rank]
  unsigned long long bar[KB_MAX_RANKS];  // barrier: [source rank] = barrier sequence number
  WinTable win[KB_MAX_RANKS];            // [source rank]
};
// halo counter ids
enum { HF_X = 0, HF_P, HF_U, HF_XF0, HF_XF1, HF_LEVEL0 = 8 };  // + 6 * level + {u, t0, t1, e, cz, dinv}
enum { HL_U = 0, HL_T0, HL_T1, HL_E, HL_CZ, HL_DINV };
// what a reduction kernel needs to finish its sum across the ranks (world <= 1: nothing to do)
struct ArLink {
  int world, rank;
  unsigned long long* seq;   // my count of all-reduces so far (device memory; identical on every rank)
  Mail* mail[KB_MAX_RANKS];  // every rank's mailbox as mapped here (mail[rank] is my own)
};

// where a producing kernel stores the boundary planes of its output besides its own memory: the halo planes of
// the z-neighbours (peer addresses, nullptr without a neighbour) and their arrival counters
struct HaloPush {
  char* lo_dst;            // the lower neighbour's upper halo plane of the same array
  char* hi_dst;            // the upper neighbour's lower halo plane
  unsigned int* lo_flag;   // ... and the counters that tell them it has arrived
  unsigned int* hi_flag;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// every thread of the CTA has issued its remote stores: make them visible system-wide, then count one arrival
__device__ __forceinline__ void halo_signal(const HaloPush& H, unsigned int n_lo, unsigned int n_hi) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (H.lo_dst && n_lo) red_add_release_sys(H.lo_flag, n_lo);
    if (H.hi_dst && n_hi) red_add_release_sys(H.hi_flag, n_hi);
  }
}
constexpr unsigned long long KB_WAIT_NS = 4000000000ull;  // a peer that stays silent this long is gone

// Wait until a cumulative arrival counter has reached `expect`; false on timeout.  Counters wrap
// (32 bit), so the comparison is on the signed difference.
__device__ __forceinline__ bool wait_counter(const unsigned int* flag, unsigned int expect) {
  if ((int)(ld_acquire_sys(flag) - expect) >= 0) return true;
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_acquire_sys(flag) - expect) < 0) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > KB_WAIT_NS) return false;
  }
  return true;
}

struct Faces {
  double w, e, n, s, f, b, diag;
};

__device__ __forceinline__ Faces faces_of(const Coef& C, unsigned c, bool x0, bool x1, bool y0, bool y1,
                                          bool z0, bool z1) {
  const bool ph = c & C_PH;
  const double sx = ph ? C.same[0][1] : C.same[0][0];
  const double sy = ph ? C.same[1][1] : C.same[1][0];
  const double sz = ph ? C.same[2][1] : C.same[2][0];
  const double wl = ph ? C.wall[1] : C.wall[0];
  Faces F;
  F.w = x0 ? 0.0 : ((c & C_W) ? C.diff[0] : sx);
  F.e = x1 ? 0.0 : ((c & C_E) ? C.diff[0] : sx);
  F.s = y1 ? 0.0 : ((c & C_S) ? C.diff[1] : sy);
  F.n = y0 ? 0.0 : ((c & C_N) ? C.diff[1] : sy);
  F.f = z0 ? 0.0 : ((c & C_F) ? C.diff[2] : sz);
  F.b = z1 ? 0.0 : ((c & C_B) ? C.diff[2] : sz);
  F.diag = ((F.w + F.e) + (F.n + F.s)) + (F.f + F.b) + ((x0 ? wl : 0.0) + (x1 ? wl : 0.0));
  return F;
}

// a_P alone (the preconditioner / relaxation divisor)
__device__ __forceinline__ double diag_of(const Coef& C, unsigned c, bool x0, bool x1, bool y0, bool y1, bool z0,
                                          bool z1) {
  return faces_of(C, c, x0, x1, y0, y1, z0, z1).diag;
}

// right-hand side b_P (non-zero only in the two wall columns)
__device__ __forceinline__ double rhs_of(const Coef& C, unsigned c, bool x0, bool x1) {
  const double wl = (c & C_PH) ? C.wall[1] : C.wall[0];
  return (x0 ? C.tl * wl : 0.0) + (x1 ? C.tr * wl : 0.0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum of N values per thread; result valid in thread 0.
// `red` must hold 32*N doubles.  Fixed reduction tree -> bit-reproducible.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* red) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < N; j++) v[j] = warp_sum(v[j]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; j++) red[wrp * N + j] = v[j];
  __syncthreads();
  if (wrp == 0) {
#pragma unroll
    for (int j = 0; j < N; j++) {
      double t = lane < nw ? red[lane * N + j] : 0.0;
      v[j] = warp_sum(t);
    }
  }
}

// ---- mbarrier / TMA (cp.async.bulk.tensor) wrappers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "KB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KB_DONE;\n\t"
      "bra KB_WAIT;\n\t"
      "KB_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

}  // namespace kb
