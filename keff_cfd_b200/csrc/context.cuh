// context.cuh -- per-GPU state of libkeffb200.  One context per GPU: a process bound to one GPU (the
// torchrun layout, keffb200_init) has one; a process that drives several (keffb200_init_multi) has one per
// device, each used by one host thread at a time -- the thread's current context is a thread-local pointer.
#pragma once
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/keffb200.h"
#include "common.cuh"

namespace kb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NCCL is bound at run time (dlopen) so that the library loads on machines without it and never
// forces a second NCCL build into a process that already carries torch's.
struct Id128 {
  char b[128];
};
struct Nccl {
  void* h = nullptr;
  void* comm = nullptr;
  int rank = 0, world = 1;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value, 128 B*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
// Peer-memory link of a z-slab group (one per context).  Built once (keffb200_init_multi inside one process,
// keffb200_dist_init between processes: the mailboxes are mapped into every rank), refreshed before every
// slab solve (the workspace windows of the neighbours, published through the mailboxes).
struct Link {
  bool p2p = false;  // mailboxes mapped: halo planes and all-reduces travel as NVLink stores, NCCL is not on the path
  bool ipc = false;  // the peers are other processes: windows are CUDA IPC mappings
  int rank = 0, world = 1;
  Mail* mail[KB_MAX_RANKS] = {};         // every rank's mailbox as mapped here (mail[rank] = my own allocation)
  unsigned long long* ar_seq = nullptr;  // device: all-reduces done (reset with the mailbox before a solve)
  unsigned long long bar_seq = 0;        // host: barriers done
  unsigned int expect[KB_HALO_FIELDS][2] = {};  // host: arrival counts my halos must have reached before their next use
  unsigned int expect_g = 0;                    // the same for all-gathers (per source rank; every source sends the same)
  // neighbours' windows as mapped here: [0] = rank-1, [1] = rank+1; pool[r] = every rank's level pool
  char* nb[2][W_COUNT] = {};
  int nb_nz[2] = {0, 0};
  char* pool[KB_MAX_RANKS] = {};
  // CUDA IPC mappings kept open between solves (re-opened when a peer re-allocates)
  struct IpcMap {
    unsigned char handle[64];
    unsigned long long gen = 0;
    void* ptr = nullptr;
  } ipcw[KB_MAX_RANKS][W_COUNT], ipc_mail[KB_MAX_RANKS];
  unsigned long long win_gen[W_COUNT] = {};  // my windows' allocation generations
  // Windows other processes have mapped must not be freed under them (CUDA IPC: undefined behaviour): a window that
  // has to grow is retired instead and freed at shutdown, after every rank has closed its mappings.
  std::vector<void*> retired;
  WinTable* h_win = nullptr;  // pinned staging of window tables
};

constexpr int KB_READY_SLOTS = 32;  // chunks of a streamed host batch

struct Ctx {
  bool ready = false;
  int device = -1, sms = 0;
  cudaStream_t st = nullptr, st2 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr, evx = nullptr, evy = nullptr, evp[3] = {nullptr, nullptr, nullptr};
  PFN_encodeTiled encode = nullptr;
  long launches = 0;
  std::string err;
  // grid-solver workspace (grow-only)
  size_t cap_field = 0;  // doubles per field incl. halo planes
  double *x = nullptr, *r = nullptr, *p = nullptr, *q = nullptr;
  float* z = nullptr;    // preconditioned residual of the multigrid path (FP32, no halo planes)
  size_t cap_z = 0;
  float* u = nullptr;    // r / a_P (FP32, pitched rows, one halo plane below and above)
  size_t cap_u = 0;
  float* xf[2] = {nullptr, nullptr};  // iterates of the voxel-grid smoothing sweeps (same layout as u; nu >= 2)
  size_t cap_xf = 0;
  float* mg_pool = nullptr;  // stored multigrid levels (grow-only)
  size_t cap_mg = 0;         // floats
  double* mg_inv = nullptr;  // dense inverse of the coarsest level
  float* mg_tab = nullptr;   // 1/d and d by (position class, code byte)
  uint8_t *code = nullptr, *mask = nullptr;
  size_t cap_aux = 0;  // doubles in qL/qR
  double *qL = nullptr, *qR = nullptr;
  double* partials = nullptr;
  Scal* S = nullptr;
  Scal* hS = nullptr;  // pinned
  // batch workspace
  size_t cap_batch = 0;
  void* batch_ws = nullptr;
  size_t cap_stage = 0;  // host-path staging of image batches (+ their temperature maps)
  uint8_t* stage = nullptr;
  size_t cap_stageT = 0;
  double* stageT = nullptr;
  Nccl nccl;
  Link link;
  bool dist_active = false;  // true only inside a slab solve: halo exchange + all-reduces enabled
  // scalar records of the last two convergence checks (pinned): the host reads check k while check k+1 is queued
  Scal* hring = nullptr;
  unsigned int* h_ready = nullptr;  // page-locked: image counts published behind the chunks of a streamed host batch
  cudaEvent_t evb[2] = {nullptr, nullptr};
  // per-device function attributes already applied (dynamic shared memory sizes)
  bool attr_batch = false, attr_ws = false;
};

constexpr int KB_MAX_GPUS = 16;
inline Ctx* ctx_table() {
  static Ctx t[KB_MAX_GPUS];
  return t;
}
inline Ctx*& ctx_current() {
  static thread_local Ctx* cur = nullptr;
  return cur;
}
// the calling thread's context (context 0 unless a fan-out worker bound another one)
inline Ctx& ctx() {
  Ctx* c = ctx_current();
  return c ? *c : ctx_table()[0];
}

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  ctx().err = buf;
  return code;
}

#define KB_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess)                                                                                  \
      return kb::fail(KEFFB200_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define KB_TRY(call)        \
  do {                      \
    int rc_ = (call);       \
    if (rc_ != 0) return rc_; \
  } while (0)

// KEFFB200_DEBUG_SYNC=1: synchronise after every launch and name the first one that faults
inline void debug_sync(const char* file, int line) {
  static const bool on = getenv("KEFFB200_DEBUG_SYNC") != nullptr;
  if (!on) return;
  cudaError_t e = cudaStreamSynchronize(ctx().st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "keffb200: launch before %s:%d failed: %s\n", file, line, cudaGetErrorString(e));
}

#define KB_LAUNCHED() (kb::ctx().launches++, kb::debug_sync(__FILE__, __LINE__))

// solver.cu
keffb200_params norm_params(const keffb200_params* p);
Coef make_coef(int nx, int ny, int nzg, const keffb200_params& P);

constexpr int PARTIAL_SLOTS = 1 << 15;  // per-CTA partial sums, 3 doubles each

}  // namespace kb
