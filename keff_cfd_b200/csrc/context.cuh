// context.cuh -- process-wide state of libkeffb200 (one process per GPU).
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
  bool dist_active = false;  // true only inside a slab solve: halo exchange + all-reduces enabled
};

inline Ctx& ctx() {
  static Ctx c;
  return c;
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

constexpr int PARTIAL_SLOTS = 1 << 15;  // per-CTA partial sums, 3 doubles each

}  // namespace kb
