// peer.cuh -- z-slab communication through peer memory (NVLink 5 / NVSwitch), no library on the data path.
//
// Every rank of a slab group owns a mailbox (common.cuh: Mail) and a workspace; its neighbours have both mapped
// (cudaDeviceEnablePeerAccess when one process drives the GPUs, CUDA IPC when there is one process per GPU).
// A halo plane travels as plain stores into the neighbour's halo plane followed by one release-add on an
// arrival counter in the neighbour's mailbox; the consumer acquires the counter before it touches the plane.
// Counters are cumulative: the host of the consuming rank knows how many arrivals each of its halos must have
// seen before its next use, because every rank runs the same kernel sequence.  Dot products are completed the
// same way (kernels.cuh: ar_sum).  Write-after-read safety: between two pushes into the same halo there is always
// an all-reduce that every rank joins after it has consumed the previous contents.
//
// No reference counterpart (the reference is single-GPU: Keff3D/keff3D.cpp runs ParallelGS3D on one volume).
#pragma once
#include <unistd.h>

#include <condition_variable>
#include <mutex>

#include "context.cuh"

namespace kb {

// ---- device side --------------------------------------------------------------------------------
constexpr int PUSH_THREADS = 256;
// CTAs of a push: a function of the plane size only, so both ends of a link count the same arrivals
inline int push_ctas(size_t bytes) { return (int)std::max<size_t>(1, std::min<size_t>(32, bytes / 32768)); }

template <typename V>
__device__ __forceinline__ void copy_span(const char* src, char* dst, size_t bytes) {
  const size_t n = bytes / sizeof(V);
  const V* s = reinterpret_cast<const V*>(src);
  V* d = reinterpret_cast<V*>(dst);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

// my first plane -> the lower neighbour's upper halo, my last plane -> the upper neighbour's lower halo
__global__ void __launch_bounds__(PUSH_THREADS) k_push_planes(const char* lo_src, char* lo_dst, const char* hi_src,
                                                              char* hi_dst, size_t bytes, int wide, unsigned int* lo_flag,
                                                              unsigned int* hi_flag) {
  if (lo_dst) {
    if (wide) copy_span<uint4>(lo_src, lo_dst, bytes);
    else copy_span<unsigned int>(lo_src, lo_dst, bytes);
  }
  if (hi_dst) {
    if (wide) copy_span<uint4>(hi_src, hi_dst, bytes);
    else copy_span<unsigned int>(hi_src, hi_dst, bytes);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (lo_dst) red_add_release_sys(lo_flag, 1u);
    if (hi_dst) red_add_release_sys(hi_flag, 1u);
  }
}

// my slice of a replicated array -> the same offset of every other rank's copy
struct GatherDst {
  char* dst[KB_MAX_RANKS];          // nullptr for my own rank
  unsigned int* flag[KB_MAX_RANKS];
  int world;
};
__global__ void __launch_bounds__(PUSH_THREADS) k_push_gather(const char* src, const GatherDst G, size_t bytes) {
  for (int r = 0; r < G.world; r++)
    if (G.dst[r]) copy_span<unsigned int>(src, G.dst[r], bytes);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < G.world && G.dst[threadIdx.x]) red_add_release_sys(G.flag[threadIdx.x], 1u);
}

__global__ void k_wait_gather(const unsigned int* gath, int world, int rank, unsigned int expect, Scal* S) {
  const int t = threadIdx.x;
  if (t >= world || t == rank) return;
  if (!wait_counter(&gath[t], expect)) {
    S->fault = 1;
    S->done = 1;
  }
}

// the consumer side of a push that is not fused into the consuming kernel
__global__ void k_wait_halo(const unsigned int* lo_flag, unsigned int lo_expect, const unsigned int* hi_flag,
                            unsigned int hi_expect, Scal* S) {
  bool ok = true;
  if (threadIdx.x == 0 && lo_flag) ok = wait_counter(lo_flag, lo_expect);
  if (threadIdx.x == 1 && hi_flag) ok = wait_counter(hi_flag, hi_expect);
  if (!ok) {
    S->fault = 1;
    S->done = 1;
  }
}

// full barrier of the group; everything the ranks stored into each other before it is visible after it
__global__ void k_barrier(const ArLink ar, unsigned long long seq, unsigned long long timeout_ns, Scal* S) {
  const int t = threadIdx.x;
  if (t >= ar.world) return;
  __threadfence_system();
  st_release_sys(&ar.mail[t]->bar[ar.rank], seq);
  const unsigned long long* mine = &ar.mail[ar.rank]->bar[t];
  const unsigned long long t0 = globaltimer_ns();
  while ((long long)(ld_acquire_sys(mine) - seq) < 0) {
    __nanosleep(128);
    if (globaltimer_ns() - t0 > timeout_ns) {
      S->fault = 1;
      break;
    }
  }
}

// my window table -> slot [rank] of every mailbox
__global__ void k_post_table(const ArLink ar) {
  const WinTable* src = &ar.mail[ar.rank]->win[ar.rank];
  const int words = sizeof(WinTable) / 4;
  for (int r = 0; r < ar.world; r++) {
    if (r == ar.rank) continue;
    unsigned int* d = reinterpret_cast<unsigned int*>(&ar.mail[r]->win[ar.rank]);
    const unsigned int* s = reinterpret_cast<const unsigned int*>(src);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = s[i];
  }
}

// ---- host side ----------------------------------------------------------------------------------
// ranks inside one process meet at a host barrier where a collective needs the host (init, shutdown)
struct HostBarrier {
  std::mutex m;
  std::condition_variable cv;
  int n = 1, count = 0;
  unsigned gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++count == n) {
      count = 0;
      gen++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

inline ArLink link_table(const Link& L) {
  ArLink a;
  memset(&a, 0, sizeof a);
  a.world = L.world;
  a.rank = L.rank;
  a.seq = L.ar_seq;
  for (int r = 0; r < L.world; r++) a.mail[r] = L.mail[r];
  return a;
}

inline int link_alloc_mail(Ctx& c) {
  Link& L = c.link;
  if (L.mail[L.rank]) return 0;
  Mail* m = nullptr;
  KB_CUDA(cudaMalloc(&m, sizeof(Mail)));
  KB_CUDA(cudaMemset(m, 0, sizeof(Mail)));
  KB_CUDA(cudaMalloc(&L.ar_seq, sizeof(unsigned long long)));
  KB_CUDA(cudaMemset(L.ar_seq, 0, sizeof(unsigned long long)));
  KB_CUDA(cudaMallocHost(&L.h_win, sizeof(WinTable) * KB_MAX_RANKS));
  L.mail[L.rank] = m;
  return 0;
}

// unmap everything other processes exported to me (must precede their frees)
inline void link_close_imports(Ctx& c) {
  Link& L = c.link;
  for (int r = 0; r < KB_MAX_RANKS; r++) {
    for (int w = 0; w < W_COUNT; w++) {
      if (L.ipcw[r][w].ptr) cudaIpcCloseMemHandle(L.ipcw[r][w].ptr);
      L.ipcw[r][w].ptr = nullptr;
    }
    if (L.ipc_mail[r].ptr) cudaIpcCloseMemHandle(L.ipc_mail[r].ptr);
    L.ipc_mail[r].ptr = nullptr;
  }
  L.p2p = false;
}

inline void link_free(Ctx& c) {
  Link& L = c.link;
  link_close_imports(c);
  if (L.p2p || L.mail[L.rank]) {
    if (L.mail[L.rank]) cudaFree(L.mail[L.rank]);
    if (L.ar_seq) cudaFree(L.ar_seq);
    if (L.h_win) cudaFreeHost(L.h_win);
  }
  for (void* p : L.retired) cudaFree(p);
  L = Link();
}

inline int link_barrier(Ctx& c, unsigned long long timeout_ns = 4 * KB_WAIT_NS) {
  Link& L = c.link;
  k_barrier<<<1, 32, 0, c.st>>>(link_table(L), ++L.bar_seq, timeout_ns, c.S);
  KB_LAUNCHED();
  return 0;
}

// free (or, while other processes may have it mapped, retire) window w of my workspace before it is re-allocated
inline void link_release_window(Ctx& c, int w, void* p) {
  if (!p) return;
  c.link.win_gen[w]++;
  if (c.link.p2p && c.link.ipc) c.link.retired.push_back(p);
  else cudaFree(p);
}

// the base address of window w in my own workspace
inline char* link_my_base(Ctx& c, int w) {
  switch (w) {
    case W_X: return (char*)c.x;
    case W_P: return (char*)c.p;
    case W_U: return (char*)c.u;
    case W_XF0: return (char*)c.xf[0];
    case W_XF1: return (char*)c.xf[1];
    default: return (char*)c.mg_pool;
  }
}

// map (or re-map) an allocation another process exported; mappings are cached by handle
inline int link_ipc_open(Link::IpcMap& m, const unsigned char* handle, unsigned long long gen, void** out) {
  if (m.ptr && m.gen == gen && memcmp(m.handle, handle, 64) == 0) {
    *out = m.ptr;
    return 0;
  }
  if (m.ptr) cudaIpcCloseMemHandle(m.ptr);
  m.ptr = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(KEFFB200_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(m.handle, handle, 64);
  m.gen = gen;
  m.ptr = p;
  *out = p;
  return 0;
}

// Publish my workspace windows and slab geometry, meet the group, map the neighbours' windows.  Collective.
// reset: also clear my arrival counters / all-reduce slots (start of a solve: nobody stores into them before the
// barrier inside this call, because every rank finishes its previous solve before it gets here).
inline int link_publish(Ctx& c, const Dom& D, bool reset) {
  Link& L = c.link;
  Mail* mine = L.mail[L.rank];
  if (reset) {
    KB_CUDA(cudaMemsetAsync(mine->ar, 0, sizeof mine->ar, c.st));
    KB_CUDA(cudaMemsetAsync(mine->halo, 0, sizeof mine->halo, c.st));
    KB_CUDA(cudaMemsetAsync(mine->gath, 0, sizeof mine->gath, c.st));
    L.expect_g = 0;
    KB_CUDA(cudaMemsetAsync(L.ar_seq, 0, sizeof(unsigned long long), c.st));
    memset(L.expect, 0, sizeof L.expect);
  }
  WinTable& t = L.h_win[L.rank];
  memset(&t, 0, sizeof t);
  t.pid = (long long)getpid();
  t.nz = D.nz;
  t.z0 = D.z0;
  for (int w = 0; w < W_COUNT; w++) {
    char* b = link_my_base(c, w);
    t.base[w] = (unsigned long long)(uintptr_t)b;
    t.gen[w] = L.win_gen[w];
    if (b && L.ipc) {
      cudaIpcMemHandle_t h;
      if (cudaIpcGetMemHandle(&h, b) == cudaSuccess) memcpy(t.handle[w], &h, sizeof h);
      else cudaGetLastError();
    }
  }
  KB_CUDA(cudaMemcpyAsync(&mine->win[L.rank], &t, sizeof t, cudaMemcpyHostToDevice, c.st));
  k_post_table<<<1, 128, 0, c.st>>>(link_table(L));
  KB_LAUNCHED();
  KB_TRY(link_barrier(c));
  KB_CUDA(cudaMemcpyAsync(L.h_win, mine->win, sizeof(WinTable) * L.world, cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaMemcpyAsync(c.hS, c.S, sizeof(Scal), cudaMemcpyDeviceToHost, c.st));
  KB_CUDA(cudaStreamSynchronize(c.st));
  if (c.hS->fault) return fail(KEFFB200_ENCCL, "slab group barrier timed out (a rank is missing)");
  const long long me = (long long)getpid();
  for (int side = 0; side < 2; side++) {
    const int r = side == 0 ? L.rank - 1 : L.rank + 1;
    for (int w = 0; w < W_COUNT; w++) L.nb[side][w] = nullptr;
    if (r < 0 || r >= L.world) continue;
    const WinTable& o = L.h_win[r];
    L.nb_nz[side] = o.nz;
    for (int w = 0; w < W_COUNT; w++) {
      if (!o.base[w]) continue;
      if (o.pid == me) L.nb[side][w] = (char*)(uintptr_t)o.base[w];
      else {
        void* p = nullptr;
        KB_TRY(link_ipc_open(L.ipcw[r][w], o.handle[w], o.gen[w], &p));
        L.nb[side][w] = (char*)p;
      }
    }
  }
  for (int r = 0; r < L.world; r++) {
    L.pool[r] = nullptr;
    if (r == L.rank) continue;
    const WinTable& o = L.h_win[r];
    if (!o.base[W_POOL]) continue;
    if (o.pid == me) L.pool[r] = (char*)(uintptr_t)o.base[W_POOL];
    else if (r == L.rank - 1) L.pool[r] = L.nb[0][W_POOL];
    else if (r == L.rank + 1) L.pool[r] = L.nb[1][W_POOL];
    else {
      void* p = nullptr;
      KB_TRY(link_ipc_open(L.ipcw[r][W_POOL], o.handle[W_POOL], o.gen[W_POOL], &p));
      L.pool[r] = (char*)p;
    }
  }
  return 0;
}

// Push the boundary planes of a plane-major array that lives in window `win` of my workspace into my
// neighbours' halo planes (plane0 = my plane 0; halo planes at -1 and nz), and wait for theirs.
// fine_grid: the array has as many planes as the slab (x, p: the lower neighbour's slab may be one plane longer or
// shorter); every other array belongs to the multigrid path, which requires equal slabs.
inline int link_exchange(Ctx& c, int field, int win, char* plane0, size_t plane_bytes, int nz, bool fine_grid = false) {
  Link& L = c.link;
  Mail* mine = L.mail[L.rank];
  const size_t off = (size_t)(plane0 - link_my_base(c, win));
  const bool lo = L.rank > 0, hi = L.rank < L.world - 1;
  char* lo_dst = lo ? L.nb[0][win] + off + (size_t)(fine_grid ? L.nb_nz[0] : nz) * plane_bytes : nullptr;
  char* hi_dst = hi ? L.nb[1][win] + off - plane_bytes : nullptr;
  if ((lo && !L.nb[0][win]) || (hi && !L.nb[1][win])) return fail(KEFFB200_ESTATE, "neighbour window %d is not mapped", win);
  const int g = push_ctas(plane_bytes);
  const int wide = ((plane_bytes | (size_t)(uintptr_t)plane0 | (size_t)(uintptr_t)lo_dst | (size_t)(uintptr_t)hi_dst) & 15) == 0;
  k_push_planes<<<g, PUSH_THREADS, 0, c.st>>>(plane0, lo_dst, plane0 + (size_t)(nz - 1) * plane_bytes, hi_dst, plane_bytes, wide,
                                              lo ? &L.mail[L.rank - 1]->halo[field][1] : nullptr,
                                              hi ? &L.mail[L.rank + 1]->halo[field][0] : nullptr);
  KB_LAUNCHED();
  if (lo) L.expect[field][0] += g;
  if (hi) L.expect[field][1] += g;
  k_wait_halo<<<1, 32, 0, c.st>>>(lo ? &mine->halo[field][0] : nullptr, L.expect[field][0], hi ? &mine->halo[field][1] : nullptr,
                                  L.expect[field][1], c.S);
  KB_LAUNCHED();
  return 0;
}

// all-gather of a replicated array of the level pool: my slice (src, inside my copy) -> every rank's copy
inline int link_gather(Ctx& c, const char* src, size_t bytes) {
  Link& L = c.link;
  Mail* mine = L.mail[L.rank];
  GatherDst G;
  memset(&G, 0, sizeof G);
  G.world = L.world;
  const size_t off = (size_t)(src - (const char*)c.mg_pool);
  for (int r = 0; r < L.world; r++) {
    if (r == L.rank) continue;
    if (!L.pool[r]) return fail(KEFFB200_ESTATE, "level pool of rank %d is not mapped", r);
    G.dst[r] = L.pool[r] + off;
    G.flag[r] = &L.mail[r]->gath[L.rank];
  }
  const int g = push_ctas(bytes);
  k_push_gather<<<g, PUSH_THREADS, 0, c.st>>>(src, G, bytes);
  KB_LAUNCHED();
  L.expect_g += (unsigned)g;
  k_wait_gather<<<1, 32, 0, c.st>>>(mine->gath, L.world, L.rank, L.expect_g, c.S);
  KB_LAUNCHED();
  return 0;
}

}  // namespace kb
