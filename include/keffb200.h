/* keffb200.h -- C ABI of the B200-native Keff-CFD hot path.
 *
 * The reference (adama-wzr/Keff-CFD) has no plugin/FFI layer: the seam is the solver call inside
 * each program's main().  Every entry point below names the reference call site it replaces.
 * Conventions: plain pointers and sizes only; the caller owns every host buffer; the library owns
 * all device memory, streams and communicator state; no exit(), no stdin, no stdout; every function
 * returns 0 on success or a negative KEFFB200_E* code, with text from keffb200_last_error().
 * One context per GPU: a process binds one GPU (keffb200_init -- the one-process-per-GPU layout) or
 * several (keffb200_init_multi); calls are made from one host thread and are not re-entrant.
 *
 * Streams: every kernel runs on the library's own non-blocking stream.  Host-buffer entry points are
 * synchronous.  Device-buffer entry points (*_dev) read their inputs on that stream: if another stream is
 * still writing them, call keffb200_wait_stream(that stream) first; they return after their outputs are
 * complete (the library stream is synchronised before the call returns).
 *
 * Structure semantics are fixed at this boundary: `solid` is one byte per cell, 1 = solid (ks),
 * anything else = fluid (kf); layout [z][y][x], x fastest -- the reference's
 * index = z*ny*nx + y*nx + x (Keff3D/keff3D.h:691).  The host maps its inputs onto it:
 * 2D JPG pixel >= 150 -> solid (Keff2D/keff2D.cpp:113-119), 3D voxel listed in the CSV -> solid
 * (Keff3D/keff3D.h:269-277, keff3D.cpp:113-117); MeshAmp replication is done by the host.
 * The physical domain is the unit square/cube, T = tl on the x = 0 face and tr on the x = 1 face,
 * all other faces adiabatic (Keff2D/keff2D.h:482-582, Keff3D/keff3D.h:666-763).
 */
#ifndef KEFFB200_H
#define KEFFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KEFFB200_OK 0
#define KEFFB200_ENOGPU (-1)   /* no CUDA device / not sm_100 */
#define KEFFB200_EINVAL (-2)   /* bad argument */
#define KEFFB200_ECUDA (-3)    /* CUDA runtime error (text in last_error) */
#define KEFFB200_ENOMEM (-4)
#define KEFFB200_ENCCL (-5)    /* NCCL missing or failed */
#define KEFFB200_ESTATE (-6)   /* init not called, or slab calls without dist_init */

/* Solver knobs.  The first six are the reference's input.txt keys that reach the path
 * (ks kf TL TR Convergence MaxIter: Keff2D/keff2D.h:142-205, Keff3D/keff3D.h:145-220).  The rest
 * have no reference counterpart; 0 selects the default. */
typedef struct keffb200_params {
  double ks, kf;          /* TCsolid, TCfluid */
  double tl, tr;          /* TempLeft, TempRight */
  double convergence;     /* ConvergeCriteria.  The reference stops when Keff changes by less than this between two
                             checks 100 sweeps apart (Keff2D/keff2D.h:825-847), 5-7e-4 away from the discrete solution
                             at the shipped 1e-5.  Here the key sets the two targets below when they are left at 0:
                             the shipped value gives the defaults, a tighter value tightens both, a looser one never
                             loosens them (the field is frozen once the targets are met, so Keff no longer changes). */
  long max_iter;          /* MAX_ITER: cap on solver iterations */
  double rel_residual;    /* stop when ||b - A T||2 / ||b||2 <= this; 0 (default) = min(1e-8, 1e-3 * convergence) */
  double flux_balance;    /* ... and |QL - QR| / |Keff| <= this; 0 (default) = min(1e-6, 0.1 * convergence) */
  int check_every;        /* iterations between host-visible convergence checks (default 32) */
  int flags;              /* KEFFB200_F_* */
} keffb200_params;

#define KEFFB200_F_NO_TMA 1      /* use the plain global-load stencil kernel (debug / odd nx) */
#define KEFFB200_F_BATCH_GLOBAL 4 /* 2D batches: force the global-memory FP64 kernel (the on-chip
                                    mixed-precision kernel is used when nx<=128, nx%4==0, ny<=128, ny%8==0) */
#define KEFFB200_F_BATCH_ADDITIVE 16 /* 2D batches: the on-chip kernel with the two-level additive preconditioner
                                    instead of the on-chip multigrid one (used anyway unless nx, ny are multiples of 16) */
#define KEFFB200_F_NO_MG 8       /* grid solver: plain diagonally preconditioned CG instead of the
                                    multigrid-preconditioned one (used anyway when nx is odd, the
                                    grid has fewer than 65 cells, ny < 2, or z-slabs are unequal) */
#define KEFFB200_F_ZERO_INIT 2   /* start from T = tl everywhere (the reference's actual T0,
                                    Keff2D/keff2D.h:584-611) instead of the linear ramp */
#define KEFFB200_F_APPLY_F32 32   /* keffb200_bench_apply_dev only: time the stencil kernel in the form the multigrid
                                    iteration launches it (input field p stored in FP32: 12 B per cell-update instead
                                    of 16); needs nx % 4 == 0 */

/* Per-solve record: the numbers the reference puts in its CSV row (keff, QL, QR, Iter:
 * Keff2D/keff2D.h:245-246) plus what its GPU variant keeps in simulationInfo
 * (Keff2DGPU/keff2D.cuh:41-55: keff, conv, gpuTime). */
typedef struct keffb200_result {
  double keff;            /* (QL + QR) / 2 / (tr - tl) */
  double q_left, q_right; /* sum of face fluxes through x = 0 and x = 1 (Q1, Q2) */
  double rel_residual;    /* TRUE ||b - A T||2 / ||b||2 of the returned field */
  double energy_residual; /* sum_cells |net flux imbalance| = reference Residual() (keff2D.h:384) */
  long iters;             /* solver iterations taken */
  int status;             /* 0 converged, 1 stopped at max_iter */
  float gpu_ms;           /* device time of the solve, CUDA events */
} keffb200_result;

/* ---- life cycle ------------------------------------------------------------------------------ */
/* Bind the process to CUDA device `device` (replaces initializeGPU's cudaSetDevice(0),
 * Keff2DGPU/keff2D.cuh:606-683; memory is allocated lazily by the solves and cached). */
int keffb200_init(int device);
/* Bind the process to GPUs 0..n_gpus-1 (n_gpus <= 0: every visible GPU), one context and one host thread per GPU
 * inside the library.  The host-buffer calls then fan out by themselves: keffb200_solve2d_batch gives every GPU a
 * contiguous share of the images (the reference's OpenMP loop over images, Keff2D/keff2D.cpp:341, with GPUs in the
 * place of threads; no communication), keffb200_solve3d cuts a large volume into z-slabs whose halo planes and
 * reductions travel as NVLink stores between the GPUs' memories (volumes below keffb200_set_slab_min_cells stay on
 * GPU 0).  The *_dev entry points and keffb200_solve2d use GPU 0. */
int keffb200_init_multi(int n_gpus);
int keffb200_device_count(void);               /* GPUs bound by init / init_multi */
void keffb200_set_slab_min_cells(long cells);  /* default 2^26; < 0 restores it */
void keffb200_shutdown(void); /* replaces unInitializeGPU, Keff2DGPU/keff2D.cuh:685-723 */
const char* keffb200_last_error(void);
const char* keffb200_version(void);
void keffb200_default_params(keffb200_params* p); /* shipped input.txt values: ks 10 kf 1 TL 0 TR 1
                                                     Convergence 1e-5 MaxIter 500000 */
/* Count of this library's kernel launches since init (bench.py's gpu_launches). */
long keffb200_launch_count(void);

/* ---- whole-path solves, HOST buffers (the reference-facing calls) ----------------------------- */
/* One 2D image.  Replaces DiscretizeMatrixCD2D + SolInitLinear + ParallelGS + the flux loop at
 * Keff2D/keff2D.cpp:131-170 (and DiscretizeMatrix2D + JacobiGPU, Keff2DGPU/keff2D.cuh:969-974).
 * T_init (nullable, ny*nx) warm start; T_out (nullable, ny*nx); qL,qR (nullable, ny) per-row fluxes. */
int keffb200_solve2d(const uint8_t* solid, int nx, int ny, const keffb200_params* p, const double* T_init,
                     double* T_out, double* qL, double* qR, keffb200_result* out);

/* A batch of independent equally-sized 2D images (n_img*ny*nx bytes).  Replaces the batch loops
 * Keff2D/keff2D.cpp:341-471 (OpenMP over images, SerialGS each) and BatchSim
 * Keff2DGPU/keff2D.cuh:1007-1168.  out[n_img].  T_out nullable (n_img*ny*nx). */
int keffb200_solve2d_batch(const uint8_t* solid, int n_img, int nx, int ny, const keffb200_params* p,
                           double* T_out, keffb200_result* out);

/* One 3D volume on this process's GPU.  Replaces DiscretizeMatrixCD3D + SolInitLinear +
 * ParallelGS3D + the flux loop at Keff3D/keff3D.cpp:134-163.  qL,qR nullable, nz*ny each. */
int keffb200_solve3d(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p,
                     const double* T_init, double* T_out, double* qL, double* qR, keffb200_result* out);

/* ---- the same solves on DEVICE buffers (inputs already resident in HBM) ----------------------- */
int keffb200_solve3d_dev(const uint8_t* d_solid, int nx, int ny, int nz, const keffb200_params* p,
                         const double* d_T_init, double* d_T_out, keffb200_result* out);
int keffb200_solve2d_batch_dev(const uint8_t* d_solid, int n_img, int nx, int ny, const keffb200_params* p,
                               double* d_T_out, keffb200_result* out /* host */);

/* ---- z-slab decomposition across processes (one process per GPU) ------------------------------ */
/* Size of the opaque communicator id that rank 0 creates and the host broadcasts to all ranks. */
#define KEFFB200_COMM_ID_BYTES 128
int keffb200_comm_id(void* id128);                       /* rank 0: ncclGetUniqueId */
/* All ranks: ncclCommInitRank, then every rank maps every other rank's mailbox through CUDA IPC.  When that
 * succeeds on all ranks the slab solve's halo planes and reductions travel as NVLink stores issued by the
 * solver's own kernels and NCCL is off the data path (keffb200_dist_peer_memory() == 1); otherwise, or with
 * KEFFB200_SLAB_NCCL set, they go through ncclSend/ncclRecv/ncclAllReduce. */
int keffb200_dist_init(int rank, int world, const void* id128);
int keffb200_dist_peer_memory(void);
void keffb200_dist_shutdown(void);
/* Solve a global nx*ny*nz_global volume whose planes [z0, z0+nz_local) live on this rank.
 * d_solid_halo: (nz_local+2) planes, plane 0 / plane nz_local+1 = the neighbouring ranks' boundary
 * planes (ignored at the global ends).  d_T_out: nz_local planes (nullable).  Collective: every rank
 * of the communicator must call it; every rank receives the same result. */
int keffb200_solve3d_slab_dev(const uint8_t* d_solid_halo, int nx, int ny, int nz_local, int z0,
                              int nz_global, const keffb200_params* p, double* d_T_out,
                              keffb200_result* out);

/* ---- pieces of the path, exposed for parity tests and roofline measurement -------------------- */
/* y = A x, the matrix-free form of the system DiscretizeMatrixCD2D/3D assembles
 * (Keff2D/keff2D.h:482-582, Keff3D/keff3D.h:666-763); b (nullable) receives the right-hand side.
 * Host buffers, nz = 1 for 2D. */
int keffb200_apply(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* x,
                   double* y, double* b);
/* Boundary-flux integration alone (Keff2D/keff2D.cpp:151-170, Keff3D/keff3D.cpp:145-163) and the
 * energy-imbalance Residual() (Keff2D/keff2D.h:384-424) of a given host field T. */
int keffb200_flux(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* T,
                  double* qL, double* qR, keffb200_result* out);
/* z = M r: one application of the multigrid V-cycle that preconditions the grid solver's CG (no
 * reference counterpart; exposed so tests can check that M is symmetric positive definite and what
 * it does to the spectrum).  Host buffers of nx*ny*nz doubles; fails with KEFFB200_EINVAL for grids
 * the multigrid path does not take (see KEFFB200_F_NO_MG). */
int keffb200_precond(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* r,
                     double* z);
/* x-direction heat-flux field of a given host field T: Qx[z][y][j], j = 0..nx, the flux through the face between
 * cells j-1 and j (j = 0 and j = nx: the fixed-temperature walls), the quantity behind printQMAP
 * (Keff2D/keff2D.h:427-479: harmonic-mean face fluxes; the reference prints "not currently available" in 3D,
 * Keff3D/keff3D.cpp:179-181).  Qx holds nz*ny*(nx+1) doubles. */
int keffb200_flux_field(const uint8_t* solid, int nx, int ny, int nz, const keffb200_params* p, const double* T,
                        double* Qx);
/* Warm start for the mesh-convergence mode: (tri)linear interpolation of a field from an sx*sy*sz mesh of the unit
 * cube onto an nx*ny*nz mesh, with the fixed-temperature walls as interpolation nodes along x.  Replaces
 * SolExpandSimple (Keff2D/keff2D.h:614-642), which replicates values without interpolating.  Host buffers. */
int keffb200_interpolate(const double* T_src, int sx, int sy, int sz, double* T_dst, int nx, int ny, int nz,
                         double tl, double tr);
/* Page-locked host memory for the buffers handed to the host-buffer entry points (optional: pageable memory works;
 * keffb200_solve2d_batch streams a batch to the GPU in chunks underneath its solve from either kind, a volume's
 * single copy is faster from page-locked memory). */
void* keffb200_host_alloc(size_t bytes);
void keffb200_host_free(void* p);

/* Timing hook: `reps` applications of the stencil kernel on device-resident data
 * (one cell-update per cell per application, 16 B/cell of compulsory traffic).  Returns the average
 * milliseconds per application, measured with CUDA events on the launching stream. */
int keffb200_bench_apply_dev(const uint8_t* d_solid, int nx, int ny, int nz, const keffb200_params* p,
                             int reps, float* ms_per_apply);

/* Region timer for benchmarks: CUDA events recorded on the library's compute stream (the stream every
 * kernel of this library is launched on), so a start/stop pair brackets exactly the device work of the
 * calls made in between, host-buffer copies included.  stop() waits for its event. */
int keffb200_timer_start(void);
int keffb200_timer_stop(float* ms);
/* Make the library's stream wait for the work queued so far on `stream` (a cudaStream_t; NULL = the legacy
 * default stream).  See "Streams" at the top. */
int keffb200_wait_stream(void* stream);

/* Synthetic input generator for benchmarks and tests (no reference counterpart): voxel i of the
 * range [first_index, first_index + n) is solid iff splitmix64(seed ^ i) < p * 2^64.  The same rule
 * is implemented in Python (keff_cfd_b200/structures.py) so host and device agree bit for bit. */
int keffb200_gen_bernoulli_dev(uint8_t* d_out, size_t n, unsigned long long first_index,
                               unsigned long long seed, double p);

#ifdef __cplusplus
}
#endif
#endif /* KEFFB200_H */
