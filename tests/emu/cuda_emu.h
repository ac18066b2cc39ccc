// TEST INFRASTRUCTURE ONLY.
//
// A minimal "CUDA on pthreads" shim: when quaff_b200/csrc/*.cu is compiled by g++ with -DQG_EMU, this
// header supplies just enough of the CUDA language and runtime for the kernels to execute on the CPU,
// one OS thread per CUDA thread, blocks run one after another.  It exists so that the kernels' index
// arithmetic, warp exchanges, shared-memory protocols and the host orchestration can be checked
// against the oracle in this GPU-less container (pytest -m "not gpu") BEFORE spending GPU minutes.
// It is NOT a fallback: quaff_b200 never loads the emulated library, bench.py never times it, and it is
// orders of magnitude slower than anything useful.  Warp collectives must be called by all 32 lanes
// (as the kernels do on the device, full mask), otherwise the emulation deadlocks.
#ifndef QG_CUDA_EMU_H
#define QG_CUDA_EMU_H
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <chrono>
#include <functional>
#include <thread>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __constant__ static

struct uint3_e { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3 (unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
struct int2 { int x, y; };
static inline uint2 make_uint2 (unsigned a, unsigned b) { uint2 r = {a, b}; return r; }
static inline uint4 make_uint4 (unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline double2 make_double2 (double a, double b) { double2 r = {a, b}; return r; }
static inline int2 make_int2 (int a, int b) { int2 r = {a, b}; return r; }

namespace qgemu {
struct BlockState {
  pthread_barrier_t block_bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<uint64_t> xchg;        // one 64-bit slot per thread
  unsigned char* dyn_smem;
  int or_flag;
};
extern thread_local uint3_e t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local BlockState* t_block;

inline unsigned lane_id () { return t_threadIdx.x & 31; }
inline unsigned warp_id () { return t_threadIdx.x >> 5; }
inline void warp_barrier () { pthread_barrier_wait (&t_block->warp_bar[warp_id ()]); }

template<class T> inline T shfl_idx (T v, int src) {
  static_assert (sizeof (T) <= 8, "shuffle of > 64 bit");
  uint64_t raw = 0; memcpy (&raw, &v, sizeof (T));
  const unsigned base = warp_id () * 32;
  t_block->xchg[base + lane_id ()] = raw;
  warp_barrier ();
  const unsigned nl = std::min (32u, t_blockDim.x - base);
  const unsigned s = ((unsigned) src) & 31;
  uint64_t got = (s < nl) ? t_block->xchg[base + s] : raw;
  warp_barrier ();
  T r; memcpy (&r, &got, sizeof (T));
  return r;
}

void launch (dim3 grid, dim3 block, size_t smem, const std::function<void ()>& body);
}  // namespace qgemu

#define threadIdx (qgemu::t_threadIdx)
#define blockIdx (qgemu::t_blockIdx)
#define blockDim (qgemu::t_blockDim)
#define gridDim (qgemu::t_gridDim)
#define warpSize 32

static inline void __syncthreads () { pthread_barrier_wait (&qgemu::t_block->block_bar); }
static inline void __syncwarp (unsigned = 0xffffffffu) { qgemu::warp_barrier (); }
static inline int __syncthreads_or (int pred) {
  if (pred) __atomic_store_n (&qgemu::t_block->or_flag, 1, __ATOMIC_RELAXED);
  __syncthreads ();
  const int r = __atomic_load_n (&qgemu::t_block->or_flag, __ATOMIC_RELAXED);
  __syncthreads ();
  if (threadIdx.x == 0) __atomic_store_n (&qgemu::t_block->or_flag, 0, __ATOMIC_RELAXED);
  __syncthreads ();
  return r;
}
template<class T> static inline T __shfl_sync (unsigned, T v, int src) { return qgemu::shfl_idx (v, src); }
template<class T> static inline T __shfl_up_sync (unsigned, T v, unsigned d) {
  const int l = (int) qgemu::lane_id (); return qgemu::shfl_idx (v, l - (int) d >= 0 ? l - (int) d : l); }
template<class T> static inline T __shfl_down_sync (unsigned, T v, unsigned d) {
  const int l = (int) qgemu::lane_id (); return qgemu::shfl_idx (v, l + (int) d < 32 ? l + (int) d : l); }
template<class T> static inline T __shfl_xor_sync (unsigned, T v, int m) { return qgemu::shfl_idx (v, (int) qgemu::lane_id () ^ m); }
static inline unsigned __ballot_sync (unsigned, int pred) {
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l) bits |= (qgemu::shfl_idx (pred ? 1 : 0, l) ? 1u : 0u) << l;
  return bits;
}
static inline int __any_sync (unsigned m, int pred) { return __ballot_sync (m, pred) != 0; }
static inline int __all_sync (unsigned m, int pred) { return __ballot_sync (m, pred) == 0xffffffffu; }

template<class T> static inline T __ldg (const T* p) { return *p; }
static inline int __popc (unsigned v) { return __builtin_popcount (v); }
static inline int __popcll (unsigned long long v) { return __builtin_popcountll (v); }
static inline int __ffs (int v) { return __builtin_ffs (v); }
static inline int __clz (int v) { return v ? __builtin_clz ((unsigned) v) : 32; }
static inline double __longlong_as_double (long long v) { double d; memcpy (&d, &v, 8); return d; }
static inline long long __double_as_longlong (double d) { long long v; memcpy (&v, &d, 8); return v; }
static inline double __dadd_rn (double a, double b) { return a + b; }
static inline double __dmul_rn (double a, double b) { return a * b; }
static inline double __ddiv_rn (double a, double b) { return a / b; }

static inline int atomicAdd (int* p, int v) { return __atomic_fetch_add (p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd (unsigned* p, unsigned v) { return __atomic_fetch_add (p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd (unsigned long long* p, unsigned long long v) { return __atomic_fetch_add (p, v, __ATOMIC_RELAXED); }
static inline double atomicAdd (double* p, double v) {
  uint64_t* ip = (uint64_t*) p; uint64_t old = __atomic_load_n (ip, __ATOMIC_RELAXED), nw;
  double od;
  do { memcpy (&od, &old, 8); const double nd = od + v; memcpy (&nw, &nd, 8); }
  while (!__atomic_compare_exchange_n (ip, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return od;
}
static inline unsigned atomicOr (unsigned* p, unsigned v) { return __atomic_fetch_or (p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMax (unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n (p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n (p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
  return old;
}
static inline int atomicMax (int* p, int v) {
  int old = __atomic_load_n (p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n (p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
  return old;
}
static inline void __threadfence () { __atomic_thread_fence (__ATOMIC_SEQ_CST); }

// ---- runtime -------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
struct qgemu_event { std::chrono::steady_clock::time_point t; };
typedef qgemu_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; int major, minor; size_t sharedMemPerBlockOptin; size_t totalGlobalMem; };
static inline const char* cudaGetErrorString (cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError () { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError () { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount (int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice (int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties (cudaDeviceProp* p, int) {
  memset (p, 0, sizeof (*p)); strcpy (p->name, "qgemu (CPU threads, test only)");
  p->multiProcessorCount = 4; p->major = 10; p->minor = 0; p->sharedMemPerBlockOptin = 227 * 1024; p->totalGlobalMem = (size_t) 8 << 30;
  return cudaSuccess;
}
static inline cudaError_t cudaMalloc (void** p, size_t n) { *p = malloc (n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template<class T> static inline cudaError_t cudaMalloc (T** p, size_t n) { return cudaMalloc ((void**) p, n); }
static inline cudaError_t cudaFree (void* p) { free (p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost (void** p, size_t n) { *p = malloc (n ? n : 1); return cudaSuccess; }
template<class T> static inline cudaError_t cudaMallocHost (T** p, size_t n) { return cudaMallocHost ((void**) p, n); }
static inline cudaError_t cudaFreeHost (void* p) { free (p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy (void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync (void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset (void* d, int v, size_t n) { memset (d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync (void* d, int v, size_t n, cudaStream_t = 0) { memset (d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags (cudaStream_t* s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate (cudaStream_t* s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy (cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2, cudaEventBlockingSync = 1 };
static inline cudaError_t cudaEventCreateWithFlags (cudaEvent_t* e, unsigned) { *e = new qgemu_event (); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize () { return cudaSuccess; }
static inline cudaError_t cudaEventCreate (cudaEvent_t* e) { *e = new qgemu_event (); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy (cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord (cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now (); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime (float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli> (b->t - a->t).count (); return cudaSuccess; }
template<class F> static inline cudaError_t cudaFuncSetAttribute (F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo (size_t* fr, size_t* tot) { *fr = (size_t) 6 << 30; *tot = (size_t) 8 << 30; return cudaSuccess; }

#endif
