// Micro-benchmark of candidate inner loops for the seeding histogram (hits / cycle / SM), synthetic read + reference.
//   A : positions in reference order, 128-bit bucket header (8 inline entries), ladder to the warp's longest bucket,
//       inactive lanes add into a per-lane dummy word (no divergence)
//   A2: as A with a predicated red.shared instead of the dummy address
//   A3: as A with `if (t < len) atomicAdd` (compiler's divergence handling)
//   B : reference tile pre-sorted by k-mer code (lanes of a warp share codes -> balanced ladder), 32-bit ring
//   B16: as B with 16-bit packed counters (larger tile)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define NK 4096
struct Args { const uint16_t* xcodes; const uint32_t* entries; int n; const uint4* hdr; unsigned long long* hits; long long* cyc; uint32_t ringmask; int tile; };

__device__ __forceinline__ uint32_t ext (const uint4 h, int t) {
  switch (t) {
    case 0: return h.x & 0x3FFFu;
    case 1: return (h.x >> 14) & 0x3FFFu;
    case 2: return __funnelshift_r (h.x, h.y, 28) & 0x3FFFu;
    case 3: return (h.y >> 10) & 0x3FFFu;
    case 4: return __funnelshift_r (h.y, h.z, 24) & 0x3FFFu;
    case 5: return (h.z >> 6) & 0x3FFFu;
    case 6: return __funnelshift_r (h.z, h.w, 20) & 0x3FFFu;
    default: return (h.w >> 2) & 0x3FFFu;
  }
}

template<int MODE>
__global__ void __launch_bounds__ (1024, 1) kA (Args a) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t* cnt = (uint32_t*) smem;
  uint4* hdr = (uint4*) (cnt + a.ringmask + 1);
  uint32_t* dummy = (uint32_t*) (hdr + NK + 1);
  for (uint32_t c = threadIdx.x; c <= a.ringmask; c += blockDim.x) cnt[c] = 0;
  for (uint32_t c = threadIdx.x; c < NK; c += blockDim.x) hdr[c] = a.hdr[c];
  if (threadIdx.x == 0) hdr[NK] = make_uint4 (0, 0, 0, 0);
  if (threadIdx.x < 32) dummy[threadIdx.x] = 0;
  __syncthreads ();
  const int lane = threadIdx.x & 31;
  const uint32_t mask = a.ringmask;
  uint32_t* mydummy = dummy + lane;
  unsigned long long hits = 0;
  const long long t0 = clock64 ();
  for (int i0 = 0; i0 < a.n; i0 += 8192) {
    uint32_t code[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { const int i = i0 + r * 1024 + threadIdx.x; code[r] = i < a.n ? a.xcodes[i] : NK; }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t i = i0 + r * 1024 + threadIdx.x;
      const uint4 h = hdr[code[r] < NK ? code[r] : NK];
      const uint32_t len = h.w >> 16;
      hits += len;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (t > 0 && !__any_sync (0xffffffffu, len > t)) break;
        const uint32_t slot = (i + ext (h, t)) & mask;
        if (MODE == 0) { uint32_t* p = (t < len) ? cnt + slot : mydummy; atomicAdd (p, 1u); }
        if (MODE == 1) { asm volatile ("{ .reg .pred p; setp.lt.u32 p, %2, %3; @p red.shared.add.u32 [%0], %1; }" :: "r" ((uint32_t) __cvta_generic_to_shared (cnt + slot)), "r" (1u), "r" ((uint32_t) t), "r" (len) : "memory"); }
        if (MODE == 2) { if (t < len) atomicAdd (cnt + slot, 1u); }
      }
    }
    __syncthreads ();
  }
  const long long t1 = clock64 ();
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync (0xffffffffu, hits, o);
  if (lane == 0) atomicAdd (a.hits, hits);
  if (threadIdx.x == 0) a.cyc[blockIdx.x] = t1 - t0;
}

// tile-sorted entries: entry = code << 16 | local position; tiles of a.tile positions, ring slides by tile
template<int BITS>
__global__ void __launch_bounds__ (1024, 1) kB (Args a) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t* cnt = (uint32_t*) smem;
  const uint32_t ringwords = BITS == 32 ? a.ringmask + 1 : (a.ringmask + 1) / 2;
  uint4* hdr = (uint4*) (cnt + ringwords);
  uint32_t* dummy = (uint32_t*) (hdr + NK + 1);
  for (uint32_t c = threadIdx.x; c < ringwords; c += blockDim.x) cnt[c] = 0;
  for (uint32_t c = threadIdx.x; c < NK; c += blockDim.x) hdr[c] = a.hdr[c];
  if (threadIdx.x < 32) dummy[threadIdx.x] = 0;
  __syncthreads ();
  const int lane = threadIdx.x & 31;
  const uint32_t mask = a.ringmask;
  uint32_t* mydummy = dummy + lane;
  unsigned long long hits = 0;
  const long long t0 = clock64 ();
  for (int tb = 0; tb < a.n; tb += a.tile) {
    const int tn = min (a.tile, a.n - tb);
    for (int e0 = 0; e0 < tn; e0 += 4096) {
      uint32_t ent[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { const int e = e0 + r * 1024 + threadIdx.x; ent[r] = e < tn ? a.entries[tb + e] : 0xFFFFFFFFu; }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t code = ent[r] >> 16;
        const uint32_t i = (uint32_t) tb + (ent[r] & 0xFFFFu);
        const uint4 h = code < NK ? hdr[code] : make_uint4 (0, 0, 0, 0);
        const uint32_t len = h.w >> 16;
        hits += len;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t > 0 && !__any_sync (0xffffffffu, len > t)) break;
          const uint32_t slot = (i + ext (h, t)) & mask;
          if (BITS == 32) { uint32_t* p = (t < len) ? cnt + slot : mydummy; atomicAdd (p, 1u); }
          else { uint32_t* p = (t < len) ? cnt + (slot >> 1) : mydummy; atomicAdd (p, (slot & 1u) ? 0x10000u : 1u); }
        }
      }
    }
    __syncthreads ();
  }
  const long long t1 = clock64 ();
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync (0xffffffffu, hits, o);
  if (lane == 0) atomicAdd (a.hits, hits);
  if (threadIdx.x == 0) a.cyc[blockIdx.x] = t1 - t0;
}

int main () {
  const int YLEN = 8000, K = 6, N = 1 << 20;
  srand (7);
  std::vector<uint8_t> y (YLEN), x (N + K);
  for (auto& v : y) v = rand () & 3;
  for (auto& v : x) v = rand () & 3;
  auto code_at = [&] (const std::vector<uint8_t>& s, int p) { uint32_t c = 0; for (int t = 0; t < K; ++t) c = c * 4 + s[p + t]; return c; };
  const int span = YLEN - K;
  std::vector<std::vector<uint32_t>> bucket (NK);
  for (int j = 0; j <= span; ++j) bucket[code_at (y, j)].push_back (span - j);
  std::vector<uint4> hdr (NK);
  double dropped = 0;
  for (int c = 0; c < NK; ++c) {
    unsigned __int128 v = 0; const size_t len = std::min<size_t> (bucket[c].size (), 8);
    dropped += bucket[c].size () - len;
    for (size_t t = 0; t < len; ++t) v |= (unsigned __int128) bucket[c][t] << (14 * t);
    v |= (unsigned __int128) len << 112;
    hdr[c] = make_uint4 ((uint32_t) v, (uint32_t) (v >> 32), (uint32_t) (v >> 64), (uint32_t) (v >> 96));
  }
  std::vector<uint16_t> xc (N);
  for (int i = 0; i < N; ++i) xc[i] = (uint16_t) code_at (x, i);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties (&p, dev);
  const int nsm = p.multiProcessorCount;
  uint16_t* dxc; uint32_t* dent; uint4* dh; unsigned long long* dhits; long long* dcyc;
  cudaMalloc (&dxc, N * 2); cudaMalloc (&dent, N * 4); cudaMalloc (&dh, NK * 16); cudaMalloc (&dhits, 8); cudaMalloc (&dcyc, 8 * nsm);
  cudaMemcpy (dxc, xc.data (), N * 2, cudaMemcpyHostToDevice); cudaMemcpy (dh, hdr.data (), NK * 16, cudaMemcpyHostToDevice);
  auto sort_tiles = [&] (int tile) {
    std::vector<uint32_t> ent (N);
    for (int tb = 0; tb < N; tb += tile) {
      const int tn = std::min (tile, N - tb);
      for (int e = 0; e < tn; ++e) ent[tb + e] = ((uint32_t) xc[tb + e] << 16) | (uint32_t) e;
      std::sort (ent.begin () + tb, ent.begin () + tb + tn);
    }
    cudaMemcpy (dent, ent.data (), N * 4, cudaMemcpyHostToDevice);
  };
  auto report = [&] (const char* name, float ms) {
    unsigned long long h = 0; cudaMemcpy (&h, dhits, 8, cudaMemcpyDeviceToHost);
    std::vector<long long> cyc (nsm); cudaMemcpy (cyc.data (), dcyc, 8 * nsm, cudaMemcpyDeviceToHost);
    double mc = 0; for (auto v : cyc) mc += v; mc /= nsm;
    cudaError_t e = cudaGetLastError ();
    printf ("%-34s %.3f ms  hits/CTA=%.0f  cycles=%.0f  hits/cycle/SM=%.2f  (event, 1.965 GHz: %.2f) %s\n", name, ms, (double) h / nsm, mc, (double) h / nsm / mc,
            (double) h / nsm / (ms * 1e-3 * 1.965e9), e == cudaSuccess ? "" : cudaGetErrorString (e));
  };
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  Args a; a.xcodes = dxc; a.entries = dent; a.n = N; a.hdr = dh; a.hits = dhits; a.cyc = dcyc; a.ringmask = 32767; a.tile = 0;
#define RUNA(M, NAME) { const size_t sm = 32768 * 4 + (NK + 1) * 16 + 128; cudaFuncSetAttribute (kA<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
    kA<M><<<nsm, 1024, sm>>> (a); cudaMemset (dhits, 0, 8); cudaEventRecord (e0); kA<M><<<nsm, 1024, sm>>> (a); cudaEventRecord (e1); cudaEventSynchronize (e1); \
    float ms; cudaEventElapsedTime (&ms, e0, e1); report (NAME, ms); }
  printf ("dropped entries (len > 8): %.0f\n", dropped);
  RUNA (0, "A  ref order, dummy address");
  RUNA (1, "A2 ref order, predicated red");
  RUNA (2, "A3 ref order, if(t<len) atomicAdd");
#define RUNB(BITS, TILE, RINGC, NAME) { sort_tiles (TILE); a.tile = TILE; a.ringmask = RINGC - 1; \
    const size_t sm = (size_t) RINGC * (BITS / 8) + (NK + 1) * 16 + 128; cudaFuncSetAttribute (kB<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
    kB<BITS><<<nsm, 1024, sm>>> (a); cudaMemset (dhits, 0, 8); cudaEventRecord (e0); kB<BITS><<<nsm, 1024, sm>>> (a); cudaEventRecord (e1); cudaEventSynchronize (e1); \
    float ms; cudaEventElapsedTime (&ms, e0, e1); report (NAME, ms); }
  RUNB (32, 8192, 32768, "B32 tile 8K sorted");
  RUNB (32, 16384, 32768, "B32 tile 16K sorted");
  RUNB (32, 24576, 32768, "B32 tile 24K sorted");
  RUNB (16, 24576, 65536, "B16 tile 24K sorted");
  RUNB (16, 49152, 65536, "B16 tile 48K sorted");
  RUNB (16, 65536 - 8192, 65536, "B16 tile 56K sorted");
  return 0;
}
