// Micro-benchmark, second series: fixed-width bucket headers for the seeding histogram (hits / cycle / SM).
//   F : each reference position does exactly W = 4 predicated shared-memory increments from a 64-bit header
//       (4 x 14-bit entries + 8-bit length), codes come as one 128-bit load per 8 positions, the ring is scanned and
//       cleared after every step like the product kernel does.  TAILS: positions whose bucket is longer than 4 are
//       finished after the 8-position group from the full bucket array.
//   L : the round-1 product loop (64-bit header with 2 entries, ladder over bpos) under the same harness, for reference.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define NK 4096
struct Args { const uint16_t* xcodes; int n; const uint2* hdr; const uint2* hdr_old; const uint16_t* bpos; const uint16_t* bstart; int ylen;
              unsigned long long* hits; long long* cyc; uint32_t ring; int threshold; uint32_t* found; };

__device__ __forceinline__ void red_if (uint32_t* cnt, uint32_t slot, bool p) {
  asm volatile ("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.shared.add.u32 [%0], %1; }" :: "r" ((uint32_t) __cvta_generic_to_shared (cnt + slot)), "r" (1u), "r" ((uint32_t) p) : "memory");
}

template<int THREADS, int MINB, int MODE>     // MODE 0: F without tails, 1: F with tails, 2: round-1 loop
__global__ void __launch_bounds__ (THREADS, MINB) kF (Args a) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t* cnt = (uint32_t*) smem;
  uint2* hdr = (uint2*) (cnt + a.ring);
  uint16_t* bpos = (uint16_t*) (hdr + NK + 1);
  uint16_t* bstart = bpos + ((a.ylen + 8) & ~7);
  const int tid = threadIdx.x, lane = tid & 31;
  for (uint32_t c = tid; c < a.ring; c += THREADS) cnt[c] = 0;
  for (uint32_t c = tid; c < NK; c += THREADS) { hdr[c] = MODE == 2 ? a.hdr_old[c] : a.hdr[c]; bstart[c] = a.bstart[c]; }
  for (int c = tid; c < a.ylen; c += THREADS) bpos[c] = a.bpos[c];
  if (tid == 0) hdr[NK] = make_uint2 (0, 0);
  __syncthreads ();
  const uint32_t mask = a.ring - 1;
  constexpr int STEP = THREADS * 8;
  unsigned long long hits = 0;
  uint32_t found = 0;
  const long long t0 = clock64 ();
  for (int i0 = 0; i0 < a.n; i0 += STEP) {
    if (MODE < 2) {
      const int ib = i0 + tid * 8;
      const uint4 cv = *(const uint4*) (a.xcodes + ib);
      const uint32_t cw[4] = { cv.x, cv.y, cv.z, cv.w };
      uint32_t tails = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        uint32_t code = (r & 1) ? (cw[r >> 1] >> 16) : (cw[r >> 1] & 0xFFFFu);
        code = code < NK ? code : NK;
        const uint32_t i = (uint32_t) (ib + r);
        const uint2 h = hdr[code];
        const uint32_t len = h.y >> 24;
        hits += len;
        red_if (cnt, (i + (h.x & 0x3FFFu)) & mask, len > 0);
        red_if (cnt, (i + ((h.x >> 14) & 0x3FFFu)) & mask, len > 1);
        red_if (cnt, (i + (__funnelshift_r (h.x, h.y, 28) & 0x3FFFu)) & mask, len > 2);
        red_if (cnt, (i + ((h.y >> 10) & 0x3FFFu)) & mask, len > 3);
        if (MODE == 1) tails |= (len > 4 ? 1u : 0u) << r;
      }
      if (MODE == 1) {
        while (tails) {
          const int r = __ffs ((int) tails) - 1;
          tails &= tails - 1;
          const uint32_t i = (uint32_t) (ib + r);
          const uint32_t code = a.xcodes[i];
          const uint32_t len = hdr[code].y >> 24;
          const uint16_t* bp = bpos + bstart[code];
          for (uint32_t t = 4; t < len; ++t) atomicAdd (&cnt[(i + bp[t]) & mask], 1u);
        }
      }
    } else {
      uint32_t code[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) { const int i = i0 + r * THREADS + tid; code[r] = a.xcodes[i]; }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (code[r] != 0xFFFFu) {
          const uint32_t i = (uint32_t) (i0 + r * THREADS + tid);
          const uint2 h = hdr[code[r]];
          const uint32_t len = h.y & 0xFFFFu;
          if (len > 0) atomicAdd (&cnt[(i + (h.x & 0xFFFFu)) & mask], 1u);
          if (len > 1) atomicAdd (&cnt[(i + (h.x >> 16)) & mask], 1u);
          if (len > 2) {
            const uint16_t* bp = bpos + (h.y >> 16);
            atomicAdd (&cnt[(i + bp[2]) & mask], 1u);
            if (len > 3) atomicAdd (&cnt[(i + bp[3]) & mask], 1u);
            if (len > 4) atomicAdd (&cnt[(i + bp[4]) & mask], 1u);
            if (len > 5) atomicAdd (&cnt[(i + bp[5]) & mask], 1u);
            for (uint32_t t = 6; t < len; ++t) atomicAdd (&cnt[(i + bp[t]) & mask], 1u);
          }
          hits += len;
        }
      }
    }
    __syncthreads ();
    // finished window: diagonals below i0 + STEP - span, here simply the STEP slots starting at (i0 & ~3)
    bool hot = false;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4* p4 = (uint4*) (cnt + (((uint32_t) i0 + (uint32_t) (q * THREADS + tid) * 4) & mask));
      const uint4 c = *p4;
      *p4 = make_uint4 (0u, 0u, 0u, 0u);
      hot = hot || (int) c.x >= a.threshold || (int) c.y >= a.threshold || (int) c.z >= a.threshold || (int) c.w >= a.threshold;
    }
    if (__syncthreads_or (hot ? 1 : 0)) ++found;
  }
  const long long t1 = clock64 ();
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync (0xffffffffu, hits, o);
  if (lane == 0) atomicAdd (a.hits, hits);
  if (tid == 0) { a.cyc[blockIdx.x] = t1 - t0; atomicAdd (a.found, found); }
}

int main () {
  const int YLEN = 8000, K = 6, N = 1 << 20;
  srand (7);
  std::vector<uint8_t> y (YLEN), x (N + K + 64);
  for (auto& v : y) v = rand () & 3;
  for (auto& v : x) v = rand () & 3;
  auto code_at = [&] (const std::vector<uint8_t>& s, int p) { uint32_t c = 0; for (int t = 0; t < K; ++t) c = c * 4 + s[p + t]; return c; };
  const int span = YLEN - K;
  std::vector<std::vector<uint32_t>> bucket (NK);
  for (int j = 0; j <= span; ++j) bucket[code_at (y, j)].push_back (span - j);
  std::vector<uint2> hdr (NK), hdr_old (NK);
  std::vector<uint16_t> bpos (YLEN + 8, 0), bstart (NK, 0);
  uint32_t run = 0; double tail_hits = 0;
  for (int c = 0; c < NK; ++c) {
    const size_t len = bucket[c].size ();
    bstart[c] = (uint16_t) run;
    for (size_t t = 0; t < len; ++t) bpos[run + t] = (uint16_t) bucket[c][t];
    uint64_t v = 0;
    for (size_t t = 0; t < std::min<size_t> (len, 4); ++t) v |= (uint64_t) bucket[c][t] << (14 * t);
    v |= (uint64_t) std::min<size_t> (len, 255) << 56;
    hdr[c] = make_uint2 ((uint32_t) v, (uint32_t) (v >> 32));
    uint32_t xx = 0;
    if (len > 0) xx = bucket[c][0];
    if (len > 1) xx |= bucket[c][1] << 16;
    hdr_old[c] = make_uint2 (xx, (run << 16) | (uint32_t) len);
    if (len > 4) tail_hits += len - 4;
    run += (uint32_t) len;
  }
  std::vector<uint16_t> xc (N + 8192, 0xFFFF);
  for (int i = 0; i < N; ++i) xc[i] = (uint16_t) code_at (x, i);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties (&p, dev);
  const int nsm = p.multiProcessorCount;
  uint16_t *dxc, *dbpos, *dbstart; uint2 *dh, *dho; unsigned long long* dhits; long long* dcyc; uint32_t* dfound;
  cudaMalloc (&dxc, (N + 8192) * 2); cudaMalloc (&dh, NK * 8); cudaMalloc (&dho, NK * 8); cudaMalloc (&dbpos, (YLEN + 8) * 2); cudaMalloc (&dbstart, NK * 2);
  cudaMalloc (&dhits, 8); cudaMalloc (&dcyc, 8 * nsm * 4); cudaMalloc (&dfound, 4);
  cudaMemcpy (dxc, xc.data (), (N + 8192) * 2, cudaMemcpyHostToDevice); cudaMemcpy (dh, hdr.data (), NK * 8, cudaMemcpyHostToDevice);
  cudaMemcpy (dho, hdr_old.data (), NK * 8, cudaMemcpyHostToDevice); cudaMemcpy (dbpos, bpos.data (), (YLEN + 8) * 2, cudaMemcpyHostToDevice);
  cudaMemcpy (dbstart, bstart.data (), NK * 2, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  Args a; a.xcodes = dxc; a.n = N; a.hdr = dh; a.hdr_old = dho; a.bpos = dbpos; a.bstart = dbstart; a.ylen = YLEN; a.hits = dhits; a.cyc = dcyc; a.threshold = 20; a.found = dfound;
  printf ("hits in tails (len > 4), fraction of positions' hits: per read index %.0f entries\n", tail_hits);
#define RUN(THREADS, MINB, MODE, RING, NAME) { a.ring = RING; const int grid = nsm * MINB; \
    const size_t sm = (size_t) RING * 4 + (NK + 1) * 8 + ((YLEN + 8) & ~7) * 2 + NK * 2 + 64; \
    cudaFuncSetAttribute (kF<THREADS, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
    kF<THREADS, MINB, MODE><<<grid, THREADS, sm>>> (a); cudaMemset (dhits, 0, 8); cudaEventRecord (e0); kF<THREADS, MINB, MODE><<<grid, THREADS, sm>>> (a); cudaEventRecord (e1); cudaEventSynchronize (e1); \
    float ms; cudaEventElapsedTime (&ms, e0, e1); \
    unsigned long long h = 0; cudaMemcpy (&h, dhits, 8, cudaMemcpyDeviceToHost); \
    std::vector<long long> cyc (grid); cudaMemcpy (cyc.data (), dcyc, 8 * grid, cudaMemcpyDeviceToHost); \
    double mc = 0; for (auto v : cyc) mc += v; mc /= grid; cudaError_t e = cudaGetLastError (); \
    printf ("%-40s thr=%4d x%d ring=%5d  %.3f ms  hits/SM=%.0f  hits/cycle/SM=%.2f (clock64) %.2f (event @1.965 GHz) %s\n", NAME, THREADS, MINB, RING, ms, (double) h / nsm, \
            (double) h / nsm / mc, (double) h / nsm / (ms * 1e-3 * 1.965e9), e == cudaSuccess ? "" : cudaGetErrorString (e)); }
  RUN (768, 2, 2, 16384, "L  round-1 loop");
  RUN (1024, 1, 2, 32768, "L  round-1 loop");
  RUN (768, 2, 0, 16384, "F4 fixed 4, no tails");
  RUN (768, 2, 1, 16384, "F4 fixed 4 + deferred tails");
  RUN (1024, 1, 0, 32768, "F4 fixed 4, no tails");
  RUN (1024, 1, 1, 32768, "F4 fixed 4 + deferred tails");
  RUN (512, 2, 1, 16384, "F4 fixed 4 + deferred tails");
  RUN (1024, 2, 1, 16384, "F4 fixed 4 + deferred tails");
  return 0;
}
