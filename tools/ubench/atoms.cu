// Micro-benchmark: shared-memory histogram increment throughput on sm_100a (hits / cycle / SM) for the access shapes
// the seeding kernel can choose between.  Not part of the product; numbers feed DESIGN.md 4.1.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template<int MODE>
__global__ void __launch_bounds__ (1024) k (int iters, uint32_t mask, unsigned long long* out, long long* cyc) {
  extern __shared__ uint32_t cnt[];
  for (uint32_t c = threadIdx.x; c <= mask; c += blockDim.x) cnt[c] = 0;
  __syncthreads ();
  uint32_t x = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64 ();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      x = x * 1664525u + 1013904223u;
      const uint32_t a = (x >> 9) & mask;
      if (MODE == 0) atomicAdd (&cnt[a], 1u);                                        // 32 lanes, random words (ATOMS.POPC.INC)
      if (MODE == 1) { if (((x >> 3) & 31u) < 12u) atomicAdd (&cnt[a], 1u); }        // ~12 of 32 lanes active
      if (MODE == 2) atomicAdd (&cnt[a >> 1], (a & 1u) ? 0x10000u : 1u);             // packed 16-bit counters (ATOMS.ADD)
      if (MODE == 3) atomicAdd (&cnt[((a & ~31u) | lane) & mask], 1u);               // conflict-free banks
      if (MODE == 4) { cnt[a] = cnt[a] + 1u; }                                       // non-atomic read-modify-write
      if (MODE == 5) atomicAdd (&cnt[a], 1u + (x >> 31));                            // generic ATOMS.ADD with a register operand
      if (MODE == 6) { acc += cnt[a]; }                                              // plain random LDS.32
      if (MODE == 7) { const uint2 h = ((const uint2*) cnt)[a >> 1]; acc += h.x + h.y; }   // random LDS.64
      if (MODE == 8) { if (((x >> 3) & 31u) < 20u) atomicAdd (&cnt[a], 1u); }        // ~20 of 32 lanes
      if (MODE == 9) { const uint32_t o = atomicAdd (&cnt[a], 1u); acc += o; }       // with return value
    }
  }
  const long long t1 = clock64 ();
  __syncthreads ();
  unsigned long long s = acc;
  for (uint32_t c = threadIdx.x; c <= mask; c += blockDim.x) s += cnt[c];
  atomicAdd (out, s);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE>
static void run (const char* name, int threads, int ctas_per_sm, uint32_t words, double frac) {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties (&p, dev);
  const int nsm = p.multiProcessorCount;
  const int grid = nsm * ctas_per_sm;
  const int iters = 4000;
  unsigned long long* out; long long* cyc;
  cudaMalloc (&out, 8); cudaMalloc (&cyc, 8 * grid); cudaMemset (out, 0, 8);
  cudaFuncSetAttribute (k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (words * 4));
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  k<MODE><<<grid, threads, words * 4>>> (100, words - 1, out, cyc);
  cudaEventRecord (e0);
  k<MODE><<<grid, threads, words * 4>>> (iters, words - 1, out, cyc);
  cudaEventRecord (e1); cudaEventSynchronize (e1);
  cudaError_t err = cudaGetLastError ();
  float ms = 0; cudaEventElapsedTime (&ms, e0, e1);
  long long* h = new long long[grid]; cudaMemcpy (h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
  double mc = 0; for (int i = 0; i < grid; ++i) mc += h[i]; mc /= grid;
  const double ops = (double) threads * ctas_per_sm * iters * 8 * frac;           // per SM
  printf ("%-28s thr=%4d x%d words=%6u  %.3f ms  cycles/CTA=%.0f  ops/cycle/SM=%.2f  (by event at 1.965GHz: %.2f) %s\n", name, threads, ctas_per_sm, words, ms, mc,
          ops / mc, ops / (ms * 1e-3 * 1.965e9), err == cudaSuccess ? "" : cudaGetErrorString (err));
  cudaFree (out); cudaFree (cyc); delete[] h;
}

int main () {
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int thr = cfg == 0 ? 1024 : cfg == 1 ? 768 : 512;
    const int cps = cfg == 0 ? 1 : 2;
    const uint32_t words = cfg == 0 ? 32768 : 16384;
    run<0> ("atomicAdd 1, 32 lanes", thr, cps, words, 1.0);
    run<1> ("atomicAdd 1, ~12 lanes", thr, cps, words, 12.0 / 32);
    run<8> ("atomicAdd 1, ~20 lanes", thr, cps, words, 20.0 / 32);
    run<2> ("packed u16 add", thr, cps, words, 1.0);
    run<3> ("atomicAdd 1, conflict-free", thr, cps, words, 1.0);
    run<4> ("non-atomic rmw", thr, cps, words, 1.0);
    run<5> ("atomicAdd reg", thr, cps, words, 1.0);
    run<9> ("atomicAdd 1 with return", thr, cps, words, 1.0);
    run<6> ("LDS.32 random", thr, cps, words, 1.0);
    run<7> ("LDS.64 random", thr, cps, words, 1.0);
  }
  return 0;
}
