// Do a big-shared-memory kernel (1 CTA/SM, 176 KB dynamic) and small one-warp CTAs of another stream share an SM?
// A spins ~2 ms per CTA, 2 waves; B spins ~0.5 ms per CTA, 16 CTAs per SM.  Reports wall time of A alone, B alone, both.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__ (1024, 1) kA (long long cyc, int* sink) {
  extern __shared__ int sm[];
  sm[threadIdx.x] = threadIdx.x;
  const long long t0 = clock64 ();
  while (clock64 () - t0 < cyc) { sm[(threadIdx.x * 7) & 1023] += 1; }
  if (sm[threadIdx.x] == -1) *sink = 1;
}
__global__ void __launch_bounds__ (32) kB (long long cyc, int* sink) {
  const long long t0 = clock64 ();
  int v = 0;
  while (clock64 () - t0 < cyc) { v += 1; }
  if (v == -1) *sink = 1;
}
int main () {
  cudaDeviceProp p; cudaGetDeviceProperties (&p, 0);
  const int nsm = p.multiProcessorCount;
  int* sink; cudaMalloc (&sink, 4);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags (&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags (&s2, cudaStreamNonBlocking);
  const size_t smA = 176 * 1024;
  cudaFuncSetAttribute (kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smA);
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 1) { cudaFuncSetAttribute (kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    if (mode == 2) { cudaFuncSetAttribute (kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                     cudaFuncSetAttribute (kA, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    float tA, tB, tAB;
    for (int rep = 0; rep < 2; ++rep) {
      cudaDeviceSynchronize (); cudaEventRecord (e0, 0);
      kA<<<nsm * 2, 1024, smA, s1>>> (4000000, sink); cudaStreamSynchronize (s1);
      cudaEventRecord (e1, 0); cudaEventSynchronize (e1); cudaEventElapsedTime (&tA, e0, e1);
      cudaDeviceSynchronize (); cudaEventRecord (e0, 0);
      kB<<<nsm * 16, 32, 0, s2>>> (1000000, sink); cudaStreamSynchronize (s2);
      cudaEventRecord (e1, 0); cudaEventSynchronize (e1); cudaEventElapsedTime (&tB, e0, e1);
      cudaDeviceSynchronize (); cudaEventRecord (e0, 0);
      kA<<<nsm * 2, 1024, smA, s1>>> (4000000, sink);
      kB<<<nsm * 16, 32, 0, s2>>> (1000000, sink);
      cudaStreamSynchronize (s1); cudaStreamSynchronize (s2);
      cudaEventRecord (e1, 0); cudaEventSynchronize (e1); cudaEventElapsedTime (&tAB, e0, e1);
    }
    printf ("mode %d (%s): A alone %.2f ms, B alone %.2f ms, A+B on two streams %.2f ms  -> %s  %s\n", mode,
            mode == 0 ? "default carveouts" : mode == 1 ? "B prefers max shared" : "A and B prefer max shared", tA, tB, tAB,
            tAB < 0.8f * (tA + tB) ? "CONCURRENT" : "serialised", cudaGetErrorString (cudaGetLastError ()));
  }
  return 0;
}
