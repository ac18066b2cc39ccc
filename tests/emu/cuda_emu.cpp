#include <mutex>
// TEST INFRASTRUCTURE ONLY -- see cuda_emu.h.
#include "cuda_emu.h"

namespace qgemu {
thread_local uint3_e t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local BlockState* t_block;

// statically declared __shared__ variables are plain statics here, so one block runs at a time -- also when several host
// threads (the contexts of a qg_pool) launch concurrently
static std::mutex g_launch_mx;

void launch (dim3 grid, dim3 block, size_t smem, const std::function<void ()>& body) {
  std::lock_guard<std::mutex> one_launch (g_launch_mx);
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  BlockState bs;
  pthread_barrier_init (&bs.block_bar, NULL, nthreads);
  bs.warp_bar.resize (nwarps);
  for (unsigned w = 0; w < nwarps; ++w)
    pthread_barrier_init (&bs.warp_bar[w], NULL, std::min (32u, nthreads - w * 32));
  bs.xchg.assign (nwarps * 32, 0);
  bs.or_flag = 0;
  bs.dyn_smem = (unsigned char*) malloc (smem + 16);
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      memset (bs.dyn_smem, 0xCD, smem);     // poison: kernels must initialise their shared memory
      std::vector<std::thread> th;
      th.reserve (nthreads);
      for (unsigned t = 0; t < nthreads; ++t)
        th.emplace_back ([&, t, bx, by] () {
          t_threadIdx = { t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
          t_blockIdx = { bx, by, 0 };
          t_blockDim = block;
          t_gridDim = grid;
          t_block = &bs;
          body ();
        });
      for (auto& x : th) x.join ();
    }
  pthread_barrier_destroy (&bs.block_bar);
  for (auto& b : bs.warp_bar) pthread_barrier_destroy (&b);
  free (bs.dyn_smem);
}
}  // namespace qgemu
