// Seeding: k-mer diagonal histogram -> seed diagonals -> band union -> runs of consecutive diagonals.
// Replaces KmerIndex (src/fastseq.cpp:240-256) + DiagonalEnvelope::initSparse (src/diagenv.cpp:20-106)
// of the reference, bit-exactly (integer work).
//
// Data layout (HBM):
//   codes[set]     uint16 per sequence position: base-4 k-mer starting there (first base most
//                  significant, fastseq.cpp:27-35), 0xFFFF where no k-mer starts.  2 B/position.
//   items[]        (pair, diagonal chunk) work units; one CTA each.
//   item_runs[]    up to run_cap [lo,hi] diagonal runs per item, ascending.
// Per CTA (shared memory): the read's bucket index (4^k+1 offsets + positions, uint16) and a ring
// of 32-bit per-diagonal counters that slides along the reference with the diagonal window, so every
// histogram increment is a shared-memory atomic and HBM only sees the 2 B/position code stream.
#ifndef QG_SEED_CUH
#define QG_SEED_CUH
#include "qg_common.cuh"

#define QG_SEED_THREADS 256
#define QG_SEED_STEP 2048            // reference positions consumed between two emit scans
#define QG_SEED_CHUNK (192 * 1024)   // diagonals owned by one work item

struct qg_seed_item {
  uint32_t pair;
  int32_t d_begin, d_end;            // diagonals [d_begin, d_end) owned by this item
};

struct qg_pair_desc {
  uint32_t xseq, yseq, xlen, ylen;
  uint64_t xoff, yoff;               // position offsets into the code / token arrays
  uint32_t item_begin, item_end;     // this pair's items (empty when the envelope is full)
  uint32_t full;                     // 1 = initFull (diagenv.cpp:11-18, 23-29)
  uint32_t run_out;                  // offset of this pair's merged runs in pair_runs
};

// ---- 2-bit packing + k-mer codes ----------------------------------------------------------------
__global__ void qg_pack_kernel (const uint8_t* __restrict__ tok, const uint64_t* __restrict__ off, const uint64_t* __restrict__ poff,
                                uint32_t nseq, uint64_t nwords, uint64_t* __restrict__ packed) {
  const uint64_t w = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nwords) {
    uint32_t lo = 0, hi = nseq;                        // last s with poff[s] <= w
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (poff[mid] <= w) lo = mid; else hi = mid; }
    const uint64_t len = off[lo + 1] - off[lo];
    const uint64_t p0 = (w - poff[lo]) * 32;
    uint64_t v = 0;
    for (int t = 0; t < 32; ++t)
      if (p0 + t < len) v |= (uint64_t) (tok[off[lo] + p0 + t] & 3) << (2 * t);
    packed[w] = v;
  }
}

__global__ void qg_codes_kernel (const uint8_t* __restrict__ tok, const uint64_t* __restrict__ off, uint32_t nseq,
                                 uint64_t total, int k, uint16_t* __restrict__ codes) {
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g < total) {
    uint32_t lo = 0, hi = nseq;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (off[mid] <= g) lo = mid; else hi = mid; }
    const uint64_t end = off[lo + 1];
    uint32_t code = 0xFFFF;
    if (g + k <= end) {
      code = 0;
      for (int t = 0; t < k; ++t) code = code * 4 + tok[g + t];
    }
    codes[g] = (uint16_t) code;
  }
}

// ---- the histogram kernel -------------------------------------------------------------------------
// shared memory: cnt[ring] u32 | boff[nk+1] u16 | bpos[ymax] u16 | seedmask[STEP/32 + 2] u32 | small state
__global__ void __launch_bounds__ (QG_SEED_THREADS)
qg_seed_kernel (const qg_seed_item* __restrict__ items, const qg_pair_desc* __restrict__ pairs,
                const uint16_t* __restrict__ xcodes, const uint16_t* __restrict__ ycodes,
                int k, int threshold, int half_band, uint32_t ring, uint32_t ymax, uint32_t run_cap,
                int2* __restrict__ item_runs, uint32_t* __restrict__ item_nruns, unsigned long long* __restrict__ hit_counter,
                uint32_t* __restrict__ overflow_flag) {
  QG_DYN_SMEM (smem);
  const uint32_t nk = 1u << (2 * k);
  uint32_t* cnt = (uint32_t*) smem;
  uint16_t* boff = (uint16_t*) (cnt + ring);
  uint16_t* bpos = boff + ((nk + 2) & ~1u);
  uint32_t* seedmask = (uint32_t*) (bpos + ((ymax + 1) & ~1u));
  __shared__ uint32_t s_warp_tot[QG_SEED_THREADS / 32];
  __shared__ int s_open_lo, s_open_hi, s_have_open;
  __shared__ uint32_t s_nruns;
  __shared__ int s_any;

  const qg_seed_item it = items[blockIdx.x];
  const qg_pair_desc pd = pairs[it.pair];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  const int nyk = ylen - k + 1;                          // read k-mer starts j in [0, nyk)
  const uint16_t* yc = ycodes + pd.yoff;
  const uint16_t* xc = xcodes + pd.xoff;
  const uint32_t mask = ring - 1;

  // -- 1. bucket index of the read: counting sort of k-mer start positions by k-mer code
  for (uint32_t c = tid; c < ring; c += QG_SEED_THREADS) cnt[c] = 0;
  if (tid == 0) { s_have_open = 0; s_nruns = 0; s_open_lo = 0; s_open_hi = 0; s_any = 0; }
  __syncthreads ();
  for (int j = tid; j < nyk; j += QG_SEED_THREADS) atomicAdd (&cnt[yc[j]], 1u);
  __syncthreads ();
  {
    // exclusive scan of cnt[0..nk) -> boff; each thread owns a contiguous slice
    const uint32_t per = (nk + QG_SEED_THREADS - 1) / QG_SEED_THREADS;
    const uint32_t b = tid * per, e = (b + per < nk) ? b + per : nk;
    uint32_t sum = 0;
    for (uint32_t c = b; c < e && c < nk; ++c) sum += cnt[c];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync (QG_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads ();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_warp_tot[w];
    uint32_t run = wbase + incl - sum;
    for (uint32_t c = b; c < e && c < nk; ++c) { const uint32_t v = cnt[c]; boff[c] = (uint16_t) run; cnt[c] = run; run += v; }
    if (tid == QG_SEED_THREADS - 1) boff[nk] = (uint16_t) nyk;
  }
  __syncthreads ();
  for (int j = tid; j < nyk; j += QG_SEED_THREADS) { const uint32_t slot = atomicAdd (&cnt[yc[j]], 1u); bpos[slot] = (uint16_t) j; }
  __syncthreads ();
  for (uint32_t c = tid; c < ring; c += QG_SEED_THREADS) cnt[c] = 0;
  __syncthreads ();

  // -- 2. slide along the reference
  const int span = ylen - k;                              // largest j
  const int d_begin = it.d_begin, d_end = it.d_end;
  int i_begin = d_begin > 0 ? d_begin : 0;
  int i_last = d_end - 1 + span;                          // inclusive
  if (i_last > xlen - k) i_last = xlen - k;
  const int min_diag = 1 - ylen, max_diag = xlen - 1;
  int emit_lo = d_begin;
  unsigned long long my_hits = 0;

  for (int i0 = i_begin; i0 <= i_last || emit_lo < d_end; i0 += QG_SEED_STEP) {
    const int i1 = (i0 + QG_SEED_STEP <= i_last + 1) ? i0 + QG_SEED_STEP : i_last + 1;    // exclusive
    for (int i = i0 + tid; i < i1; i += QG_SEED_THREADS) {
      const uint32_t code = xc[i];
      if (code != 0xFFFFu) {
        const uint32_t b0 = boff[code], b1 = boff[code + 1];
        for (uint32_t b = b0; b < b1; ++b) {
          const int d = i - (int) bpos[b];
          if (d >= d_begin && d < d_end) { atomicAdd (&cnt[(uint32_t) (d + ylen) & mask], 1u); ++my_hits; }
        }
      }
    }
    __syncthreads ();
    // diagonals below i1 - span can receive no further hits
    int emit_hi = (i1 > i_last) ? d_end : i1 - span;
    if (emit_hi > d_end) emit_hi = d_end;
    if (emit_hi > emit_lo) {
      const int ngroups = (emit_hi - emit_lo + 31) / 32;   // <= STEP/32 + 1 except for the final flush
      for (int g0 = 0; g0 < ngroups; g0 += QG_SEED_STEP / 32) {
        const int gcount = (ngroups - g0 < QG_SEED_STEP / 32) ? ngroups - g0 : QG_SEED_STEP / 32;
        for (int g = wid; g < gcount; g += QG_SEED_THREADS / 32) {
          const int d = emit_lo + (g0 + g) * 32 + lane;
          uint32_t c = 0;
          if (d < emit_hi) { const uint32_t idx = (uint32_t) (d + ylen) & mask; c = cnt[idx]; cnt[idx] = 0; }
          const uint32_t m = __ballot_sync (QG_FULL_MASK, d < emit_hi && (int) c >= threshold && c > 0);
          if (lane == 0) { seedmask[g] = m; if (m) s_any = 1; }
        }
        __syncthreads ();
        if (tid == 0 && s_any) {
          s_any = 0;
          // seeds in ascending order -> union of [seed-half, seed+half] clipped to the matrix (diagenv.cpp:79-84)
          int open_lo = s_open_lo, open_hi = s_open_hi, have = s_have_open;
          uint32_t nr = s_nruns;
          for (int g = 0; g < gcount; ++g) {
            uint32_t m = seedmask[g];
            while (m) {
              const int b = __ffs ((int) m) - 1;
              m &= m - 1;
              const int seed = emit_lo + (g0 + g) * 32 + b;
              int lo = seed - half_band, hi = seed + half_band;
              if (lo < min_diag) lo = min_diag;
              if (hi > max_diag) hi = max_diag;
              if (have && lo <= open_hi + 1) { if (hi > open_hi) open_hi = hi; }
              else {
                if (have) { if (nr < run_cap) item_runs[(size_t) blockIdx.x * run_cap + nr] = make_int2 (open_lo, open_hi); else *overflow_flag = 1; ++nr; }
                open_lo = lo; open_hi = hi; have = 1;
              }
            }
          }
          s_open_lo = open_lo; s_open_hi = open_hi; s_have_open = have; s_nruns = nr;
        }
        __syncthreads ();
      }
      emit_lo = emit_hi;
    }
    if (i1 > i_last && emit_lo >= d_end) break;
  }
  if (tid == 0) {
    uint32_t nr = s_nruns;
    if (s_have_open) { if (nr < run_cap) item_runs[(size_t) blockIdx.x * run_cap + nr] = make_int2 (s_open_lo, s_open_hi); else *overflow_flag = 1; ++nr; }
    item_nruns[blockIdx.x] = nr < run_cap ? nr : run_cap;
  }
  // hit statistics
  for (int o = 16; o > 0; o >>= 1) my_hits += __shfl_down_sync (QG_FULL_MASK, my_hits, o);
  if (lane == 0 && my_hits) atomicAdd (hit_counter, my_hits);
}

// ---- merge item runs of a pair, add diagonal 0, count cells ---------------------------------------
// One thread per pair (runs per pair are few).  pair_runs[pd.run_out ...] receives the final, sorted,
// maximal runs; pair_info[p] = {n_runs, n_diagonals}; pair_cu[p] = iterated cells.
__device__ __forceinline__ unsigned long long qg_run_cells (int lo, int hi, int xlen, int ylen) {
  unsigned long long cu = 0;
  for (int d = lo; d <= hi; ++d) {
    const int jlo = (1 - d > 1) ? 1 - d : 1;
    const int jhi = (xlen - d < ylen) ? xlen - d : ylen;
    if (jhi >= jlo) cu += (unsigned long long) (jhi - jlo + 1);
  }
  return cu;
}

__global__ void qg_envelope_finalize_kernel (const qg_pair_desc* __restrict__ pairs, uint32_t npairs,
                                             const int2* __restrict__ item_runs, const uint32_t* __restrict__ item_nruns, uint32_t run_cap,
                                             int2* __restrict__ pair_runs, uint2* __restrict__ pair_info, unsigned long long* __restrict__ pair_cu) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_desc pd = pairs[p];
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  int2* out = pair_runs + pd.run_out;
  uint32_t n = 0, ndiag = 0;
  unsigned long long cu = 0;
  if (pd.full) {
    out[0] = make_int2 (1 - ylen, xlen - 1);
    n = 1;
  } else {
    // stream the item runs (ascending) with the always-present run [0,0] (diagenv.cpp:52-54) merged in
    bool zero_done = false, have = false;
    int lo = 0, hi = 0;
    uint32_t it = pd.item_begin, r = 0;
    while (true) {
      int2 cur; bool got = false;
      while (it < pd.item_end && r >= item_nruns[it]) { ++it; r = 0; }
      const bool more = it < pd.item_end;
      if (more) cur = item_runs[(size_t) it * run_cap + r];
      if (!zero_done && (!more || cur.x > 0)) { cur = make_int2 (0, 0); zero_done = true; got = true; }
      else if (more) { ++r; got = true; }
      if (!got) break;
      if (have && cur.x <= hi + 1) { if (cur.y > hi) hi = cur.y; }
      else { if (have) out[n++] = make_int2 (lo, hi); lo = cur.x; hi = cur.y; have = true; }
    }
    if (have) out[n++] = make_int2 (lo, hi);
  }
  for (uint32_t t = 0; t < n; ++t) { ndiag += (uint32_t) (out[t].y - out[t].x + 1); cu += qg_run_cells (out[t].x, out[t].y, xlen, ylen); }
  pair_info[p] = make_uint2 (n, ndiag);
  pair_cu[p] = cu;
}

#endif
