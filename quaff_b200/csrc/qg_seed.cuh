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

#ifndef QG_SEED_THREADS
#define QG_SEED_THREADS 768
#endif
#ifndef QG_SEED_MINB
#define QG_SEED_MINB 2             /* two CTAs per SM: caps registers at 40 */
#endif
#define QG_SEED_PPT 8                // reference positions per thread per step
#define QG_SEED_STEP (QG_SEED_THREADS * QG_SEED_PPT)   // reference positions consumed between two emit scans
#define QG_SEED_CHUNK (192 * 1024)   // diagonals owned by one work item

struct qg_seed_item {
  uint32_t pair;
  int32_t d_begin, d_end;            // diagonals [d_begin, d_end) owned by this item
};

struct qg_pair_desc {
  uint32_t xseq, yseq, xlen, ylen;
  uint64_t xoff, yoff;               // position offsets into the code / token arrays
  uint32_t item_begin, item_end;     // this pair's items (empty when the envelope is full)
  uint32_t full;                     // 1 = initFull (diagenv.cpp:11-18, 23-29); 2 = memory-guided, runs already in pair_runs
  uint32_t run_out;                  // offset of this pair's merged runs in pair_runs
  uint32_t run_cap;                  // memory-guided mode: capacity of that region
  uint32_t pad_;
  uint64_t idx_off;                  // general seeding path: offset of the read's sorted (k-mer, position) index
  uint64_t count_off;                // memory-guided mode: offset of this pair's per-diagonal counts
  uint64_t bits_off;                 // memory-guided mode: word offset of this pair's three bitmaps
  uint64_t hist_off;                 // memory-guided mode: offset of this pair's count histogram
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
// shared memory: cnt[ring] u32 | hdr[nk] uint2 | bpos[ymax] u16 | seedmask[STEP/32 + 2] u32
//   hdr[code] = { first two bucket entries (16 bit each), start << 16 | length }   one LDS.64 per reference position
//   bpos holds, bucket by bucket, span - j (span = yLen - k) for every k-mer start j of the read, so that a hit of
//   reference position i lands on ring slot (i + bpos) & (ring-1) = (d + span) & (ring-1) with one add and one and.
// A finished window of diagonals is read and cleared with 128-bit accesses; the seeds are only collected if some
// counter of the block reached the threshold (block-wide OR), otherwise the window is just zeroed.
template<bool COUNTS>
__global__ void __launch_bounds__ (QG_SEED_THREADS, QG_SEED_MINB)
qg_seed_kernel (const qg_seed_item* __restrict__ items, const qg_pair_desc* __restrict__ pairs,
                const uint16_t* __restrict__ xcodes, const uint16_t* __restrict__ ycodes,
                int k, int threshold, int half_band, uint32_t ring, uint32_t ymax, uint32_t run_cap,
                int2* __restrict__ item_runs, uint32_t* __restrict__ item_nruns, unsigned long long* __restrict__ hit_counter,
                uint32_t* __restrict__ overflow_flag, uint32_t* __restrict__ counts_out) {
  QG_DYN_SMEM (smem);
  const uint32_t nk = 1u << (2 * k);
  uint32_t* cnt = (uint32_t*) smem;
  uint2* hdr = (uint2*) (cnt + ring);
  uint16_t* bpos = (uint16_t*) (hdr + nk);
  uint32_t* seedmask = (uint32_t*) (bpos + ((ymax + 2) & ~1u));
  __shared__ uint32_t s_warp_tot[QG_SEED_THREADS / 32];
  __shared__ int s_open_lo, s_open_hi, s_have_open;
  __shared__ uint32_t s_nruns;

  const qg_seed_item it = items[blockIdx.x];
  const qg_pair_desc pd = pairs[it.pair];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  const int nyk = ylen - k + 1;                          // read k-mer starts j in [0, nyk)
  const int span = ylen - k;                             // largest j
  const uint16_t* yc = ycodes + pd.yoff;
  const uint16_t* xc = xcodes + pd.xoff;
  const uint32_t mask = ring - 1;

  // -- 1. bucket index of the read: counting sort of k-mer start positions by k-mer code
  for (uint32_t c = tid; c < ring; c += QG_SEED_THREADS) cnt[c] = 0;
  if (tid == 0) { s_have_open = 0; s_nruns = 0; s_open_lo = 0; s_open_hi = 0; }
  __syncthreads ();
  for (int j = tid; j < nyk; j += QG_SEED_THREADS) atomicAdd (&cnt[yc[j]], 1u);
  __syncthreads ();
  const uint32_t per = (nk + QG_SEED_THREADS - 1) / QG_SEED_THREADS;
  const uint32_t cb = tid * per, ce = (cb + per < nk) ? cb + per : nk;
  {
    // exclusive scan of cnt[0..nk) -> bucket starts; each thread owns a contiguous slice
    uint32_t sum = 0;
    for (uint32_t c = cb; c < ce; ++c) sum += cnt[c];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync (QG_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads ();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_warp_tot[w];
    uint32_t run = wbase + incl - sum;
    for (uint32_t c = cb; c < ce; ++c) { const uint32_t v = cnt[c]; hdr[c] = make_uint2 (0u, (run << 16) | v); cnt[c] = run; run += v; }
  }
  __syncthreads ();
  for (int j = tid; j < nyk; j += QG_SEED_THREADS) { const uint32_t slot = atomicAdd (&cnt[yc[j]], 1u); bpos[slot] = (uint16_t) (span - j); }
  __syncthreads ();
  for (uint32_t c = cb; c < ce; ++c) {                   // the first two entries of every bucket ride in its header
    const uint32_t e = hdr[c].y, st = e >> 16, len = e & 0xFFFFu;
    uint32_t x = 0;
    if (len > 0) x = bpos[st];
    if (len > 1) x |= (uint32_t) bpos[st + 1] << 16;
    hdr[c].x = x;
  }
  for (uint32_t c = tid; c < ring; c += QG_SEED_THREADS) cnt[c] = 0;
  __syncthreads ();

  // -- 2. slide along the reference
  const int d_begin = it.d_begin, d_end = it.d_end;
  const int i_begin = d_begin > 0 ? d_begin : 0;
  int i_last = d_end - 1 + span;                          // inclusive
  if (i_last > xlen - k) i_last = xlen - k;
  const int min_diag = 1 - ylen, max_diag = xlen - 1;
  const uint32_t dlen = (uint32_t) (d_end - d_begin);
  int emit_lo = d_begin;                                  // (emit_lo + span) is always a multiple of 4
  uint32_t my_hits = 0;
  unsigned long long my_hits64 = 0;

  for (int i0 = i_begin; i0 <= i_last || emit_lo < d_end; i0 += QG_SEED_STEP) {
    const int i1 = (i0 + QG_SEED_STEP <= i_last + 1) ? i0 + QG_SEED_STEP : i_last + 1;    // exclusive
    uint32_t code[QG_SEED_PPT];
#pragma unroll
    for (int r = 0; r < QG_SEED_PPT; ++r) { const int i = i0 + r * QG_SEED_THREADS + tid; code[r] = (i < i1) ? (uint32_t) xc[i] : 0xFFFFu; }
#pragma unroll
    for (int r = 0; r < QG_SEED_PPT; ++r) {
      if (code[r] != 0xFFFFu) {
        const int i = i0 + r * QG_SEED_THREADS + tid;
        const uint2 h = hdr[code[r]];
        const uint32_t len = h.y & 0xFFFFu;
        if (i >= d_begin + span && i < d_end) {           // every diagonal this position can hit belongs to the item
          // bucket lengths are Poisson(~2): two entries come with the header, the next four are straight-line code,
          // the rare longer tails loop
          if (len > 0) atomicAdd (&cnt[((uint32_t) i + (h.x & 0xFFFFu)) & mask], 1u);
          if (len > 1) atomicAdd (&cnt[((uint32_t) i + (h.x >> 16)) & mask], 1u);
          if (len > 2) {
            const uint16_t* bp = bpos + (h.y >> 16);
            atomicAdd (&cnt[((uint32_t) i + bp[2]) & mask], 1u);
            if (len > 3) atomicAdd (&cnt[((uint32_t) i + bp[3]) & mask], 1u);
            if (len > 4) atomicAdd (&cnt[((uint32_t) i + bp[4]) & mask], 1u);
            if (len > 5) atomicAdd (&cnt[((uint32_t) i + bp[5]) & mask], 1u);
            for (uint32_t t = 6; t < len; ++t) atomicAdd (&cnt[((uint32_t) i + bp[t]) & mask], 1u);
          }
          my_hits += len;
        } else {
          const uint16_t* bp = bpos + (h.y >> 16);
          const uint32_t ib = (uint32_t) (i - span - d_begin);     // d - d_begin = ib + bp[t]
          for (uint32_t t = 0; t < len; ++t) {
            const uint32_t v = bp[t];
            if (ib + v < dlen) { atomicAdd (&cnt[((uint32_t) i + v) & mask], 1u); ++my_hits; }
          }
        }
      }
    }
    if (my_hits > 0x40000000u) { my_hits64 += my_hits; my_hits = 0; }
    __syncthreads ();
    // diagonals below i1 - span can receive no further hits; windows end on a 4-slot boundary except the last one
    int emit_hi = (i1 > i_last) ? d_end : (i1 - span) - (i1 & 3);
    if (emit_hi > d_end) emit_hi = d_end;
    if (emit_hi > emit_lo) {
      for (int base = emit_lo; base < emit_hi; base += QG_SEED_STEP) {
        uint32_t nib[QG_SEED_PPT / 4];
        bool hot = false;
#pragma unroll
        for (int q = 0; q < QG_SEED_PPT / 4; ++q) {
          const int d0 = base + (q * QG_SEED_THREADS + tid) * 4;
          nib[q] = 0;
          if (d0 < emit_hi) {
            uint4* p4 = (uint4*) (cnt + ((uint32_t) (d0 + span) & mask));
            const uint4 c = *p4;
            *p4 = make_uint4 (0u, 0u, 0u, 0u);
            const uint32_t cc[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (d0 + e < emit_hi) {
                if (COUNTS) counts_out[pd.count_off + (uint64_t) (d0 + e + span)] = cc[e];     // memory-guided mode: raw counts only
                else if ((int) cc[e] >= threshold && cc[e] > 0) nib[q] |= 1u << e;
              }
            }
          }
          hot = hot || nib[q] != 0;
        }
        if (__syncthreads_or (hot ? 1 : 0)) {
          for (int w = tid; w < QG_SEED_STEP / 32; w += QG_SEED_THREADS) seedmask[w] = 0;
          __syncthreads ();
#pragma unroll
          for (int q = 0; q < QG_SEED_PPT / 4; ++q)
            if (nib[q]) { const uint32_t bit = (uint32_t) (q * QG_SEED_THREADS + tid) * 4; atomicOr (&seedmask[bit >> 5], nib[q] << (bit & 31)); }
          __syncthreads ();
          if (tid == 0) {
            // seeds in ascending order -> union of [seed-half, seed+half] clipped to the matrix (diagenv.cpp:79-84)
            int open_lo = s_open_lo, open_hi = s_open_hi, have = s_have_open;
            uint32_t nr = s_nruns;
            for (int g = 0; g < QG_SEED_STEP / 32; ++g) {
              uint32_t m = seedmask[g];
              while (m) {
                const int bb = __ffs ((int) m) - 1;
                m &= m - 1;
                const int seed = base + g * 32 + bb;
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
      }
      __syncthreads ();
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
  my_hits64 += my_hits;
  for (int o = 16; o > 0; o >>= 1) my_hits64 += __shfl_down_sync (QG_FULL_MASK, my_hits64, o);
  if (lane == 0 && my_hits64) atomicAdd (hit_counter, my_hits64);
}

// ---- the tile-sorted histogram kernel (k = 5, 6; reads up to 16 380 + k bases) ------------------------------------
// The reference side is static, so it is prepared once per sequence set: every tile of QG_TILE_POS consecutive
// positions is sorted by k-mer code (qg_sort_tiles_kernel; entry = code << 16 | position in the tile, 4 B/position).
// Lanes of a warp then look up neighbouring codes -- the header loads become near-broadcasts instead of 32 random
// 8-byte reads -- and walk buckets of equal length.  The bucket header holds the first FOUR entries (4 x 14 bit + 8-bit
// length), issued as four predicated shared-memory increments without a branch; longer buckets (4.8 % of the codes at
// 8 kb) continue from the full bucket array under a warp vote.  What bounds the kernel is the shared-memory pipe
// (wavefronts: bank conflicts of a random scatter), measured in tools/ubench/seedloop2.cu; DESIGN.md 4.1.
// One CTA of QG_TSEED_THREADS threads per SM (the ring alone is 128 KB) at <= 40 registers, which leaves register file
// and warp slots for the DP kernels of the other contexts to run on the same SMs.
#ifdef QG_EMU                        /* CPU-thread shim of the tests: small tiles, so that short inputs cross tile and item boundaries */
#define QG_TILE_POS 1024
#define QG_TSEED_RING 4096u
#define QG_TSEED_THREADS 128
#else
#define QG_TILE_POS 16384
#define QG_TSEED_RING 32768u
#define QG_TSEED_THREADS 1024
#endif
#define QG_TSEED_ESTEP (QG_TSEED_THREADS * 8)      // diagonals one pass of the emit scan covers

struct qg_tile_job { uint64_t off; uint32_t len; uint32_t pad_; };   // one tile: positions [off, off + len) of the flat code array

__global__ void __launch_bounds__ (256)
qg_sort_tiles_kernel (const qg_tile_job* __restrict__ jobs, const uint16_t* __restrict__ codes, uint32_t nk, uint32_t* __restrict__ sorted) {
  __shared__ uint32_t cnt[4096 + 1];
  __shared__ uint32_t s_warp_tot[8];
  const qg_tile_job jb = jobs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (uint32_t c = tid; c <= nk; c += 256) cnt[c] = 0;
  __syncthreads ();
  for (uint32_t p = tid; p < jb.len; p += 256) { uint32_t c = codes[jb.off + p]; if (c > nk) c = nk; atomicAdd (&cnt[c], 1u); }
  __syncthreads ();
  const uint32_t per = (nk + 1 + 255) / 256;
  const uint32_t cb = tid * per, ce = (cb + per < nk + 1) ? cb + per : nk + 1;
  uint32_t sum = 0;
  for (uint32_t c = cb; c < ce; ++c) sum += cnt[c];
  uint32_t incl = sum;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync (QG_FULL_MASK, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp_tot[wid] = incl;
  __syncthreads ();
  uint32_t run = incl - sum;
  for (int w = 0; w < wid; ++w) run += s_warp_tot[w];
  for (uint32_t c = cb; c < ce; ++c) { const uint32_t v = cnt[c]; cnt[c] = run; run += v; }
  __syncthreads ();
  for (uint32_t p = tid; p < jb.len; p += 256) {
    uint32_t c = codes[jb.off + p]; if (c > nk) c = nk;            // nk: no k-mer starts here (the empty header)
    const uint32_t slot = atomicAdd (&cnt[c], 1u);
    sorted[jb.off + slot] = (c << 16) | (p << 2);                  // position pre-scaled to a byte offset into the counter ring
  }
}

// One shared-memory increment at byte offset `boff` of the ring if `pred`, else on the warp's dummy word: ptxas turns
// a PREDICATED shared atomic into a divergent branch around it (BSSY / BRA / ATOMS / BSYNC, measured: 8 instructions per
// slot), a select between two addresses costs one.  (A reduction predicated at the PTX level -- "@p red.shared.add" -- is
// turned into the same branch by ptxas 12.9.)
#define QG_RED_IF(cnt, boff, pred, dummy_off) atomicAdd ((uint32_t*) ((char*) (cnt) + ((pred) ? (boff) : (dummy_off))), 1u)

// shared memory: cnt[QG_TSEED_RING] u32 | hdr[nk + 1] uint2 | bstart[nk + 2] u16 | bpos[ymax + 2] u16 | seedmask[ESTEP/32 + 2] u32 | dummy[32] u32
//   hdr[code]  = the first four bucket entries as 16-bit fields holding 4 * (span - j), the byte offset a hit adds to the ring
//                address; 0xFFFF = empty slot; bit 0 of field 3 = "the bucket has more than four entries"; hdr[nk] = all empty
//                (positions without a k-mer)
//   bpos       = span - j for every k-mer start j of the read, bucket by bucket; bstart[code] = first entry of the bucket
//   ring slot of diagonal d = (d + span + off) & (ring - 1); off in 0..3 makes the item's first diagonal 16-byte aligned
// The hit statistic is the sum of the counters the emit scan reads (every hit of the item is one increment of one of them).
template<bool COUNTS>
__global__ void
#if defined(QG_TSEED_MAXNREG) && !defined(QG_EMU)
__maxnreg__ (QG_TSEED_MAXNREG)       /* leaves register file next to the one resident CTA for the DP kernels of the other contexts */
#else
__launch_bounds__ (QG_TSEED_THREADS, 1)
#endif
qg_seed_tile_kernel (const qg_seed_item* __restrict__ items, const qg_pair_desc* __restrict__ pairs,
                     const uint32_t* __restrict__ xsorted, const uint16_t* __restrict__ ycodes,
                     int k, int threshold, int half_band, uint32_t ymax, uint32_t run_cap,
                     int2* __restrict__ item_runs, uint32_t* __restrict__ item_nruns, unsigned long long* __restrict__ hit_counter,
                     uint32_t* __restrict__ overflow_flag, uint32_t* __restrict__ counts_out) {
  QG_DYN_SMEM (smem);
  constexpr int T = QG_TSEED_THREADS;
  const uint32_t nk = 1u << (2 * k);
  const uint32_t ring = QG_TSEED_RING, mask = ring - 1;
  uint32_t* cnt = (uint32_t*) smem;
  uint2* hdr = (uint2*) (cnt + ring);
  uint16_t* bstart = (uint16_t*) (hdr + nk + 1);
  uint16_t* bpos = bstart + ((nk + 2 + 3) & ~3u);
  uint32_t* seedmask = (uint32_t*) (bpos + ((ymax + 2 + 1) & ~1u));
  const uint32_t dummy_off = (uint32_t) ((char*) (seedmask + QG_TSEED_ESTEP / 32 + 2) - (char*) cnt) + 4u * (threadIdx.x >> 5);   // one word per warp: ATOMS.POPC.INC folds the lanes that share an address into one access
  __shared__ uint32_t s_warp_tot[T / 32];
  __shared__ int s_open_lo, s_open_hi, s_have_open;
  __shared__ uint32_t s_nruns;

  const qg_seed_item it = items[blockIdx.x];
  const qg_pair_desc pd = pairs[it.pair];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  const int nyk = ylen - k + 1, span = ylen - k;
  const uint16_t* yc = ycodes + pd.yoff;

  // -- 1. bucket index of the read (counting sort of its k-mer starts by code)
  for (uint32_t c = tid; c < ring; c += T) cnt[c] = 0;
  if (tid == 0) { s_have_open = 0; s_nruns = 0; s_open_lo = 0; s_open_hi = 0; }
  __syncthreads ();
  for (int j = tid; j < nyk; j += T) atomicAdd (&cnt[yc[j]], 1u);
  __syncthreads ();
  const uint32_t per = (nk + T - 1) / T;
  const uint32_t cb = tid * per < nk ? tid * per : nk, ce = (cb + per < nk) ? cb + per : nk;
  {
    uint32_t sum = 0;
    for (uint32_t c = cb; c < ce; ++c) sum += cnt[c];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync (QG_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads ();
    uint32_t run = incl - sum;
    for (int w = 0; w < wid; ++w) run += s_warp_tot[w];
    for (uint32_t c = cb; c < ce; ++c) { const uint32_t v = cnt[c]; bstart[c] = (uint16_t) run; cnt[c] = run; run += v; }
    if (tid == 0) { bstart[nk] = (uint16_t) nyk; bstart[nk + 1] = (uint16_t) nyk; }
  }
  __syncthreads ();
  for (int j = tid; j < nyk; j += T) { const uint32_t slot = atomicAdd (&cnt[yc[j]], 1u); bpos[slot] = (uint16_t) (span - j); }
  __syncthreads ();
  for (uint32_t c = cb; c < ce; ++c) {
    const uint32_t st = bstart[c], len = (uint32_t) bstart[c + 1] - st;
    uint32_t f[4];
    for (uint32_t t = 0; t < 4; ++t) f[t] = t < len ? (uint32_t) bpos[st + t] << 2 : 0xFFFFu;
    if (len > 4) f[3] |= 1u;
    hdr[c] = make_uint2 (f[0] | (f[1] << 16), f[2] | (f[3] << 16));
  }
  if (tid == 0) hdr[nk] = make_uint2 (0xFFFFFFFFu, 0xFFFFFFFFu);
  for (uint32_t c = tid; c < ring; c += T) cnt[c] = 0;
  __syncthreads ();

  // -- 2. tile by tile along the reference
  const int d_begin = it.d_begin, d_end = it.d_end;
  const int i_begin = d_begin > 0 ? d_begin : 0;
  int i_last = d_end - 1 + span;                          // inclusive
  if (i_last > xlen - k) i_last = xlen - k;
  const int min_diag = 1 - ylen, max_diag = xlen - 1;
  const uint32_t dlen = (uint32_t) (d_end - d_begin);
  const int off = (4 - ((d_begin + span) & 3)) & 3;
  const int spo = span + off;                             // (d_begin + spo) is a multiple of 4
  int emit_lo = d_begin;
  const uint32_t mask4 = mask << 2, dlen4 = dlen << 2;
  uint32_t my_hits = 0;                                   // per thread: sums of at most (chunk / T) windows of counters

  for (int tile = i_begin / QG_TILE_POS; tile <= i_last / QG_TILE_POS; ++tile) {
    const int tb = tile * QG_TILE_POS;
    const int tn = (xlen - tb < QG_TILE_POS) ? xlen - tb : QG_TILE_POS;
    const bool interior = tb >= d_begin + span && tb + tn <= d_end;      // every hit of every position belongs to the item
    const uint32_t* xs = xsorted + pd.xoff + (uint64_t) tb;
    const uint32_t tbo4 = (uint32_t) (tb + off) << 2;
    const uint32_t ibase4 = (uint32_t) (tb - span - d_begin) << 2;       // 4 (d - d_begin) = ibase4 + 4 position + field
    const bool whole = tn == QG_TILE_POS;                                // QG_TILE_POS is a multiple of 4 T: no bounds tests
    for (int e0 = 0; e0 < tn; e0 += 4 * T) {
      uint32_t ent[4];
      if (whole) {
        const uint32_t* xe = xs + e0 + tid;
#pragma unroll
        for (int r = 0; r < 4; ++r) ent[r] = xe[r * T];
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) { const int e = e0 + r * T + tid; ent[r] = e < tn ? xs[e] : (nk << 16); }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t code = ent[r] >> 16, pos4 = ent[r] & 0xFFFFu;
        const uint32_t io4 = tbo4 + pos4;
        const uint2 h = hdr[code];
        const uint32_t f0 = h.x & 0xFFFFu, f1 = h.x >> 16, f2 = h.y & 0xFFFFu, f3 = h.y >> 16;
        const bool more = (f3 & 1u) && f3 != 0xFFFFu;
        if (interior) {
          QG_RED_IF (cnt, (io4 + f0) & mask4, f0 != 0xFFFFu, dummy_off);
          QG_RED_IF (cnt, (io4 + f1) & mask4, f1 != 0xFFFFu, dummy_off);
          QG_RED_IF (cnt, (io4 + f2) & mask4, f2 != 0xFFFFu, dummy_off);
          QG_RED_IF (cnt, (io4 + f3) & mask4, f3 != 0xFFFFu, dummy_off);
        } else {
          const uint32_t ib4 = ibase4 + pos4;
          QG_RED_IF (cnt, (io4 + f0) & mask4, f0 != 0xFFFFu && ib4 + f0 < dlen4, dummy_off);
          QG_RED_IF (cnt, (io4 + f1) & mask4, f1 != 0xFFFFu && ib4 + f1 < dlen4, dummy_off);
          QG_RED_IF (cnt, (io4 + f2) & mask4, f2 != 0xFFFFu && ib4 + f2 < dlen4, dummy_off);
          QG_RED_IF (cnt, (io4 + f3) & mask4, f3 != 0xFFFFu && ib4 + (f3 & ~3u) < dlen4, dummy_off);
        }
        if (__any_sync (QG_FULL_MASK, more)) {            // buckets longer than four: continue from the full bucket array
          uint32_t st = 0, en = 0;
          if (more) { st = (uint32_t) bstart[code] + 4; en = bstart[code + 1]; }
          const uint32_t ib4 = ibase4 + pos4;
          do {
            const bool act = st < en;
            const uint32_t v4 = (uint32_t) bpos[act ? st : 0u] << 2;
            QG_RED_IF (cnt, (io4 + v4) & mask4, act && (interior || ib4 + v4 < dlen4), dummy_off);
            st += act ? 1u : 0u;
          } while (__any_sync (QG_FULL_MASK, st < en));
        }
      }
    }
    __syncthreads ();
    // diagonals below i1 - span can receive no further hits; windows end on a 4-slot boundary except the last one
    const int i1 = (tb + QG_TILE_POS <= i_last + 1) ? tb + QG_TILE_POS : i_last + 1;    // exclusive
    int emit_hi = (i1 > i_last) ? d_end : (i1 - span) - ((i1 + off) & 3);
    if (emit_hi > d_end) emit_hi = d_end;
    if (emit_hi > emit_lo) {
      for (int base = emit_lo; base < emit_hi; base += QG_TSEED_ESTEP) {
        uint32_t nib[2];
        bool hot = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d0 = base + (q * T + tid) * 4;
          nib[q] = 0;
          if (d0 < emit_hi) {
            uint4* p4 = (uint4*) (cnt + ((uint32_t) (d0 + spo) & mask));
            const uint4 c = *p4;
            *p4 = make_uint4 (0u, 0u, 0u, 0u);
            my_hits += c.x + c.y + c.z + c.w;
            const uint32_t cxy = c.x > c.y ? c.x : c.y, czw = c.z > c.w ? c.z : c.w, cmax = cxy > czw ? cxy : czw;
            if (COUNTS || (cmax > 0 && (int) cmax >= threshold)) {         // rare without COUNTS: some diagonal of the four is a seed
              const uint32_t cc[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (d0 + e < emit_hi) {
                  if (COUNTS) counts_out[pd.count_off + (uint64_t) (d0 + e + span)] = cc[e];     // memory-guided mode: raw counts only
                  else if ((int) cc[e] >= threshold && cc[e] > 0) nib[q] |= 1u << e;
                }
              }
            }
          }
          hot = hot || nib[q] != 0;
        }
        if (__syncthreads_or (hot ? 1 : 0)) {
          for (int w = tid; w < QG_TSEED_ESTEP / 32; w += T) seedmask[w] = 0;
          __syncthreads ();
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (nib[q]) { const uint32_t bit = (uint32_t) (q * T + tid) * 4; atomicOr (&seedmask[bit >> 5], nib[q] << (bit & 31)); }
          __syncthreads ();
          if (tid == 0) {
            // seeds in ascending order -> union of [seed-half, seed+half] clipped to the matrix (diagenv.cpp:79-84)
            int open_lo = s_open_lo, open_hi = s_open_hi, have = s_have_open;
            uint32_t nr = s_nruns;
            for (int g = 0; g < QG_TSEED_ESTEP / 32; ++g) {
              uint32_t m = seedmask[g];
              while (m) {
                const int bb = __ffs ((int) m) - 1;
                m &= m - 1;
                const int seed = base + g * 32 + bb;
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
      }
      __syncthreads ();
      emit_lo = emit_hi;
    }
  }
  if (tid == 0) {
    uint32_t nr = s_nruns;
    if (s_have_open) { if (nr < run_cap) item_runs[(size_t) blockIdx.x * run_cap + nr] = make_int2 (s_open_lo, s_open_hi); else *overflow_flag = 1; ++nr; }
    item_nruns[blockIdx.x] = nr < run_cap ? nr : run_cap;
  }
  unsigned long long hits64 = my_hits;
  for (int o = 16; o > 0; o >>= 1) hits64 += __shfl_down_sync (QG_FULL_MASK, hits64, o);
  if (lane == 0 && hits64) atomicAdd (hit_counter, hits64);
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
  if (pd.full == 2) {
    n = pair_info[p].x;                                     // runs were written by qg_memtier_kernel
  } else if (pd.full) {
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


// ---- memory-guided threshold (-kmatchmb / -kmatchmax): diagenv.cpp:62-96 with kmerThreshold < 0 -----------------
// Count tiers are visited in descending order; a tier's seeds (all diagonals with exactly that count) are accepted
// only if the storage diagonals of the enlarged envelope still fit: |storageDiags| * min(xLen,yLen) * cellSize < maxSize,
// tested BEFORE accepting (so a budget smaller than one band leaves diagonal 0 alone).  One CTA per pair; rarely used,
// so clarity over speed: two passes over the pair's per-diagonal counts per non-empty tier.
__global__ void __launch_bounds__ (256)
qg_memtier_kernel (const qg_pair_desc* __restrict__ pairs, const uint32_t* __restrict__ pair_ids, const uint32_t* __restrict__ counts,
                   uint32_t* __restrict__ hist_all, uint32_t* __restrict__ bits_all, int k, int half_band, int fixed_threshold,
                   unsigned long long cell_size, unsigned long long max_size,
                   int2* __restrict__ pair_runs, uint2* __restrict__ pair_info, uint32_t* __restrict__ overflow_flag) {
  __shared__ uint32_t s_cmax, s_new;
  __shared__ unsigned long long s_nstorage;
  __shared__ int s_stop;
  const uint32_t p = pair_ids[blockIdx.x];
  const qg_pair_desc pd = pairs[p];
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  const int span = ylen - k;
  const int ndiag = (xlen - k) + span + 1;                  // diagonals that can receive hits: d = t - span
  const uint32_t* cnt = counts + pd.count_off;
  uint32_t* hist = hist_all + pd.hist_off;                  // [ylen + 2]
  const int nbits = xlen + ylen + 3;                        // bit index = d + ylen + 1, d in [-ylen-1, xlen+1]
  const int nwords = (nbits + 31) / 32;
  uint32_t* accepted = bits_all + pd.bits_off;
  uint32_t* trial = accepted + nwords;
  uint32_t* env = trial + nwords;
  const int min_diag = 1 - ylen, max_diag = xlen - 1;
  const unsigned long long diag_size = (unsigned long long) (xlen < ylen ? xlen : ylen) * cell_size;
  const int tid = threadIdx.x;

  for (int w = tid; w < 3 * nwords; w += blockDim.x) accepted[w] = 0;
  for (int c = tid; c < ylen + 2; c += blockDim.x) hist[c] = 0;
  if (tid == 0) { s_cmax = 0; s_new = 0; s_nstorage = 1; s_stop = 0; }
  __syncthreads ();
  for (int t = tid; t < ndiag; t += blockDim.x) { const uint32_t c = cnt[t]; if (c) { atomicAdd (&hist[c], 1u); atomicMax (&s_cmax, c); } }
  if (tid == 0) { const int b = 0 + ylen + 1; accepted[b >> 5] |= 1u << (b & 31); env[b >> 5] |= 1u << (b & 31); }   // diags = storageDiags = {0}
  __syncthreads ();
  const uint32_t cmax = s_cmax;
  if (fixed_threshold >= 0) {
    // plain -kmatchn threshold on the general path's global counts (diagenv.cpp:62-96 with kmerThreshold >= 0)
    for (int t = tid; t < ndiag; t += blockDim.x)
      if (cnt[t] != 0 && cnt[t] >= (uint32_t) fixed_threshold) {   // only diagonals that received a hit exist in diagKmerCount (diagenv.cpp:33-46)
        const int seed = t - span;
        const int lo = seed - half_band > min_diag ? seed - half_band : min_diag;
        const int hi = seed + half_band < max_diag ? seed + half_band : max_diag;
        for (int d = lo; d <= hi; ++d) { const int b = d + ylen + 1; atomicOr (&env[b >> 5], 1u << (b & 31)); }
      }
  } else
  for (uint32_t c = cmax; c >= 1; --c) {
    if (hist[c] == 0) continue;
    for (int t = tid; t < ndiag; t += blockDim.x)
      if (cnt[t] == c) {
        const int seed = t - span;
        const int lo = (seed - half_band > min_diag ? seed - half_band : min_diag) - 1;
        const int hi = (seed + half_band < max_diag ? seed + half_band : max_diag) + 1;
        for (int d = lo; d <= hi; ++d) {
          const int b = d + ylen + 1; const uint32_t bit = 1u << (b & 31);
          if (!(accepted[b >> 5] & bit)) { const uint32_t old = atomicOr (&trial[b >> 5], bit); if (!(old & bit)) atomicAdd (&s_new, 1u); }
        }
      }
    __syncthreads ();
    if (tid == 0 && (s_nstorage + s_new) * diag_size >= max_size) s_stop = 1;      // diagenv.cpp:87-89
    __syncthreads ();
    if (s_stop) break;
    for (int w = tid; w < nwords; w += blockDim.x) { accepted[w] |= trial[w]; trial[w] = 0; }
    for (int t = tid; t < ndiag; t += blockDim.x)
      if (cnt[t] == c) {
        const int seed = t - span;
        const int lo = seed - half_band > min_diag ? seed - half_band : min_diag;
        const int hi = seed + half_band < max_diag ? seed + half_band : max_diag;
        for (int d = lo; d <= hi; ++d) { const int b = d + ylen + 1; atomicOr (&env[b >> 5], 1u << (b & 31)); }
      }
    __syncthreads ();
    if (tid == 0) { s_nstorage += s_new; s_new = 0; }
    __syncthreads ();
  }
  __syncthreads ();
  if (tid == 0) {
    // envelope bitmap -> maximal runs, ascending
    int2* out = pair_runs + pd.run_out;
    uint32_t n = 0; int run_lo = 0; bool open = false;
    for (int w = 0; w < nwords; ++w) {
      const uint32_t v = env[w];
      if (v == 0 && !open) continue;
      for (int b = 0; b < 32; ++b) {
        const bool set = (v >> b) & 1u;
        const int d = w * 32 + b - ylen - 1;
        if (set && !open) { run_lo = d; open = true; }
        else if (!set && open) { if (n < pd.run_cap) out[n] = make_int2 (run_lo, d - 1); else *overflow_flag = 2; ++n; open = false; }
      }
    }
    if (open) { if (n < pd.run_cap) out[n] = make_int2 (run_lo, nwords * 32 - 1 - ylen - 1); else *overflow_flag = 2; ++n; }
    pair_info[p] = make_uint2 (n < pd.run_cap ? n : pd.run_cap, 0);
  }
}


// ---- general seeding path: any k <= 32, any read length ----------------------------------------------------------
// Used when the shared-memory kernel above does not apply (k outside 5..7, or a read whose bucket index and counter
// ring do not fit one CTA's shared memory).  The read's k-mer starts are radix-sorted by k-mer (the GPU form of
// KmerIndex's map, fastseq.cpp:240-256); every reference position binary-searches its k-mer and increments the
// pair's per-diagonal counters in HBM (diagenv.cpp:33-41); qg_memtier_kernel turns the counters into runs.
__global__ void qg_codes64_kernel (const uint8_t* __restrict__ tok, const uint64_t* __restrict__ off, uint32_t nseq,
                                   uint64_t total, int k, unsigned long long* __restrict__ codes) {
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g < total) {
    uint32_t lo = 0, hi = nseq;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (off[mid] <= g) lo = mid; else hi = mid; }
    const uint64_t end = off[lo + 1];
    unsigned long long code = 0;
    if (g + k <= end) for (int t = 0; t < k; ++t) code = code * 4 + (tok[g + t] & 3);
    codes[g] = code;                                       // positions without a k-mer are never read (bounds come from the lengths)
  }
}

struct qg_index_job { uint64_t yoff; uint64_t idx_off; uint32_t n; uint32_t pad_; };   // one per distinct read: n = yLen - k + 1 k-mer starts

__global__ void qg_index_fill_kernel (const qg_index_job* __restrict__ jobs, const unsigned long long* __restrict__ ycodes,
                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const qg_index_job jb = jobs[blockIdx.y];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < jb.n; j += gridDim.x * blockDim.x) {
    keys[jb.idx_off + j] = ycodes[jb.yoff + j];
    vals[jb.idx_off + j] = j;
  }
}

__global__ void __launch_bounds__ (256)
qg_seed_general_kernel (const qg_seed_item* __restrict__ items, const qg_pair_desc* __restrict__ pairs,
                        const unsigned long long* __restrict__ xcodes, const unsigned long long* __restrict__ keys,
                        const uint32_t* __restrict__ vals, int k, uint32_t* __restrict__ counts, unsigned long long* __restrict__ hit_counter) {
  const qg_seed_item it = items[blockIdx.x];               // d_begin / d_end: reference positions [i_begin, i_end)
  const qg_pair_desc pd = pairs[it.pair];
  const uint32_t n = pd.ylen - (uint32_t) k + 1;
  const unsigned long long* ky = keys + pd.idx_off;
  const uint32_t* vy = vals + pd.idx_off;
  uint32_t* cnt = counts + pd.count_off;
  const int span = (int) pd.ylen - k;
  unsigned long long hits = 0;
  for (int i = it.d_begin + (int) threadIdx.x; i < it.d_end; i += (int) blockDim.x) {
    const unsigned long long code = xcodes[pd.xoff + (uint64_t) i];
    uint32_t lo = 0, hi = n;                               // first entry with key >= code
    while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (ky[mid] < code) lo = mid + 1; else hi = mid; }
    for (uint32_t e = lo; e < n && ky[e] == code; ++e) { atomicAdd (&cnt[i - (int) vy[e] + span], 1u); ++hits; }
  }
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync (0xffffffffu, hits, o);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd (hit_counter, hits);
}

#endif
