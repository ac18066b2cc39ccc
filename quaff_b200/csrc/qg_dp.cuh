// Banded pair-HMM fills for `align` / `train`: Viterbi (+ pointers), Forward, Backward (+ E-step counts).
// Reference semantics: src/qmodel.cpp:1343-1391 (Forward), :1393-1510 (Backward, transCount),
// :1512-1560 (Viterbi fill), :1562-1646 (traceback).  All arithmetic is FP64 with the reference's own
// expression order (compiled with -fmad=false), so Viterbi cells, scores and paths and Forward cells
// are bit-identical to the reference; Backward/counts differ only through CUDA's exp().
//
// Work decomposition: the envelope of a pair is a sorted set of diagonals d = i - j; a maximal run of
// consecutive diagonals is a SEGMENT and is independent of the other runs (the halo diagonals between
// them stay -inf, diagenv.cpp:108-118).  A segment of width W is filled by NW warps (NW = 1 unless
// W > 32 R): virtual lane v owns the R adjacent diagonals dlo + R v .. dlo + R v + R-1, keeps the previous
// row's M/I/D of those diagonals in registers, and at macro-step u fills row j = u - v, so that
//   M(d,j) <- (d,  j-1)   own registers
//   I(d,j) <- (d+1,j-1)   own registers, or the right neighbour's first cell of this macro-step
//   D(d,j) <- (d-1,j)     own registers, or the left neighbour's last cell of the previous macro-step
// i.e. two 2-double shuffles per macro-step per lane; cross-warp edges go through shared memory.
#ifndef QG_DP_CUH
#define QG_DP_CUH
#include "qg_common.cuh"

#ifdef QG_EMU
#define QG_ALIGN(n) __attribute__ ((aligned (n)))
#else
#define QG_ALIGN(n) __align__ (n)
#endif

// per-row parameters of one read, gathered once from the model tables:
//   e[t]  = match[t][yMatchKmer[j-1]].q[yQual[j-1]]      (matchEmitScore, qmodel.h:399-402)
//   ins   = insert[yTok[j-1]].q[yQual[j-1]]              (cachedInsertEmitScore[j], qmodel.cpp:1319-1320)
//   m2m   = m2m[c(j-1)], m2i = m2i[c(j-1)], m2d = m2d[c(j)]   (m2*Score, qmodel.h:392-395)
// row 0 carries m2e[c(yLen)] in .m2m; row yLen+1 is a zero filler so that j+1 is always loadable.
struct QG_ALIGN (16) qg_rowp { double e[4]; double ins, m2m, m2i, m2d; };

#define QG_MAX_NW 32

struct qg_fill_args {
  const qg_segment* segs;
  const uint64_t* xpacked;           // 2-bit packed tokens of the x set
  const uint64_t* xpoff;             // word offset per x sequence
  const qg_rowp* rp;                 // row parameters, indexed seg.rp_off + j
  const double* lse;                 // log-sum-exp table (Forward/Backward)
  double i2i, i2m, d2d, d2m;
  int local;
  uint32_t* trace;                   // Viterbi: 4 bit per cell, one u32 per (macro-step, virtual lane)
  double* store;                     // Forward matrix, [j][state][slot] per segment (MODE 2 / Backward)
  double* endvals;                   // per segment slot: M(i,yLen)+m2e (Forward) ; per segment {score, i} (Viterbi)
  double* rowacc;                    // Backward: per segment per row 8 doubles
  const double* pair_z;              // Backward: Forward result per pair
  double* seg_scal;                  // Backward: per segment 12 doubles (d2m,i2m,i2i,d2d,m2e,s2m[4],pad)
};

// ---- table log-sum-exp, the reference's arithmetic verbatim (logsumexp.cpp:34-59, 84-103) ---------
__device__ __forceinline__ double qg_lse_unary (const double* __restrict__ tab, double x) {
  if (x >= 10.0 || isnan (x) || isinf (x)) return 0;
  const int n = (int) (x / .0001);
  const double dx = x - (n * .0001);
  const double f0 = tab[n], f1 = tab[n + 1];
  const double df = f1 - f0;
  return f0 + df * (dx / .0001);
}
__device__ __forceinline__ double qg_lse (const double* __restrict__ tab, double a, double b) {
  double mx, diff;
  if (a == b) { mx = a; diff = 0; }
  else if (a < b) { mx = b; diff = b - a; }
  else { mx = a; diff = a - b; }
  return mx + qg_lse_unary (tab, diff);
}

// ---- 2-bit token access ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t qg_xword (const uint64_t* __restrict__ xw, int nxw, int wi) {
  return (wi >= 0 && wi < nxw) ? xw[wi] : 0ull;
}
// 32 consecutive tokens starting at (possibly negative / past-the-end) position p
__device__ __forceinline__ uint64_t qg_fetch32 (const uint64_t* __restrict__ xw, int nxw, int p) {
  const int wi = p >> 5;                                   // arithmetic shift = floor
  const int sh = (p & 31) * 2;
  const uint64_t lo = qg_xword (xw, nxw, wi);
  if (sh == 0) return lo;
  const uint64_t hi = qg_xword (xw, nxw, wi + 1);
  return (lo >> sh) | (hi << (64 - sh));
}
__device__ __forceinline__ int qg_tok (const uint64_t* __restrict__ xw, int nxw, int p) {
  return (int) ((qg_xword (xw, nxw, p >> 5) >> ((p & 31) * 2)) & 3);
}
__device__ __forceinline__ double qg_sel4 (const double* e, int t) {
  const double lo = (t & 1) ? e[1] : e[0];
  const double hi = (t & 1) ? e[3] : e[2];
  return (t & 2) ? hi : lo;
}

// ---- row parameter gather -------------------------------------------------------------------------------
// One CTA per distinct read of the call.  FastSeq::kmers (fastseq.cpp:85-99): the context k-mer ENDING at
// each position, left-padded with the read's most frequent token (first maximum on ties).
struct qg_rp_job { uint32_t yseq, ylen; uint64_t yoff, rp_off; };

__global__ void qg_rowparams_kernel (const qg_rp_job* __restrict__ jobs, const uint8_t* __restrict__ ytok, const uint8_t* __restrict__ yqual,
                                     const double* __restrict__ match, const double* __restrict__ insert, const double* __restrict__ gap,
                                     int match_k, int gap_k, qg_rowp* __restrict__ rp, double2* __restrict__ rps) {
  __shared__ unsigned s_count[4];
  __shared__ int s_mf;
  const qg_rp_job jb = jobs[blockIdx.x];
  const uint8_t* tok = ytok + jb.yoff;
  const uint8_t* ql = yqual ? yqual + jb.yoff : nullptr;
  const int ylen = (int) jb.ylen;
  const uint64_t nK = 1ull << (2 * match_k), nG = 1ull << (2 * gap_k);
  if (threadIdx.x < 4) s_count[threadIdx.x] = 0;
  __syncthreads ();
  unsigned c[4] = {0, 0, 0, 0};
  for (int p = threadIdx.x; p < ylen; p += blockDim.x) ++c[tok[p] & 3];
  for (int t = 0; t < 4; ++t) if (c[t]) atomicAdd (&s_count[t], c[t]);
  __syncthreads ();
  if (threadIdx.x == 0) { int b = 0; for (int t = 1; t < 4; ++t) if (s_count[t] > s_count[b]) b = t; s_mf = b; }
  __syncthreads ();
  const int mf = s_mf;
  const double* m2m = gap, *m2i = gap + nG, *m2d = gap + 2 * nG, *m2e = gap + 3 * nG;
  qg_rowp* out = rp + jb.rp_off;
  for (int j = threadIdx.x; j <= ylen + 1; j += blockDim.x) {
    qg_rowp r;
    if (j >= 1 && j <= ylen) {
      uint64_t mk = 0, gcur = 0, gprev = 0;                 // k-mers ending at base j (1-based) and j-1
      for (int t = match_k - 1; t >= 0; --t) { const int q = j - 1 - t; mk = mk * 4 + (q >= 0 ? tok[q] : mf); }
      for (int t = gap_k - 1; t >= 0; --t) { const int q = j - 1 - t; gcur = gcur * 4 + (q >= 0 ? tok[q] : mf); }
      if (j >= 2) for (int t = gap_k - 1; t >= 0; --t) { const int q = j - 2 - t; gprev = gprev * 4 + (q >= 0 ? tok[q] : mf); }
      const int q = ql ? ql[j - 1] : QG_NQUAL;
      for (int t = 0; t < 4; ++t) r.e[t] = match[((uint64_t) t * nK + mk) * QG_NQ1 + q];
      r.ins = insert[(uint64_t) tok[j - 1] * QG_NQ1 + q];
      r.m2m = m2m[gprev]; r.m2i = m2i[gprev]; r.m2d = m2d[gcur];
    } else {
      for (int t = 0; t < 4; ++t) r.e[t] = 0;
      r.ins = 0; r.m2m = 0; r.m2i = 0; r.m2d = 0;
      if (j == 0) {
        uint64_t gl = 0;                                    // c(yLen); c(0) = 0 for an empty read
        if (ylen >= 1) for (int t = gap_k - 1; t >= 0; --t) { const int q = ylen - 1 - t; gl = gl * 4 + (q >= 0 ? tok[q] : mf); }
        r.m2m = m2e[gl];
      }
    }
    out[j] = r;
    if (rps) {                                              // structure-of-arrays copy for qg_vit_kernel
      double2* o2 = rps + 4 * jb.rp_off + j;
      const uint64_t rows = (uint64_t) ylen + 2;
      o2[0] = make_double2 (r.e[0], r.e[1]); o2[rows] = make_double2 (r.e[2], r.e[3]);
      o2[2 * rows] = make_double2 (r.ins, r.m2m); o2[3 * rows] = make_double2 (r.m2i, r.m2d);
    }
  }
}

// ---- Viterbi / Forward fill ----------------------------------------------------------------------------------
// MODE 0: Viterbi with pointers; MODE 1: Forward; MODE 2: Forward, storing the matrix for Backward.
// Pointer nibble of a cell: bits 0-1 = source of Match (0 M, 1 I, 2 D, 3 Start), bit 2 = source of Insert
// (0 M, 1 I), bit 3 = source of Delete (0 M, 1 D), chosen with the traceback's candidate order and strict '>'
// (qmodel.cpp:1590-1594, 1604-1605, 1614-1615) on the same rounded sums the traceback compares.
template<int R, int MODE, bool MULTI>
__global__ void __launch_bounds__ (MULTI ? 1024 : 32)
qg_fill_kernel (const qg_fill_args a) {
  __shared__ double sA[QG_MAX_NW][2], sB[QG_MAX_NW][2];
  __shared__ double s_best[QG_MAX_NW];
  __shared__ int s_besti[QG_MAX_NW];
  const qg_segment sg = a.segs[blockIdx.x];
  const int NW = (int) blockDim.x >> 5;
  const int vl = threadIdx.x, lane = vl & 31, wid = vl >> 5;
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, width = (int) sg.width;
  const int SW = 32 * NW * R;                               // slots of this segment
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowp* rp = a.rp + sg.rp_off;
  const double i2i = a.i2i, i2m = a.i2m, d2d = a.d2d, d2m = a.d2m;
  const bool local = a.local != 0;
  const double m2e = rp[0].m2m;
  const int s0 = R * vl;                                    // my first slot
  const int d0 = sg.dlo + s0;                               // my first diagonal

  double M[R], I[R], D[R];
#pragma unroll
  for (int c = 0; c < R; ++c) { M[c] = QG_NEG_INF; I[c] = QG_NEG_INF; D[c] = QG_NEG_INF; }
  double leftM = QG_NEG_INF, leftD = QG_NEG_INF;
  double bestEnd = QG_NEG_INF; int bestI = 0;
  uint64_t win = 0; int pw = 0; bool have_win = false;

  const int total = ylen + 32 * NW - 1;
  // row parameters are fetched one macro-step ahead: the 64 B load of row j+1 is in flight while row j is computed
  qg_rowp Pnext = rp[(1 - vl) < 0 ? 0 : (1 - vl)];
  for (int u = 1; u <= total; ++u) {
    const int j = u - vl;
    const bool active = (j >= 1) && (j <= ylen);
    const qg_rowp P = Pnext;
    { const int jn = j + 1; Pnext = rp[jn < 0 ? 0 : (jn > ylen + 1 ? ylen + 1 : jn)]; }
    const int p0 = d0 + j - 1;                              // x index (0-based) of cell 0: i - 1
    if (active && (!have_win || p0 < pw || p0 + R > pw + 32)) { win = qg_fetch32 (xw, nxw, p0); pw = p0; have_win = true; }
    const uint64_t wsh = win >> (2 * ((p0 - pw) & 31));
    const bool startRow = (j == 1);
    const bool endRow = (j == ylen);
    unsigned tword = 0;
    double rM = QG_NEG_INF, rI = QG_NEG_INF;                // right neighbour's first cell at row j-1

#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1 && R > 1) {
        // ---- exchange 1: first cells (row j-1 of the right neighbour) travel one lane to the left
        const double m0 = M[0], i0 = I[0];                  // already holds this macro-step's new values
        rM = __shfl_down_sync (QG_FULL_MASK, m0, 1);
        rI = __shfl_down_sync (QG_FULL_MASK, i0, 1);
        if (MULTI) {
          if (lane == 0) { sA[wid][0] = m0; sA[wid][1] = i0; }
          __syncthreads ();
          if (lane == 31) { if (wid + 1 < NW) { rM = sA[wid + 1][0]; rI = sA[wid + 1][1]; } else { rM = QG_NEG_INF; rI = QG_NEG_INF; } }
        } else {
          if (lane == 31) { rM = QG_NEG_INF; rI = QG_NEG_INF; }
        }
      }
      const int i = d0 + c + j;
      const bool ok = active && (s0 + c < width) && (i >= 1) && (i <= xlen);
      const int tk = (int) ((wsh >> (2 * c)) & 3);
      const double E = qg_sel4 (P.e, tk);
      // sources
      const double mM = M[c], mI = I[c], mD = D[c];                         // (i-1, j-1)
      const double iM = (c + 1 < R) ? M[(c + 1) % R] : rM;                  // (i,   j-1)
      const double iI = (c + 1 < R) ? I[(c + 1) % R] : rI;
      const double dM = (c > 0) ? M[(c + R - 1) % R] : leftM;               // (i-1, j) -- already this row's values
      const double dD = (c > 0) ? D[(c + R - 1) % R] : leftD;
      double nM, nI, nD;
      unsigned ptr = 0;
      if (MODE == 0) {
        const double cM = (mM + P.m2m) + E, cI = (mI + i2m) + E, cD = (mD + d2m) + E;
        nM = cM;
        if (cI > nM) { nM = cI; ptr = 1; }
        if (cD > nM) { nM = cD; ptr = 2; }
        if (startRow && (i == 1 || local) && E > nM) { nM = E; ptr = 3; }
        const double aM = (iM + P.m2i) + P.ins, aI = (iI + i2i) + P.ins;
        nI = aM;
        if (aI > nI) { nI = aI; ptr |= 4; }
        const double bM = dM + P.m2d, bD = dD + d2d;
        nD = bM;
        if (bD > nD) { nD = bD; ptr |= 8; }
      } else {
        double mat = qg_lse (a.lse, qg_lse (a.lse, mM + P.m2m, mD + d2m), mI + i2m);
        if (startRow && (i == 1 || local)) mat = qg_lse (a.lse, mat, 0.0);
        nM = mat + E;
        nI = P.ins + qg_lse (a.lse, iI + i2i, iM + P.m2i);
        nD = qg_lse (a.lse, dD + d2d, dM + P.m2d);
      }
      if (!ok) { nM = QG_NEG_INF; nI = QG_NEG_INF; nD = QG_NEG_INF; ptr = 0; }
      M[c] = nM; I[c] = nI; D[c] = nD;
      tword |= ptr << (4 * c);
      if (endRow) {
        const bool isEnd = ok && (i == xlen || local);
        if (MODE == 0) {
          if (isEnd) { const double e = nM + m2e; if (e >= bestEnd) { bestEnd = e; bestI = i; } }   // ascending i: ties -> largest i
        } else {
          a.endvals[sg.aux_off + s0 + c] = isEnd ? nM + m2e : QG_NEG_INF;
        }
      }
    }
    // ---- exchange 2: last cells (row j) travel one lane to the right, for the next macro-step
    {
      const double mL = M[R - 1], dL = D[R - 1];
      leftM = __shfl_up_sync (QG_FULL_MASK, mL, 1);
      leftD = __shfl_up_sync (QG_FULL_MASK, dL, 1);
      if (MULTI) {
        if (lane == 31) { sB[wid][0] = mL; sB[wid][1] = dL; }
        __syncthreads ();
        if (lane == 0) { if (wid > 0) { leftM = sB[wid - 1][0]; leftD = sB[wid - 1][1]; } else { leftM = QG_NEG_INF; leftD = QG_NEG_INF; } }
      } else {
        if (lane == 0) { leftM = QG_NEG_INF; leftD = QG_NEG_INF; }
      }
    }
    if (MODE == 0) a.trace[sg.trace_off + (uint64_t) u * (32 * NW) + vl] = tword;
    if (MODE == 2 && active) {
      // skewed layout [macro-step][virtual lane][state][c]: one contiguous 24 R byte record per lane, a warp writes one
      // contiguous block per macro-step, and Backward (mirrored) reads exactly one such block per macro-step
      double* st = a.store + sg.store_off + ((uint64_t) u * (32 * NW) + vl) * (3 * R);
#pragma unroll
      for (int c = 0; c < R; ++c) { st[c] = M[c]; st[R + c] = I[c]; st[2 * R + c] = D[c]; }
    }
  }

  if (MODE == 0) {
    // segment result: max over end cells, ties -> largest i (qmodel.cpp:1568-1574)
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync (QG_FULL_MASK, bestEnd, o);
      const int oi = __shfl_down_sync (QG_FULL_MASK, bestI, o);
      if (ob > bestEnd || (ob == bestEnd && oi > bestI)) { bestEnd = ob; bestI = oi; }
    }
    if (MULTI) {
      if (lane == 0) { s_best[wid] = bestEnd; s_besti[wid] = bestI; }
      __syncthreads ();
      if (vl == 0) for (int w = 1; w < NW; ++w) if (s_best[w] > bestEnd || (s_best[w] == bestEnd && s_besti[w] > bestI)) { bestEnd = s_best[w]; bestI = s_besti[w]; }
    }
    if (vl == 0) { a.endvals[2 * sg.aux_off] = bestEnd; a.endvals[2 * sg.aux_off + 1] = (double) bestI; }
  }
}


// ---- per-pair reductions and traceback ---------------------------------------------------------------------------
struct qg_pair_dp {
  uint32_t seg_begin, seg_end;       // this pair's segments, ascending dlo
  uint32_t xlen, ylen;
  uint64_t path_off;                 // start of this pair's path scratch region
  uint32_t path_cap;
  uint32_t want_path;
};

// Forward result: end = lse(end, mat(i,yLen) + m2e) folded over ascending i (qmodel.cpp:1381-1383)
__global__ void qg_forward_finalize_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                            const double* __restrict__ endvals, const double* __restrict__ lse, double* __restrict__ result) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  double end = QG_NEG_INF;
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_segment sg = segs[s];
    for (uint32_t t = 0; t < sg.width; ++t) end = qg_lse (lse, end, endvals[sg.aux_off + t]);
  }
  result[p] = end;
}

// Viterbi result per pair: the end cell (max, ties -> largest i, qmodel.cpp:1565-1575) over the pair's segments
__global__ void qg_pair_score_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const double* __restrict__ seg_end,
                                      double* __restrict__ score, uint32_t* __restrict__ x_end) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  double best = QG_NEG_INF; int bi = 0;
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const double sc = seg_end[2 * (uint64_t) s]; const int si = (int) seg_end[2 * (uint64_t) s + 1];
    if (sc > best || (sc == best && si > bi)) { best = sc; bi = si; }
  }
  score[p] = best;
  x_end[p] = (best > QG_NEG_INF) ? (uint32_t) bi : 0u;
}

// Follow the pointers from the end cell.  One thread per pair; ops are written back-to-front into the pair's
// scratch region (state sequence of qmodel.cpp:1579-1622).
__global__ void qg_traceback_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                     const uint32_t* __restrict__ trace, const double* __restrict__ score, const uint32_t* __restrict__ x_end,
                                     uint32_t* __restrict__ x_start, uint8_t* __restrict__ path_scratch, uint32_t* __restrict__ path_len,
                                     uint32_t* __restrict__ err_flag) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  x_start[p] = 0; path_len[p] = 0;
  if (!pd.want_path || !(score[p] > QG_NEG_INF)) return;
  int i = (int) x_end[p], j = (int) pd.ylen;
  int state = 1;                                           // 0 Start, 1 Match, 2 Insert, 3 Delete
  uint32_t n = 0;
  uint32_t cs = pd.seg_begin;
  uint8_t* buf = path_scratch + pd.path_off;
  while (state != 0) {
    const int d = i - j;
    qg_segment sg = segs[cs];
    if (d < sg.dlo || d >= sg.dlo + (int) sg.width) {
      bool found = false;
      for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
        const qg_segment t = segs[s];
        if (d >= t.dlo && d < t.dlo + (int) t.width) { cs = s; sg = t; found = true; break; }
      }
      if (!found) { *err_flag = 1; break; }
    }
    const int slot = d - sg.dlo, R = (int) sg.R;
    const int vl = slot / R, c = slot - vl * R;
    const uint32_t word = sg.half ? (uint32_t) ((const uint16_t*) (trace + sg.trace_off))[(uint64_t) (j + vl) * 32 + vl]
                                  : trace[sg.trace_off + (uint64_t) (j + vl) * (32 * sg.nwarps) + vl];
    const uint32_t nib = (word >> (4 * c)) & 15u;
    if (n >= pd.path_cap) { *err_flag = 2; break; }
    if (state == 1) {
      buf[pd.path_cap - 1 - n] = QG_OP_MATCH; ++n; --i; --j;
      const uint32_t src = nib & 3u;
      state = (src == 0) ? 1 : (src == 1) ? 2 : (src == 2) ? 3 : 0;
    } else if (state == 2) {
      buf[pd.path_cap - 1 - n] = QG_OP_INSERT; ++n; --j;
      state = (nib & 4u) ? 2 : 1;
    } else {
      buf[pd.path_cap - 1 - n] = QG_OP_DELETE; ++n; --i;
      state = (nib & 8u) ? 3 : 1;
    }
    if (i < 0 || j < 0) { *err_flag = 3; break; }
  }
  x_start[p] = (uint32_t) (i + 1);
  path_len[p] = n;
}

// Warp-cooperative traceback for single-warp segments: the pointer words of 32 consecutive macro-steps (32 x 128 B,
// each row one coalesced load) are staged in shared memory, then lane 0 walks inside the tile.  A step never increases
// the macro-step index u = j + lane-of-slot and lowers it by at most one, so a tile serves >= 32 steps: one DRAM
// round trip per 32+ steps instead of one per step.
#define QG_TB_WARPS 4
__global__ void __launch_bounds__ (32 * QG_TB_WARPS)
qg_traceback_warp_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                          const uint32_t* __restrict__ trace, const double* __restrict__ score, const uint32_t* __restrict__ x_end,
                          uint32_t* __restrict__ x_start, uint8_t* __restrict__ path_scratch, uint32_t* __restrict__ path_len,
                          uint32_t* __restrict__ err_flag) {
  __shared__ uint32_t s_tile[QG_TB_WARPS][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t p = blockIdx.x * QG_TB_WARPS + w;
  const bool have = p < npairs;
  const qg_pair_dp pd = pairs[have ? p : 0];
  bool go = have && pd.want_path && (score[have ? p : 0] > QG_NEG_INF);
  if (have && lane == 0) { x_start[p] = 0; path_len[p] = 0; }
  int i = go ? (int) x_end[p] : 0, j = (int) pd.ylen, state = go ? 1 : 0;
  uint32_t n = 0, cs = pd.seg_begin;
  uint8_t* buf = path_scratch + pd.path_off;
  int fail = 0;
  while (__shfl_sync (QG_FULL_MASK, (int) (state != 0 && !fail), 0)) {
    // lane 0 walks and decides the segment and the tile; everybody follows
    qg_segment sg = segs[cs];
    if (lane == 0 && state != 0 && !fail) {
      const int d = i - j;
      if (d < sg.dlo || d >= sg.dlo + (int) sg.width) {
        bool found = false;
        for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
          const qg_segment t = segs[s];
          if (d >= t.dlo && d < t.dlo + (int) t.width) { cs = s; found = true; break; }
        }
        if (!found) fail = 1;
      }
    }
    cs = __shfl_sync (QG_FULL_MASK, cs, 0);
    sg = segs[cs];
    const int R = (int) sg.R;
    int u_top = 0;
    if (lane == 0 && state != 0 && !fail) u_top = j + (i - j - sg.dlo) / R;
    u_top = __shfl_sync (QG_FULL_MASK, u_top, 0);
    const bool multi = sg.nwarps != 1;                      // more than one warp, or a narrow segment (nwarps == 0): direct loads
    const bool narrow = sg.nwarps == 0;                     // one u32 per row, nibble = slot (qg_vit_narrow_kernel)
    if (!multi) {
      // 32 loads in flight per lane either way: the word size is decided outside the loop
      if (sg.half) {
        const uint16_t* t16 = (const uint16_t*) (trace + sg.trace_off) + lane;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) { const int u = u_top - r; s_tile[w][r][lane] = (u >= 0) ? (uint32_t) t16[(uint64_t) u * 32] : 0u; }
      } else {
        const uint32_t* t32 = trace + sg.trace_off + lane;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) { const int u = u_top - r; s_tile[w][r][lane] = (u >= 0) ? t32[(uint64_t) u * 32] : 0u; }
      }
    }
    __syncwarp ();
    if (lane == 0 && state != 0 && !fail) {
      // (vl, c) = lane and cell of the current diagonal, kept incrementally: Match stays on the diagonal, Insert moves to
      // d + 1, Delete to d - 1 (no division in the dependent chain)
      int slot = (i - j) - sg.dlo;
      int vl = slot / R, c = slot - vl * R;
      const int width = (int) sg.width;
      while (state != 0) {
        if (slot < 0 || slot >= width) { fail = 1; break; }                 // cannot happen: paths do not cross the halo
        const int u = j + vl;
        uint32_t word;
        if (narrow) word = trace[sg.trace_off + (uint64_t) j];
        else if (multi) word = trace[sg.trace_off + (uint64_t) u * (32 * sg.nwarps) + vl];
        else {
          if (u < u_top - 31) break;                                       // next tile
          word = s_tile[w][u_top - u][vl];
        }
        const uint32_t nib = (word >> (4 * c)) & 15u;
        if (n >= pd.path_cap) { fail = 2; break; }
        if (state == 1) {
          buf[pd.path_cap - 1 - n] = QG_OP_MATCH; ++n; --i; --j;
          const uint32_t src = nib & 3u;
          state = (src == 0) ? 1 : (src == 1) ? 2 : (src == 2) ? 3 : 0;
        } else if (state == 2) {
          buf[pd.path_cap - 1 - n] = QG_OP_INSERT; ++n; --j;
          ++slot; if (++c == R) { c = 0; ++vl; }
          state = (nib & 4u) ? 2 : 1;
        } else {
          buf[pd.path_cap - 1 - n] = QG_OP_DELETE; ++n; --i;
          --slot; if (--c < 0) { c = R - 1; --vl; }
          state = (nib & 8u) ? 3 : 1;
        }
        if (i < 0 || j < 0) { fail = 3; break; }
      }
    }
    __syncwarp ();
  }
  if (have && lane == 0) {
    if (fail) *err_flag = (uint32_t) fail;
    if (go) { x_start[p] = (uint32_t) (i + 1); path_len[p] = n; }
  }
}

// gather the back-to-front scratch paths into one contiguous 5'->3' array
__global__ void qg_path_gather_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const uint8_t* __restrict__ path_scratch,
                                       const uint32_t* __restrict__ path_len, const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out) {
  const uint32_t p = blockIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  const uint32_t n = path_len[p];
  const uint8_t* src = path_scratch + pd.path_off + (pd.path_cap - n);
  uint8_t* dst = out + out_off[p];
  for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) dst[t] = src[t];
}

#endif
