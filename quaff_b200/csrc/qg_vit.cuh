// Viterbi fill for single-warp segments (runs of up to 256 diagonals: every segment of the banded `align` workload).
// Same lane / diagonal mapping, FP64 expressions, pointer rules and trace layout as qg_fill_kernel<R,0,false>
// (qg_dp.cuh), i.e. the same bits; what differs is the instruction count of the steady state:
//   * the macro-steps are split into a general phase (some lane is before row 1, on the start row, on the end row or
//     past it, or some real cell falls outside 1 <= i <= xLen) and a fast phase in which none of that can happen, so the
//     start / end / out-of-matrix logic is not in the inner loop (qmodel.cpp:1526-1551 has no such tests either);
//   * in the fast phase only the first padding diagonal next to the run has to be held at -inf, and only its Delete
//     state can become finite (its Match and Insert sources are padding cells);
//   * row parameters come from a structure-of-arrays copy ([4][rows] double2 per read): the 32 lanes of a warp read
//     32 consecutive rows, so each of the four loads of a macro-step is one contiguous 512 B request instead of 32
//     16 B pieces at a 64 B stride.
#ifndef QG_VIT_CUH
#define QG_VIT_CUH
#include "qg_dp.cuh"

struct qg_vit_args {
  const qg_segment* segs;
  const uint64_t* xpacked;
  const uint64_t* xpoff;
  const double2* rps;                // row parameters, SoA: read block at 4 * rp_off, field pair q at + q * (yLen + 2)
  double i2i, i2m, d2d, d2m;
  int local;
  uint32_t* trace;
  double* endvals;                   // per segment {score, i}
};

template<int R>
struct qg_vit_lane {
  double M[R], I[R], D[R];
  double leftM, leftD, bestEnd;
  int bestI;
  uint64_t win; int pw; bool have_win;
  double2 Pn[4];                     // row parameters of the next macro-step (prefetched)
  // constants
  const double2* rq; int rows;
  const uint64_t* xw; int nxw;
  double i2i, i2m, d2d, d2m, m2e;
  int vl, s0, d0, xlen, ylen, width, padc;
  bool local;

  template<bool GEN>
  __device__ __forceinline__ unsigned step (int u) {
    const int j = u - vl;
    const bool active = GEN ? ((j >= 1) && (j <= ylen)) : true;
    const double2 q0 = Pn[0], q1 = Pn[1], q2 = Pn[2], q3 = Pn[3];
    const double e[4] = { q0.x, q0.y, q1.x, q1.y };
    const double ins = q2.x, m2m = q2.y, m2i = q3.x, m2d = q3.y;
    {
      int jn = j + 1;
      if (GEN) jn = jn < 0 ? 0 : (jn > ylen + 1 ? ylen + 1 : jn);
#pragma unroll
      for (int q = 0; q < 4; ++q) Pn[q] = rq[q * rows + jn];
    }
    const int p0 = d0 + j - 1;                              // x index (0-based) of cell 0: i - 1
    if (active && (!have_win || p0 < pw || p0 + R > pw + 32)) { win = qg_fetch32 (xw, nxw, p0); pw = p0; have_win = true; }
    const uint64_t wsh = win >> (2 * ((p0 - pw) & 31));
    const bool startRow = GEN && (j == 1);
    const bool endRow = GEN && (j == ylen);
    unsigned tword = 0;
    double rM = QG_NEG_INF, rI = QG_NEG_INF;                // right neighbour's first cell at row j-1
#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1) {
        // exchange 1: first cells (row j-1 of the right neighbour) travel one lane to the left
        rM = __shfl_down_sync (QG_FULL_MASK, M[0], 1);
        rI = __shfl_down_sync (QG_FULL_MASK, I[0], 1);
        if (vl == 31) { rM = QG_NEG_INF; rI = QG_NEG_INF; }
      }
      const int i = d0 + c + j;
      const int tk = (int) ((wsh >> (2 * c)) & 3);
      const double E = qg_sel4 (e, tk);
      const double mM = M[c], mI = I[c], mD = D[c];                         // (i-1, j-1)
      const double iM = (c + 1 < R) ? M[(c + 1) % R] : rM;                  // (i,   j-1)
      const double iI = (c + 1 < R) ? I[(c + 1) % R] : rI;
      const double dM = (c > 0) ? M[(c + R - 1) % R] : leftM;               // (i-1, j) -- already this row's values
      const double dD = (c > 0) ? D[(c + R - 1) % R] : leftD;
      unsigned ptr = 0;
      const double cM = (mM + m2m) + E, cI = (mI + i2m) + E, cD = (mD + d2m) + E;
      double nM = cM;
      if (cI > nM) { nM = cI; ptr = 1; }
      if (cD > nM) { nM = cD; ptr = 2; }
      if (GEN) { if (startRow && (i == 1 || local) && E > nM) { nM = E; ptr = 3; } }
      const double aM = (iM + m2i) + ins, aI = (iI + i2i) + ins;
      double nI = aM;
      if (aI > nI) { nI = aI; ptr |= 4; }
      const double bM = dM + m2d, bD = dD + d2d;
      double nD = bM;
      if (bD > nD) { nD = bD; ptr |= 8; }
      if (GEN) {
        const bool ok = active && (s0 + c < width) && (i >= 1) && (i <= xlen);
        if (!ok) { nM = QG_NEG_INF; nI = QG_NEG_INF; nD = QG_NEG_INF; ptr = 0; }
        M[c] = nM; I[c] = nI; D[c] = nD;
        if (endRow) {
          const bool isEnd = ok && (i == xlen || local);
          if (isEnd) { const double en = nM + m2e; if (en >= bestEnd) { bestEnd = en; bestI = i; } }   // ascending i: ties -> largest i
        }
      } else {
        // every real cell is inside the matrix; of the padding diagonals only the first one has a finite source
        if (c == padc) { nD = QG_NEG_INF; ptr &= 7u; }
        M[c] = nM; I[c] = nI; D[c] = nD;
      }
      tword |= ptr << (4 * c);
    }
    // exchange 2: last cells (row j) travel one lane to the right, for the next macro-step
    leftM = __shfl_up_sync (QG_FULL_MASK, M[R - 1], 1);
    leftD = __shfl_up_sync (QG_FULL_MASK, D[R - 1], 1);
    if (vl == 0) { leftM = QG_NEG_INF; leftD = QG_NEG_INF; }
    return tword;
  }
};

template<int R>
__global__ void __launch_bounds__ (32)
qg_vit_kernel (const qg_vit_args a) {
  const qg_segment sg = a.segs[blockIdx.x];
  qg_vit_lane<R> L;
  L.vl = threadIdx.x;
  L.xlen = (int) sg.xlen; L.ylen = (int) sg.ylen; L.width = (int) sg.width;
  L.xw = a.xpacked + a.xpoff[sg.xseq];
  L.nxw = (L.xlen + 31) >> 5;
  L.rows = L.ylen + 2;
  L.rq = a.rps + 4 * sg.rp_off;
  L.i2i = a.i2i; L.i2m = a.i2m; L.d2d = a.d2d; L.d2m = a.d2m;
  L.local = a.local != 0;
  L.m2e = L.rq[2 * L.rows].y;                               // row 0 carries m2e[c(yLen)] in .m2m
  L.s0 = R * L.vl; L.d0 = sg.dlo + L.s0;
  L.padc = (L.width >= L.s0 && L.width < L.s0 + R) ? L.width - L.s0 : -1;
#pragma unroll
  for (int c = 0; c < R; ++c) { L.M[c] = QG_NEG_INF; L.I[c] = QG_NEG_INF; L.D[c] = QG_NEG_INF; }
  L.leftM = QG_NEG_INF; L.leftD = QG_NEG_INF; L.bestEnd = QG_NEG_INF; L.bestI = 0;
  L.win = 0; L.pw = 0; L.have_win = false;
  {
    const int j1 = (1 - L.vl) < 0 ? 0 : (1 - L.vl);
#pragma unroll
    for (int q = 0; q < 4; ++q) L.Pn[q] = L.rq[q * L.rows + j1];
  }
  const int ylen = L.ylen, xlen = L.xlen;
  const int total = ylen + 31;
  // fast phase: all lanes on rows 2..yLen-1 (u in [33, yLen-1]) and every real cell inside 1 <= i <= xLen:
  // i = dlo + u + (R-1) v + c is smallest for slot 0 and largest for the last real slot
  const int v_last = (L.width - 1) / R, c_last = (L.width - 1) - v_last * R;
  int f_lo = 33, f_hi = ylen - 1;
  if (1 - sg.dlo > f_lo) f_lo = 1 - sg.dlo;
  { const long long h = (long long) xlen - sg.dlo - (long long) (R - 1) * v_last - c_last; if (h < f_hi) f_hi = (int) h; }
  if (f_hi < f_lo) { f_lo = total + 1; f_hi = total; }      // no fast phase
  int u = 1;
  if (R <= 4) {                                             // 16-bit pointer words (segment flag `half`, set by the plan for exactly these kernels)
    uint16_t* tr = (uint16_t*) (a.trace + sg.trace_off) + L.vl;
    for (; u < f_lo && u <= total; ++u) tr[(uint64_t) u * 32] = (uint16_t) L.template step<true> (u);
    for (; u <= f_hi; ++u) tr[(uint64_t) u * 32] = (uint16_t) L.template step<false> (u);
    for (; u <= total; ++u) tr[(uint64_t) u * 32] = (uint16_t) L.template step<true> (u);
  } else {
    uint32_t* tr = a.trace + sg.trace_off + L.vl;
    for (; u < f_lo && u <= total; ++u) tr[(uint64_t) u * 32] = L.template step<true> (u);
    for (; u <= f_hi; ++u) tr[(uint64_t) u * 32] = L.template step<false> (u);
    for (; u <= total; ++u) tr[(uint64_t) u * 32] = L.template step<true> (u);
  }

  // segment result: max over end cells, ties -> largest i (qmodel.cpp:1568-1574)
  double bestEnd = L.bestEnd; int bestI = L.bestI;
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_down_sync (QG_FULL_MASK, bestEnd, o);
    const int oi = __shfl_down_sync (QG_FULL_MASK, bestI, o);
    if (ob > bestEnd || (ob == bestEnd && oi > bestI)) { bestEnd = ob; bestI = oi; }
  }
  if (L.vl == 0) { a.endvals[2 * sg.aux_off] = bestEnd; a.endvals[2 * sg.aux_off + 1] = (double) bestI; }
}


// ---- narrow segments: runs of 1..4 diagonals (the always-present diagonal 0 of every pair, diagenv.cpp:52-54) ----
// A warp per run would spend yLen + 31 macro-steps on a handful of cells per row.  Here one THREAD fills one run, row by
// row, cells in ascending i (so Delete sees this row's left neighbour and Insert the previous row's right neighbour);
// same expressions and pointer rules.  Pointer layout of these segments (nwarps == 0): one u32 per row j at
// trace_off + j, nibble c = slot; four rows are written as one 16 B store.
template<int W>
__global__ void __launch_bounds__ (64)
qg_vit_narrow_kernel (const qg_vit_args a, uint32_t nseg) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  const qg_segment sg = a.segs[t];
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, dlo = sg.dlo;
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const int rows = ylen + 2;
  const double2* rq = a.rps + 4 * sg.rp_off;
  const double i2i = a.i2i, i2m = a.i2m, d2d = a.d2d, d2m = a.d2m;
  const bool local = a.local != 0;
  const double m2e = rq[2 * rows].y;
  double M[W], I[W], D[W];
#pragma unroll
  for (int c = 0; c < W; ++c) { M[c] = QG_NEG_INF; I[c] = QG_NEG_INF; D[c] = QG_NEG_INF; }
  double bestEnd = QG_NEG_INF; int bestI = 0;
  uint64_t win = 0; int pw = 0; bool have_win = false;
  uint32_t* tr = a.trace + sg.trace_off;
  uint32_t acc[4] = {0u, 0u, 0u, 0u};
  // a single thread is one long dependent chain and there are few of them: row parameters are fetched four rows ahead
  double2 ring[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int jr = (1 + k > ylen + 1) ? ylen + 1 : 1 + k;
#pragma unroll
    for (int q = 0; q < 4; ++q) ring[k][q] = rq[q * rows + jr];
  }
  for (int j0 = 1; j0 <= ylen; j0 += 4) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = j0 + k;
    if (j > ylen) break;
    const double2 q0 = ring[k][0], q1 = ring[k][1], q2 = ring[k][2], q3 = ring[k][3];
    const double e[4] = { q0.x, q0.y, q1.x, q1.y };
    const double ins = q2.x, m2m = q2.y, m2i = q3.x, m2d = q3.y;
    {
      const int jr = (j + 4 > ylen + 1) ? ylen + 1 : j + 4;                 // row yLen+1 exists (zero filler)
#pragma unroll
      for (int q = 0; q < 4; ++q) ring[k][q] = rq[q * rows + jr];
    }
    const int p0 = dlo + j - 1;
    if (!have_win || p0 < pw || p0 + W > pw + 32) { win = qg_fetch32 (xw, nxw, p0); pw = p0; have_win = true; }
    const uint64_t wsh = win >> (2 * ((p0 - pw) & 31));
    const bool startRow = (j == 1), endRow = (j == ylen);
    unsigned tword = 0;
    double leftM = QG_NEG_INF, leftD = QG_NEG_INF;                          // halo diagonal dlo - 1
#pragma unroll
    for (int c = 0; c < W; ++c) {
      const int i = dlo + c + j;
      const bool ok = (i >= 1) && (i <= xlen);
      const double E = qg_sel4 (e, (int) ((wsh >> (2 * c)) & 3));
      const double mM = M[c], mI = I[c], mD = D[c];
      const double iM = (c + 1 < W) ? M[(c + 1) % W] : QG_NEG_INF;          // halo diagonal dlo + W
      const double iI = (c + 1 < W) ? I[(c + 1) % W] : QG_NEG_INF;
      unsigned ptr = 0;
      const double cM = (mM + m2m) + E, cI = (mI + i2m) + E, cD = (mD + d2m) + E;
      double nM = cM;
      if (cI > nM) { nM = cI; ptr = 1; }
      if (cD > nM) { nM = cD; ptr = 2; }
      if (startRow && (i == 1 || local) && E > nM) { nM = E; ptr = 3; }
      const double aM = (iM + m2i) + ins, aI = (iI + i2i) + ins;
      double nI = aM;
      if (aI > nI) { nI = aI; ptr |= 4; }
      const double bM = leftM + m2d, bD = leftD + d2d;
      double nD = bM;
      if (bD > nD) { nD = bD; ptr |= 8; }
      if (!ok) { nM = QG_NEG_INF; nI = QG_NEG_INF; nD = QG_NEG_INF; ptr = 0; }
      M[c] = nM; I[c] = nI; D[c] = nD;
      leftM = nM; leftD = nD;
      tword |= ptr << (4 * c);
      if (endRow) {
        const bool isEnd = ok && (i == xlen || local);
        if (isEnd) { const double en = nM + m2e; if (en >= bestEnd) { bestEnd = en; bestI = i; } }
      }
    }
    acc[j & 3] = tword;
    if ((j & 3) == 3 || j == ylen) *(uint4*) (tr + (j & ~3)) = make_uint4 (acc[0], acc[1], acc[2], acc[3]);
  }
  }
  a.endvals[2 * sg.aux_off] = bestEnd; a.endvals[2 * sg.aux_off + 1] = (double) bestI;
}

#endif
