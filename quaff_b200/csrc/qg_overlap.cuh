// Read-vs-read overlap: QuaffOverlapScores (src/qoverlap.cpp:9-75), QuaffOverlapViterbiMatrix fill
// (src/qoverlap.cpp:77-160) and traceback (src/qoverlap.cpp:162-229), with the accessor swaps of
// src/qoverlap.h:46-51 resolved once on the host ("eff" = effective value the reference's code reads):
//   effI2M = i2i   effI2I = i2m   effI2D = i2d   effD2M = d2i   effD2I = d2m   effD2D = d2d
// Same lane mapping as qg_dp.cuh; the neighbour exchanges carry three values because Insert and Delete
// both read all three states of their source cell.
#ifndef QG_OVERLAP_CUH
#define QG_OVERLAP_CUH
#include "qg_dp.cuh"

// ---- emission table matchMinusInsert[iK][jK].logSymQualPairProb[xq][yq] (qoverlap.cpp:51-74) -------------------
// one thread per entry; the 4-term log-sum-exp over the hidden reference base folds left from -inf, as the reference
__global__ void qg_overlap_pair_table_kernel (const double* __restrict__ match, const double* __restrict__ insert, const double* __restrict__ lse,
                                              double lrb0, double lrb1, double lrb2, double lrb3, int match_k, int y_comp, int with_qual,
                                              double* __restrict__ table) {
  const uint64_t nK = 1ull << (2 * match_k);
  const uint64_t nq = with_qual ? QG_NQUAL : 1;
  const uint64_t total = nK * nK * nq * nq;
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const double lrb[4] = {lrb0, lrb1, lrb2, lrb3};
  if (with_qual) {
    const int jk = (int) (g % QG_NQUAL), ik = (int) ((g / QG_NQUAL) % QG_NQUAL);
    const uint64_t j = (g / (QG_NQUAL * QG_NQUAL)) % nK, i = g / (QG_NQUAL * QG_NQUAL * nK);
    double mij = QG_NEG_INF;
    for (int r = 0; r < 4; ++r) {
      const int yr = y_comp ? 3 - r : r;
      mij = qg_lse (lse, mij, (lrb[r] + match[((uint64_t) r * nK + i) * QG_NQ1 + ik]) + match[((uint64_t) yr * nK + j) * QG_NQ1 + jk]);
    }
    table[g] = (mij - insert[(i & 3) * QG_NQ1 + ik]) - insert[(j & 3) * QG_NQ1 + jk];
  } else {
    // logSymPairProb: accumulated over ik (outer) and jk (inner) in loop order (qoverlap.cpp:58-71)
    const uint64_t j = g % nK, i = g / nK;
    double acc = QG_NEG_INF;
    for (int ik = 0; ik < QG_NQUAL; ++ik)
      for (int jk = 0; jk < QG_NQUAL; ++jk) {
        double mij = QG_NEG_INF;
        for (int r = 0; r < 4; ++r) {
          const int yr = y_comp ? 3 - r : r;
          mij = qg_lse (lse, mij, (lrb[r] + match[((uint64_t) r * nK + i) * QG_NQ1 + ik]) + match[((uint64_t) yr * nK + j) * QG_NQ1 + jk]);
        }
        acc = qg_lse (lse, acc, (mij - insert[(i & 3) * QG_NQ1 + QG_NQUAL]) - insert[(j & 3) * QG_NQ1 + QG_NQUAL]);
      }
    table[g] = acc;
  }
}

// ---- per-pair context arrays -------------------------------------------------------------------------------------
struct qg_opair {
  uint32_t xseq, yseq, xlen, ylen;
  uint64_t xoff, yoff;               // token offsets in the READS set
  uint64_t xa_off, yb_off;           // offsets of this pair's xa[xlen] / yb[ylen] emission index terms
  uint64_t gx_off, gy_off;           // offsets of the padded gap-context arrays gx[xlen+1], gy[ylen+1]
  uint32_t y_comp, pad_;
};

__device__ __forceinline__ int qg_most_frequent (const uint8_t* __restrict__ tok, int len, bool complement, unsigned* s_count) {
  // block-wide; s_count[4] in shared memory
  if (threadIdx.x < 4) s_count[threadIdx.x] = 0;
  __syncthreads ();
  unsigned c[4] = {0, 0, 0, 0};
  for (int p = threadIdx.x; p < len; p += blockDim.x) ++c[complement ? 3 - (tok[p] & 3) : (tok[p] & 3)];
  for (int t = 0; t < 4; ++t) if (c[t]) atomicAdd (&s_count[t], c[t]);
  __syncthreads ();
  int b = 0;
  for (int t = 1; t < 4; ++t) if (s_count[t] > s_count[b]) b = t;
  __syncthreads ();
  return b;
}

// xa[i] = xMatchKmer[i] * nK * Q2 + xQ[i] * QG_NQUAL ; yb[j] = yMatchKmer[j] * Q2 + yQ[j]  (Q2 = 94*94 with qualities, 1 without)
// gx[0] = gy[0] = 0 (the dummy entry, qoverlap.cpp:104-106), gx[i] = gap k-mer ending at x base i, times nG; gy[j] likewise (not scaled).
// When y is complemented, y's k-mers are those of the OTHER strand read in its own direction, re-reversed
// (qoverlap.cpp:91-98): at y position j the context k-mer "ending" there covers complement(y[j]), complement(y[j+1]), ...
__global__ void qg_overlap_prep_kernel (const qg_opair* __restrict__ pairs, const uint8_t* __restrict__ tok, const uint8_t* __restrict__ qual,
                                        int match_k, int gap_k, int with_qual,
                                        uint32_t* __restrict__ xa, uint32_t* __restrict__ yb, uint32_t* __restrict__ gx, uint32_t* __restrict__ gy,
                                        const double* __restrict__ insert, double* __restrict__ ins_sums,
                                        const double* __restrict__ match, double lrb0, double lrb1, double lrb2, double lrb3,
                                        double* __restrict__ ea, double* __restrict__ eb) {
  __shared__ unsigned s_count[4];
  const qg_opair pd = pairs[blockIdx.x];
  const uint8_t* xt = tok + pd.xoff; const uint8_t* yt = tok + pd.yoff;
  const uint8_t* xq = with_qual ? qual + pd.xoff : nullptr; const uint8_t* yq = with_qual ? qual + pd.yoff : nullptr;
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  const uint32_t nK = 1u << (2 * match_k), nG = 1u << (2 * gap_k);
  const uint32_t Q2 = with_qual ? QG_NQUAL * QG_NQUAL : 1;
  const int mfx = qg_most_frequent (xt, xlen, false, s_count);
  const int mfy = qg_most_frequent (yt, ylen, pd.y_comp != 0, s_count);   // most frequent token of revcomp(y) = of complement(y)
  uint32_t* XA = xa + pd.xa_off; uint32_t* YB = yb + pd.yb_off; uint32_t* GX = gx + pd.gx_off; uint32_t* GY = gy + pd.gy_off;
  for (int i = threadIdx.x; i < xlen; i += blockDim.x) {
    uint32_t mk = 0, gk = 0;
    for (int t = match_k - 1; t >= 0; --t) { const int q = i - t; mk = mk * 4 + (q >= 0 ? xt[q] : mfx); }
    for (int t = gap_k - 1; t >= 0; --t) { const int q = i - t; gk = gk * 4 + (q >= 0 ? xt[q] : mfx); }
    XA[i] = mk * nK * Q2 + (with_qual ? xq[i] * QG_NQUAL : 0);
    GX[i + 1] = gk * nG;
    if (ea) {                                              // on-the-fly emission (qg_overlap_fill_kernel): the x factors of qoverlap.cpp:60-66
      double* A = ea + 5 * (pd.xa_off + i);
      const double lrb[4] = {lrb0, lrb1, lrb2, lrb3};
      for (int r = 0; r < 4; ++r) A[r] = lrb[r] + match[((uint64_t) r * nK + mk) * QG_NQ1 + xq[i]];
      A[4] = insert[(mk & 3) * QG_NQ1 + xq[i]];
    }
  }
  for (int j = threadIdx.x; j < ylen; j += blockDim.x) {
    uint32_t mk = 0, gk = 0;
    if (!pd.y_comp) {
      for (int t = match_k - 1; t >= 0; --t) { const int q = j - t; mk = mk * 4 + (q >= 0 ? yt[q] : mfy); }
      for (int t = gap_k - 1; t >= 0; --t) { const int q = j - t; gk = gk * 4 + (q >= 0 ? yt[q] : mfy); }
    } else {
      // position j of y is position ylen-1-j of revcomp(y); its k-mer there ends at that base and starts k-1 earlier on
      // that strand, i.e. covers y positions j, j+1, .., j+k-1 complemented, the LAST base of the k-mer being y[j]
      for (int t = match_k - 1; t >= 0; --t) { const int q = j + t; mk = mk * 4 + (q < ylen ? 3 - yt[q] : mfy); }
      for (int t = gap_k - 1; t >= 0; --t) { const int q = j + t; gk = gk * 4 + (q < ylen ? 3 - yt[q] : mfy); }
    }
    YB[j] = mk * Q2 + (with_qual ? yq[j] : 0);
    GY[j + 1] = gk;
    if (eb) {
      double* B = eb + 5 * (pd.yb_off + j);
      for (int r = 0; r < 4; ++r) B[r] = match[((uint64_t) (pd.y_comp ? 3 - r : r) * nK + mk) * QG_NQ1 + yq[j]];
      B[4] = insert[(mk & 3) * QG_NQ1 + yq[j]];
    }
  }
  if (threadIdx.x == 0) {
    GX[0] = 0; GY[0] = 0;
    // xInsertScore / yInsertScore: sequential sums in sequence order (qoverlap.cpp:108-116)
    double sx = 0, sy = 0;
    for (int i = 0; i < xlen; ++i) sx += insert[(uint32_t) xt[i] * QG_NQ1 + (with_qual ? xq[i] : QG_NQUAL)];
    for (int j = 0; j < ylen; ++j) sy += insert[(uint32_t) (pd.y_comp ? 3 - yt[j] : yt[j]) * QG_NQ1 + (with_qual ? yq[j] : QG_NQUAL)];
    ins_sums[2 * (uint64_t) blockIdx.x] = sx; ins_sums[2 * (uint64_t) blockIdx.x + 1] = sy;
  }
}

struct qg_ofill_args {
  const qg_segment* segs;
  const qg_opair* pairs;
  const uint32_t *xa, *yb, *gx, *gy;
  const double* table;               // emission table for strand 0 / 1
  const double* table1;
  const double *ea, *eb;             // non-null: no table, the emission is evaluated per cell from 5 doubles per position of x and of y
  const double *m2m, *m2i, *m2d;     // [nG][nG]
  const double* lse;
  double effI2M, effI2I, effI2D, effD2M, effD2I, effD2D;
  int nG;
  unsigned long long* trace;         // 8 bit per cell, one u64 per (macro-step, virtual lane)
  double* lastrow;                   // per segment slot: mat(i, yLen)
  double* lastcol;                   // per segment, per row j: mat(xLen, j)   (pre-filled with -inf)
};

__global__ void qg_fill_neginf_kernel (double* __restrict__ p, uint64_t n) {
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = QG_NEG_INF;
}

// pointer byte: bits 0-1 Match source (0 M, 1 I, 2 D, 3 Start); bits 2-3 Insert source (0 M, 1 I, 2 D);
// bits 4-5 Delete source (0 M, 1 I, 2 D) -- candidate order and strict '>' of qoverlap.cpp:197-228.  NB the
// Delete pointer compares ins + i2d (what the traceback reads) although the fill adds ins + effD2I (qoverlap.cpp:148).
template<int R, bool MULTI>
__global__ void __launch_bounds__ (MULTI ? 1024 : 32)
qg_overlap_fill_kernel (const qg_ofill_args a) {
  __shared__ double sA[QG_MAX_NW][3], sB[QG_MAX_NW][3];
  const qg_segment sg = a.segs[blockIdx.x];
  const qg_opair pd = a.pairs[sg.pair];
  const int NW = (int) blockDim.x >> 5;
  const int vl = threadIdx.x, lane = vl & 31, wid = vl >> 5;
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, width = (int) sg.width;
  const int SW = 32 * NW * R;
  const uint32_t* XA = a.xa + pd.xa_off; const uint32_t* YB = a.yb + pd.yb_off;
  const uint32_t* GX = a.gx + pd.gx_off; const uint32_t* GY = a.gy + pd.gy_off;
  const double* table = pd.y_comp ? a.table1 : a.table;
  const bool ctx_gap = a.nG > 1;
  const double m2m0 = a.m2m[0], m2i0 = a.m2i[0], m2d0 = a.m2d[0];
  const double I2M = a.effI2M, I2I = a.effI2I, I2D = a.effI2D, D2M = a.effD2M, D2I = a.effD2I, D2D = a.effD2D;
  const int s0 = R * vl, d0 = sg.dlo + s0;

  double M[R], I[R], D[R];
#pragma unroll
  for (int c = 0; c < R; ++c) { M[c] = QG_NEG_INF; I[c] = QG_NEG_INF; D[c] = QG_NEG_INF; }
  double lM = QG_NEG_INF, lI = QG_NEG_INF, lD = QG_NEG_INF;

  const int total = ylen + 32 * NW - 1;
  for (int u = 1; u <= total; ++u) {
    const int j = u - vl;
    const bool active = (j >= 1) && (j <= ylen);
    const uint32_t ybj = active ? YB[j - 1] : 0;
    double B[5] = {0, 0, 0, 0, 0};
    if (a.ea && active) { const double* pb = a.eb + 5 * (pd.yb_off + (uint64_t) (j - 1)); for (int r = 0; r < 5; ++r) B[r] = pb[r]; }
    const uint32_t gyj = (active && ctx_gap) ? GY[j] : 0, gyjm1 = (active && ctx_gap) ? GY[j - 1] : 0;
    unsigned long long tword = 0;
    double rM = QG_NEG_INF, rI = QG_NEG_INF, rD = QG_NEG_INF;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1) {
        const double m0 = M[0], i0 = I[0], dd0 = D[0];
        rM = __shfl_down_sync (QG_FULL_MASK, m0, 1); rI = __shfl_down_sync (QG_FULL_MASK, i0, 1); rD = __shfl_down_sync (QG_FULL_MASK, dd0, 1);
        if (MULTI) {
          if (lane == 0) { sA[wid][0] = m0; sA[wid][1] = i0; sA[wid][2] = dd0; }
          __syncthreads ();
          if (lane == 31) { if (wid + 1 < NW) { rM = sA[wid + 1][0]; rI = sA[wid + 1][1]; rD = sA[wid + 1][2]; } else { rM = rI = rD = QG_NEG_INF; } }
        } else {
          if (lane == 31) { rM = rI = rD = QG_NEG_INF; }
        }
      }
      const int i = d0 + c + j;
      const bool ok = active && (s0 + c < width) && (i >= 1) && (i <= xlen);
      double E = 0, tm2m = m2m0, tm2i = m2i0, tm2d = m2d0;
      if (ok) {
        if (a.ea) {
          // matchMinusInsert evaluated in place: the same left fold over the hidden reference base as the table entry
          // (qoverlap.cpp:60-72), so the value has the same bits; lifts the 16^K * 94^2 table limit on K
          const double* pa = a.ea + 5 * (pd.xa_off + (uint64_t) (i - 1));
          double mij = QG_NEG_INF;
#pragma unroll
          for (int r = 0; r < 4; ++r) mij = qg_lse (a.lse, mij, pa[r] + B[r]);
          E = (mij - pa[4]) - B[4];
        } else E = table[(uint64_t) XA[i - 1] + ybj];
        if (ctx_gap) {
          tm2m = a.m2m[GX[i - 1] + gyjm1];                   // m2mScore(i-1, j-1)
          tm2i = a.m2i[GX[i] + gyjm1];                       // m2iScore(i,   j-1)
          tm2d = a.m2d[GX[i - 1] + gyj];                     // m2dScore(i-1, j)
        }
      }
      const double mM = M[c], mI = I[c], mD = D[c];                                   // (i-1, j-1)
      const double iM = (c + 1 < R) ? M[(c + 1) % R] : rM, iI = (c + 1 < R) ? I[(c + 1) % R] : rI, iD = (c + 1 < R) ? D[(c + 1) % R] : rD;   // (i, j-1)
      const double dM = (c > 0) ? M[(c + R - 1) % R] : lM, dI = (c > 0) ? I[(c + R - 1) % R] : lI, dD = (c > 0) ? D[(c + R - 1) % R] : lD;   // (i-1, j)
      unsigned ptr = 0;
      // Match
      const double cM = (mM + tm2m) + E, cI = (mI + I2M) + E, cD = (mD + D2M) + E;
      double nM = cM;
      if (cI > nM) { nM = cI; ptr = 1; }
      if (cD > nM) { nM = cD; ptr = 2; }
      if ((j == 1 || i == 1) && E > nM) { nM = E; ptr = 3; }
      // Insert: value = max(lse(ins + i2i', del + d2i'), mat + m2i); pointer = argmax of the three plain candidates
      const double aM = iM + tm2i, aI = iI + I2I, aD = iD + D2I;
      { double b = aM; unsigned q = 0; if (aI > b) { b = aI; q = 1; } if (aD > b) { b = aD; q = 2; } ptr |= q << 2; }
      const double li = qg_lse (a.lse, aI, aD);
      double nI = (li > aM) ? li : aM;
      // Delete: value = max(lse(del + d2d, ins + d2i'), mat + m2d); pointer candidates use ins + i2d
      const double bM = dM + tm2d, bD = dD + D2D, bIfill = dI + D2I, bItrace = dI + I2D;
      { double b = bM; unsigned q = 0; if (bItrace > b) { b = bItrace; q = 1; } if (bD > b) { b = bD; q = 2; } ptr |= q << 4; }
      const double ld = qg_lse (a.lse, bD, bIfill);
      double nD = (ld > bM) ? ld : bM;
      if (!ok) { nM = QG_NEG_INF; nI = QG_NEG_INF; nD = QG_NEG_INF; ptr = 0; }
      M[c] = nM; I[c] = nI; D[c] = nD;
      tword |= (unsigned long long) ptr << (8 * c);
      if (j == ylen) a.lastrow[sg.aux_off + s0 + c] = ok ? nM : QG_NEG_INF;
      if (ok && i == xlen) a.lastcol[sg.acc_off + j] = nM;
    }
    {
      const double mL = M[R - 1], iL = I[R - 1], dL = D[R - 1];
      lM = __shfl_up_sync (QG_FULL_MASK, mL, 1); lI = __shfl_up_sync (QG_FULL_MASK, iL, 1); lD = __shfl_up_sync (QG_FULL_MASK, dL, 1);
      if (MULTI) {
        if (lane == 31) { sB[wid][0] = mL; sB[wid][1] = iL; sB[wid][2] = dL; }
        __syncthreads ();
        if (lane == 0) { if (wid > 0) { lM = sB[wid - 1][0]; lI = sB[wid - 1][1]; lD = sB[wid - 1][2]; } else { lM = lI = lD = QG_NEG_INF; } }
      } else {
        if (lane == 0) { lM = lI = lD = QG_NEG_INF; }
      }
    }
    a.trace[sg.trace_off + (uint64_t) u * (32 * NW) + vl] = tword;
  }
}

// end cell (qoverlap.cpp:164-182): mat(xLen,yLen) first, then row yLen by decreasing i, then column xLen by decreasing j,
// replaced only on strict '>'; result = end + xInsertScore + yInsertScore (qoverlap.cpp:159).  Then the state walk.
__global__ void qg_overlap_traceback_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                             const double* __restrict__ lastrow, const double* __restrict__ lastcol, const double* __restrict__ ins_sums,
                                             const unsigned long long* __restrict__ trace,
                                             double* __restrict__ score, uint32_t* __restrict__ coords4,
                                             uint8_t* __restrict__ path_scratch, uint32_t* __restrict__ path_len, uint32_t* __restrict__ err_flag) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  const int xlen = (int) pd.xlen, ylen = (int) pd.ylen;
  double best = QG_NEG_INF; int xe = xlen, ye = ylen;
  // mat(xLen, yLen): diagonal xlen - ylen
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_segment sg = segs[s]; const int d = xlen - ylen;
    if (d >= sg.dlo && d < sg.dlo + (int) sg.width) best = lastrow[sg.aux_off + (d - sg.dlo)];
  }
  for (uint32_t s = pd.seg_end; s-- > pd.seg_begin; ) {      // row yLen, i descending = diagonals descending
    const qg_segment sg = segs[s];
    for (int t = (int) sg.width - 1; t >= 0; --t) {
      const int i = sg.dlo + t + ylen;
      if (i < 1 || i > xlen) continue;
      const double sc = lastrow[sg.aux_off + t];
      if (sc > best) { best = sc; xe = i; ye = ylen; }
    }
  }
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {     // column xLen, j descending = diagonals ascending
    const qg_segment sg = segs[s];
    for (int t = 0; t < (int) sg.width; ++t) {
      const int j = xlen - (sg.dlo + t);
      if (j < 1 || j > ylen) continue;
      const double sc = lastcol[sg.acc_off + j];
      if (sc > best) { best = sc; xe = xlen; ye = j; }
    }
  }
  score[p] = (best + ins_sums[2 * (uint64_t) p]) + ins_sums[2 * (uint64_t) p + 1];
  coords4[4 * p] = coords4[4 * p + 1] = coords4[4 * p + 2] = coords4[4 * p + 3] = 0;
  path_len[p] = 0;
  if (!pd.want_path || !(best > QG_NEG_INF)) return;
  int i = xe, j = ye, state = 1;
  uint32_t n = 0, cs = pd.seg_begin;
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
    const unsigned long long word = trace[sg.trace_off + (uint64_t) (j + vl) * (32 * sg.nwarps) + vl];
    const uint32_t byte = (uint32_t) ((word >> (8 * c)) & 255ull);
    if (n >= pd.path_cap) { *err_flag = 2; break; }
    uint32_t src;
    if (state == 1) { buf[pd.path_cap - 1 - n] = QG_OP_MATCH; ++n; --i; --j; src = byte & 3u; state = (src == 0) ? 1 : (src == 1) ? 2 : (src == 2) ? 3 : 0; }
    else if (state == 2) { buf[pd.path_cap - 1 - n] = QG_OP_INSERT; ++n; --j; src = (byte >> 2) & 3u; state = (src == 0) ? 1 : (src == 1) ? 2 : 3; }
    else { buf[pd.path_cap - 1 - n] = QG_OP_DELETE; ++n; --i; src = (byte >> 4) & 3u; state = (src == 0) ? 1 : (src == 1) ? 2 : 3; }
    if (i < 0 || j < 0) { *err_flag = 3; break; }
  }
  coords4[4 * p] = (uint32_t) (i + 1); coords4[4 * p + 1] = (uint32_t) xe;
  coords4[4 * p + 2] = (uint32_t) (j + 1); coords4[4 * p + 3] = (uint32_t) ye;
  path_len[p] = n;
}

#endif
