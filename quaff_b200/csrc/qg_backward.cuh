// Backward fill + E-step counts (src/qmodel.cpp:1393-1510) in "pull" form.
//
// The reference pushes from each destination cell into its sources, B_src = lse(B_src, t + B_dst), visiting
// rows downwards and columns right-to-left.  For a given source cell the pushes therefore arrive in a fixed
// order -- the Match destination (i+1,j+1), then the Insert destination (i,j+1), then the Delete destination
// (i+1,j), then (last row only) the end transition -- and folding the same candidates in that order
// reproduces every B cell bit for bit.  Pulling lets each lane own its cells outright: the count of a
// transition is exp((F_src + candidate) - Z) with the source's own Forward values, read once from the stored
// Forward matrix.
//
// Mapping: the Forward kernel's, mirrored (diagonals right-to-left, rows bottom-up), so the dependency
// pattern is identical: B(d,j) <- (d, j+1) own registers; (d-1, j+1) own registers or the mirrored right
// neighbour's first cell of this macro-step; (d+1, j) own registers or the mirrored left neighbour's last
// cell of the previous macro-step.  Per-row count sums travel with the row from lane to lane and are written
// once, by the last lane, so row accumulation needs no atomics and is deterministic.
#ifndef QG_BACKWARD_CUH
#define QG_BACKWARD_CUH
#include "qg_dp.cuh"

struct qg_rowrec { double cnt[4]; double ins, m2m, m2i, m2d; };   // 64 B: one source row's count sums

template<int R, bool MULTI>
__global__ void __launch_bounds__ (MULTI ? 1024 : 32)
qg_backward_kernel (const qg_fill_args a) {
  __shared__ double sA[QG_MAX_NW], sB[QG_MAX_NW];
  __shared__ double sRec[QG_MAX_NW][8];
  const qg_segment sg = a.segs[blockIdx.x];
  const int NW = (int) blockDim.x >> 5;
  const int vl = threadIdx.x, lane = vl & 31, wid = vl >> 5;
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, width = (int) sg.width;
  const int SW = 32 * NW * R;
  const int lastvl = 32 * NW - 1;
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowp* rp = a.rp + sg.rp_off;
  const double* stbase = a.store + sg.store_off;
  const double i2i = a.i2i, i2m = a.i2m, d2d = a.d2d, d2m = a.d2m;
  const bool local = a.local != 0;
  const double m2e = rp[0].m2m;
  const double Z = a.pair_z[sg.pair];
  const bool zok = Z > QG_NEG_INF;                          // no path at all: every count is exp(-inf - -inf) in the reference; we emit zeros
  qg_rowrec* rowacc = (qg_rowrec*) a.rowacc + sg.acc_off;

  double bM[R], bI[R];
#pragma unroll
  for (int c = 0; c < R; ++c) { bM[c] = QG_NEG_INF; bI[c] = QG_NEG_INF; }
  double leftD = QG_NEG_INF;
  double s_d2m = 0, s_i2m = 0, s_i2i = 0, s_d2d = 0, s_m2e = 0, s2m[4] = {0, 0, 0, 0};
  qg_rowrec rec;
#pragma unroll
  for (int t = 0; t < 4; ++t) rec.cnt[t] = 0;
  rec.ins = rec.m2m = rec.m2i = rec.m2d = 0;
  qg_rowp Pn = rp[ylen + 1];                                // parameters of row j+1 (zero filler above the last row)

  const int total = ylen + 32 * NW - 1;
  for (int u = 1; u <= total; ++u) {
    const int j = ylen + 1 - (u - vl);                      // actual row
    const bool active = (j >= 1) && (j <= ylen);
    const int jj = j < 0 ? 0 : (j > ylen + 1 ? ylen + 1 : j);
    const qg_rowp Pc = rp[jj];
    if (j == ylen) Pn = rp[ylen + 1];

    // ---- the row's running sums arrive from the mirrored left neighbour
    {
      qg_rowrec in;
#pragma unroll
      for (int t = 0; t < 4; ++t) in.cnt[t] = __shfl_up_sync (QG_FULL_MASK, rec.cnt[t], 1);
      in.ins = __shfl_up_sync (QG_FULL_MASK, rec.ins, 1);
      in.m2m = __shfl_up_sync (QG_FULL_MASK, rec.m2m, 1);
      in.m2i = __shfl_up_sync (QG_FULL_MASK, rec.m2i, 1);
      in.m2d = __shfl_up_sync (QG_FULL_MASK, rec.m2d, 1);
      if (MULTI) {
        if (lane == 31) { for (int t = 0; t < 4; ++t) sRec[wid][t] = rec.cnt[t]; sRec[wid][4] = rec.ins; sRec[wid][5] = rec.m2m; sRec[wid][6] = rec.m2i; sRec[wid][7] = rec.m2d; }
        __syncthreads ();
        if (lane == 0 && wid > 0) { for (int t = 0; t < 4; ++t) in.cnt[t] = sRec[wid - 1][t]; in.ins = sRec[wid - 1][4]; in.m2m = sRec[wid - 1][5]; in.m2i = sRec[wid - 1][6]; in.m2d = sRec[wid - 1][7]; }
      }
      if (vl == 0) { for (int t = 0; t < 4; ++t) in.cnt[t] = 0; in.ins = in.m2m = in.m2i = in.m2d = 0; }
      rec = in;
    }

    double rightI = QG_NEG_INF;
    double nD[R];
#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1) {
        // ---- exchange 1: first cells' new B_I (row j+1 of the mirrored right neighbour) travel one lane left
        const double i0 = bI[0];
        rightI = __shfl_down_sync (QG_FULL_MASK, i0, 1);
        if (MULTI) {
          if (lane == 0) sA[wid] = i0;
          __syncthreads ();
          if (lane == 31) rightI = (wid + 1 < NW) ? sA[wid + 1] : QG_NEG_INF;
        } else {
          if (lane == 31) rightI = QG_NEG_INF;
        }
      }
      const int s = SW - 1 - (R * vl + c);                  // actual slot
      const int d = sg.dlo + s;
      const int i = d + j;
      const bool ok = active && (s < width) && (i >= 1) && (i <= xlen);
      const int tn = qg_tok (xw, nxw, i);                   // x[i]   (0-based): the base emitted by the Match destination (i+1, j+1)
      const int tc = qg_tok (xw, nxw, i - 1);               // x[i-1]: the base of this cell
      const double En = qg_sel4 (Pn.e, tn);
      const double srcM = bM[c];                            // B_M(i+1, j+1)
      const double srcI = (c + 1 < R) ? bI[(c + 1) % R] : rightI;          // B_I(i, j+1)
      const double srcD = (c > 0) ? nD[(c + R - 1) % R] : leftD;           // B_D(i+1, j)
      const double cM = (Pn.m2m + En) + srcM, cIM = (i2m + En) + srcM, cDM = (d2m + En) + srcM;
      const double cI = (Pn.m2i + Pn.ins) + srcI, cII = (i2i + Pn.ins) + srcI;
      const double cD = Pc.m2d + srcD, cDD = d2d + srcD;
      const bool isEnd = (j == ylen) && (i == xlen || local);
      double BM = qg_lse (a.lse, qg_lse (a.lse, cM, cI), cD);
      if (isEnd) BM = qg_lse (a.lse, BM, m2e + 0.0);
      double BI = qg_lse (a.lse, cIM, cII);
      double BD = qg_lse (a.lse, cDM, cDD);
      if (!ok) { BM = QG_NEG_INF; BI = QG_NEG_INF; BD = QG_NEG_INF; }
      if (ok && zok) {
        // Forward wrote cell (slot s, row j) at macro-step j + fvl by virtual lane fvl = s / R, position s % R
        const int fvl = s / R, fc = s - fvl * R;
        const double* st = stbase + ((uint64_t) (j + fvl) * (32 * NW) + fvl) * (3 * R);
        const double fM = st[fc], fI = st[R + fc], fD = st[2 * R + fc];
        const double n_m2m = exp ((fM + cM) - Z), n_i2m = exp ((fI + cIM) - Z), n_d2m = exp ((fD + cDM) - Z);
        const double n_m2i = exp ((fM + cI) - Z), n_i2i = exp ((fI + cII) - Z);
        const double n_m2d = exp ((fM + cD) - Z), n_d2d = exp ((fD + cDD) - Z);
        const double nm = n_m2m + n_d2m + n_i2m;
        rec.cnt[0] += (tn == 0) ? nm : 0.0; rec.cnt[1] += (tn == 1) ? nm : 0.0;
        rec.cnt[2] += (tn == 2) ? nm : 0.0; rec.cnt[3] += (tn == 3) ? nm : 0.0;
        rec.ins += n_m2i + n_i2i;
        rec.m2m += n_m2m; rec.m2i += n_m2i; rec.m2d += n_m2d;
        s_d2m += n_d2m; s_i2m += n_i2m; s_i2i += n_i2i; s_d2d += n_d2d;
        if (isEnd) s_m2e += exp ((fM + (m2e + 0.0)) - Z);
        if (j == 1 && (i == 1 || local)) {
          const double ns = exp ((0.0 + (qg_sel4 (Pc.e, tc) + BM)) - Z);
          s2m[0] += (tc == 0) ? ns : 0.0; s2m[1] += (tc == 1) ? ns : 0.0;
          s2m[2] += (tc == 2) ? ns : 0.0; s2m[3] += (tc == 3) ? ns : 0.0;
        }
      }
      if (j == 1) a.endvals[sg.aux_off + s] = (ok && (i == 1 || local)) ? qg_sel4 (Pc.e, tc) + BM : QG_NEG_INF;
      bM[c] = BM; bI[c] = BI; nD[c] = BD;
    }

    // ---- exchange 2: last cells' new B_D (row j) travel one lane to the right, for the next macro-step
    {
      const double dL = nD[R - 1];
      leftD = __shfl_up_sync (QG_FULL_MASK, dL, 1);
      if (MULTI) {
        if (lane == 31) sB[wid] = dL;
        __syncthreads ();
        if (lane == 0) leftD = (wid > 0) ? sB[wid - 1] : QG_NEG_INF;
      } else {
        if (lane == 0) leftD = QG_NEG_INF;
      }
    }
    if (vl == lastvl && active) rowacc[j] = rec;
    Pn = Pc;
  }

  // ---- row-independent sums: warp reduce, then one atomic per warp into the segment's slots
  double sc[9] = {s_d2m, s_i2m, s_i2i, s_d2d, s_m2e, s2m[0], s2m[1], s2m[2], s2m[3]};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    double v = sc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync (QG_FULL_MASK, v, o);
    if (lane == 0 && v != 0.0) atomicAdd (&a.seg_scal[12 * sg.seg_id + t], v);
  }
}

// Backward result: start = lse(start, E(i,1) + B_M(i,1)) folded over DESCENDING i (qmodel.cpp:1440-1446)
__global__ void qg_backward_finalize_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                             const double* __restrict__ endvals, const double* __restrict__ lse, double* __restrict__ result) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  double start = QG_NEG_INF;
  for (uint32_t s = pd.seg_end; s-- > pd.seg_begin; ) {
    const qg_segment sg = segs[s];
    for (uint32_t t = sg.width; t-- > 0; ) start = qg_lse (lse, start, endvals[sg.aux_off + t]);
  }
  result[p] = start;
}

// ---- scatter the per-row sums into the pair's QuaffCounts table ---------------------------------------------------
__device__ __forceinline__ uint64_t qg_ctx_kmer (const uint8_t* __restrict__ tok, int mf, int last, int k) {
  uint64_t v = 0;                                           // k-mer of length k ENDING at 0-based position `last`
  for (int t = k - 1; t >= 0; --t) { const int q = last - t; v = v * 4 + (q >= 0 ? tok[q] : mf); }
  return v;
}

// one CTA per pair; table layout = include/quaffgpu.h (QuaffCounts)
__global__ void qg_counts_scatter_kernel (const qg_pair_dp* __restrict__ pairs, const qg_segment* __restrict__ segs,
                                          const uint8_t* __restrict__ ytok, const uint8_t* __restrict__ yqual, const uint64_t* __restrict__ yoffs,
                                          const qg_rowrec* __restrict__ rowacc_all, const double* __restrict__ seg_scal,
                                          int match_k, int gap_k, uint64_t ncounts, double* __restrict__ tables) {
  __shared__ unsigned s_count[4];
  __shared__ int s_mf;
  const qg_pair_dp pd = pairs[blockIdx.x];
  if (pd.seg_begin == pd.seg_end) return;
  const qg_segment s0 = segs[pd.seg_begin];
  const uint8_t* tok = ytok + yoffs[s0.yseq];
  const uint8_t* ql = yqual + yoffs[s0.yseq];
  const int ylen = (int) pd.ylen;
  const uint64_t nK = 1ull << (2 * match_k), nG = 1ull << (2 * gap_k);
  double* T = tables + (uint64_t) blockIdx.x * ncounts;
  double* Tmatch = T;
  double* Tins = Tmatch + 4 * nK * QG_NQUAL;
  double* Tm2m = Tins + 4 * QG_NQUAL;
  double* Tm2i = Tm2m + nG; double* Tm2d = Tm2i + nG; double* Tm2e = Tm2d + nG; double* Tsc = Tm2e + nG;
  if (threadIdx.x < 4) s_count[threadIdx.x] = 0;
  __syncthreads ();
  unsigned c[4] = {0, 0, 0, 0};
  for (int p = threadIdx.x; p < ylen; p += blockDim.x) ++c[tok[p] & 3];
  for (int t = 0; t < 4; ++t) if (c[t]) atomicAdd (&s_count[t], c[t]);
  __syncthreads ();
  if (threadIdx.x == 0) { int b = 0; for (int t = 1; t < 4; ++t) if (s_count[t] > s_count[b]) b = t; s_mf = b; }
  __syncthreads ();
  const int mf = s_mf;
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_segment sg = segs[s];
    const qg_rowrec* ra = rowacc_all + sg.acc_off;
    for (int j = 1 + (int) threadIdx.x; j <= ylen; j += blockDim.x) {
      const qg_rowrec r = ra[j];
      const uint64_t g = qg_ctx_kmer (tok, mf, j - 1, gap_k);                 // c(j)
      if (r.m2m != 0.0) atomicAdd (&Tm2m[g], r.m2m);
      if (r.m2i != 0.0) atomicAdd (&Tm2i[g], r.m2i);
      if (r.m2d != 0.0) atomicAdd (&Tm2d[g], r.m2d);
      if (j + 1 <= ylen) {                                                    // emissions of destination row j+1
        const uint64_t mk = qg_ctx_kmer (tok, mf, j, match_k);
        const int q = ql[j];
        for (int t = 0; t < 4; ++t) if (r.cnt[t] != 0.0) atomicAdd (&Tmatch[((uint64_t) t * nK + mk) * QG_NQUAL + q], r.cnt[t]);
        if (r.ins != 0.0) atomicAdd (&Tins[(uint64_t) tok[j] * QG_NQUAL + q], r.ins);
      }
    }
    if (threadIdx.x == 0) {
      const double* sc = seg_scal + 12 * sg.seg_id;
      // QuaffCounts scalar order: d2d, d2m, i2i, i2m
      if (sc[3] != 0.0) atomicAdd (&Tsc[0], sc[3]);
      if (sc[0] != 0.0) atomicAdd (&Tsc[1], sc[0]);
      if (sc[2] != 0.0) atomicAdd (&Tsc[2], sc[2]);
      if (sc[1] != 0.0) atomicAdd (&Tsc[3], sc[1]);
      if (sc[4] != 0.0) atomicAdd (&Tm2e[qg_ctx_kmer (tok, mf, ylen - 1, gap_k)], sc[4]);
      const uint64_t mk1 = qg_ctx_kmer (tok, mf, 0, match_k);                 // Start -> Match lands on row 1
      for (int t = 0; t < 4; ++t) if (sc[5 + t] != 0.0) atomicAdd (&Tmatch[((uint64_t) t * nK + mk1) * QG_NQUAL + ql[0]], sc[5 + t]);
    }
  }
}

// out[k] (+)= sum_p w[p] * tables[p][k], pairs in order (one thread per table entry)
__global__ void qg_counts_reduce_kernel (const double* __restrict__ tables, const double* __restrict__ w, uint32_t npairs, uint64_t ncounts, double* __restrict__ out) {
  const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncounts) return;
  double acc = out[k];
  for (uint32_t p = 0; p < npairs; ++p) acc += (w ? w[p] : 1.0) * tables[(uint64_t) p * ncounts + k];
  out[k] = acc;
}

#endif
