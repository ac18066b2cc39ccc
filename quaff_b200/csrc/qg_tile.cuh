// Envelopes with very wide runs (more than 8192 consecutive diagonals: `-kmatchoff` on long sequences, BASELINE
// config 5) do not fit the one-CTA-per-run scheme of qg_dp.cuh.  They are filled in (i, j) space instead: the matrix is
// cut into tiles of QG_TCW columns x QG_TRH rows; a tile depends on its left, upper and upper-left neighbours only, so
// all tiles of one anti-diagonal a + b = w of ALL pairs of the batch run in one launch (one warp per tile), w = 0, 1, ...
// Inside a tile lane l owns QG_TRC adjacent columns and fills row j0 + u - l at step u: everything flows left-to-right,
// one 3-double shuffle per step, no two-phase exchange.  Tile edges go through two small HBM buffers (last column per
// column block, last row per row block).  Same FP64 arithmetic, pointer rules and tie order as qg_dp.cuh (Viterbi,
// qmodel.cpp:1512-1646) and the same table log-sum-exp (Forward, qmodel.cpp:1343-1391): results stay bit-identical.
#ifndef QG_TILE_CUH
#define QG_TILE_CUH
#include "qg_dp.cuh"

#define QG_TRC 8                       /* columns per lane  */
#define QG_TCW (32 * QG_TRC)           /* columns per tile  */
#define QG_TRH 256                     /* rows per tile     */
#define QG_TILE_WORDS ((QG_TRH + 32) * 32)

struct qg_wseg {                       // one run of a "wide" pair
  uint32_t pair, xseq, xlen, ylen;
  int32_t dlo, dhi;
  uint32_t nCB, nRB;
  uint64_t rp_off;
  uint64_t table_off;                  // int64 [nCB][nRB]: id of the tile, -1 if the run does not touch it
  uint64_t end_off;                    // Forward: doubles [xlen+1], M(i,yLen)+m2e per column
};

// id: index in pair order; it addresses the tile's pointer block (id * QG_TILE_WORDS), its last-column slice
// (id * QG_TRH * 3 doubles), its last-row slice (id * QG_TCW * 3 doubles) and its tile_best slot.
// left / up / diag: ids of the tiles (a-1,b), (a,b-1), (a-1,b-1) of the same run, -1 where the run does not touch them
struct qg_tile { uint32_t seg, a, b, id; int32_t left, up, diag, pad_; };

struct qg_tile_args {
  const qg_wseg* segs;
  const qg_tile* tiles;
  const uint64_t* xpacked;
  const uint64_t* xpoff;
  const qg_rowp* rp;
  const double* lse;
  double i2i, i2m, d2d, d2m;
  int local;
  uint32_t* trace;
  double* colbuf;                      // [tile][QG_TRH][3]  last column of every tile
  double* rowbuf;                      // [tile][QG_TCW][3]  last row of every tile
  double* tile_best;                   // Viterbi: {score, i} per tile of the last row block
  double* endvals;                     // Forward
};

template<int MODE>                     // 0 Viterbi (+ pointers), 1 Forward (log space, reference table)
__global__ void __launch_bounds__ (32)
qg_tile_kernel (const qg_tile_args a, uint32_t tile_begin) {
  const qg_tile tl = a.tiles[tile_begin + blockIdx.x];
  const qg_wseg ws = a.segs[tl.seg];
  const int lane = threadIdx.x;
  const int xlen = (int) ws.xlen, ylen = (int) ws.ylen;
  const int dlo = ws.dlo, dhi = ws.dhi;
  const uint64_t* xw = a.xpacked + a.xpoff[ws.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowp* rp = a.rp + ws.rp_off;
  const double i2i = a.i2i, i2m = a.i2m, d2d = a.d2d, d2m = a.d2m;
  const bool local = a.local != 0;
  const double m2e = rp[0].m2m;
  const int i0 = (int) tl.a * QG_TCW + 1 + lane * QG_TRC;        // my first column (1-based)
  const int j0 = (int) tl.b * QG_TRH + 1;                        // first row of the tile
  const int jend = (j0 + QG_TRH - 1 < ylen) ? j0 + QG_TRH - 1 : ylen;   // last row of the tile
  const int ti0 = (int) tl.a * QG_TCW + 1;                       // first column of the tile
  // edge slices, addressed relative to the tile: column slices by j - j0, row slices by i - ti0
  const double* colIn = a.colbuf + ((int64_t) (tl.left < 0 ? 0 : tl.left) * (QG_TRH * 3) - (int64_t) j0 * 3);    // column ti0 - 1
  double* colOut = a.colbuf + ((int64_t) tl.id * (QG_TRH * 3) - (int64_t) j0 * 3);
  const double* rowIn = a.rowbuf + ((int64_t) (tl.up < 0 ? 0 : tl.up) * (QG_TCW * 3) - (int64_t) ti0 * 3);       // row j0 - 1
  double* rowOut = a.rowbuf + ((int64_t) tl.id * (QG_TCW * 3) - (int64_t) ti0 * 3);
  const double* diagIn = a.rowbuf + (uint64_t) (tl.diag < 0 ? 0 : tl.diag) * (QG_TCW * 3) + (uint64_t) (QG_TCW - 1) * 3;   // (ti0 - 1, j0 - 1)

  // row j0-1 of my columns, and of the column to my left
  double M[QG_TRC], I[QG_TRC], D[QG_TRC];
  int tok[QG_TRC];
#pragma unroll
  for (int c = 0; c < QG_TRC; ++c) {
    const int i = i0 + c, jt = j0 - 1;
    const bool have = tl.up >= 0 && i <= xlen && (i - jt) >= dlo && (i - jt) <= dhi;
    M[c] = have ? rowIn[(int64_t) i * 3] : QG_NEG_INF;
    I[c] = have ? rowIn[(int64_t) i * 3 + 1] : QG_NEG_INF;
    D[c] = have ? rowIn[(int64_t) i * 3 + 2] : QG_NEG_INF;
    tok[c] = qg_tok (xw, nxw, i - 1);
  }
  double tlM, tlI, tlD;                                          // (i0-1, j0-1)
  {
    const int i = i0 - 1, jt = j0 - 1;
    // lane 0's upper-left cell lies in the tile (a-1, b-1); the other lanes' in the tile above
    const double* src = (lane == 0) ? diagIn : rowIn + (int64_t) i * 3;
    const bool have = (lane == 0 ? tl.diag >= 0 : tl.up >= 0) && i >= 1 && jt >= 1 && (i - jt) >= dlo && (i - jt) <= dhi;
    tlM = have ? src[0] : QG_NEG_INF;
    tlI = have ? src[1] : QG_NEG_INF;
    tlD = have ? src[2] : QG_NEG_INF;
  }
  double l1M = QG_NEG_INF, l1I = QG_NEG_INF, l1D = QG_NEG_INF;   // left column at my current row (from the left lane / colIn)
  double pvM = QG_NEG_INF, pvI = QG_NEG_INF, pvD = QG_NEG_INF;   // ... and at my previous row
  double bestEnd = QG_NEG_INF; int bestI = 0;

  const int total = QG_TRH + 31;
  for (int u = 0; u < total; ++u) {
    const int j = j0 + u - lane;
    const bool active = (j >= j0) && (j <= jend);
    const int jj = j < 0 ? 0 : (j > ylen + 1 ? ylen + 1 : j);
    const qg_rowp P = rp[jj];
    // left column (i0-1) at my row j: lane 0 reads the previous tile's last column, the others received it from the left
    // lane at the end of the previous step (that lane is one row ahead); at row j-1: what was "row j" one step ago
    if (lane == 0) {
      const int i = i0 - 1;
      const bool have = active && tl.left >= 0 && (i - j) >= dlo && (i - j) <= dhi;
      l1M = have ? colIn[(int64_t) j * 3] : QG_NEG_INF;
      l1I = have ? colIn[(int64_t) j * 3 + 1] : QG_NEG_INF;
      l1D = have ? colIn[(int64_t) j * 3 + 2] : QG_NEG_INF;
    }
    double l0M = pvM, l0I = pvI, l0D = pvD;
    if (u == lane) { l0M = tlM; l0I = tlI; l0D = tlD; }          // first row of the tile: (i0-1, j0-1) comes from the row above
    const bool startRow = (j == 1), endRow = (j == ylen);
    unsigned tword = 0;
    double dgM = l0M, dgI = l0I, dgD = l0D;                      // (i-1, j-1) of the current column
    double lfM = l1M, lfD = l1D;                                 // (i-1, j)
#pragma unroll
    for (int c = 0; c < QG_TRC; ++c) {
      const int i = i0 + c;
      const bool ok = active && (i <= xlen) && (i - j) >= dlo && (i - j) <= dhi;
      const double E = qg_sel4 (P.e, tok[c]);
      const double upM = M[c], upI = I[c], upD = D[c];           // (i, j-1)
      double nM, nI, nD;
      unsigned ptr = 0;
      if (MODE == 0) {
        const double cM = (dgM + P.m2m) + E, cI = (dgI + i2m) + E, cD = (dgD + d2m) + E;
        nM = cM;
        if (cI > nM) { nM = cI; ptr = 1; }
        if (cD > nM) { nM = cD; ptr = 2; }
        if (startRow && (i == 1 || local) && E > nM) { nM = E; ptr = 3; }
        const double aM = (upM + P.m2i) + P.ins, aI = (upI + i2i) + P.ins;
        nI = aM;
        if (aI > nI) { nI = aI; ptr |= 4; }
        const double bM = lfM + P.m2d, bD = lfD + d2d;
        nD = bM;
        if (bD > nD) { nD = bD; ptr |= 8; }
      } else {
        double mat = qg_lse (a.lse, qg_lse (a.lse, dgM + P.m2m, dgD + d2m), dgI + i2m);
        if (startRow && (i == 1 || local)) mat = qg_lse (a.lse, mat, 0.0);
        nM = mat + E;
        nI = P.ins + qg_lse (a.lse, upI + i2i, upM + P.m2i);
        nD = qg_lse (a.lse, lfD + d2d, lfM + P.m2d);
      }
      if (!ok) { nM = QG_NEG_INF; nI = QG_NEG_INF; nD = QG_NEG_INF; ptr = 0; }
      // the old values of this column are the next column's (i-1, j-1); the new ones its (i-1, j)
      dgM = upM; dgI = upI; dgD = upD;
      lfM = nM; lfD = nD;
      if (active) { M[c] = nM; I[c] = nI; D[c] = nD; }
      tword |= ptr << (4 * c);
      if (endRow && active) {
        const bool isEnd = ok && (i == xlen || local);
        if (MODE == 0) { if (isEnd) { const double e = nM + m2e; if (e >= bestEnd) { bestEnd = e; bestI = i; } } }
        else if (i <= xlen) a.endvals[ws.end_off + i] = isEnd ? nM + m2e : QG_NEG_INF;
      }
    }
    if (MODE == 0) a.trace[(uint64_t) tl.id * QG_TILE_WORDS + (uint64_t) u * 32 + lane] = tword;
    // my last column at row j goes to the right lane (its row j at the next step) / to the column buffer of the next tile
    {
      const double sM = active ? M[QG_TRC - 1] : QG_NEG_INF, sI = active ? I[QG_TRC - 1] : QG_NEG_INF, sD = active ? D[QG_TRC - 1] : QG_NEG_INF;
      const double rM = __shfl_up_sync (QG_FULL_MASK, sM, 1), rI = __shfl_up_sync (QG_FULL_MASK, sI, 1), rD = __shfl_up_sync (QG_FULL_MASK, sD, 1);
      pvM = l1M; pvI = l1I; pvD = l1D;
      if (lane > 0) { l1M = rM; l1I = rI; l1D = rD; }
      if (lane == 31 && active) { colOut[(int64_t) j * 3] = sM; colOut[(int64_t) j * 3 + 1] = sI; colOut[(int64_t) j * 3 + 2] = sD; }
    }
    if (active && j == jend) {
#pragma unroll
      for (int c = 0; c < QG_TRC; ++c) {
        const int i = i0 + c;
        if (i <= xlen) { rowOut[(int64_t) i * 3] = M[c]; rowOut[(int64_t) i * 3 + 1] = I[c]; rowOut[(int64_t) i * 3 + 2] = D[c]; }
      }
    }
  }
  if (MODE == 0) {
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync (QG_FULL_MASK, bestEnd, o);
      const int oi = __shfl_down_sync (QG_FULL_MASK, bestI, o);
      if (ob > bestEnd || (ob == bestEnd && oi > bestI)) { bestEnd = ob; bestI = oi; }
    }
    if (lane == 0) { a.tile_best[2 * (uint64_t) tl.id] = bestEnd; a.tile_best[2 * (uint64_t) tl.id + 1] = (double) bestI; }
  }
}

// per wide pair: Viterbi end cell over its tiles; Forward fold over its columns
struct qg_wpair { uint32_t seg_begin, seg_end, tile_begin, tile_end, xlen, ylen; uint64_t path_off; uint32_t path_cap, want_path; };

__global__ void qg_wide_score_kernel (const qg_wpair* __restrict__ pairs, uint32_t npairs, const double* __restrict__ tile_best,
                                      double* __restrict__ score, uint32_t* __restrict__ x_end) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_wpair pd = pairs[p];
  double best = QG_NEG_INF; int bi = 0;
  for (uint32_t t = pd.tile_begin; t < pd.tile_end; ++t) {
    const double sc = tile_best[2 * (uint64_t) t]; const int si = (int) tile_best[2 * (uint64_t) t + 1];
    if (sc > best || (sc == best && si > bi)) { best = sc; bi = si; }
  }
  score[p] = best;
  x_end[p] = (best > QG_NEG_INF) ? (uint32_t) bi : 0u;
}

__global__ void qg_wide_forward_finalize_kernel (const qg_wpair* __restrict__ pairs, uint32_t npairs, const qg_wseg* __restrict__ segs,
                                                 const double* __restrict__ endvals, const double* __restrict__ lse, double* __restrict__ result) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_wpair pd = pairs[p];
  // ascending i over the whole last row: the runs of a pair are disjoint in i at row yLen, ascending with dlo
  double end = QG_NEG_INF;
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_wseg ws = segs[s];
    int ilo = ws.dlo + (int) ws.ylen, ihi = ws.dhi + (int) ws.ylen;
    if (ilo < 1) ilo = 1;
    if (ihi > (int) ws.xlen) ihi = (int) ws.xlen;
    for (int i = ilo; i <= ihi; ++i) end = qg_lse (lse, end, endvals[ws.end_off + i]);
  }
  result[p] = end;
}

__global__ void qg_wide_traceback_kernel (const qg_wpair* __restrict__ pairs, uint32_t npairs, const qg_wseg* __restrict__ segs,
                                          const long long* __restrict__ tables, const uint32_t* __restrict__ trace,
                                          const double* __restrict__ score, const uint32_t* __restrict__ x_end,
                                          uint32_t* __restrict__ x_start, uint8_t* __restrict__ path_scratch, uint32_t* __restrict__ path_len,
                                          uint32_t* __restrict__ err_flag) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_wpair pd = pairs[p];
  x_start[p] = 0; path_len[p] = 0;
  if (!pd.want_path || !(score[p] > QG_NEG_INF)) return;
  int i = (int) x_end[p], j = (int) pd.ylen, state = 1;
  uint32_t n = 0, cs = pd.seg_begin;
  uint8_t* buf = path_scratch + pd.path_off;
  while (state != 0) {
    const int d = i - j;
    qg_wseg ws = segs[cs];
    if (d < ws.dlo || d > ws.dhi) {
      bool found = false;
      for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) { const qg_wseg t = segs[s]; if (d >= t.dlo && d <= t.dhi) { cs = s; ws = t; found = true; break; } }
      if (!found) { *err_flag = 1; break; }
    }
    const int ta = (i - 1) / QG_TCW, tb = (j - 1) / QG_TRH;
    const long long toff = tables[ws.table_off + (uint64_t) ta * ws.nRB + tb];
    if (toff < 0) { *err_flag = 4; break; }
    const uint64_t tbase = (uint64_t) toff * QG_TILE_WORDS;
    const int ci = (i - 1) - ta * QG_TCW, ln = ci / QG_TRC, c = ci - ln * QG_TRC;
    const int u = (j - 1 - tb * QG_TRH) + ln;
    const uint32_t nib = (trace[tbase + (uint64_t) u * 32 + ln] >> (4 * c)) & 15u;
    if (n >= pd.path_cap) { *err_flag = 2; break; }
    if (state == 1) { buf[pd.path_cap - 1 - n] = QG_OP_MATCH; ++n; --i; --j; const uint32_t src = nib & 3u; state = (src == 0) ? 1 : (src == 1) ? 2 : (src == 2) ? 3 : 0; }
    else if (state == 2) { buf[pd.path_cap - 1 - n] = QG_OP_INSERT; ++n; --j; state = (nib & 4u) ? 2 : 1; }
    else { buf[pd.path_cap - 1 - n] = QG_OP_DELETE; ++n; --i; state = (nib & 8u) ? 3 : 1; }
    if (i < 0 || j < 0) { *err_flag = 3; break; }
  }
  x_start[p] = (uint32_t) (i + 1);
  path_len[p] = n;
}

__global__ void qg_wide_path_gather_kernel (const qg_wpair* __restrict__ pairs, uint32_t npairs, const uint8_t* __restrict__ path_scratch,
                                            const uint32_t* __restrict__ path_len, const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out) {
  const uint32_t p = blockIdx.x;
  if (p >= npairs) return;
  const qg_wpair pd = pairs[p];
  const uint32_t n = path_len[p];
  const uint8_t* src = path_scratch + pd.path_off + (pd.path_cap - n);
  uint8_t* dst = out + out_off[p];
  for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) dst[t] = src[t];
}

#endif
