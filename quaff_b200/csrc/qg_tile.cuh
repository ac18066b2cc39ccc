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
#include "qg_backward.cuh"

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
  uint64_t end_off;                    // Forward: doubles [xlen+1], M(i,yLen)+m2e per column; Backward: E(i,1) + B_M(i,1) per column
  uint64_t acc_off;                    // Backward: row offset of this run's per-row count sums (qg_rowrec [ylen+2])
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
  // Forward-Backward (count / train on wide runs)
  double* fstore;                      // Forward cells, [tile][row in tile][column in tile][M, I, D]; null = not kept
  const qg_tile* btiles;               // Backward: tiles of the REVERSED matrix (i' = xLen+1-i, j' = yLen+1-j), wavefront order
  const long long* tables;             // per run [nCB][nRB]: Forward tile id of a block (where a cell's stored values are)
  const double* pair_z;                // Forward log-likelihood per pair
  double* rowacc;                      // qg_rowrec per run row
  double* seg_scal;                    // 12 doubles per run
};
#define QG_TILE_CELLS3 ((uint64_t) QG_TRH * QG_TCW * 3)

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
      if (MODE == 1 && a.fstore && active) {
        double* fs = a.fstore + (uint64_t) tl.id * QG_TILE_CELLS3 + ((uint64_t) (j - j0) * QG_TCW + (uint64_t) (i - ti0)) * 3;
        fs[0] = nM; fs[1] = nI; fs[2] = nD;
      }
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

// ---- Backward + counts on tiles (qmodel.cpp:1393-1510, pull form as in qg_backward.cuh) ---------------------------------
// The Backward recurrence B(i,j) <- (i+1,j+1), (i,j+1), (i+1,j) is a Forward-shaped recurrence in the reversed coordinates
// i' = xLen+1-i, j' = yLen+1-j: (i'-1,j'-1), (i',j'-1), (i'-1,j').  The same tile scheme is therefore run over the reversed
// matrix (its own tile list, band [xLen-yLen-dhi, xLen-yLen-dlo]); a cell's Forward values are fetched from the Forward
// tiles' store through the run's block table.  Candidate order, table log-sum-exp and count expressions are those of
// qg_backward_kernel; the per-row count sums are added to the run's row records with FP64 atomics (several tiles share a row:
// the sums are not order-deterministic, unlike the banded kernel's; the parity bar for counts is 1e-4 relative).
__global__ void __launch_bounds__ (32)
qg_tile_backward_kernel (const qg_tile_args a, uint32_t tile_begin) {
  const qg_tile tl = a.btiles[tile_begin + blockIdx.x];
  const qg_wseg ws = a.segs[tl.seg];
  const int lane = threadIdx.x;
  const int xlen = (int) ws.xlen, ylen = (int) ws.ylen;
  const int dlo = (xlen - ylen) - ws.dhi, dhi = (xlen - ylen) - ws.dlo;     // band of the reversed matrix
  const uint64_t* xw = a.xpacked + a.xpoff[ws.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowp* rp = a.rp + ws.rp_off;
  const double i2i = a.i2i, i2m = a.i2m, d2d = a.d2d, d2m = a.d2m;
  const bool local = a.local != 0;
  const double m2e = rp[0].m2m;
  const double Z = a.pair_z[ws.pair];
  const bool zok = Z > QG_NEG_INF;
  const long long* tab = a.tables + ws.table_off;
  qg_rowrec* rowacc = (qg_rowrec*) a.rowacc + ws.acc_off;
  const int i0 = (int) tl.a * QG_TCW + 1 + lane * QG_TRC;        // my first reversed column
  const int j0 = (int) tl.b * QG_TRH + 1;                        // first reversed row of the tile
  const int jend = (j0 + QG_TRH - 1 < ylen) ? j0 + QG_TRH - 1 : ylen;
  const int ti0 = (int) tl.a * QG_TCW + 1;
  const double* colIn = a.colbuf + ((int64_t) (tl.left < 0 ? 0 : tl.left) * (QG_TRH * 3) - (int64_t) j0 * 3);
  double* colOut = a.colbuf + ((int64_t) tl.id * (QG_TRH * 3) - (int64_t) j0 * 3);
  const double* rowIn = a.rowbuf + ((int64_t) (tl.up < 0 ? 0 : tl.up) * (QG_TCW * 3) - (int64_t) ti0 * 3);
  double* rowOut = a.rowbuf + ((int64_t) tl.id * (QG_TCW * 3) - (int64_t) ti0 * 3);
  const double* diagIn = a.rowbuf + (uint64_t) (tl.diag < 0 ? 0 : tl.diag) * (QG_TCW * 3) + (uint64_t) (QG_TCW - 1) * 3;

  double M[QG_TRC], I[QG_TRC], D[QG_TRC];                        // B of reversed row j'-1 (= row j+1) of my columns
  int tokn[QG_TRC], tokc[QG_TRC];
#pragma unroll
  for (int c = 0; c < QG_TRC; ++c) {
    const int ip = i0 + c, jt = j0 - 1;
    const bool have = tl.up >= 0 && ip <= xlen && (ip - jt) >= dlo && (ip - jt) <= dhi;
    M[c] = have ? rowIn[(int64_t) ip * 3] : QG_NEG_INF;
    I[c] = have ? rowIn[(int64_t) ip * 3 + 1] : QG_NEG_INF;
    D[c] = have ? rowIn[(int64_t) ip * 3 + 2] : QG_NEG_INF;
    const int i = xlen + 1 - ip;                                 // the column itself
    tokn[c] = qg_tok (xw, nxw, i);                               // x[i]  : base of the Match destination (i+1, j+1)
    tokc[c] = qg_tok (xw, nxw, i - 1);                           // x[i-1]: base of this column
  }
  double tlM, tlI, tlD;
  {
    const int ip = i0 - 1, jt = j0 - 1;
    const double* src = (lane == 0) ? diagIn : rowIn + (int64_t) ip * 3;
    const bool have = (lane == 0 ? tl.diag >= 0 : tl.up >= 0) && ip >= 1 && jt >= 1 && (ip - jt) >= dlo && (ip - jt) <= dhi;
    tlM = have ? src[0] : QG_NEG_INF; tlI = have ? src[1] : QG_NEG_INF; tlD = have ? src[2] : QG_NEG_INF;
  }
  double l1M = QG_NEG_INF, l1I = QG_NEG_INF, l1D = QG_NEG_INF;
  double pvM = QG_NEG_INF, pvI = QG_NEG_INF, pvD = QG_NEG_INF;
  double s_d2m = 0, s_i2m = 0, s_i2i = 0, s_d2d = 0, s_m2e = 0, s2m[4] = {0, 0, 0, 0};

  const int total = QG_TRH + 31;
  for (int u = 0; u < total; ++u) {
    const int jp = j0 + u - lane;                                // reversed row
    const bool active = (jp >= j0) && (jp <= jend);
    const int j = ylen + 1 - jp;                                 // the row itself
    const int jc = j < 0 ? 0 : (j > ylen + 1 ? ylen + 1 : j);
    const qg_rowp Pc = rp[jc];
    const qg_rowp Pn = rp[jc + 1 > ylen + 1 ? ylen + 1 : jc + 1];
    if (lane == 0) {
      const int ip = i0 - 1;
      const bool have = active && tl.left >= 0 && (ip - jp) >= dlo && (ip - jp) <= dhi;
      l1M = have ? colIn[(int64_t) jp * 3] : QG_NEG_INF;
      l1I = have ? colIn[(int64_t) jp * 3 + 1] : QG_NEG_INF;
      l1D = have ? colIn[(int64_t) jp * 3 + 2] : QG_NEG_INF;
    }
    double l0M = pvM, l0I = pvI, l0D = pvD;
    if (u == lane) { l0M = tlM; l0I = tlI; l0D = tlD; }
    double dgM = l0M;                                            // B_M(i+1, j+1)
    double lfD = l1D;                                            // B_D(i+1, j)
    qg_rowrec rec;
#pragma unroll
    for (int t = 0; t < 4; ++t) rec.cnt[t] = 0;
    rec.ins = rec.m2m = rec.m2i = rec.m2d = 0;
    const int tb = (j - 1) / QG_TRH;                             // Forward row block of this row
#pragma unroll
    for (int c = 0; c < QG_TRC; ++c) {
      const int ip = i0 + c;
      const int i = xlen + 1 - ip;
      const bool ok = active && (ip <= xlen) && (ip - jp) >= dlo && (ip - jp) <= dhi;
      const int tn = tokn[c], tc = tokc[c];
      const double En = qg_sel4 (Pn.e, tn);
      const double upM = M[c], upI = I[c];                       // B(i, j+1)
      const double srcM = dgM, srcI = upI, srcD = lfD;
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
        const int ta = (i - 1) / QG_TCW;
        const long long fid = tab[(uint64_t) ta * ws.nRB + tb];
        const double* fs = a.fstore + (uint64_t) (fid < 0 ? 0 : fid) * QG_TILE_CELLS3 + ((uint64_t) (j - 1 - tb * QG_TRH) * QG_TCW + (uint64_t) (i - 1 - ta * QG_TCW)) * 3;
        const double fM = fs[0], fI = fs[1], fD = fs[2];
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
      if (active && j == 1 && i >= 1) a.endvals[ws.end_off + i] = (ok && (i == 1 || local)) ? qg_sel4 (Pc.e, tc) + BM : QG_NEG_INF;
      dgM = upM;                                                 // this column's old row is the next column's (i+1, j+1)
      lfD = BD;
      if (active) { M[c] = BM; I[c] = BI; D[c] = BD; }
    }
    if (active && zok) {
      double* ra = (double*) &rowacc[j];
#pragma unroll
      for (int t = 0; t < 4; ++t) if (rec.cnt[t] != 0.0) atomicAdd (ra + t, rec.cnt[t]);
      if (rec.ins != 0.0) atomicAdd (ra + 4, rec.ins);
      if (rec.m2m != 0.0) atomicAdd (ra + 5, rec.m2m);
      if (rec.m2i != 0.0) atomicAdd (ra + 6, rec.m2i);
      if (rec.m2d != 0.0) atomicAdd (ra + 7, rec.m2d);
    }
    {
      const double sM = active ? M[QG_TRC - 1] : QG_NEG_INF, sI = active ? I[QG_TRC - 1] : QG_NEG_INF, sD = active ? D[QG_TRC - 1] : QG_NEG_INF;
      const double rM = __shfl_up_sync (QG_FULL_MASK, sM, 1), rI = __shfl_up_sync (QG_FULL_MASK, sI, 1), rD = __shfl_up_sync (QG_FULL_MASK, sD, 1);
      pvM = l1M; pvI = l1I; pvD = l1D;
      if (lane > 0) { l1M = rM; l1I = rI; l1D = rD; }
      if (lane == 31 && active) { colOut[(int64_t) jp * 3] = sM; colOut[(int64_t) jp * 3 + 1] = sI; colOut[(int64_t) jp * 3 + 2] = sD; }
    }
    if (active && jp == jend) {
#pragma unroll
      for (int c = 0; c < QG_TRC; ++c) {
        const int ip = i0 + c;
        if (ip <= xlen) { rowOut[(int64_t) ip * 3] = M[c]; rowOut[(int64_t) ip * 3 + 1] = I[c]; rowOut[(int64_t) ip * 3 + 2] = D[c]; }
      }
    }
  }
  double sc[9] = {s_d2m, s_i2m, s_i2i, s_d2d, s_m2e, s2m[0], s2m[1], s2m[2], s2m[3]};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    double v = sc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync (QG_FULL_MASK, v, o);
    if (lane == 0 && v != 0.0) atomicAdd (&a.seg_scal[12 * (uint64_t) tl.seg + t], v);
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

// Backward result: start = lse(start, E(i,1) + B_M(i,1)) folded over DESCENDING i (qmodel.cpp:1440-1446)
__global__ void qg_wide_backward_finalize_kernel (const qg_wpair* __restrict__ pairs, uint32_t npairs, const qg_wseg* __restrict__ segs,
                                                  const double* __restrict__ endvals, const double* __restrict__ lse, double* __restrict__ result) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_wpair pd = pairs[p];
  double start = QG_NEG_INF;
  for (uint32_t s = pd.seg_end; s-- > pd.seg_begin; ) {
    const qg_wseg ws = segs[s];
    int ilo = ws.dlo + 1, ihi = ws.dhi + 1;                      // row 1: i = d + 1
    if (ilo < 1) ilo = 1;
    if (ihi > (int) ws.xlen) ihi = (int) ws.xlen;
    for (int i = ihi; i >= ilo; --i) start = qg_lse (lse, start, endvals[ws.end_off + i]);
  }
  result[p] = start;
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
