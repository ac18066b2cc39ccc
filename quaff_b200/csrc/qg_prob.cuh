// Probability-space Forward / Backward ("fast" mode of train): the same recurrences as qg_dp.cuh / qg_backward.cuh
// (src/qmodel.cpp:1343-1510) evaluated on probabilities instead of log-probabilities, 9 FP64 FMA/MUL per Forward cell
// instead of four table log-sum-exps.  It is NOT bit-identical to the reference (whose log_sum_exp is a table
// approximation, logsumexp.cpp:84-103); it agrees with it to ~1e-8 relative, well inside the stated tolerances
// (Forward log-likelihood 1e-3 nats per read, counts 1e-4 relative).  The exact log-space kernels remain available
// (qg_set_option QG_OPT_FB_EXACT) and are always used for runs wider than one warp.
//
// Range control, two levels:
//  * per read row j the emissions are divided by exp(s_j), s_j = max(log e[0..3][j], log ins[j]) -- a constant of the
//    read alone, summed exactly in FP64 (S_j); true F(i,j) = f(i,j) exp(S_j), true B(i,j) = b(i,j) exp(S_yLen - S_j),
//    and a posterior count is f * candidate * exp(S_yLen) / Z with no per-row factor left;
//  * per lane a block-floating exponent: the lane's registers hold value * 2^-ex; neighbours exchange (values, ex) and
//    align by exact power-of-two multiplications; a lane renormalises itself when its largest value leaves [2^-300, 2^300].
#ifndef QG_PROB_CUH
#define QG_PROB_CUH
#include "qg_backward.cuh"

// probability-space row parameters: pe[t] = exp(e[t] - s), pins = exp(ins - s), transitions as probabilities
struct QG_ALIGN (16) qg_rowq { double pe[4]; double pins, pm2m, pm2i, pm2d; };

// one thread per row of a job's read, from the log-space rows; also S_j (prefix sums are done by one thread per read:
// 8 k sequential adds, once per read per call)
__global__ void qg_rowq_kernel (const qg_rp_job* __restrict__ jobs, const qg_rowp* __restrict__ rp, qg_rowq* __restrict__ rq, double* __restrict__ rs,
                                double2* __restrict__ rqs) {
  const qg_rp_job jb = jobs[blockIdx.x];
  const int ylen = (int) jb.ylen;
  const qg_rowp* in = rp + jb.rp_off;
  qg_rowq* out = rq + jb.rp_off;
  double2* soa = rqs + 4 * jb.rp_off;                       // structure-of-arrays copy: field pair q of row j at [q * (yLen + 2) + j]
  const int rows = ylen + 2;
  double* s = rs + jb.rp_off;
  for (int j = threadIdx.x; j <= ylen + 1; j += blockDim.x) {
    const qg_rowp r = in[j];
    qg_rowq q;
    if (j >= 1 && j <= ylen) {
      double m = r.ins;
      for (int t = 0; t < 4; ++t) m = (r.e[t] > m) ? r.e[t] : m;
      if (!(m > QG_NEG_INF)) m = 0;
      for (int t = 0; t < 4; ++t) q.pe[t] = exp (r.e[t] - m);
      q.pins = exp (r.ins - m);
      q.pm2m = exp (r.m2m); q.pm2i = exp (r.m2i); q.pm2d = exp (r.m2d);
      s[j] = m;
    } else {
      for (int t = 0; t < 4; ++t) q.pe[t] = 0;
      q.pins = 0; q.pm2i = 0; q.pm2d = 0;
      q.pm2m = (j == 0) ? exp (r.m2m) : 0;                 // row 0 carries m2e
      s[j] = 0;
    }
    out[j] = q;
    soa[j] = make_double2 (q.pe[0], q.pe[1]); soa[rows + j] = make_double2 (q.pe[2], q.pe[3]);
    soa[2 * rows + j] = make_double2 (q.pins, q.pm2m); soa[3 * rows + j] = make_double2 (q.pm2i, q.pm2d);
  }
  __syncthreads ();
  if (threadIdx.x == 0) { double acc = 0; for (int j = 1; j <= ylen; ++j) { acc += s[j]; s[j] = acc; } s[ylen + 1] = acc; }
}

__device__ __forceinline__ double qg_pow2 (int k) {          // exact 2^k, flushing to 0 below the normal range
  if (k < -1022) return 0.0;
  if (k > 1023) k = 1023;
  return __longlong_as_double ((long long) (1023 + k) << 52);
}
__device__ __forceinline__ int qg_exponent (double v) {      // floor(log2 v) for positive normal v
  return (int) ((__double_as_longlong (v) >> 52) & 0x7ff) - 1023;
}

// log_sum_exp in probability space WITH the reference's cut-off: log_sum_exp_unary returns 0 for |a-b| >= 10
// (logsumexp.cpp:88-89), i.e. the smaller term is dropped once it is below e^-10 of the larger one.  Over an 8 kb
// read that truncation is worth ~0.1 nat, far more than the 1e-3 nats parity bar, so it is reproduced here.
#define QG_LSE_CUT 4.5399929762484854e-05                    /* exp(-10) */
__device__ __forceinline__ double qg_psum (double a, double b) {
  // one compare orders the pair (fmax + fmin cost twice as many instructions); a, b >= 0, never NaN
  const bool agtb = a > b;
  const double hi = agtb ? a : b, lo = agtb ? b : a;
  return (lo <= hi * QG_LSE_CUT) ? hi : hi + lo;
}
// high word of a non-negative double: ordered like the value itself (up to the low word), 0 only for 0 / denormals
__device__ __forceinline__ int qg_hi (double v) { return (int) (__double_as_longlong (v) >> 32); }
__device__ __forceinline__ int qg_imax (int a, int b) { return a > b ? a : b; }
#define QG_HI_1EM90 0x2D400000                              /* below ~2^-299 */
#define QG_HI_1EP90 0x52A00000                              /* above ~2^299  */

struct qg_prob_args {
  const qg_segment* segs;
  const uint64_t* xpacked;
  const uint64_t* xpoff;
  const qg_rowq* rq;
  const double2* rqs;                // the same rows as four arrays of double2 per read (32 lanes read 32 consecutive rows: one contiguous request per field pair)
  const double* rs;                  // S_j per read row (same indexing as rq)
  double pi2i, pi2m, pd2d, pd2m;
  int local;
  double* store;                     // Forward: [j][state][slot] mantissas per segment (mode store)
  int* store_ex;                     // Forward: [j][lane] exponents per segment
  double* endvals;                   // per segment slot: end mantissa (Forward) / start mantissa (Backward)
  int* endex;                        // per segment slot: its exponent
  double* rowacc;
  const double* pair_zm;             // Backward: normalised Forward result per pair = zm * 2^ze
  const int* pair_ze;
  double* seg_scal;
  int do_store;
};

__device__ __forceinline__ qg_rowq qg_load_rowq (const double2* __restrict__ soa, int rows, int j) {
  const double2 a = soa[j], b = soa[rows + j], c = soa[2 * rows + j], d = soa[3 * rows + j];
  qg_rowq q;
  q.pe[0] = a.x; q.pe[1] = a.y; q.pe[2] = b.x; q.pe[3] = b.y; q.pins = c.x; q.pm2m = c.y; q.pm2i = d.x; q.pm2d = d.y;
  return q;
}

#define QG_RESCALE_ALL(f) do { _Pragma ("unroll") for (int c_ = 0; c_ < R; ++c_) { M[c_] *= (f); I[c_] *= (f); D[c_] *= (f); } } while (0)

template<int R>
__global__ void __launch_bounds__ (32)
qg_forward_prob_kernel (const qg_prob_args a) {
  const qg_segment sg = a.segs[blockIdx.x];
  const int lane = threadIdx.x;
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, width = (int) sg.width;
  const int SW = 32 * R;
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowq* rq = a.rq + sg.rp_off;
  const double2* rqs = a.rqs + 4 * sg.rp_off;
  const int rows = ylen + 2;
  const double i2i = a.pi2i, i2m = a.pi2m, d2d = a.pd2d, d2m = a.pd2m;
  const bool local = a.local != 0;
  const double m2e = rq[0].pm2m;
  const int s0 = R * lane, d0 = sg.dlo + s0;

  double M[R], I[R], D[R];
#pragma unroll
  for (int c = 0; c < R; ++c) { M[c] = 0; I[c] = 0; D[c] = 0; }
  double lM = 0, lD = 0;
  int ex = 0;
  bool zero = true;                                         // all of M/I/D are zero
  uint64_t win = 0; int pw = 0; bool have_win = false;

  const int total = ylen + 31;
  qg_rowq Pnext = qg_load_rowq (rqs, rows, (1 - lane) < 0 ? 0 : (1 - lane));     // fetched one macro-step ahead
  for (int u = 1; u <= total; ++u) {
    const int j = u - lane;
    const bool active = (j >= 1) && (j <= ylen);
    const qg_rowq P = Pnext;
    { const int jn = j + 1; Pnext = qg_load_rowq (rqs, rows, jn < 0 ? 0 : (jn > ylen + 1 ? ylen + 1 : jn)); }
    const int p0 = d0 + j - 1;
    if (active && (!have_win || p0 < pw || p0 + R > pw + 32)) { win = qg_fetch32 (xw, nxw, p0); pw = p0; have_win = true; }
    const uint64_t wsh = win >> (2 * ((p0 - pw) & 31));
    const bool startRow = (j == 1);
    if (startRow && zero && lM == 0 && lD == 0) ex = 0;
    const double startv = startRow ? qg_pow2 (-ex) : 0.0;
    double rM = 0, rI = 0;

#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1) {
        // ---- exchange 1: first cells (row j-1 of the right neighbour) travel one lane to the left, with their exponent
        const double m0 = M[0], i0 = I[0];
        rM = __shfl_down_sync (QG_FULL_MASK, m0, 1);
        rI = __shfl_down_sync (QG_FULL_MASK, i0, 1);
        int exR = __shfl_down_sync (QG_FULL_MASK, ex, 1);
        if (lane == 31) { rM = 0; rI = 0; }
        if (rM != 0 || rI != 0) {
          int any = 0;
#pragma unroll
          for (int c2 = 0; c2 < R; ++c2) any |= qg_hi (M[c2]) | qg_hi (I[c2]) | qg_hi (D[c2]);
          if (any == 0) ex = exR;
          else if (exR > ex) { const double f = qg_pow2 (ex - exR); QG_RESCALE_ALL (f); ex = exR; }
          const double g = qg_pow2 (exR - ex);
          rM *= g; rI *= g;
        }
      }
      const int i = d0 + c + j;
      const bool ok = active && (s0 + c < width) && (i >= 1) && (i <= xlen);
      const int tk = (int) ((wsh >> (2 * c)) & 3);
      const double E = qg_sel4 (P.pe, tk);
      const double mM = M[c], mI = I[c], mD = D[c];
      const double iM = (c + 1 < R) ? M[(c + 1) % R] : rM;
      const double iI = (c + 1 < R) ? I[(c + 1) % R] : rI;
      const double dM = (c > 0) ? M[(c + R - 1) % R] : lM;
      const double dD = (c > 0) ? D[(c + R - 1) % R] : lD;
      double nM = qg_psum (qg_psum (mM * P.pm2m, mD * d2m), mI * i2m);
      if (startRow && (i == 1 || local)) nM = qg_psum (nM, startv);
      nM *= E;
      double nI = P.pins * qg_psum (iI * i2i, iM * P.pm2i);
      double nD = qg_psum (dD * d2d, dM * P.pm2d);
      if (!ok) { nM = 0; nI = 0; nD = 0; }
      M[c] = nM; I[c] = nI; D[c] = nD;
    }
    // ---- renormalise this lane if its values drifted (largest value by its high word: all values are >= 0)
    int mxh = 0;
#pragma unroll
    for (int c = 0; c < R; ++c) mxh = qg_imax (mxh, qg_imax (qg_hi (M[c]), qg_imax (qg_hi (I[c]), qg_hi (D[c]))));
    zero = (mxh == 0);
    if (!zero && (mxh < QG_HI_1EM90 || mxh > QG_HI_1EP90)) {
      const int k = (mxh >> 20) - 1023;
      const double f = qg_pow2 (-k);
      QG_RESCALE_ALL (f);
      ex += k;
    }
    if (j == ylen) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        const int i = d0 + c + j;
        const bool isEnd = (s0 + c < width) && (i >= 1) && (i <= xlen) && (i == xlen || local);
        a.endvals[sg.aux_off + s0 + c] = isEnd ? M[c] * m2e : 0.0;
        a.endex[sg.aux_off + s0 + c] = ex;
      }
    }
    if (a.do_store && active && s0 < width) {                 // lanes beyond the run hold padding only (a 65-diagonal band uses 22 of 32 lanes): nothing to keep
      // skewed layout [macro-step][lane][state][c] (see qg_dp.cuh); exponents [macro-step][lane]
      double* st = a.store + sg.store_off + ((uint64_t) u * 32 + lane) * (3 * R);
#pragma unroll
      for (int c = 0; c < R; ++c) { st[c] = M[c]; st[R + c] = I[c]; st[2 * R + c] = D[c]; }
      a.store_ex[(sg.acc_off + 32 * sg.seg_id) * 32 + (uint64_t) u * 32 + lane] = ex;
    }
    // ---- exchange 2: last cells (row j) travel one lane to the right, for the next macro-step
    {
      const double mL = M[R - 1], dL = D[R - 1];
      lM = __shfl_up_sync (QG_FULL_MASK, mL, 1);
      lD = __shfl_up_sync (QG_FULL_MASK, dL, 1);
      const int exL = __shfl_up_sync (QG_FULL_MASK, ex, 1);
      if (lane == 0) { lM = 0; lD = 0; }
      if (lM != 0 || lD != 0) {
        if (zero) ex = exL;
        else if (exL > ex) { const double f = qg_pow2 (ex - exL); QG_RESCALE_ALL (f); ex = exL; }
        const double g = qg_pow2 (exL - ex);
        lM *= g; lD *= g;
      }
    }
  }
}

// ---- Forward on an isolated diagonal (a run of width 1, e.g. the always-present diagonal 0 of every pair) ------------
// Its neighbours are halo diagonals, so Insert and Delete stay 0 and Match is one product chain
// M(j) = (M(j-1) * m2m_j) * E_j, started on row 1.  A warp per such run would spend yLen + 31 macro-steps on one cell per
// row; one THREAD does the chain here (same multiplications in the same order, power-of-two renormalisation: the same
// value).  Forward-only calls (the E-step's gate pass); the store for Backward still comes from the warp kernel.
__global__ void __launch_bounds__ (64)
qg_forward_prob_diag_kernel (const qg_prob_args a, uint32_t nseg) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  const qg_segment sg = a.segs[t];
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, dlo = sg.dlo;
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowq* rq = a.rq + sg.rp_off;
  const bool local = a.local != 0;
  const double m2e = rq[0].pm2m;
  double M = 0; int ex = 0;
  uint64_t win = 0; int pw = 0; bool have_win = false;
  double pe_ring[4][4], m2m_ring[4];                        // row parameters four rows ahead (one long dependent chain)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int jr = (1 + k > ylen + 1) ? ylen + 1 : 1 + k;
#pragma unroll
    for (int q = 0; q < 4; ++q) pe_ring[k][q] = rq[jr].pe[q];
    m2m_ring[k] = rq[jr].pm2m;
  }
  for (int j0 = 1; j0 <= ylen; j0 += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + k;
      if (j > ylen) break;
      const double pe[4] = { pe_ring[k][0], pe_ring[k][1], pe_ring[k][2], pe_ring[k][3] };
      const double pm2m = m2m_ring[k];
      {
        const int jr = (j + 4 > ylen + 1) ? ylen + 1 : j + 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) pe_ring[k][q] = rq[jr].pe[q];
        m2m_ring[k] = rq[jr].pm2m;
      }
      const int i = dlo + j, p0 = i - 1;
      if (!have_win || p0 < pw || p0 + 1 > pw + 32) { win = qg_fetch32 (xw, nxw, p0); pw = p0; have_win = true; }
      const int tk = (int) ((win >> (2 * ((p0 - pw) & 31))) & 3);
      const bool ok = (i >= 1) && (i <= xlen);
      double nM = M * pm2m;
      if (j == 1 && (i == 1 || local)) nM = qg_psum (nM, 1.0);             // ex == 0 on the start row
      nM *= qg_sel4 (pe, tk);
      if (!ok) nM = 0;
      M = nM;
      const int mh = qg_hi (M);
      if (mh != 0 && (mh < QG_HI_1EM90 || mh > QG_HI_1EP90)) { const int kk = (mh >> 20) - 1023; M *= qg_pow2 (-kk); ex += kk; }
    }
  }
  const int iend = dlo + ylen;
  const bool isEnd = (iend >= 1) && (iend <= xlen) && (iend == xlen || local);
  a.endvals[sg.aux_off] = isEnd ? M * m2e : 0.0;
  a.endex[sg.aux_off] = ex;
}

// Forward result per pair: log( sum_slot m * 2^e ) + S_yLen; also the normalised sum (zm, ze) for Backward
__global__ void qg_forward_prob_finalize_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                                 const double* __restrict__ endvals, const int* __restrict__ endex, const double* __restrict__ rs,
                                                 double* __restrict__ result, double* __restrict__ zm, int* __restrict__ ze) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  int emax = -(1 << 30);
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_segment sg = segs[s];
    for (uint32_t t = 0; t < sg.width; ++t) if (endvals[sg.aux_off + t] > 0 && endex[sg.aux_off + t] > emax) emax = endex[sg.aux_off + t];
  }
  double sum = 0;
  double S = 0;
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {          // ascending i, as the reference folds `end` (qmodel.cpp:1381-1383)
    const qg_segment sg = segs[s];
    S = rs[sg.rp_off + sg.ylen];
    for (uint32_t t = 0; t < sg.width; ++t) { const double v = endvals[sg.aux_off + t]; if (v > 0) sum = qg_psum (sum, v * qg_pow2 (endex[sg.aux_off + t] - emax)); }
  }
  if (sum > 0) { result[p] = (log (sum) + emax * 0.69314718055994530942) + S; zm[p] = sum; ze[p] = emax; }
  else { result[p] = QG_NEG_INF; zm[p] = 0; ze[p] = 0; }
}

// ---- Backward + counts, probability space, pull form (see qg_backward.cuh for the structure) -------------------------
#ifndef QG_BWD_MINB
#define QG_BWD_MINB 12
#endif
template<int R>
__global__ void __launch_bounds__ (32, QG_BWD_MINB)
qg_backward_prob_kernel (const qg_prob_args a) {
  const qg_segment sg = a.segs[blockIdx.x];
  const int lane = threadIdx.x;
  const int xlen = (int) sg.xlen, ylen = (int) sg.ylen, width = (int) sg.width;
  const int SW = 32 * R;
  const uint64_t* xw = a.xpacked + a.xpoff[sg.xseq];
  const int nxw = (xlen + 31) >> 5;
  const qg_rowq* rq = a.rq + sg.rp_off;
  const double* stbase = a.store + sg.store_off;
  const int* stex = a.store_ex + (sg.acc_off + 32 * sg.seg_id) * 32;
  const double i2i = a.pi2i, i2m = a.pi2m, d2d = a.pd2d, d2m = a.pd2m;
  const bool local = a.local != 0;
  const double m2e = rq[0].pm2m;
  const double zm = a.pair_zm[sg.pair];
  const int ze = a.pair_ze[sg.pair];
  const bool zok = zm > 0;
  const double rzm = zok ? 1.0 / zm : 0.0;
  qg_rowrec* rowacc = (qg_rowrec*) a.rowacc + sg.acc_off;
  const int flane = 31 - lane;                              // the Forward lane that owns my slots
  if (width == 1) {
    // An isolated diagonal is a single path (its neighbours are halo diagonals: Insert and Delete stay 0), so the
    // posterior of every transition on it is the same number w = (run's Forward value) / Z and Backward is closed form:
    // every source row j < yLen counts w for Match->Match in context c(j) and w for the emission of row j+1, the end
    // transition and the start emission count w once.  Next to a real alignment w underflows to exactly 0 (the run is
    // ~e^-9000 of Z; the reference's exp() underflows as well) and nothing is written (the count rows are pre-zeroed).
    // The run's Backward start value is its path probability, which is what the Forward kernel left in slot 0.
    const double fv = a.endvals[sg.aux_off];
    const int fe = a.endex[sg.aux_off];
    const double w = (zok && fv != 0.0) ? (fv * rzm) * qg_pow2 (fe - ze) : 0.0;
    if (w != 0.0) {
      for (int j = 1 + lane; j < ylen; j += 32) {
        qg_rowrec r;
        const int tn = qg_tok (xw, nxw, sg.dlo + j);          // x base of the destination cell (i + 1, j + 1), i = dlo + j
#pragma unroll
        for (int t = 0; t < 4; ++t) r.cnt[t] = (t == tn) ? w : 0.0;
        r.ins = 0; r.m2m = w; r.m2i = 0; r.m2d = 0;
        rowacc[j] = r;
      }
      if (lane == 0) {
        a.seg_scal[12 * sg.seg_id + 4] = w;                                         // m2e
        a.seg_scal[12 * sg.seg_id + 5 + qg_tok (xw, nxw, sg.dlo)] = w;               // start emission: x base of cell (dlo + 1, 1)
      }
    }
    return;
  }
  double M[R], I[R], D[R];                                  // B_M, B_I of row j+1 (then of row j); D[] is this row's B_D chain
#pragma unroll
  for (int c = 0; c < R; ++c) { M[c] = 0; I[c] = 0; D[c] = 0; }
  double lD = 0;
  int ex = 0;
  bool zero = true;
  double s_d2m = 0, s_i2m = 0, s_i2i = 0, s_d2d = 0, s_m2e = 0, s2m[4] = {0, 0, 0, 0};
  qg_rowrec rec;
#pragma unroll
  for (int t = 0; t < 4; ++t) rec.cnt[t] = 0;
  rec.ins = rec.m2m = rec.m2i = rec.m2d = 0;
  qg_rowq Pn = rq[ylen + 1];

  const int total = ylen + 31;
  // the stored Forward record of a macro-step (3R mantissas + exponent, written by the Forward lane that owns my slots at
  // its macro-step yLen + 32 - u) is fetched one macro-step ahead: a warp is one long dependent chain and there are
  // few warps per SM, so an HBM round trip per step would otherwise be fully exposed
  double pf[3 * R]; int pfex;
  const bool fkept = R * flane < width;                       // the Forward lane wrote its records only if it owns real diagonals
  {
    const uint64_t fr = (uint64_t) (ylen + 32 - 1) * 32 + flane;
    const double* s1 = stbase + fr * (3 * R);
#pragma unroll
    for (int t = 0; t < 3 * R; ++t) pf[t] = fkept ? s1[t] : 0.0;
    pfex = fkept ? stex[fr] : 0;
  }
  for (int u = 1; u <= total; ++u) {
    double st[3 * R];
#pragma unroll
    for (int t = 0; t < 3 * R; ++t) st[t] = pf[t];
    const int stex_cur = pfex;
    if (u < total && fkept) {
      const uint64_t fr = (uint64_t) (ylen + 32 - (u + 1)) * 32 + flane;
      const double* s1 = stbase + fr * (3 * R);
#pragma unroll
      for (int t = 0; t < 3 * R; ++t) pf[t] = s1[t];
      pfex = stex[fr];
    }
    const int j = ylen + 1 - (u - lane);
    const bool active = (j >= 1) && (j <= ylen);
    const int jj = j < 0 ? 0 : (j > ylen + 1 ? ylen + 1 : j);
    const qg_rowq Pc = rq[jj];                              // array-of-structures here: the SoA form costs this kernel registers (168, spills) and time (measured 54 -> 58 ms)
    if (j == ylen) { Pn = rq[ylen + 1]; if (zero && lD == 0) ex = 0; }
    {
      qg_rowrec in;
#pragma unroll
      for (int t = 0; t < 4; ++t) in.cnt[t] = __shfl_up_sync (QG_FULL_MASK, rec.cnt[t], 1);
      in.ins = __shfl_up_sync (QG_FULL_MASK, rec.ins, 1);
      in.m2m = __shfl_up_sync (QG_FULL_MASK, rec.m2m, 1);
      in.m2i = __shfl_up_sync (QG_FULL_MASK, rec.m2i, 1);
      in.m2d = __shfl_up_sync (QG_FULL_MASK, rec.m2d, 1);
      if (lane == 0) { for (int t = 0; t < 4; ++t) in.cnt[t] = 0; in.ins = in.m2m = in.m2i = in.m2d = 0; }
      rec = in;
    }
    // the Forward warp wrote this row's cells of my slots at its macro-step j + flane = yLen + 32 - u: one block per step
    const int exF = (active && zok) ? stex_cur : 0;
    double K = rzm * qg_pow2 (exF + ex - ze);               // count = f * candidate * K (true scale)
    double rI = 0;
    double nBM[R], nBI[R];
#pragma unroll
    for (int c = 0; c < R; ++c) {
      if (c == R - 1) {
        const double i0 = nBI[0];
        rI = __shfl_down_sync (QG_FULL_MASK, i0, 1);
        const int exR = __shfl_down_sync (QG_FULL_MASK, ex, 1);
        if (lane == 31) rI = 0;
        if (rI != 0) {
          int any = qg_hi (lD);
#pragma unroll
          for (int c2 = 0; c2 < R; ++c2) { any |= qg_hi (M[c2]) | qg_hi (I[c2]); if (c2 < c) any |= qg_hi (nBM[c2]) | qg_hi (nBI[c2]) | qg_hi (D[c2]); }
          if (any == 0) ex = exR;
          else if (exR > ex) {
            const double f = qg_pow2 (ex - exR);
#pragma unroll
            for (int c2 = 0; c2 < R; ++c2) { M[c2] *= f; I[c2] *= f; if (c2 < c) { nBM[c2] *= f; nBI[c2] *= f; D[c2] *= f; } }
            lD *= f; ex = exR;
          }
          rI *= qg_pow2 (exR - ex);
          K = rzm * qg_pow2 (exF + ex - ze);
        }
      }
      const int s = SW - 1 - (R * lane + c);
      const int i = sg.dlo + s + j;
      const bool ok = active && (s < width) && (i >= 1) && (i <= xlen);
      const int tn = qg_tok (xw, nxw, i);
      const double En = qg_sel4 (Pn.pe, tn);
      const double srcM = M[c];
      const double srcI = (c + 1 < R) ? I[(c + 1) % R] : rI;
      const double srcD = (c > 0) ? D[(c + R - 1) % R] : lD;
      const double eM = En * srcM;
      const double cM = Pn.pm2m * eM, cIM = i2m * eM, cDM = d2m * eM;
      const double eI = Pn.pins * srcI;
      const double cI = Pn.pm2i * eI, cII = i2i * eI;
      const double cD = Pc.pm2d * srcD, cDD = d2d * srcD;
      const bool isEnd = (j == ylen) && (i == xlen || local);
      const double cE = m2e * qg_pow2 (-ex);
      double BM = qg_psum (qg_psum (cM, cI), cD);
      if (isEnd) BM = qg_psum (BM, cE);
      double BI = qg_psum (cIM, cII);
      double BD = qg_psum (cDM, cDD);
      if (!ok) { BM = 0; BI = 0; BD = 0; }
      if (ok && zok) {
        const int fc = R - 1 - c;                           // my mirrored cell c is the Forward lane's cell R-1-c
        const double fM = st[fc] * K, fI = st[R + fc] * K, fD = st[2 * R + fc] * K;
        const double n_m2m = fM * cM, n_i2m = fI * cIM, n_d2m = fD * cDM;
        const double n_m2i = fM * cI, n_i2i = fI * cII;
        const double n_m2d = fM * cD, n_d2d = fD * cDD;
        const double nm = n_m2m + n_d2m + n_i2m;
        rec.cnt[0] += (tn == 0) ? nm : 0.0; rec.cnt[1] += (tn == 1) ? nm : 0.0;
        rec.cnt[2] += (tn == 2) ? nm : 0.0; rec.cnt[3] += (tn == 3) ? nm : 0.0;
        rec.ins += n_m2i + n_i2i;
        rec.m2m += n_m2m; rec.m2i += n_m2i; rec.m2d += n_m2d;
        s_d2m += n_d2m; s_i2m += n_i2m; s_i2i += n_i2i; s_d2d += n_d2d;
        if (isEnd) s_m2e += fM * cE;
        if (j == 1 && (i == 1 || local)) {
          const int tc = qg_tok (xw, nxw, i - 1);
          const double ns = (qg_sel4 (Pc.pe, tc) * BM) * (rzm * qg_pow2 (ex - ze));     // Start has f = 1, exponent 0
          s2m[0] += (tc == 0) ? ns : 0.0; s2m[1] += (tc == 1) ? ns : 0.0;
          s2m[2] += (tc == 2) ? ns : 0.0; s2m[3] += (tc == 3) ? ns : 0.0;
        }
      }
      nBM[c] = BM; nBI[c] = BI; D[c] = BD;
    }
#pragma unroll
    for (int c = 0; c < R; ++c) { M[c] = nBM[c]; I[c] = nBI[c]; }
    int mxh = 0;
#pragma unroll
    for (int c = 0; c < R; ++c) mxh = qg_imax (mxh, qg_imax (qg_hi (M[c]), qg_imax (qg_hi (I[c]), qg_hi (D[c]))));
    zero = (mxh == 0);
    if (!zero && (mxh < QG_HI_1EM90 || mxh > QG_HI_1EP90)) {
      const int k = (mxh >> 20) - 1023;
      const double f = qg_pow2 (-k);
      QG_RESCALE_ALL (f);
      ex += k;
    }
    if (j == 1) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        const int s = SW - 1 - (R * lane + c);
        const int i = sg.dlo + s + j;
        const bool ok = active && (s < width) && (i >= 1) && (i <= xlen) && (i == 1 || local);
        a.endvals[sg.aux_off + s] = ok ? qg_sel4 (Pc.pe, qg_tok (xw, nxw, i - 1)) * M[c] : 0.0;
        a.endex[sg.aux_off + s] = ex;
      }
    }
    {
      const double dL = D[R - 1];
      lD = __shfl_up_sync (QG_FULL_MASK, dL, 1);
      const int exL = __shfl_up_sync (QG_FULL_MASK, ex, 1);
      if (lane == 0) lD = 0;
      if (lD != 0) {
        if (zero) ex = exL;
        else if (exL > ex) { const double f = qg_pow2 (ex - exL); QG_RESCALE_ALL (f); ex = exL; }
        lD *= qg_pow2 (exL - ex);
      }
    }
    if (lane == 31 && active) rowacc[j] = rec;
    Pn = Pc;
  }
  double sc[9] = {s_d2m, s_i2m, s_i2i, s_d2d, s_m2e, s2m[0], s2m[1], s2m[2], s2m[3]};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    double v = sc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync (QG_FULL_MASK, v, o);
    if (lane == 0 && v != 0.0) atomicAdd (&a.seg_scal[12 * sg.seg_id + t], v);
  }
}

// Backward result per pair: log( sum over row-1 start cells ) + S_yLen
__global__ void qg_backward_prob_finalize_kernel (const qg_pair_dp* __restrict__ pairs, uint32_t npairs, const qg_segment* __restrict__ segs,
                                                  const double* __restrict__ endvals, const int* __restrict__ endex, const double* __restrict__ rs,
                                                  double* __restrict__ result) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const qg_pair_dp pd = pairs[p];
  int emax = -(1 << 30);
  for (uint32_t s = pd.seg_begin; s < pd.seg_end; ++s) {
    const qg_segment sg = segs[s];
    for (uint32_t t = 0; t < sg.width; ++t) if (endvals[sg.aux_off + t] > 0 && endex[sg.aux_off + t] > emax) emax = endex[sg.aux_off + t];
  }
  double sum = 0, S = 0;
  for (uint32_t s = pd.seg_end; s-- > pd.seg_begin; ) {           // descending i, as the reference folds `start` (qmodel.cpp:1440-1446)
    const qg_segment sg = segs[s];
    S = rs[sg.rp_off + sg.ylen];
    for (uint32_t t = sg.width; t-- > 0; ) { const double v = endvals[sg.aux_off + t]; if (v > 0) sum = qg_psum (sum, v * qg_pow2 (endex[sg.aux_off + t] - emax)); }
  }
  result[p] = (sum > 0) ? (log (sum) + emax * 0.69314718055994530942) + S : QG_NEG_INF;
}

#endif
