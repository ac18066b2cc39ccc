/* TEST INFRASTRUCTURE ONLY -- see quaff_oracle.h.  Plain-C restatement of ihh/quaff's banded
 * pair-HMM hot path; checker for the CUDA library, never part of the product path.
 * All file:line citations are relative to /root/reference. */
#define _GNU_SOURCE
#include "quaff_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NEG_INF (-INFINITY)

void qo_free (void* p) { free (p); }

static uint64_t ipow4 (int k) { uint64_t n = 1; while (k-- > 0) n *= 4; return n; }   /* fastseq.cpp:37-42 */

/* ------------------------------------------------------------------------------------------
 * log-sum-exp: logsumexp.cpp:7-28 (table), :34-59 (binary), :84-103 (unary, interpolated)
 * ---------------------------------------------------------------------------------------- */
#define LSE_MAX 10
#define LSE_PRECISION .0001
#define LSE_ENTRIES (((int) (LSE_MAX / LSE_PRECISION)) + 1)
static double* lse_table = NULL;

static void lse_init (void) {
  if (lse_table) return;
  double* t = (double*) malloc (sizeof(double) * LSE_ENTRIES);
  for (int n = 0; n < LSE_ENTRIES; ++n) {
    const double x = n * LSE_PRECISION;
    t[n] = log (1. + exp (-x));                       /* log_sum_exp_unary_slow, logsumexp.cpp:105-107 */
  }
  lse_table = t;
}

const double* qo_lse_table (int* n_entries) {
  lse_init();
  if (n_entries) *n_entries = LSE_ENTRIES;
  return lse_table;
}

double qo_lse_unary (double x) {
  lse_init();
  if (x >= LSE_MAX || isnan (x) || isinf (x)) return 0;
  if (x < 0) return -x;
  const int n = (int) (x / LSE_PRECISION);
  const double dx = x - (n * LSE_PRECISION);
  const double f0 = lse_table[n], f1 = lse_table[n+1];
  const double df = f1 - f0;
  return f0 + df * (dx / LSE_PRECISION);
}

double qo_lse (double a, double b) {
  double mx, diff;
  if (a == b) { mx = a; diff = 0; }
  else if (a < b) { mx = b; diff = b - a; }
  else { mx = a; diff = a - b; }
  return mx + qo_lse_unary (diff);
}
static double lse3 (double a, double b, double c) { return qo_lse (qo_lse (a, b), c); }   /* logsumexp.cpp:61-63 */

/* ------------------------------------------------------------------------------------------
 * sequence encodings: fastseq.cpp:27-35 (makeKmer), :85-99 (kmers)
 * ---------------------------------------------------------------------------------------- */
static uint64_t make_kmer (int k, const uint8_t* tok) {
  uint64_t kmer = 0, mul = 1;
  for (int j = 0; j < k; ++j) { kmer += mul * tok[k - j - 1]; mul *= 4; }
  return kmer;
}

void qo_context_kmers (const uint8_t* tok, uint32_t len, int k, uint32_t* out) {
  if (k == 0) { for (uint32_t p = 0; p < len; ++p) out[p] = 0; return; }
  int count[4] = {0,0,0,0};
  for (uint32_t p = 0; p < len; ++p) ++count[tok[p]];
  int best = 0;
  for (int t = 1; t < 4; ++t) if (count[t] > count[best]) best = t;   /* max_element: first max wins */
  uint8_t* padded = (uint8_t*) malloc (len + k);
  for (int p = 0; p < k - 1; ++p) padded[p] = (uint8_t) best;
  memcpy (padded + k - 1, tok, len);
  for (uint32_t p = 0; p < len; ++p) out[p] = (uint32_t) make_kmer (k, padded + p);
  free (padded);
}

/* ------------------------------------------------------------------------------------------
 * model tables: negbinom.cpp:30-32, qmodel.cpp:87-93 (SymQualScores), :296-325 (QuaffScores)
 * ---------------------------------------------------------------------------------------- */
double qo_log_negbinom (int k, double p, double r) {
  /* gsl_ran_negative_binomial_pdf(k,p,n) = exp(lnG(k+n) - lnG(n) - lnG(k+1)) p^n (1-p)^k */
  const double f = lgamma (k + r), a = lgamma (r), b = lgamma (k + 1.0);
  const double pdf = exp (f - a - b) * pow (p, r) * pow (1 - p, (double) k);
  return log (pdf);
}

static void symqual_scores (const qo_symqual* d, double* out95) {
  const double lsp = log (d->p);
  for (int k = 0; k < QO_NQUAL; ++k) out95[k] = lsp + qo_log_negbinom (k, d->q, d->r);
  out95[QO_NQUAL] = lsp;
}

void qo_scores_from_params (const qo_params* qp, qo_scores* out) {
  const uint64_t nK = ipow4 (qp->match_k), nG = ipow4 (qp->gap_k);
  out->match_k = qp->match_k;
  out->gap_k = qp->gap_k;
  for (int i = 0; i < 4; ++i) {
    symqual_scores (&qp->insert[i], out->insert + i * QO_NQ1);
    for (uint64_t j = 0; j < nK; ++j)
      symqual_scores (&qp->match[i * nK + j], out->match + (i * nK + j) * QO_NQ1);
  }
  for (uint64_t j = 0; j < nG; ++j) {
    out->m2m[j] = log (1 - qp->begin_insert[j]) + log (1 - qp->begin_delete[j]);
    out->m2i[j] = log (qp->begin_insert[j]);
    out->m2d[j] = log (1 - qp->begin_insert[j]) + log (qp->begin_delete[j]);
    out->m2e[j] = log (qp->begin_insert[j]);          /* sic, qmodel.cpp:317 */
  }
  out->d2d = log (qp->extend_delete);
  out->d2m = log (1 - qp->extend_delete);
  out->i2i = log (qp->extend_insert);
  out->i2m = log (1 - qp->extend_insert);
}

double qo_null_loglike (const qo_null_params* np, const qo_seq* s) {   /* qmodel.cpp:1875-1890 */
  double ll = s->len * log (np->null_emit) + log (1. - np->null_emit);
  for (uint32_t i = 0; i < s->len; ++i) {
    ll += log (np->null[s->tok[i]].p);
    if (s->qual)
      ll += qo_log_negbinom (s->qual[i], np->null[s->tok[i]].q, np->null[s->tok[i]].r);
  }
  return ll;
}

/* ------------------------------------------------------------------------------------------
 * envelope: fastseq.cpp:240-256 (KmerIndex), diagenv.cpp:11-18 (initFull), :20-106 (initSparse),
 * :108-133 (initStorage)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t xLen, yLen;
  int nd; int32_t* diags;       /* envelope diagonals, ascending */
  int ns; int32_t* sdiags;      /* storage diagonals = envelope +-1, ascending (diagenv.cpp:109-115) */
  int32_t* rank;                /* rank[yLen + d] = index into sdiags or -1   (storageIndex, :116-118) */
} env_t;

static void env_free (env_t* e) { free (e->diags); free (e->sdiags); free (e->rank); }

static void env_init_storage (env_t* e) {
  const int64_t span = (int64_t) e->xLen + e->yLen + 1;
  e->rank = (int32_t*) malloc (sizeof(int32_t) * span);
  uint8_t* mark = (uint8_t*) calloc (span + 2, 1);
  for (int n = 0; n < e->nd; ++n)
    for (int t = -1; t <= 1; ++t) {
      const int64_t idx = (int64_t) e->yLen + e->diags[n] + t;
      if (idx >= 0 && idx < span) mark[idx] = 1;
    }
  e->ns = 0;
  for (int64_t idx = 0; idx < span; ++idx) e->ns += mark[idx];
  e->sdiags = (int32_t*) malloc (sizeof(int32_t) * (e->ns + 1));
  int r = 0;
  for (int64_t idx = 0; idx < span; ++idx) {
    if (mark[idx]) { e->sdiags[r] = (int32_t) (idx - e->yLen); e->rank[idx] = r++; }
    else e->rank[idx] = -1;
  }
  free (mark);
}

static void env_init_full (env_t* e) {
  e->nd = (int) (e->xLen + e->yLen - 1);
  e->diags = (int32_t*) malloc (sizeof(int32_t) * (e->nd + 1));
  for (int n = 0; n < e->nd; ++n) e->diags[n] = 1 - (int32_t) e->yLen + n;
  env_init_storage (e);
}

typedef struct { uint64_t kmer; uint32_t pos; } kmer_pos;
static int cmp_kmer_pos (const void* a, const void* b) {
  const kmer_pos *p = (const kmer_pos*) a, *q = (const kmer_pos*) b;
  if (p->kmer != q->kmer) return p->kmer < q->kmer ? -1 : 1;
  return p->pos < q->pos ? -1 : (p->pos > q->pos ? 1 : 0);
}
typedef struct { uint32_t count; int32_t diag; } count_diag;
static int cmp_count_desc (const void* a, const void* b) {
  const count_diag *p = (const count_diag*) a, *q = (const count_diag*) b;
  if (p->count != q->count) return p->count > q->count ? -1 : 1;
  return p->diag < q->diag ? -1 : (p->diag > q->diag ? 1 : 0);
}

static int env_init_sparse (env_t* e, const qo_seq* x, const qo_seq* y, const qo_config* cfg, uint64_t cellSize) {
  const uint32_t k = (uint32_t) cfg->kmer_len;
  const uint32_t xLen = e->xLen, yLen = e->yLen;
  if (cfg->kmer_threshold >= 0) {                                   /* diagenv.cpp:23-29 */
    const uint32_t minLen = 2 * (k + (uint32_t) cfg->kmer_threshold);
    if (xLen < minLen || yLen < minLen) { env_init_full (e); return 0; }
  }
  if (xLen < k || yLen < k) return -1;     /* reference underflows here (fastseq.cpp:247, diagenv.cpp:35) */

  /* KmerIndex of the read y: every k-mer start j in [0, yLen-k]  (fastseq.cpp:247-248) */
  const uint32_t ny = yLen - k + 1;
  kmer_pos* idx = (kmer_pos*) malloc (sizeof(kmer_pos) * ny);
  for (uint32_t j = 0; j < ny; ++j) { idx[j].kmer = make_kmer ((int) k, y->tok + j); idx[j].pos = j; }
  qsort (idx, ny, sizeof(kmer_pos), cmp_kmer_pos);

  /* diagKmerCount[i-j]++ for every matching (i,j)  (diagenv.cpp:33-40) */
  const int64_t span = (int64_t) xLen + yLen + 1;
  uint32_t* cnt = (uint32_t*) calloc (span, sizeof(uint32_t));      /* index yLen + d */
  for (uint32_t i = 0; i + k <= xLen; ++i) {
    const uint64_t km = make_kmer ((int) k, x->tok + i);
    uint32_t lo = 0, hi = ny;
    while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (idx[mid].kmer < km) lo = mid + 1; else hi = mid; }
    for (uint32_t t = lo; t < ny && idx[t].kmer == km; ++t)
      ++cnt[(int64_t) yLen + (int64_t) i - (int64_t) idx[t].pos];
  }
  free (idx);

  /* countDistrib: diagonals grouped by count, visited in descending count (diagenv.cpp:42-44, 68-70) */
  size_t nz = 0;
  for (int64_t t = 0; t < span; ++t) nz += cnt[t] != 0;
  count_diag* cd = (count_diag*) malloc (sizeof(count_diag) * (nz + 1));
  nz = 0;
  for (int64_t t = 0; t < span; ++t)
    if (cnt[t]) { cd[nz].count = cnt[t]; cd[nz].diag = (int32_t) (t - yLen); ++nz; }
  free (cnt);
  qsort (cd, nz, sizeof(count_diag), cmp_count_desc);

  uint8_t* inDiags = (uint8_t*) calloc (span + 2, 1);      /* index yLen + d */
  uint8_t* inStorage = (uint8_t*) calloc (span + 4, 1);    /* index yLen + d + 1 (d may reach minDiag-1.. maxDiag+1) */
  size_t nStorage = 0;
  inDiags[yLen] = 1;                                        /* diags.insert(0), :52-54 */
  inStorage[yLen + 1] = 1; nStorage = 1;                    /* storageDiags.insert(0) */
  const int half = (int) ((unsigned int) cfg->band_size / 2);
  const uint64_t diagSize = (uint64_t) (xLen < yLen ? xLen : yLen) * cellSize;
  const int minDiag = 1 - (int) yLen, maxDiag = (int) xLen - 1;

  size_t t0 = 0;
  while (t0 < nz) {
    const uint32_t tierCount = cd[t0].count;
    size_t t1 = t0;
    while (t1 < nz && cd[t1].count == tierCount) ++t1;
    if (cfg->kmer_threshold >= 0 && tierCount < (uint32_t) cfg->kmer_threshold) break;   /* :72-73 */
    /* trial union (moreDiags / moreStorageDiags, :75-85) */
    size_t moreStorage = nStorage;
    for (size_t t = t0; t < t1; ++t) {
      const int seed = cd[t].diag;
      const int dMin = (minDiag > seed - half) ? minDiag : seed - half;
      const int dMax = (maxDiag < seed + half) ? maxDiag : seed + half;
      for (int d = dMin - 1; d <= dMax + 1; ++d)
        if (!inStorage[(int64_t) yLen + d + 1]) { inStorage[(int64_t) yLen + d + 1] = 2; ++moreStorage; }
    }
    if (cfg->kmer_threshold < 0 && moreStorage * diagSize >= cfg->max_size) break;       /* :87-89 */
    for (size_t t = t0; t < t1; ++t) {                      /* accept the tier (:92-94) */
      const int seed = cd[t].diag;
      const int dMin = (minDiag > seed - half) ? minDiag : seed - half;
      const int dMax = (maxDiag < seed + half) ? maxDiag : seed + half;
      for (int d = dMin; d <= dMax; ++d) inDiags[(int64_t) yLen + d] = 1;
      for (int d = dMin - 1; d <= dMax + 1; ++d) inStorage[(int64_t) yLen + d + 1] = 1;
    }
    nStorage = moreStorage;
    t0 = t1;
  }
  free (cd);
  free (inStorage);

  e->nd = 0;
  for (int64_t t = 0; t < span; ++t) e->nd += inDiags[t];
  e->diags = (int32_t*) malloc (sizeof(int32_t) * (e->nd + 1));
  int n = 0;
  for (int64_t t = 0; t < span; ++t) if (inDiags[t]) e->diags[n++] = (int32_t) (t - yLen);   /* :104 */
  free (inDiags);
  env_init_storage (e);
  return 0;
}

static int make_envelope (env_t* e, const qo_seq* x, const qo_seq* y, const qo_config* cfg, uint64_t cellSize) {
  memset (e, 0, sizeof(*e));                                /* qmodel.cpp:1049-1056 */
  e->xLen = x->len; e->yLen = y->len;
  if (cfg->sparse) return env_init_sparse (e, x, y, cfg, cellSize);
  env_init_full (e);
  return 0;
}

/* row iteration bounds: envelope diagonals with 1 <= d + j <= xLen (diagenv.h:75-85) */
static void row_range (const env_t* e, uint32_t j, int* b, int* en) {
  int lo = 0, hi = e->nd;                                   /* upper_bound(diags, -j) */
  while (lo < hi) { const int mid = (lo + hi) / 2; if (e->diags[mid] <= -(int) j) lo = mid + 1; else hi = mid; }
  *b = lo;
  lo = 0; hi = e->nd;                                       /* upper_bound(diags, xLen - j) */
  while (lo < hi) { const int mid = (lo + hi) / 2; if (e->diags[mid] <= (int) e->xLen - (int) j) lo = mid + 1; else hi = mid; }
  *en = lo;
}

static uint64_t env_cell_updates (const env_t* e) {
  uint64_t cu = 0;
  for (uint32_t j = 1; j <= e->yLen; ++j) { int b, en; row_range (e, j, &b, &en); cu += (uint64_t) (en - b); }
  return cu;
}

int qo_envelope (const qo_seq* x, const qo_seq* y, const qo_config* cfg, uint64_t cell_size,
                 int32_t** diags, uint64_t* cell_updates) {
  env_t e;
  if (make_envelope (&e, x, y, cfg, cell_size) != 0) return -1;
  *diags = (int32_t*) malloc (sizeof(int32_t) * (e.nd + 1));
  memcpy (*diags, e.diags, sizeof(int32_t) * e.nd);
  if (cell_updates) *cell_updates = env_cell_updates (&e);
  const int nd = e.nd;
  env_free (&e);
  return nd;
}

/* ------------------------------------------------------------------------------------------
 * DP matrix container: qmodel.h:357-387, qmodel.cpp:1243-1253 (all cells -inf), :1308-1324
 * Storage here is rectangular (storage diagonal rank x row), which addresses a superset of the
 * reference's compact storage; cells the reference never touches stay -inf in both.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const env_t* env;
  double* cell;                  /* [ns][yLen+1][3] */
  double start, end, result;
} dpm_t;

static void dpm_init (dpm_t* m, const env_t* e) {
  m->env = e;
  const size_t n = (size_t) e->ns * (e->yLen + 1) * 3;
  m->cell = (double*) malloc (sizeof(double) * (n + 1));
  for (size_t t = 0; t < n; ++t) m->cell[t] = NEG_INF;
  m->start = m->end = m->result = NEG_INF;
}
static void dpm_free (dpm_t* m) { free (m->cell); }

static double dummy_cell;
static inline double* cellp (const dpm_t* m, int64_t i, int64_t j, int state) {
  const env_t* e = m->env;
  const int64_t idx = (int64_t) e->yLen + i - j;
  if (i < 0 || i > (int64_t) e->xLen || j < 0 || j > (int64_t) e->yLen || idx < 0 || idx > (int64_t) e->xLen + e->yLen || e->rank[idx] < 0) {
    dummy_cell = NEG_INF;                                   /* getStorageIndexSafe -> dummy, qmodel.h:372-375 */
    return &dummy_cell;
  }
  return m->cell + ((size_t) e->rank[idx] * (e->yLen + 1) + (size_t) j) * 3 + state;
}
#define MAT(m,i,j) (*cellp (m, i, j, 0))
#define INS(m,i,j) (*cellp (m, i, j, 1))
#define DEL(m,i,j) (*cellp (m, i, j, 2))

static double* dump_cells (const dpm_t* m, uint64_t* n_cells) {
  const env_t* e = m->env;
  const uint64_t cu = env_cell_updates (e);
  double* out = (double*) malloc (sizeof(double) * (cu * 3 + 1));
  uint64_t n = 0;
  for (uint32_t j = 1; j <= e->yLen; ++j) {
    int b, en; row_range (e, j, &b, &en);
    for (int t = b; t < en; ++t) {
      const int64_t i = (int64_t) e->diags[t] + j;
      out[n++] = MAT (m, i, j); out[n++] = INS (m, i, j); out[n++] = DEL (m, i, j);
    }
  }
  *n_cells = cu;
  return out;
}

/* per-pair context arrays: qmodel.cpp:1308-1324, accessors qmodel.h:399-414 */
typedef struct {
  const qo_scores* qs;
  const qo_seq *x, *y;
  uint32_t *yMatchKmer, *yIndelKmer;   /* yIndelKmer padded with a leading 0 entry (qmodel.cpp:1323) */
  double* insEmit;                     /* cachedInsertEmitScore[j], j = 1..yLen */
  uint64_t nK;
} ctx_t;

static void ctx_init (ctx_t* c, const qo_seq* x, const qo_seq* y, const qo_scores* qs) {
  c->qs = qs; c->x = x; c->y = y;
  c->nK = ipow4 (qs->match_k);
  c->yMatchKmer = (uint32_t*) malloc (sizeof(uint32_t) * (y->len + 1));
  c->yIndelKmer = (uint32_t*) malloc (sizeof(uint32_t) * (y->len + 2));
  qo_context_kmers (y->tok, y->len, qs->match_k, c->yMatchKmer);
  c->yIndelKmer[0] = 0;
  qo_context_kmers (y->tok, y->len, qs->gap_k, c->yIndelKmer + 1);
  c->insEmit = (double*) malloc (sizeof(double) * (y->len + 2));
  c->insEmit[0] = NEG_INF;
  for (uint32_t j = 1; j <= y->len; ++j)
    c->insEmit[j] = qs->insert[y->tok[j-1] * QO_NQ1 + (y->qual ? y->qual[j-1] : QO_NQUAL)];
}
static void ctx_free (ctx_t* c) { free (c->yMatchKmer); free (c->yIndelKmer); free (c->insEmit); }
static inline double matchEmit (const ctx_t* c, int64_t i, int64_t j) {
  return c->qs->match[((size_t) c->x->tok[i-1] * c->nK + c->yMatchKmer[j-1]) * QO_NQ1 + (c->y->qual ? c->y->qual[j-1] : QO_NQUAL)];
}
#define M2M(c,j) ((c)->qs->m2m[(c)->yIndelKmer[j]])
#define M2I(c,j) ((c)->qs->m2i[(c)->yIndelKmer[j]])
#define M2D(c,j) ((c)->qs->m2d[(c)->yIndelKmer[j]])
#define M2E(c,j) ((c)->qs->m2e[(c)->yIndelKmer[j]])

/* ------------------------------------------------------------------------------------------
 * Viterbi fill (qmodel.cpp:1512-1560) and traceback (qmodel.cpp:1562-1646)
 * ---------------------------------------------------------------------------------------- */
static void viterbi_fill (dpm_t* m, const ctx_t* c, const qo_config* cfg) {
  const env_t* e = m->env;
  const uint32_t xLen = e->xLen, yLen = e->yLen;
  const qo_scores* qs = c->qs;
  m->start = 0;
  for (uint32_t j = 1; j <= yLen; ++j) {
    int b, en; row_range (e, j, &b, &en);
    for (int t = b; t < en; ++t) {
      const int64_t i = (int64_t) e->diags[t] + j;
      double mat = fmax (fmax (MAT (m, i-1, j-1) + M2M (c, j-1), DEL (m, i-1, j-1) + qs->d2m), INS (m, i-1, j-1) + qs->i2m);
      if (j == 1 && (i == 1 || cfg->local)) mat = fmax (mat, m->start);
      mat += matchEmit (c, i, j);
      MAT (m, i, j) = mat;
      INS (m, i, j) = c->insEmit[j] + fmax (INS (m, i, j-1) + qs->i2i, MAT (m, i, j-1) + M2I (c, j-1));
      DEL (m, i, j) = fmax (DEL (m, i-1, j) + qs->d2d, MAT (m, i-1, j) + M2D (c, j));
      if (j == yLen && (i == xLen || cfg->local)) m->end = fmax (m->end, MAT (m, i, j) + M2E (c, j));
    }
  }
  m->result = m->end;
}

enum { ST_START = 0, ST_MATCH = 1, ST_INSERT = 2, ST_DELETE = 3 };
#define UPDATE_MAX(cur, curIdx, cand, candIdx) do { const double cand_ = (cand); if (cand_ > cur) { cur = cand_; curIdx = candIdx; } } while (0)   /* qmodel.cpp:1294-1299 */

int qo_viterbi (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                double* result, uint32_t* x_start, uint32_t* x_end,
                uint8_t** path, uint32_t* path_len, double** cells, uint64_t* n_cells) {
  env_t e;
  if (make_envelope (&e, x, y, cfg, 24) != 0) return -1;       /* cellSize() = 3 doubles, qmodel.h:386 */
  dpm_t m; dpm_init (&m, &e);
  ctx_t c; ctx_init (&c, x, y, qs);
  viterbi_fill (&m, &c, cfg);
  *result = m.result;
  if (cells) *cells = dump_cells (&m, n_cells);
  if (path) { *path = NULL; *path_len = 0; }
  if (x_start) { *x_start = 0; *x_end = 0; }
  int rc = 0;
  if (path && m.result > NEG_INF) {
    const uint32_t xLen = e.xLen, yLen = e.yLen;
    uint32_t xEnd = xLen;
    if (cfg->local) {                                           /* qmodel.cpp:1565-1575 */
      double best = NEG_INF;
      for (uint32_t iEnd = xLen; iEnd > 0; --iEnd) {
        const double sc = MAT (&m, iEnd, yLen) + M2E (&c, yLen);
        if (iEnd == xLen || sc > best) { best = sc; xEnd = iEnd; }
      }
    }
    int64_t i = xEnd, j = yLen;
    uint8_t* rev = (uint8_t*) malloc ((size_t) xLen + yLen + 2);
    uint32_t n = 0;
    int state = ST_MATCH;
    while (state != ST_START) {                                 /* qmodel.cpp:1579-1622 */
      double src = NEG_INF, emit;
      switch (state) {
      case ST_MATCH:
        emit = matchEmit (&c, i, j);
        --i; --j;
        rev[n++] = 0;
        UPDATE_MAX (src, state, MAT (&m, i, j) + M2M (&c, j) + emit, ST_MATCH);
        UPDATE_MAX (src, state, INS (&m, i, j) + qs->i2m + emit, ST_INSERT);
        UPDATE_MAX (src, state, DEL (&m, i, j) + qs->d2m + emit, ST_DELETE);
        if (j == 0 && (i == 0 || cfg->local)) UPDATE_MAX (src, state, emit, ST_START);
        if (!(src == MAT (&m, i+1, j+1))) rc = -2;             /* Assert "Traceback error", :1594 */
        break;
      case ST_INSERT:
        emit = c.insEmit[j];
        --j;
        rev[n++] = 1;
        UPDATE_MAX (src, state, MAT (&m, i, j) + M2I (&c, j) + emit, ST_MATCH);
        UPDATE_MAX (src, state, INS (&m, i, j) + qs->i2i + emit, ST_INSERT);
        if (!(src == INS (&m, i, j+1))) rc = -2;
        break;
      default:
        --i;
        rev[n++] = 2;
        UPDATE_MAX (src, state, MAT (&m, i, j) + M2D (&c, j), ST_MATCH);
        UPDATE_MAX (src, state, DEL (&m, i, j) + qs->d2d, ST_DELETE);
        if (!(src == DEL (&m, i+1, j))) rc = -2;
        break;
      }
      if (rc != 0 || n > xLen + yLen) { rc = -2; break; }
    }
    *x_start = (uint32_t) (i + 1);
    *x_end = xEnd;
    *path = (uint8_t*) malloc (n + 1);
    for (uint32_t t = 0; t < n; ++t) (*path)[t] = rev[n - 1 - t];
    *path_len = n;
    free (rev);
  }
  ctx_free (&c); dpm_free (&m); env_free (&e);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * Forward (qmodel.cpp:1343-1391)
 * ---------------------------------------------------------------------------------------- */
static void forward_fill (dpm_t* m, const ctx_t* c, const qo_config* cfg) {
  const env_t* e = m->env;
  const uint32_t xLen = e->xLen, yLen = e->yLen;
  const qo_scores* qs = c->qs;
  m->start = 0;
  for (uint32_t j = 1; j <= yLen; ++j) {
    int b, en; row_range (e, j, &b, &en);
    for (int t = b; t < en; ++t) {
      const int64_t i = (int64_t) e->diags[t] + j;
      double mat = lse3 (MAT (m, i-1, j-1) + M2M (c, j-1), DEL (m, i-1, j-1) + qs->d2m, INS (m, i-1, j-1) + qs->i2m);
      if (j == 1 && (i == 1 || cfg->local)) mat = qo_lse (mat, m->start);
      mat += matchEmit (c, i, j);
      MAT (m, i, j) = mat;
      INS (m, i, j) = c->insEmit[j] + qo_lse (INS (m, i, j-1) + qs->i2i, MAT (m, i, j-1) + M2I (c, j-1));
      DEL (m, i, j) = qo_lse (DEL (m, i-1, j) + qs->d2d, MAT (m, i-1, j) + M2D (c, j));
      if (j == yLen && (i == xLen || cfg->local)) m->end = qo_lse (m->end, MAT (m, i, yLen) + M2E (c, yLen));
    }
  }
  m->result = m->end;
}

int qo_forward (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                double* result, double** cells, uint64_t* n_cells) {
  env_t e;
  if (make_envelope (&e, x, y, cfg, 48) != 0) return -1;       /* 2*cellSize(), qmodel.cpp:2249 */
  dpm_t m; dpm_init (&m, &e);
  ctx_t c; ctx_init (&c, x, y, qs);
  forward_fill (&m, &c, cfg);
  *result = m.result;
  if (cells) *cells = dump_cells (&m, n_cells);
  ctx_free (&c); dpm_free (&m); env_free (&e);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Backward + E-step counts (qmodel.cpp:1393-1503), transCount (qmodel.cpp:1505-1510)
 * ---------------------------------------------------------------------------------------- */
size_t qo_counts_size (int match_k, int gap_k) {
  return 4 * ipow4 (match_k) * QO_NQUAL + 4 * QO_NQUAL + 4 * ipow4 (gap_k) + 4;
}

typedef struct {
  double *match, *insert, *m2m, *m2i, *m2d, *m2e, *scal;   /* views into the flat buffer; scal = d2d,d2m,i2i,i2m */
} counts_view;

static void counts_view_init (counts_view* v, double* flat, int match_k, int gap_k) {
  const uint64_t nK = ipow4 (match_k), nG = ipow4 (gap_k);
  v->match = flat;
  v->insert = v->match + 4 * nK * QO_NQUAL;
  v->m2m = v->insert + 4 * QO_NQUAL;
  v->m2i = v->m2m + nG;
  v->m2d = v->m2i + nG;
  v->m2e = v->m2d + nG;
  v->scal = v->m2e + nG;
}

static double trans_count (double* backSrc, double fwdSrc, double trans, double backDest, double Z) {
  const double transBackDest = trans + backDest;
  const double count = exp (fwdSrc + transBackDest - Z);
  *backSrc = qo_lse (*backSrc, transBackDest);
  return count;
}

static int backward_fill (dpm_t* bm, const dpm_t* fm, const ctx_t* c, const qo_config* cfg, double* counts_flat) {
  const env_t* e = bm->env;
  const uint32_t xLen = e->xLen, yLen = e->yLen;
  const qo_scores* qs = c->qs;
  if (!c->y->qual) return -3;                                   /* Require hasQual, qmodel.cpp:1398 */
  counts_view cv; counts_view_init (&cv, counts_flat, qs->match_k, qs->gap_k);
  memset (counts_flat, 0, sizeof(double) * qo_counts_size (qs->match_k, qs->gap_k));
  const double Z = fm->result;
  const double fwdStart = 0;
  bm->end = 0;
  for (uint32_t j = yLen; j > 0; --j) {
    int b, en; row_range (e, j, &b, &en);
    for (int t = en - 1; t >= b; --t) {                         /* reverse iterator, diagenv.h:107-123 */
      const int64_t i = (int64_t) e->diags[t] + j;
      if (j == yLen && (i == xLen || cfg->local))
        cv.m2e[c->yIndelKmer[yLen]] += trans_count (&MAT (bm, i, yLen), MAT (fm, i, yLen), M2E (c, yLen), bm->end, Z);

      const double matEmit = matchEmit (c, i, j);
      const double matDest = MAT (bm, i, j);
      double* matCount = &cv.match[((size_t) c->x->tok[i-1] * c->nK + c->yMatchKmer[j-1]) * QO_NQUAL + c->y->qual[j-1]];

      const double m2m = trans_count (&MAT (bm, i-1, j-1), MAT (fm, i-1, j-1), M2M (c, j-1) + matEmit, matDest, Z);
      cv.m2m[c->yIndelKmer[j-1]] += m2m;
      *matCount += m2m;
      const double d2m = trans_count (&DEL (bm, i-1, j-1), DEL (fm, i-1, j-1), qs->d2m + matEmit, matDest, Z);
      cv.scal[1] += d2m;
      *matCount += d2m;
      const double i2m = trans_count (&INS (bm, i-1, j-1), INS (fm, i-1, j-1), qs->i2m + matEmit, matDest, Z);
      cv.scal[3] += i2m;
      *matCount += i2m;
      if (j == 1 && (i == 1 || cfg->local)) {
        const double s2m = trans_count (&bm->start, fwdStart, matEmit, matDest, Z);
        *matCount += s2m;
      }

      const double insEmit = c->insEmit[j];
      const double insDest = INS (bm, i, j);
      double* insCount = &cv.insert[(size_t) c->y->tok[j-1] * QO_NQUAL + c->y->qual[j-1]];
      const double m2i = trans_count (&MAT (bm, i, j-1), MAT (fm, i, j-1), M2I (c, j-1) + insEmit, insDest, Z);
      cv.m2i[c->yIndelKmer[j-1]] += m2i;
      *insCount += m2i;
      const double i2i = trans_count (&INS (bm, i, j-1), INS (fm, i, j-1), qs->i2i + insEmit, insDest, Z);
      cv.scal[2] += i2i;
      *insCount += i2i;

      const double delDest = DEL (bm, i, j);
      const double m2d = trans_count (&MAT (bm, i-1, j), MAT (fm, i-1, j), M2D (c, j), delDest, Z);
      cv.m2d[c->yIndelKmer[j]] += m2d;
      const double d2d = trans_count (&DEL (bm, i-1, j), DEL (fm, i-1, j), qs->d2d, delDest, Z);
      cv.scal[0] += d2d;
    }
  }
  bm->result = bm->start;
  return 0;
}

int qo_backward (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                 double* fwd_result, double* back_result, double* counts_flat,
                 double** cells, uint64_t* n_cells) {
  env_t e;
  if (make_envelope (&e, x, y, cfg, 48) != 0) return -1;
  dpm_t fm; dpm_init (&fm, &e);
  dpm_t bm; dpm_init (&bm, &e);
  ctx_t c; ctx_init (&c, x, y, qs);
  forward_fill (&fm, &c, cfg);
  const int rc = backward_fill (&bm, &fm, &c, cfg, counts_flat);
  *fwd_result = fm.result;
  *back_result = bm.result;
  if (cells) *cells = dump_cells (&bm, n_cells);
  ctx_free (&c); dpm_free (&fm); dpm_free (&bm); env_free (&e);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * E-step per read: QuaffCountingTask::run (qmodel.cpp:2238-2271); QuaffParamCounts(QuaffCounts)
 * (qmodel.cpp:407-417); addWeighted (qmodel.cpp:1656-1673); totals (qmodel.cpp:2416-2422)
 * ---------------------------------------------------------------------------------------- */
static void param_counts_add_weighted (double* dst, const double* qcounts, double w, int match_k, int gap_k) {
  const uint64_t nK = ipow4 (match_k), nG = ipow4 (gap_k);
  const size_t nEmit = 4 * nK * QO_NQUAL + 4 * QO_NQUAL;
  counts_view cv; counts_view_init (&cv, (double*) qcounts, match_k, gap_k);
  for (size_t t = 0; t < nEmit; ++t) dst[t] += w * qcounts[t];
  double* bINo = dst + nEmit; double* bIYes = bINo + nG; double* bDNo = bIYes + nG; double* bDYes = bDNo + nG;
  double* ext = bDYes + nG;
  for (uint64_t g = 0; g < nG; ++g) {
    bINo[g] += w * (cv.m2m[g] + cv.m2d[g]);
    bIYes[g] += w * (cv.m2i[g] + cv.m2e[g]);
    bDNo[g] += w * cv.m2m[g];
    bDYes[g] += w * cv.m2d[g];
  }
  ext[0] += w * cv.scal[3];   /* extendInsertNo  = i2m */
  ext[1] += w * cv.scal[2];   /* extendInsertYes = i2i */
  ext[2] += w * cv.scal[1];   /* extendDeleteNo  = d2m */
  ext[3] += w * cv.scal[0];   /* extendDeleteYes = d2d */
}

int qo_estep (const qo_seq* xs, int nx, const qo_seq* ys, int ny, const qo_scores* qs,
              const qo_null_params* np, int use_null, const qo_config* cfg,
              uint32_t* sort_order, uint32_t* sort_len, double* y_loglike, double* param_counts_flat) {
  const size_t nC = qo_counts_size (qs->match_k, qs->gap_k);
  memset (param_counts_flat, 0, sizeof(double) * nC);
  double* xyLL = (double*) malloc (sizeof(double) * nx);
  double* xyCounts = (double*) malloc (sizeof(double) * nC * nx);
  double* yCounts = (double*) malloc (sizeof(double) * nC);
  uint32_t* order = (uint32_t*) malloc (sizeof(uint32_t) * nx);
  int rc = 0;
  for (int m = 0; m < ny && rc == 0; ++m) {
    const qo_seq* y = &ys[m];
    const double yNull = use_null ? qo_null_loglike (np, y) : NEG_INF;
    double yLL = yNull;
    for (int n = 0; n < nx; ++n) xyLL[n] = NEG_INF;
    memset (xyCounts, 0, sizeof(double) * nC * nx);
    for (uint32_t s = 0; s < sort_len[m]; ++s) {
      const uint32_t n = sort_order[(size_t) m * nx + s];
      env_t e;
      if (make_envelope (&e, &xs[n], y, cfg, 48) != 0) { rc = -1; break; }
      dpm_t fm; dpm_init (&fm, &e);
      ctx_t c; ctx_init (&c, &xs[n], y, qs);
      forward_fill (&fm, &c, cfg);
      xyLL[n] = fm.result;
      if (xyLL[n] >= yLL - 20) {                               /* MAX_TRAINING_LOG_DELTA, qmodel.cpp:23, :2252 */
        dpm_t bm; dpm_init (&bm, &e);
        rc = backward_fill (&bm, &fm, &c, cfg, xyCounts + nC * n);
        dpm_free (&bm);
      }
      yLL = qo_lse (yLL, xyLL[n]);
      ctx_free (&c); dpm_free (&fm); env_free (&e);
    }
    memset (yCounts, 0, sizeof(double) * nC);
    for (int n = 0; n < nx; ++n)
      param_counts_add_weighted (yCounts, xyCounts + nC * n, exp (xyLL[n] - yLL), qs->match_k, qs->gap_k);
    for (size_t t = 0; t < nC; ++t) param_counts_flat[t] += yCounts[t];
    y_loglike[m] = yLL;
    /* orderedIndices ascending (stable sort by value, util.h), reversed, cut at first F < yLL-20 */
    for (int n = 0; n < nx; ++n) order[n] = (uint32_t) n;
    for (int a = 1; a < nx; ++a) {                              /* insertion sort = stable */
      const uint32_t v = order[a]; int b = a - 1;
      while (b >= 0 && xyLL[order[b]] > xyLL[v]) { order[b+1] = order[b]; --b; }
      order[b+1] = v;
    }
    uint32_t len = 0;
    for (int a = nx - 1; a >= 0; --a) {
      if (xyLL[order[a]] < yLL - 20) break;
      sort_order[(size_t) m * nx + len++] = order[a];
    }
    sort_len[m] = len;
  }
  free (xyLL); free (xyCounts); free (yCounts); free (order);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * Overlap model (qoverlap.cpp:9-75)
 * ---------------------------------------------------------------------------------------- */
void qo_overlap_scores_from_params (const qo_params* qp, int y_complemented, qo_overlap_scores* out) {
  const uint64_t nK = ipow4 (qp->match_k), nG = ipow4 (qp->gap_k);
  out->match_k = qp->match_k; out->gap_k = qp->gap_k; out->y_complemented = y_complemented;
  double* gapOpen = (double*) malloc (sizeof(double) * nG);
  double sumPGapIsInsert = 0, sumGapAdjacent = 0;
  for (uint64_t j = 0; j < nG; ++j) {                           /* qoverlap.cpp:23-32 */
    const double readInsertProb = qp->begin_insert[j];
    const double readDeleteProb = (1 - qp->begin_insert[j]) * qp->begin_delete[j];
    gapOpen[j] = readInsertProb + readDeleteProb;
    const double pGapIsInsert = readInsertProb / gapOpen[j];
    const double gapAdjacentProb = pGapIsInsert * readInsertProb + (1 - pGapIsInsert) * gapOpen[j] / (1 - qp->extend_delete * (1 - gapOpen[j]));
    sumPGapIsInsert += pGapIsInsert;                            /* accumulate(..., 0.) in index order */
    sumGapAdjacent += gapAdjacentProb;
  }
  for (uint64_t i = 0; i < nG; ++i)
    for (uint64_t j = 0; j < nG; ++j) {                         /* :34-39 */
      out->m2m[i*nG+j] = log (1 - gapOpen[i]) + log (1 - gapOpen[j]);
      out->m2i[i*nG+j] = log (gapOpen[i]);
      out->m2d[i*nG+j] = log (1 - gapOpen[i]) + log (gapOpen[j]);
    }
  const double pGapIsInsert = sumPGapIsInsert / nG;
  const double meanGapLength = pGapIsInsert / qp->extend_insert + (1 - pGapIsInsert) / qp->extend_delete;
  const double gapExtendProb = 1 / meanGapLength;
  const double gapAdjacentProb = sumGapAdjacent / nG;
  out->i2i = out->d2d = log (gapExtendProb);
  out->i2d = out->d2i = log (1 - gapExtendProb) + log (gapAdjacentProb);
  out->i2m = out->d2m = log (1 - gapExtendProb) + log (1 - gapAdjacentProb);
  free (gapOpen);

  qo_scores qs;
  qs.match = (double*) malloc (sizeof(double) * 4 * nK * QO_NQ1);
  qs.insert = out->insert;
  qs.m2m = (double*) malloc (sizeof(double) * nG * 4);
  qs.m2i = qs.m2m + nG; qs.m2d = qs.m2i + nG; qs.m2e = qs.m2d + nG;
  qo_scores_from_params (qp, &qs);
  const double* ins = out->insert;

  for (uint64_t t = 0; t < nK * nK * QO_NQUAL; ++t) { out->x_only[t] = NEG_INF; out->y_only[t] = NEG_INF; }
  for (uint64_t t = 0; t < nK * nK; ++t) out->none[t] = NEG_INF;
  for (uint64_t i = 0; i < nK; ++i) {                           /* :51-74, loops in the same order */
    const int iSuffix = (int) (i % 4);
    for (uint64_t j = 0; j < nK; ++j) {
      const int jSuffix = (int) (j % 4);
      double* pair = out->pair + (i*nK+j) * QO_NQUAL * QO_NQUAL;
      double* xo = out->x_only + (i*nK+j) * QO_NQUAL;
      double* yo = out->y_only + (i*nK+j) * QO_NQUAL;
      double* no = out->none + (i*nK+j);
      for (int ik = 0; ik < QO_NQUAL; ++ik)
        for (int jk = 0; jk < QO_NQUAL; ++jk) {
          double mij = NEG_INF;
          for (int r = 0; r < 4; ++r) {
            const int yr = y_complemented ? 3 - r : r;
            mij = qo_lse (mij, log (qp->ref_base[r]) + qs.match[((size_t) r * nK + i) * QO_NQ1 + ik] + qs.match[((size_t) yr * nK + j) * QO_NQ1 + jk]);
          }
          pair[ik * QO_NQUAL + jk] = mij - ins[iSuffix * QO_NQ1 + ik] - ins[jSuffix * QO_NQ1 + jk];
          xo[ik] = qo_lse (xo[ik], mij - ins[iSuffix * QO_NQ1 + ik] - ins[jSuffix * QO_NQ1 + QO_NQUAL]);
          yo[jk] = qo_lse (yo[jk], mij - ins[iSuffix * QO_NQ1 + QO_NQUAL] - ins[jSuffix * QO_NQ1 + jk]);
          *no = qo_lse (*no, mij - ins[iSuffix * QO_NQ1 + QO_NQUAL] - ins[jSuffix * QO_NQ1 + QO_NQUAL]);
        }
    }
  }
  free (qs.match); free (qs.m2m);
}

/* ------------------------------------------------------------------------------------------
 * Overlap Viterbi fill (qoverlap.cpp:77-160) and traceback (qoverlap.cpp:162-290).
 * Accessor resolution (qoverlap.h:46-51): i2mScore()=i2i, i2iScore()=i2m, i2dScore()=i2d,
 * d2mScore()=d2i, d2iScore()=d2m, d2dScore()=d2d.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const qo_overlap_scores* os;
  const qo_seq *x, *y;
  uint8_t* yTok;
  uint32_t *xMatchKmer, *yMatchKmer, *xIndelKmer, *yIndelKmer;   /* indel k-mers padded with leading 0 */
  uint64_t nK, nG;
} octx_t;

static void reversed_revcomp_kmers (const uint8_t* tok, uint32_t len, int k, uint32_t* out) {
  /* kmers of y.revcomp(), reversed back into y's coordinates (qoverlap.cpp:91-98) */
  uint8_t* rc = (uint8_t*) malloc (len + 1);
  uint32_t* km = (uint32_t*) malloc (sizeof(uint32_t) * (len + 1));
  for (uint32_t t = 0; t < len; ++t) rc[t] = (uint8_t) (3 - tok[len - 1 - t]);
  qo_context_kmers (rc, len, k, km);
  for (uint32_t t = 0; t < len; ++t) out[t] = km[len - 1 - t];
  free (rc); free (km);
}

static void octx_init (octx_t* c, const qo_seq* x, const qo_seq* y, const qo_overlap_scores* os) {
  c->os = os; c->x = x; c->y = y;
  c->nK = ipow4 (os->match_k); c->nG = ipow4 (os->gap_k);
  c->yTok = (uint8_t*) malloc (y->len + 1);
  c->xMatchKmer = (uint32_t*) malloc (sizeof(uint32_t) * (x->len + 1));
  c->yMatchKmer = (uint32_t*) malloc (sizeof(uint32_t) * (y->len + 1));
  c->xIndelKmer = (uint32_t*) malloc (sizeof(uint32_t) * (x->len + 2));
  c->yIndelKmer = (uint32_t*) malloc (sizeof(uint32_t) * (y->len + 2));
  qo_context_kmers (x->tok, x->len, os->match_k, c->xMatchKmer);
  c->xIndelKmer[0] = 0;
  qo_context_kmers (x->tok, x->len, os->gap_k, c->xIndelKmer + 1);
  c->yIndelKmer[0] = 0;
  if (os->y_complemented) {
    for (uint32_t t = 0; t < y->len; ++t) c->yTok[t] = (uint8_t) (3 - y->tok[t]);
    reversed_revcomp_kmers (y->tok, y->len, os->match_k, c->yMatchKmer);
    reversed_revcomp_kmers (y->tok, y->len, os->gap_k, c->yIndelKmer + 1);
  } else {
    memcpy (c->yTok, y->tok, y->len);
    qo_context_kmers (y->tok, y->len, os->match_k, c->yMatchKmer);
    qo_context_kmers (y->tok, y->len, os->gap_k, c->yIndelKmer + 1);
  }
}
static void octx_free (octx_t* c) { free (c->yTok); free (c->xMatchKmer); free (c->yMatchKmer); free (c->xIndelKmer); free (c->yIndelKmer); }

static inline double o_emit (const octx_t* c, int64_t i, int64_t j) {     /* qoverlap.h:52-61 */
  const size_t kk = (size_t) c->xMatchKmer[i-1] * c->nK + c->yMatchKmer[j-1];
  if (c->x->qual)
    return c->y->qual ? c->os->pair[(kk * QO_NQUAL + c->x->qual[i-1]) * QO_NQUAL + c->y->qual[j-1]]
                      : c->os->x_only[kk * QO_NQUAL + c->x->qual[i-1]];
  return c->y->qual ? c->os->y_only[kk * QO_NQUAL + c->y->qual[j-1]] : c->os->none[kk];
}
#define OM2M(c,i,j) ((c)->os->m2m[(size_t) (c)->xIndelKmer[i] * (c)->nG + (c)->yIndelKmer[j]])
#define OM2I(c,i,j) ((c)->os->m2i[(size_t) (c)->xIndelKmer[i] * (c)->nG + (c)->yIndelKmer[j]])
#define OM2D(c,i,j) ((c)->os->m2d[(size_t) (c)->xIndelKmer[i] * (c)->nG + (c)->yIndelKmer[j]])

typedef struct { char* s; size_t n, cap; } cbuf;     /* grows at the FRONT logically: we append, reverse at the end */
static void cb_push (cbuf* b, char ch) {
  if (b->n + 1 >= b->cap) { b->cap = b->cap ? b->cap * 2 : 256; b->s = (char*) realloc (b->s, b->cap); }
  b->s[b->n++] = ch;
}

int qo_overlap_viterbi (const qo_seq* x, const qo_seq* y, const qo_overlap_scores* os, const qo_config* cfg,
                        double* result, uint32_t* coords4, char** xrow, char** yrow,
                        double** cells, uint64_t* n_cells) {
  static const char alph[] = "ACGT";
  env_t e;
  if (make_envelope (&e, x, y, cfg, 24) != 0) return -1;
  dpm_t m; dpm_init (&m, &e);
  octx_t c; octx_init (&c, x, y, os);
  const uint32_t xLen = e.xLen, yLen = e.yLen;
  /* swapped accessors resolved here */
  const double I2M = os->i2i, I2I = os->i2m, I2D = os->i2d, D2M = os->d2i, D2I = os->d2m, D2D = os->d2d;

  double xInsertScore = 0, yInsertScore = 0;                     /* qoverlap.cpp:108-116 */
  for (uint32_t i = 0; i < xLen; ++i) xInsertScore += os->insert[x->tok[i] * QO_NQ1 + (x->qual ? x->qual[i] : QO_NQUAL)];
  for (uint32_t j = 0; j < yLen; ++j) yInsertScore += os->insert[c.yTok[j] * QO_NQ1 + (y->qual ? y->qual[j] : QO_NQUAL)];

  m.start = 0;
  for (uint32_t j = 1; j <= yLen; ++j) {                         /* :122-157 */
    int b, en; row_range (&e, j, &b, &en);
    for (int t = b; t < en; ++t) {
      const int64_t i = (int64_t) e.diags[t] + j;
      double mat = fmax (fmax (MAT (&m, i-1, j-1) + OM2M (&c, i-1, j-1), DEL (&m, i-1, j-1) + D2M), INS (&m, i-1, j-1) + I2M);
      if (j == 1 || i == 1) mat = fmax (mat, m.start);
      mat += o_emit (&c, i, j);
      MAT (&m, i, j) = mat;
      INS (&m, i, j) = fmax (qo_lse (INS (&m, i, j-1) + I2I, DEL (&m, i, j-1) + D2I), MAT (&m, i, j-1) + OM2I (&c, i, j-1));
      DEL (&m, i, j) = fmax (qo_lse (DEL (&m, i-1, j) + D2D, INS (&m, i-1, j) + D2I), MAT (&m, i-1, j) + OM2D (&c, i-1, j));
      if (j == yLen || i == xLen) m.end = fmax (m.end, MAT (&m, i, j));
    }
  }
  m.result = m.end + xInsertScore + yInsertScore;
  *result = m.result;
  if (cells) *cells = dump_cells (&m, n_cells);
  if (xrow) { *xrow = NULL; *yrow = NULL; }
  if (coords4) coords4[0] = coords4[1] = coords4[2] = coords4[3] = 0;
  int rc = 0;

  if (xrow && m.result > NEG_INF) {
    uint32_t xEnd = xLen, yEnd = yLen;                           /* :164-182 */
    double best = MAT (&m, xLen, yLen);
    for (uint32_t iEnd = xLen; iEnd > 0; --iEnd) { const double sc = MAT (&m, iEnd, yLen); if (sc > best) { best = sc; xEnd = iEnd; yEnd = yLen; } }
    for (uint32_t jEnd = yLen; jEnd > 0; --jEnd) { const double sc = MAT (&m, xLen, jEnd); if (sc > best) { best = sc; xEnd = xLen; yEnd = jEnd; } }
    int64_t i = xEnd, j = yEnd;
    /* rows are built back-to-front; xr/yr hold the finished part REVERSED, delRun/insRun the
       pending deleted x chars / inserted y chars, also reversed (most recent = leftmost last) */
    cbuf xr = {0,0,0}, yr = {0,0,0}, delRun = {0,0,0}, insRun = {0,0,0};
    int state = ST_MATCH;
    size_t guard = 0;
    while (state != ST_START) {
      double src = NEG_INF, emit;
      switch (state) {
      case ST_MATCH:
        emit = o_emit (&c, i, j);
        --i; --j;
        cb_push (&xr, alph[x->tok[i]]);
        cb_push (&yr, alph[y->tok[j]]);
        UPDATE_MAX (src, state, MAT (&m, i, j) + OM2M (&c, i, j) + emit, ST_MATCH);
        UPDATE_MAX (src, state, INS (&m, i, j) + I2M + emit, ST_INSERT);
        UPDATE_MAX (src, state, DEL (&m, i, j) + D2M + emit, ST_DELETE);
        if (j == 0 || i == 0) UPDATE_MAX (src, state, emit, ST_START);
        if (!(src == MAT (&m, i+1, j+1))) rc = -2;
        break;
      case ST_INSERT:
        --j;
        cb_push (&insRun, alph[y->tok[j]]);
        UPDATE_MAX (src, state, MAT (&m, i, j) + OM2I (&c, i, j), ST_MATCH);
        UPDATE_MAX (src, state, INS (&m, i, j) + I2I, ST_INSERT);
        UPDATE_MAX (src, state, DEL (&m, i, j) + D2I, ST_DELETE);
        break;
      default:
        --i;
        cb_push (&delRun, alph[x->tok[i]]);
        UPDATE_MAX (src, state, MAT (&m, i, j) + OM2D (&c, i, j), ST_MATCH);
        UPDATE_MAX (src, state, INS (&m, i, j) + I2D, ST_INSERT);
        UPDATE_MAX (src, state, DEL (&m, i, j) + D2D, ST_DELETE);
        break;
      }
      if (state == ST_MATCH || state == ST_START) {
        /* NB the reference tests `state == Match` only (qoverlap.cpp:231): a run that ends in Start
           is dropped.  Start can only be entered from the Match case, after which both runs are
           already empty, so the two conditions coincide. */
        /* squash (qoverlap.cpp:231-267).  Final left-to-right layout of the pending block is
             [shared x over shared y][extra deleted x over gaps][gaps over extra inserted y]
           where "shared" are the FIRST sharedLen chars of each run in left-to-right order. */
        const size_t insLen = insRun.n, delLen = delRun.n;
        const size_t shared = insLen < delLen ? insLen : delLen;
        /* left-to-right order of a run = reverse of the push order: l2r[t] = run.s[n-1-t] */
        /* we append to reversed rows, so emit the block right-to-left: */
        for (size_t t = insLen; t-- > shared; ) { cb_push (&xr, '-'); cb_push (&yr, insRun.s[insLen - 1 - t]); }
        for (size_t t = delLen; t-- > shared; ) { cb_push (&xr, delRun.s[delLen - 1 - t]); cb_push (&yr, '-'); }
        for (size_t t = shared; t-- > 0; ) { cb_push (&xr, delRun.s[delLen - 1 - t]); cb_push (&yr, insRun.s[insLen - 1 - t]); }
        insRun.n = delRun.n = 0;
      }
      if (rc != 0 || ++guard > (size_t) xLen + yLen + 2) { rc = -2; break; }
    }
    coords4[0] = (uint32_t) (i + 1); coords4[1] = xEnd; coords4[2] = (uint32_t) (j + 1); coords4[3] = yEnd;
    *xrow = (char*) malloc (xr.n + 1); *yrow = (char*) malloc (yr.n + 1);
    for (size_t t = 0; t < xr.n; ++t) { (*xrow)[t] = xr.s[xr.n - 1 - t]; (*yrow)[t] = yr.s[yr.n - 1 - t]; }
    (*xrow)[xr.n] = 0; (*yrow)[yr.n] = 0;
    free (xr.s); free (yr.s); free (delRun.s); free (insRun.s);
  }
  octx_free (&c); dpm_free (&m); env_free (&e);
  return rc;
}
