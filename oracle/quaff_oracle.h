/* TEST INFRASTRUCTURE ONLY.
 *
 * quaff_oracle: a plain-C, single-threaded CPU restatement of the banded pair-HMM hot path of
 * ihh/quaff.  It exists to CHECK the CUDA library; it is never linked into,
 * imported by or executed from the product (quaff_b200/, libquaffgpu.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_ref.py, run wherever
 * oracle/_ref/libquaffref.so exists) against the unmodified reference compiled from
 * /root/reference, and against the committed golden vectors in tests/golden/ everywhere else.
 * The reference's own golden files (data/c8f30-self-{align,overlap,counts}.json) are reproduced
 * byte-for-byte by the _ref build and numerically by this file (tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef QUAFF_ORACLE_H
#define QUAFF_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QO_NQUAL 94            /* FastSeq::qualScoreRange, fastseq.cpp:69 */
#define QO_NQ1   95            /* 94 quality bins + the marginal logSymProb slot */

typedef struct { double p, q, r; } qo_symqual;   /* SymQualDist, qmodel.h:88-96 */

typedef struct {               /* QuaffParams, qmodel.h:154-170 */
  int match_k, gap_k;          /* matchContext.kmerLen (1 + suborder), indelContext.kmerLen */
  double ref_base[4];
  const double *begin_insert, *begin_delete;   /* [4^gap_k] */
  double extend_insert, extend_delete;
  qo_symqual insert[4];
  const qo_symqual* match;     /* [4][4^match_k] */
} qo_params;

typedef struct {               /* QuaffNullParams, qmodel.h:172-186 */
  double null_emit;
  qo_symqual null[4];
} qo_null_params;

typedef struct {               /* QuaffScores, qmodel.h:181-191 -- flattened, caller-allocated */
  int match_k, gap_k;
  double* match;               /* [4][4^match_k][95] */
  double* insert;              /* [4][95] */
  double *m2m, *m2i, *m2d, *m2e;   /* [4^gap_k] */
  double d2d, d2m, i2i, i2m;
} qo_scores;

typedef struct {               /* the QuaffDPConfig members the DP reads, qmodel.h:280-290 */
  int sparse, kmer_len, kmer_threshold, band_size, local;
  uint64_t max_size;
} qo_config;

typedef struct {               /* one sequence: tokens 0..3, quality scores 0..93 or NULL */
  const uint8_t* tok;
  const uint8_t* qual;
  uint32_t len;
} qo_seq;

void   qo_free (void* p);

/* logsumexp.cpp:20-28, 34-59, 84-103 */
double qo_lse (double a, double b);
double qo_lse_unary (double x);
const double* qo_lse_table (int* n_entries);

/* fastseq.cpp:27-35, 85-99 : context k-mer ending at each position */
void   qo_context_kmers (const uint8_t* tok, uint32_t len, int k, uint32_t* out);

/* negbinom.cpp:30-32 via GSL's formula */
double qo_log_negbinom (int k, double p, double r);
/* qmodel.cpp:87-93, 296-325 */
void   qo_scores_from_params (const qo_params* qp, qo_scores* out);
/* qmodel.cpp:1875-1890 */
double qo_null_loglike (const qo_null_params* np, const qo_seq* s);

/* fastseq.cpp:240-256 + diagenv.cpp:11-106 ; returns number of diagonals, *diags malloc'd.
 * cell_updates = sum over rows of iterated envelope cells (SURVEY 8d "CU"). */
int    qo_envelope (const qo_seq* x, const qo_seq* y, const qo_config* cfg, uint64_t cell_size,
                    int32_t** diags, uint64_t* cell_updates);

/* qmodel.cpp:1512-1654.  path: malloc'd op codes 0=M 1=I 2=D, 5'->3'; x_start/x_end 1-based closed.
 * cells (optional): malloc'd [n_cells][3] in row-major envelope iteration order. */
int    qo_viterbi (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                   double* result, uint32_t* x_start, uint32_t* x_end,
                   uint8_t** path, uint32_t* path_len, double** cells, uint64_t* n_cells);

/* qmodel.cpp:1343-1391 */
int    qo_forward (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                   double* result, double** cells, uint64_t* n_cells);

/* qmodel.cpp:1393-1510 ; counts_flat in the QuaffCounts layout of include/quaffgpu.h */
int    qo_backward (const qo_seq* x, const qo_seq* y, const qo_scores* qs, const qo_config* cfg,
                    double* fwd_result, double* back_result, double* counts_flat,
                    double** cells, uint64_t* n_cells);
size_t qo_counts_size (int match_k, int gap_k);

/* qmodel.cpp:2238-2271, 407-417, 1656-1673, 2416-2422.
 * sort_order in/out [ny][nx] with lengths sort_len[ny]; param_counts_flat in QuaffParamCounts layout */
int    qo_estep (const qo_seq* xs, int nx, const qo_seq* ys, int ny, const qo_scores* qs,
                 const qo_null_params* np, int use_null, const qo_config* cfg,
                 uint32_t* sort_order, uint32_t* sort_len, double* y_loglike, double* param_counts_flat);

/* qoverlap.cpp:9-75.  Caller-allocated outputs; pair table is [nK][nK][94][94]. */
typedef struct {
  int match_k, gap_k, y_complemented;
  double *m2m, *m2i, *m2d;          /* [4^G][4^G] */
  double i2m, i2i, i2d, d2m, d2i, d2d;   /* as STORED (not through the swapped accessors) */
  double* pair;                     /* matchMinusInsert[iK][jK].logSymQualPairProb[xq][yq] */
  double *x_only, *y_only;          /* [nK][nK][94] */
  double* none;                     /* [nK][nK] */
  double insert[4 * QO_NQ1];        /* xInsert = yInsert = QuaffScores.insert */
} qo_overlap_scores;
void   qo_overlap_scores_from_params (const qo_params* qp, int y_complemented, qo_overlap_scores* out);

/* qoverlap.cpp:77-302.  y is the sequence AS STORED in the overlap read set (already the reverse
 * complement when y_complemented).  xrow/yrow: malloc'd gapped rows over the alphabet "ACGT-"
 * (0..3 tokens as letters), after the reference's indel squashing; coords4 = xs,xe,ys,ye. */
int    qo_overlap_viterbi (const qo_seq* x, const qo_seq* y, const qo_overlap_scores* os, const qo_config* cfg,
                           double* result, uint32_t* coords4, char** xrow, char** yrow,
                           double** cells, uint64_t* n_cells);

#ifdef __cplusplus
}
#endif
#endif
