/* quaffgpu.h -- C ABI of libquaffgpu.so: the B200 (sm_100a) implementation of ihh/quaff's banded
 * pair-HMM hot path.  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * The reference has no plugin/FFI interface; the narrowest seams are three member functions whose
 * bodies are "spawn the thread pool, drain the scheduler, join" (SURVEY.md section 8b):
 *   seam A  QuaffAligner::align          src/qmodel.cpp:2624   -> qg_align_reads
 *   seam B  QuaffOverlapAligner::align   src/qoverlap.cpp:312  -> qg_overlap_reads
 *   seam C  QuaffTrainer::getCounts      src/qmodel.cpp:2005   -> qg_estep
 * Below those, one entry point per reference object on the path, which is what the parity tests
 * drive (file:line = the reference interface each one replaces, relative to /root/reference):
 *   qg_envelopes          QuaffDPConfig::makeEnvelope  qmodel.cpp:1049 / DiagonalEnvelope::initSparse diagenv.cpp:20
 *   qg_viterbi            QuaffViterbiMatrix ctor + alignment()   qmodel.cpp:1512, 1562
 *   qg_forward            QuaffForwardMatrix ctor                 qmodel.cpp:1343
 *   qg_backward_counts    QuaffBackwardMatrix ctor (-> QuaffCounts) qmodel.cpp:1393
 *   qg_overlap_viterbi    QuaffOverlapViterbiMatrix ctor + alignment()  qoverlap.cpp:77, 162
 *   qg_scores_from_params QuaffScores ctor  qmodel.cpp:296      (host arithmetic, built once, uploaded)
 *   qg_null_loglike       QuaffNullParams::logLikelihood qmodel.cpp:1875 (host, O(L) per read)
 *
 * Conventions: every call returns 0 on success or a QG_ERR_* code, with a message available from
 * qg_last_error(); no exception crosses the boundary; a context is driven by one host thread and
 * owns one GPU (one process per GPU; multi-GPU = one context per rank, reads sharded by the caller);
 * calls are synchronous at return.  Inputs are caller-owned and copied; outputs are either
 * caller-allocated flat arrays or library-allocated (documented per call) and released with qg_free.
 * There is NO CPU fallback: without a CUDA device qg_create fails.
 */
#ifndef QUAFFGPU_H
#define QUAFFGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_OK                 0
#define QG_ERR_CUDA           1   /* a CUDA runtime call failed (message has the CUDA error string)      */
#define QG_ERR_INVALID        2   /* bad argument (null pointer, index out of range, token > 3, ...)     */
#define QG_ERR_UNSUPPORTED    3   /* legal for the reference but outside what this build implements      */
#define QG_ERR_NO_DEVICE      4   /* no usable CUDA device: there is no CPU path                         */
#define QG_ERR_STATE          5   /* call order (e.g. no model / sequences uploaded yet)                 */
#define QG_ERR_PRECONDITION   6   /* input on which the reference itself crashes (sequence shorter than k) */

#define QG_NQUAL 94               /* FastSeq::qualScoreRange (fastseq.cpp:69)                             */
#define QG_NQ1   95               /* 94 quality bins + the marginal (logSymProb) slot used by -noquals    */

#define QG_REFS  0                /* sequence set "x": references (align/train)                           */
#define QG_READS 1                /* sequence set "y": reads; overlap uses this set for both x and y      */

/* path op codes (5' -> 3'), from the traceback's state sequence */
#define QG_OP_MATCH  0
#define QG_OP_INSERT 1            /* read base against a gap in the reference row */
#define QG_OP_DELETE 2            /* reference base against a gap in the read row */

typedef struct qg_ctx qg_ctx;

/* the QuaffDPConfig members the DP reads (qmodel.h:280-290) */
typedef struct {
  int32_t  sparse;                /* 0 = -kmatchoff (full DP)                                             */
  int32_t  kmer_len;              /* -kmatch                                                              */
  int32_t  kmer_threshold;        /* -kmatchn ; < 0 = memory-guided mode (-kmatchmb / -kmatchmax)          */
  int32_t  band_size;             /* -kmatchband                                                          */
  int32_t  local;                 /* 0 = -global                                                          */
  uint64_t max_size;              /* effectiveMaxSize() in bytes, memory-guided mode only                  */
} qg_dpconfig;

/* QuaffScores (qmodel.h:181-191), flattened, FP64 */
typedef struct {
  int32_t match_k, gap_k;         /* matchContext.kmerLen, indelContext.kmerLen                            */
  const double* match;            /* [4][4^match_k][95]  ref base x read context k-mer x (quality | marginal) */
  const double* insert;           /* [4][95]                                                              */
  const double *m2m, *m2i, *m2d, *m2e;   /* [4^gap_k] each                                                 */
  double d2d, d2m, i2i, i2m;
} qg_align_model;

/* QuaffParams values (qmodel.h:154-170): three doubles (p, q, r) per SymQualDist */
typedef struct {
  int32_t match_k, gap_k;
  double ref_base[4];
  const double *begin_insert, *begin_delete;   /* [4^gap_k] */
  double extend_insert, extend_delete;
  const double* insert_pqr;       /* [4][3]              */
  const double* match_pqr;        /* [4][4^match_k][3]   */
} qg_params;

/* flattened QuaffCounts (qmodel.h:205-212):
 *   match[4][4^K][94], insert[4][94], m2m[4^G], m2i[4^G], m2d[4^G], m2e[4^G], d2d, d2m, i2i, i2m   */
size_t qg_counts_size (int match_k, int gap_k);
/* flattened QuaffParamCounts (qmodel.h:214-238) has the same length:
 *   match, insert, beginInsertNo[4^G], beginInsertYes, beginDeleteNo, beginDeleteYes,
 *   extendInsertNo, extendInsertYes, extendDeleteNo, extendDeleteYes                               */

/* ---- context ---------------------------------------------------------------------------------- */
int  qg_create (qg_ctx** out, int device);
void qg_destroy (qg_ctx* ctx);
const char* qg_last_error (const qg_ctx* ctx);   /* ctx may be NULL: last qg_create failure          */
void qg_free (void* p);
int  qg_abi_version (void);

/* options */
#define QG_OPT_FB_EXACT 1         /* 1: Forward/Backward with the reference's table log-sum-exp in log space (bit-exact
                                     Forward); 0 (default): probability-space FP64 kernels, ~1e-8 relative to the reference */
int  qg_set_option (qg_ctx* ctx, int option, int64_t value);

/* ---- inputs ----------------------------------------------------------------------------------- */
/* tok: concatenated tokens in {0,1,2,3}; qual: concatenated quality scores in 0..93 or NULL when the
 * set carries no qualities; offsets[n+1].  Replaces whatever the set held before.                   */
int qg_set_seqs (qg_ctx* ctx, int which, size_t n, const uint8_t* tok, const uint8_t* qual, const uint64_t* offsets);
int qg_set_align_model (qg_ctx* ctx, const qg_align_model* model);

/* host helpers (no GPU work): QuaffScores tables from QuaffParams; the caller allocates the outputs
 * with the sizes documented in qg_align_model                                                        */
int qg_scores_from_params (const qg_params* qp, double* match, double* insert,
                           double* m2m, double* m2i, double* m2d, double* m2e, double* scal4 /* d2d,d2m,i2i,i2m */);
/* null_pqr[4][3]; qual may be NULL */
double qg_null_loglike (double null_emit, const double* null_pqr, const uint8_t* tok, const uint8_t* qual, uint64_t len);

/* ---- seeding: k-mer diagonal envelope ----------------------------------------------------------- */
/* For pair p: x = refs[xi[p]], y = reads[yi[p]] (overlap: pass x_set = QG_READS).
 * diags_out: library-allocated sorted diagonal lists, pair p in [diag_offsets[p], diag_offsets[p+1]);
 * cell_updates (optional, n_pairs): iterated envelope cells per pair (SURVEY 8d "CU").               */
int qg_envelopes (qg_ctx* ctx, const qg_dpconfig* cfg, uint64_t cell_size, int x_set,
                  size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                  int32_t** diags_out, uint64_t* diag_offsets /* n_pairs+1 */, uint64_t* cell_updates);

/* ---- align: Viterbi ----------------------------------------------------------------------------- */
/* score[p] = QuaffViterbiMatrix::result (before the null-model adjustment), -inf if no path.
 * want_path (optional, n_pairs): nonzero = run the traceback for that pair; NULL = all pairs.
 * path_out: library-allocated op codes, pair p in [path_offsets[p], path_offsets[p+1]) (empty when
 * not requested or score = -inf); x_start/x_end: 1-based closed reference interval.                 */
int qg_viterbi (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                const uint8_t* want_path, double* score, uint32_t* x_start, uint32_t* x_end,
                uint8_t** path_out, uint64_t* path_offsets /* n_pairs+1 */);

/* ---- train: Forward, Backward + E-step counts ---------------------------------------------------- */
int qg_forward (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                double* loglike);
/* Runs Forward (kept on the device) then Backward for every listed pair.
 * weights (optional): posterior weight per pair, exp(F - yLL); NULL = 1.
 * counts_sum (optional): sum_p weights[p] * QuaffCounts_p, qg_counts_size doubles.
 * counts_per_pair (optional): unweighted QuaffCounts of each pair, n_pairs * qg_counts_size doubles. */
int qg_backward_counts (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                        const double* weights, double* fwd_loglike, double* back_loglike,
                        double* counts_sum, double* counts_per_pair);

/* ---- seam A: QuaffAligner::align (qmodel.cpp:2624) ---------------------------------------------- */
/* Every read of the READS set against every sequence of the REFS set (which already holds the reverse
 * complements when both strands are wanted); keeps, per read, the best-scoring reference, earliest
 * reference index on ties (qmodel.cpp:2773-2775).  null_loglike[n_reads] is subtracted from the
 * Viterbi score (qmodel.cpp:1648-1654).  Outputs are per read; best_ref = UINT32_MAX when no
 * reference gave a finite score.  path_out as in qg_viterbi, for the best pair of each read.       */
int qg_align_reads (qg_ctx* ctx, const qg_dpconfig* cfg, const double* null_loglike,
                    uint32_t* best_ref, double* score, uint32_t* x_start, uint32_t* x_end,
                    uint8_t** path_out, uint64_t* path_offsets /* n_reads+1 */);

/* the same over reads [first_read, first_read + n_reads) of the READS set (outputs sized n_reads): lets a caller that
 * keeps a large read set resident on the device drive it batch by batch                               */
int qg_align_reads_range (qg_ctx* ctx, const qg_dpconfig* cfg, size_t first_read, size_t n_reads, const double* null_loglike,
                          uint32_t* best_ref, double* score, uint32_t* x_start, uint32_t* x_end,
                          uint8_t** path_out, uint64_t* path_offsets /* n_reads+1 */);

/* ---- seam C: QuaffTrainer::getCounts (qmodel.cpp:2005), one E-step over this rank's reads --------- */
/* sort_order: in/out [n_reads][n_refs] with lengths sort_len[n_reads] (qmodel.cpp:2247, 2264-2270);
 * null_loglike[n_reads] used when use_null.  y_loglike[n_reads]; param_counts (QuaffParamCounts layout)
 * and *loglike_sum are this rank's partial sums -- the caller all-reduces them across ranks.         */
int qg_estep (qg_ctx* ctx, const qg_dpconfig* cfg, int use_null, const double* null_loglike,
              uint32_t* sort_order, uint32_t* sort_len, double* y_loglike,
              double* param_counts, double* loglike_sum);

/* ---- overlap ------------------------------------------------------------------------------------- */
/* QuaffOverlapScores (qoverlap.h:20-32) is derived on the device from these factors, once per strand */
typedef struct {
  int32_t match_k, gap_k;
  const double* match;            /* QuaffScores.match  [4][4^K][95] */
  const double* insert;           /* QuaffScores.insert [4][95]      */
  double log_ref_base[4];         /* log(refBase[r])                  */
  const double *begin_insert, *begin_delete;   /* [4^G] probabilities */
  double extend_insert, extend_delete;
} qg_overlap_model;
int qg_set_overlap_model (qg_ctx* ctx, const qg_overlap_model* model);
/* both x and y index the READS set; y_complemented[p] as in QuaffOverlapTask (qoverlap.cpp:457).
 * score = QuaffOverlapViterbiMatrix::result; coords4[p] = x_start,x_end,y_start,y_end (1-based closed);
 * path_out: raw state path (QG_OP_*), before the reference's indel squashing (qoverlap.cpp:231-267),
 * which qg_overlap_rows applies on the host when building the two gapped rows.                      */
int qg_overlap_viterbi (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                        const uint8_t* y_complemented, const uint8_t* want_path,
                        double* score, uint32_t* coords4, uint8_t** path_out, uint64_t* path_offsets);
/* host string assembly: op path -> gapped rows ("ACGT-"), NUL-terminated, library-allocated          */
int qg_overlap_rows (const uint8_t* x_tok, const uint8_t* y_tok, const uint32_t* coords4,
                     const uint8_t* path, uint64_t path_len, char** xrow, char** yrow);
/* seam B: all pairs nx < ny over n_originals originals + their reverse complements (qoverlap.cpp:528-547);
 * null_loglike[n] per stored sequence (the caller passes nullLL(revcomp) = nullLL of the stored strand).
 * Outputs per pair in the reference's task order; library-allocated pair list (xi, yi).             */
int qg_overlap_reads (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_originals, const double* null_loglike,
                      size_t* n_pairs_out, uint32_t** xi_out, uint32_t** yi_out,
                      double** score_out, uint32_t** coords4_out, uint8_t** path_out, uint64_t** path_offsets_out);

/* ---- several contexts / several GPUs behind one handle ------------------------- */
/* One process drives every device: the reads shard over `contexts_per_device` contexts on each of `devices`
 * (one host thread + one stream each); every context holds the reference set and the model.  This is what
 * `quaff ... -gpu 0,1,2,3` / `-gpu all` binds (host/quaff_gpu_seams.cpp); it replaces the reference's thread pool
 * (qmodel.cpp:2650-2668 for align, qmodel.cpp:2013-2029 for the E-step).                                  */
typedef struct qg_pool qg_pool;
int  qg_device_count (void);
/* blocking: driver + primary-context initialisation of the listed devices (seconds on a multi-GPU box); meant to be called
 * from a side thread while the host is still reading its inputs, so that qg_create / qg_pool_create find them ready   */
int  qg_init_devices (const int* devices, int n_devices);
int  qg_pool_create (qg_pool** out, const int* devices, int n_devices, int contexts_per_device);
void qg_pool_destroy (qg_pool* pool);
const char* qg_pool_last_error (const qg_pool* pool);     /* pool may be NULL: last qg_pool_create failure */
int  qg_pool_size (const qg_pool* pool);                  /* number of contexts                             */
qg_ctx* qg_pool_context (qg_pool* pool, int i);           /* for qg_get_stats                               */
int  qg_pool_set_refs (qg_pool* pool, size_t n, const uint8_t* tok, const uint64_t* offsets);
int  qg_pool_set_align_model (qg_pool* pool, const qg_align_model* model);
int  qg_pool_set_option (qg_pool* pool, int option, int64_t value);
/* seam A over a whole read set held in HOST memory: chunks of `chunk_reads` reads (0 = default) go to the contexts as
 * they become free; `on_chunk` is called once per chunk, on the worker's thread (concurrently with other chunks, in
 * no particular order), with the outputs of qg_align_reads for reads [first_read, first_read + n_reads); the buffers
 * are valid during the call only.                                                                      */
typedef void (*qg_chunk_fn) (void* user, int worker, size_t first_read, size_t n_reads,
                             const uint32_t* best_ref, const double* score, const uint32_t* x_start, const uint32_t* x_end,
                             const uint8_t* paths, const uint64_t* path_offsets /* n_reads+1, relative to paths */);
int  qg_pool_align_reads (qg_pool* pool, const qg_dpconfig* cfg, size_t n_reads, const uint8_t* tok, const uint8_t* qual,
                          const uint64_t* offsets, const double* null_loglike, size_t chunk_reads,
                          qg_chunk_fn on_chunk, void* user);
/* seam C over a whole read set: contiguous read ranges per context, the partial counts and log-likelihoods summed on
 * the host in context order.  Arguments as qg_estep, for all n_reads reads (sort_order is [n_reads][n_refs]).      */
int  qg_pool_estep (qg_pool* pool, const qg_dpconfig* cfg, int use_null, size_t n_refs, size_t n_reads,
                    const uint8_t* tok, const uint8_t* qual, const uint64_t* offsets, const double* null_loglike,
                    uint32_t* sort_order, uint32_t* sort_len, double* y_loglike, double* param_counts, double* loglike_sum);

/* seam B over the pool: every context holds the whole read set and the overlap model; the scheduler's pair list is cut into
 * contiguous ranges balanced by cells.  Outputs as qg_overlap_reads (library-allocated, pair order).                  */
int  qg_pool_overlap_reads (qg_pool* pool, const qg_dpconfig* cfg, const qg_overlap_model* model,
                            size_t n_seqs, const uint8_t* tok, const uint8_t* qual, const uint64_t* offsets,
                            size_t n_originals, const double* null_loglike,
                            size_t* n_pairs_out, uint32_t** xi_out, uint32_t** yi_out,
                            double** score_out, uint32_t** coords4_out, uint8_t** path_out, uint64_t** path_offsets_out);

/* ---- instrumentation ------------------------------------------------------------------------------ */
typedef struct {
  double ms_seed, ms_envelope, ms_prep, ms_viterbi, ms_traceback, ms_forward, ms_backward, ms_overlap, ms_h2d, ms_d2h;
  uint64_t kernel_launches;       /* kernels of this library launched since the last reset             */
  uint64_t cell_updates;          /* envelope cells filled (Viterbi/Forward/Backward/overlap, summed)   */
  uint64_t kmer_hits;             /* histogram increments done by the seeding kernel                    */
  uint64_t trace_bytes, fwd_store_bytes;
  uint64_t n_pairs, n_segments;
} qg_stats;
int qg_get_stats (qg_ctx* ctx, qg_stats* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* QUAFFGPU_H */
