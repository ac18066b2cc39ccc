// Shared declarations for libquaffgpu: context, device buffers, launch/err macros.
#ifndef QG_COMMON_CUH
#define QG_COMMON_CUH

#ifdef QG_EMU
#include "cuda_emu.h"          // tests/emu: CPU-thread shim, test infrastructure only
#define QG_LAUNCH(kern, grid, block, smem, stream, ...) \
  qgemu::launch (dim3 (grid), dim3 (block), (size_t) (smem), [=] () { kern (__VA_ARGS__); })
#define QG_DYN_SMEM(name) unsigned char* name = qgemu::t_block->dyn_smem
#else
#include <cuda_runtime.h>
#define QG_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { const dim3 qg_grid_ (grid); if (qg_grid_.x && qg_grid_.y && qg_grid_.z) kern<<<qg_grid_, dim3 (block), (size_t) (smem), (stream)>>> (__VA_ARGS__); } while (0)   /* an empty grid is a no-op, not an error */
#define QG_DYN_SMEM(name) extern __shared__ __align__ (16) unsigned char name[]
#endif

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/quaffgpu.h"

#define QG_NEG_INF (-INFINITY)
#define QG_FULL_MASK 0xffffffffu

// ---- error plumbing -----------------------------------------------------------------------------
struct qg_error { int code; std::string msg; };

#define QG_FAIL(ctx, c, ...) do { char buf_[512]; snprintf (buf_, sizeof (buf_), __VA_ARGS__); \
    (ctx)->err.code = (c); (ctx)->err.msg = buf_; return (c); } while (0)
#define QG_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    QG_FAIL (ctx, QG_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString (e_)); } while (0)
#define QG_TRY(expr) do { int rc_ = (expr); if (rc_ != QG_OK) return rc_; } while (0)

// ---- a growable device buffer -------------------------------------------------------------------
struct qg_dbuf {
  void* p = nullptr;
  size_t cap = 0;
  template<class T> T* as () const { return (T*) p; }
};

// ---- sequence set on the device ------------------------------------------------------------------
struct qg_seqset {
  size_t n = 0;
  uint64_t total = 0;
  bool has_qual = false;
  std::vector<uint64_t> off;         // host copy of offsets [n+1]
  std::vector<uint8_t> h_tok, h_qual; // host copies (the host drivers need tokens for row assembly)
  qg_dbuf d_tok, d_qual, d_off;      // uint8 tokens / quals, uint64 offsets
  qg_dbuf d_packed, d_poff;          // 2-bit packed tokens (uint64 words, 32 tokens each) + word offsets [n+1]
  std::vector<uint64_t> poff;
  // k-mer codes (uint16 per position, 0xFFFF = no k-mer starts here), cached for one k
  int codes_k = 0;
  qg_dbuf d_codes;
  // tiles of QG_TILE_POS positions sorted by k-mer code (uint32 per position: code << 16 | position in tile), for one k
  int sorted_k = 0;
  qg_dbuf d_sorted;
  uint32_t max_len = 0;
  uint32_t len (size_t i) const { return (uint32_t) (off[i + 1] - off[i]); }
};

// ---- model on the device ------------------------------------------------------------------------
struct qg_model_dev {
  bool set = false;
  int match_k = 1, gap_k = 0;
  uint64_t nK = 4, nG = 1;
  qg_dbuf d_match, d_insert, d_gap;  // match [4][nK][95]; insert [4][95]; gap = m2m|m2i|m2d|m2e each [nG]
  double d2d = 0, d2m = 0, i2i = 0, i2m = 0;
  std::vector<double> h_m2e;
};

struct qg_overlap_dev {
  bool set = false;
  int match_k = 1, gap_k = 0;
  uint64_t nK = 4, nG = 1;
  qg_dbuf d_match, d_insert;         // factors, as in qg_overlap_model
  double log_ref_base[4];
  // derived per strand s (0 = same strand, 1 = y complemented)
  qg_dbuf d_pair[2], d_xonly[2], d_yonly[2], d_none[2];
  qg_dbuf d_m2m, d_m2i, d_m2d;       // [nG][nG]
  double i2m, i2i, i2d, d2m, d2i, d2d;   // as stored by the reference (before accessor swapping)
  bool built[2] = {false, false};
};

// one DP work unit: a maximal run of consecutive envelope diagonals of one pair
struct qg_segment {
  uint32_t pair;                     // index into the call's pair list
  int32_t  dlo;                      // first diagonal
  uint32_t width;                    // number of diagonals
  uint32_t xlen, ylen;
  uint32_t xseq, yseq;               // sequence indices
  uint32_t nwarps;                   // warps cooperating on the segment (1 unless width > 32*R)
  uint32_t R;                        // diagonals per lane
  uint32_t half;                     // Viterbi pointers of this segment are 16-bit words (R <= 4 nibbles: qg_vit_kernel), not 32-bit
  uint64_t trace_off;                // word offset of this segment's pointer block
  uint64_t store_off;                // double offset of this segment's stored Forward matrix
  uint64_t rp_off;                   // row offset of the read's row-parameter block
  uint64_t aux_off;                  // offset into the per-segment end-value area (Viterbi: pair-order segment index)
  uint64_t acc_off;                  // Backward: row offset of this segment's per-row count sums
  uint64_t seg_id;                   // pair-order index of the segment
};

struct qg_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = 0;
  cudaEvent_t ev[2] = {0, 0};
  cudaStream_t side[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // launch classes of one stage run concurrently on these
  cudaEvent_t ev_fork = 0, ev_join[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  qg_error err;
  qg_seqset seqs[2];
  qg_model_dev model;
  qg_overlap_dev omodel;
  qg_dbuf d_lse;                     // the 100001-entry FP64 log-sum-exp table (logsumexp.cpp:20-28)
  qg_stats stats;
  // scratch, grown on demand and reused across calls
  qg_dbuf scratch[48];
  int fb_exact = 0;                  // QG_OPT_FB_EXACT
  cudaEvent_t ev_sync = nullptr;     // blocking host waits (qg_sync)
  // stage timers do not block: the event pairs wait here until the next host synchronisation resolves them
  struct pending_timer { cudaEvent_t a, b; double* acc; };
  std::vector<pending_timer> timers;
  std::vector<cudaEvent_t> ev_pool;
  void* h_pinned = nullptr;          // pinned staging for device-to-host copies
  size_t h_pinned_cap = 0, h_pinned_used = 0;
  struct pending_fetch { void* dst; size_t off, bytes; };
  std::vector<pending_fetch> fetches;  // copies queued into h_pinned, delivered by qg_fetch_wait after ONE host synchronisation
};

static inline int qg_reserve (qg_ctx* ctx, qg_dbuf& b, size_t bytes) {
  if (bytes <= b.cap) return QG_OK;
  if (b.p) { QG_CUDA (ctx, cudaFree (b.p)); b.p = nullptr; b.cap = 0; }
  size_t want = bytes + bytes / 8 + 256;
  QG_CUDA (ctx, cudaMalloc (&b.p, want));
  b.cap = want;
  return QG_OK;
}

static inline uint64_t qg_pow4 (int k) { uint64_t n = 1; while (k-- > 0) n *= 4; return n; }

#endif
