// libquaffgpu: C ABI (include/quaffgpu.h) and host orchestration of the sm_100a kernels.
// Compiled by nvcc for sm_100a (see quaff_b200/build.py).  The same translation unit also compiles
// under g++ with -DQG_EMU against tests/emu/cuda_emu.h, which is test infrastructure only.
#include "qg_common.cuh"
#include "qg_seed.cuh"
#include "qg_dp.cuh"
#include "qg_backward.cuh"
#include "qg_overlap.cuh"
#include "qg_prob.cuh"
#include "qg_tile.cuh"
#include "qg_vit.cuh"
#ifndef QG_EMU
#include <cub/device/device_segmented_radix_sort.cuh>
#endif
#include <map>
#include <functional>
#include <numeric>

static thread_local qg_error g_create_error;     // per thread: qg_last_error (NULL) reports the calling thread's last failed qg_create

enum {
  SC_PAIRDESC = 0, SC_ITEMS, SC_ITEMRUNS, SC_ITEMNRUNS, SC_PAIRRUNS, SC_PAIRINFO, SC_PAIRCU, SC_FLAGS,
  SC_SEGS, SC_RPJOBS, SC_RP, SC_TRACE, SC_ENDVALS, SC_PAIRDP, SC_OUT0, SC_OUT1, SC_OUT2, SC_OUT3, SC_PATHSCR, SC_PATHOUT,
  SC_STORE, SC_ROWACC, SC_MISC0, SC_MISC1, SC_RQ, SC_RS, SC_ENDEX, SC_STOREEX, SC_ZM, SC_ZE,
  SC_RPS, SC_XC64, SC_YC64, SC_KEYS0, SC_KEYS1, SC_VALS0, SC_VALS1, SC_IDXJOBS, SC_SEGOFF, SC_SORTTMP, SC_TILEJOBS
};

// ---- small helpers -----------------------------------------------------------------------------------
// QG_HOST_TRACE=1: host wall-clock between the marked points of an align call, on stderr (diagnostics only)
#include <chrono>
static void qg_htrace (const char* label) {
  static const bool on = getenv ("QG_HOST_TRACE") != nullptr;
  if (!on) return;
  static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now ();
  const auto now = std::chrono::steady_clock::now ();
  fprintf (stderr, "[qg %p] %-14s +%.2f ms\n", (void*) &last, label, std::chrono::duration<double, std::milli> (now - last).count ());
  last = now;
}

// Host waits sleep on a blocking event instead of spinning in cudaStreamSynchronize: several contexts (host threads) share
// one GPU, and spinning waiters burn the cores (and any CPU quota) the other contexts' host work needs.
static cudaEvent_t qg_event_get (qg_ctx* ctx) {
  if (!ctx->ev_pool.empty ()) { cudaEvent_t e = ctx->ev_pool.back (); ctx->ev_pool.pop_back (); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate (&e);
  return e;
}
// the stream has just been synchronised: every recorded timer event is complete
static void qg_resolve_timers (qg_ctx* ctx) {
  for (auto& t : ctx->timers) {
    float ms = 0;
    if (t.a && t.b && cudaEventElapsedTime (&ms, t.a, t.b) == cudaSuccess) *t.acc += ms;
    if (t.a) ctx->ev_pool.push_back (t.a);
    if (t.b) ctx->ev_pool.push_back (t.b);
  }
  ctx->timers.clear ();
}
static cudaError_t qg_sync (qg_ctx* ctx) {
  static const bool spin = getenv ("QG_SPIN_SYNC") != nullptr;
  cudaError_t e;
  if (spin) e = cudaStreamSynchronize (ctx->stream);
  else {
    e = cudaEventRecord (ctx->ev_sync, ctx->stream);
    if (e != cudaSuccess) return e;
    e = cudaEventSynchronize (ctx->ev_sync);
  }
  if (e == cudaSuccess) qg_resolve_timers (ctx);
  return e;
}
static int qg_upload (qg_ctx* ctx, qg_dbuf& b, const void* src, size_t bytes) {
  QG_TRY (qg_reserve (ctx, b, bytes));
  if (bytes) QG_CUDA (ctx, cudaMemcpyAsync (b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return QG_OK;
}
// Device-to-host results: qg_fetch queues a copy into the pinned staging buffer, qg_fetch_wait synchronises ONCE and hands
// every queued piece to its destination (a pageable cudaMemcpyAsync would stall the host per call).
static int qg_fetch_wait (qg_ctx* ctx) {
  QG_CUDA (ctx, qg_sync (ctx));
  for (const auto& f : ctx->fetches) memcpy (f.dst, (const char*) ctx->h_pinned + f.off, f.bytes);
  ctx->fetches.clear ();
  ctx->h_pinned_used = 0;
  return QG_OK;
}
static int qg_fetch (qg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!bytes) return QG_OK;
  const size_t need = (bytes + 255) & ~(size_t) 255;
  if (ctx->h_pinned_used + need > ctx->h_pinned_cap) {
    if (!ctx->fetches.empty ()) QG_TRY (qg_fetch_wait (ctx));      // deliver what is queued before the buffer moves
    if (need > ctx->h_pinned_cap) {
      if (ctx->h_pinned) { cudaFreeHost (ctx->h_pinned); ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0; }
      const size_t want = std::max<size_t> (need + need / 4, (size_t) 1 << 20);
      QG_CUDA (ctx, cudaMallocHost (&ctx->h_pinned, want));
      ctx->h_pinned_cap = want;
    }
  }
  QG_CUDA (ctx, cudaMemcpyAsync ((char*) ctx->h_pinned + ctx->h_pinned_used, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  qg_ctx::pending_fetch f; f.dst = dst; f.off = ctx->h_pinned_used; f.bytes = bytes;
  ctx->fetches.push_back (f);
  ctx->h_pinned_used += need;
  return QG_OK;
}
// a malloc'd host buffer that is freed on every return path unless it is handed to the caller (release)
struct qg_hostbuf {
  uint8_t* p = nullptr;
  ~qg_hostbuf () { free (p); }
  uint8_t* release () { uint8_t* r = p; p = nullptr; return r; }
  qg_hostbuf () = default;
  qg_hostbuf (const qg_hostbuf&) = delete;
  qg_hostbuf& operator= (const qg_hostbuf&) = delete;
};

static int qg_download (qg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  QG_TRY (qg_fetch (ctx, dst, src, bytes));
  return qg_fetch_wait (ctx);
}
static int qg_check_launch (qg_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError ();
  if (e != cudaSuccess) QG_FAIL (ctx, QG_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString (e));
  ctx->stats.kernel_launches += 1;
  return QG_OK;
}
// device time of a stage, by CUDA events on the launching stream; never blocks the host (resolved at the next qg_sync)
struct qg_timer {
  qg_ctx* ctx; double* acc; cudaEvent_t a;
  qg_timer (qg_ctx* c, double* acc_) : ctx (c), acc (acc_), a (qg_event_get (c)) { if (a) cudaEventRecord (a, ctx->stream); }
  ~qg_timer () {
    cudaEvent_t b = qg_event_get (ctx);
    if (b) cudaEventRecord (b, ctx->stream);
    qg_ctx::pending_timer t; t.a = a; t.b = b; t.acc = acc;
    ctx->timers.push_back (t);
  }
};
// the launch classes of one DP stage (different R / warp counts) are independent: run them side by side
static int qg_fork (qg_ctx* ctx) { QG_CUDA (ctx, cudaEventRecord (ctx->ev_fork, ctx->stream)); return QG_OK; }
static int qg_side (qg_ctx* ctx, int k, cudaStream_t* out) {
  *out = ctx->side[k & 7];
  if (k < 8) QG_CUDA (ctx, cudaStreamWaitEvent (*out, ctx->ev_fork, 0));
  return QG_OK;
}
static int qg_join (qg_ctx* ctx, int n) {
  for (int i = 0; i < n && i < 8; ++i) {
    QG_CUDA (ctx, cudaEventRecord (ctx->ev_join[i], ctx->side[i]));
    QG_CUDA (ctx, cudaStreamWaitEvent (ctx->stream, ctx->ev_join[i], 0));
  }
  return QG_OK;
}
static size_t qg_env_size (const char* name, size_t dflt) {
  const char* v = getenv (name);
  return (v && *v) ? (size_t) strtoull (v, nullptr, 10) : dflt;
}

// ---- context ---------------------------------------------------------------------------------------------
extern "C" int qg_abi_version (void) { return 1; }

extern "C" const char* qg_last_error (const qg_ctx* ctx) { return ctx ? ctx->err.msg.c_str () : g_create_error.msg.c_str (); }

extern "C" void qg_free (void* p) { free (p); }

extern "C" size_t qg_counts_size (int match_k, int gap_k) {
  return 4 * (size_t) qg_pow4 (match_k) * QG_NQUAL + 4 * QG_NQUAL + 4 * (size_t) qg_pow4 (gap_k) + 4;
}

extern "C" int qg_create (qg_ctx** out, int device) {
  if (!out) return QG_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount (&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_create_error.code = QG_ERR_NO_DEVICE;
    g_create_error.msg = std::string ("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString (e) : "device count 0") +
                         "); libquaffgpu has no CPU path";
    return QG_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { g_create_error.code = QG_ERR_INVALID; g_create_error.msg = "device index out of range"; return QG_ERR_INVALID; }
  qg_ctx* ctx = new qg_ctx ();
  ctx->device = device;
  memset (&ctx->stats, 0, sizeof (ctx->stats));
  auto fail = [&] (const char* what, cudaError_t ce) {
    g_create_error.code = QG_ERR_CUDA; g_create_error.msg = std::string (what) + ": " + cudaGetErrorString (ce);
    delete ctx; return QG_ERR_CUDA; };
  if ((e = cudaSetDevice (device)) != cudaSuccess) return fail ("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties (&prop, device)) != cudaSuccess) return fail ("cudaGetDeviceProperties", e);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin > 2048 ? prop.sharedMemPerBlockOptin - 1024 : prop.sharedMemPerBlockOptin;   // dynamic budget: leave room for the kernels' static shared memory
  if ((e = cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail ("cudaStreamCreate", e);
  if ((e = cudaEventCreateWithFlags (&ctx->ev[0], cudaEventBlockingSync)) != cudaSuccess) return fail ("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags (&ctx->ev[1], cudaEventBlockingSync)) != cudaSuccess) return fail ("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags (&ctx->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return fail ("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags (&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail ("cudaEventCreate", e);
  for (int i = 0; i < 8; ++i) {
    if ((e = cudaStreamCreateWithFlags (&ctx->side[i], cudaStreamNonBlocking)) != cudaSuccess) return fail ("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags (&ctx->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return fail ("cudaEventCreate", e);
  }
  // the reference's log-sum-exp table, built with the host libm exactly as logsumexp.cpp:20-28 does
  {
    const int n = ((int) (10 / .0001)) + 1;
    std::vector<double> tab (n + 1);
    for (int t = 0; t < n; ++t) { const double x = t * .0001; tab[t] = log (1. + exp (-x)); }
    tab[n] = 0;
    if (qg_upload (ctx, ctx->d_lse, tab.data (), sizeof (double) * (n + 1)) != QG_OK || qg_sync (ctx) != cudaSuccess) {
      g_create_error = ctx->err; delete ctx; return QG_ERR_CUDA; }
  }
  *out = ctx;
  return QG_OK;
}

extern "C" void qg_destroy (qg_ctx* ctx) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx) return;
  cudaSetDevice (ctx->device);
  qg_sync (ctx);
  auto rel = [] (qg_dbuf& b) { if (b.p) cudaFree (b.p); b.p = nullptr; b.cap = 0; };
  for (auto& s : ctx->seqs) { rel (s.d_tok); rel (s.d_qual); rel (s.d_off); rel (s.d_packed); rel (s.d_poff); rel (s.d_codes); rel (s.d_sorted); }
  rel (ctx->model.d_match); rel (ctx->model.d_insert); rel (ctx->model.d_gap);
  rel (ctx->omodel.d_match); rel (ctx->omodel.d_insert); rel (ctx->omodel.d_m2m); rel (ctx->omodel.d_m2i); rel (ctx->omodel.d_m2d);
  for (int s = 0; s < 2; ++s) { rel (ctx->omodel.d_pair[s]); rel (ctx->omodel.d_xonly[s]); rel (ctx->omodel.d_yonly[s]); rel (ctx->omodel.d_none[s]); }
  rel (ctx->d_lse);
  if (ctx->h_pinned) cudaFreeHost (ctx->h_pinned);
  for (auto& b : ctx->scratch) rel (b);
  cudaEventDestroy (ctx->ev[0]); cudaEventDestroy (ctx->ev[1]);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy (e);
  cudaEventDestroy (ctx->ev_fork); if (ctx->ev_sync) cudaEventDestroy (ctx->ev_sync);
  for (int i = 0; i < 8; ++i) { cudaStreamDestroy (ctx->side[i]); cudaEventDestroy (ctx->ev_join[i]); }
  cudaStreamDestroy (ctx->stream);
  delete ctx;
}

extern "C" int qg_set_option (qg_ctx* ctx, int option, int64_t value) {
  if (!ctx) return QG_ERR_INVALID;
  if (option == QG_OPT_FB_EXACT) { ctx->fb_exact = value ? 1 : 0; return QG_OK; }
  QG_FAIL (ctx, QG_ERR_INVALID, "unknown option %d", option);
}

extern "C" int qg_get_stats (qg_ctx* ctx, qg_stats* out, int reset) {
  if (!ctx || !out) return QG_ERR_INVALID;
  cudaSetDevice (ctx->device);
  QG_CUDA (ctx, qg_sync (ctx));                              // resolves the pending stage timers
  *out = ctx->stats;
  if (reset) memset (&ctx->stats, 0, sizeof (ctx->stats));
  return QG_OK;
}

// ---- inputs ----------------------------------------------------------------------------------------------
extern "C" int qg_set_seqs (qg_ctx* ctx, int which, size_t n, const uint8_t* tok, const uint8_t* qual, const uint64_t* offsets) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx) return QG_ERR_INVALID;
  if (which != QG_REFS && which != QG_READS) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: unknown sequence set %d", which);
  if (!offsets || (n && !tok)) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: null pointer");
  if (n >= 0xFFFFFFFFull) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: too many sequences");
  qg_seqset& s = ctx->seqs[which];
  s.n = n;
  s.off.assign (offsets, offsets + n + 1);
  s.total = s.off[n] - s.off[0];
  if (s.off[0] != 0) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: offsets[0] must be 0");
  s.max_len = 0;
  for (size_t i = 0; i < n; ++i) {
    if (s.off[i + 1] < s.off[i]) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: offsets not monotone at %zu", i);
    const uint64_t l = s.off[i + 1] - s.off[i];
    if (l > 0x7FFFFFF0ull) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: sequence %zu too long", i);
    s.max_len = std::max (s.max_len, (uint32_t) l);
  }
  for (uint64_t t = 0; t < s.total; ++t)
    if (tok[t] > 3) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: token %u at position %llu is not in {0,1,2,3} (the reference aborts on non-ACGT, fastseq.cpp:76-79)", tok[t], (unsigned long long) t);
  s.has_qual = qual != nullptr;
  if (qual) for (uint64_t t = 0; t < s.total; ++t) if (qual[t] >= QG_NQUAL) QG_FAIL (ctx, QG_ERR_INVALID, "qg_set_seqs: quality score %u out of range", qual[t]);
  s.h_tok.assign (tok, tok + s.total);
  if (qual) s.h_qual.assign (qual, qual + s.total); else s.h_qual.clear ();
  s.codes_k = 0; s.sorted_k = 0;
  // packed layout: each sequence starts on a word boundary and is followed by one zero word
  s.poff.resize (n + 1);
  uint64_t w = 0;
  for (size_t i = 0; i < n; ++i) { s.poff[i] = w; w += (s.off[i + 1] - s.off[i] + 31) / 32 + 1; }
  s.poff[n] = w;
  {
    qg_timer tm (ctx, &ctx->stats.ms_h2d);
    QG_TRY (qg_upload (ctx, s.d_tok, tok, s.total));
    if (qual) QG_TRY (qg_upload (ctx, s.d_qual, qual, s.total));
    QG_TRY (qg_upload (ctx, s.d_off, s.off.data (), sizeof (uint64_t) * (n + 1)));
    QG_TRY (qg_upload (ctx, s.d_poff, s.poff.data (), sizeof (uint64_t) * (n + 1)));
    QG_TRY (qg_reserve (ctx, s.d_packed, sizeof (uint64_t) * (w + 1)));
  }
  if (w) {
    qg_timer tm (ctx, &ctx->stats.ms_prep);
    QG_LAUNCH (qg_pack_kernel, (unsigned) ((w + 255) / 256), 256, 0, ctx->stream,
               s.d_tok.as<uint8_t> (), s.d_off.as<uint64_t> (), s.d_poff.as<uint64_t> (), (uint32_t) n, w, s.d_packed.as<uint64_t> ());
    QG_TRY (qg_check_launch (ctx, "qg_pack_kernel"));
  }
  QG_CUDA (ctx, qg_sync (ctx));
  return QG_OK;
}

static int qg_ensure_codes (qg_ctx* ctx, int which, int k) {
  qg_seqset& s = ctx->seqs[which];
  if (s.codes_k == k) return QG_OK;
  QG_TRY (qg_reserve (ctx, s.d_codes, sizeof (uint16_t) * (s.total + 1)));
  if (s.total) {
    qg_timer tm (ctx, &ctx->stats.ms_prep);
    QG_LAUNCH (qg_codes_kernel, (unsigned) ((s.total + 255) / 256), 256, 0, ctx->stream,
               s.d_tok.as<uint8_t> (), s.d_off.as<uint64_t> (), (uint32_t) s.n, s.total, k, s.d_codes.as<uint16_t> ());
    QG_TRY (qg_check_launch (ctx, "qg_codes_kernel"));
  }
  s.codes_k = k;
  return QG_OK;
}

// the static side of the tile-sorted seeding kernel: every QG_TILE_POS positions of every sequence sorted by k-mer code
static int qg_ensure_sorted (qg_ctx* ctx, int which, int k) {
  qg_seqset& s = ctx->seqs[which];
  QG_TRY (qg_ensure_codes (ctx, which, k));
  if (s.sorted_k == k) return QG_OK;
  std::vector<qg_tile_job> jobs;
  for (size_t i = 0; i < s.n; ++i)
    for (uint64_t b = s.off[i]; b < s.off[i + 1]; b += QG_TILE_POS) {
      qg_tile_job jb; jb.off = b; jb.len = (uint32_t) std::min<uint64_t> (QG_TILE_POS, s.off[i + 1] - b); jb.pad_ = 0;
      jobs.push_back (jb);
    }
  QG_TRY (qg_reserve (ctx, s.d_sorted, sizeof (uint32_t) * (s.total + 1)));
  if (!jobs.empty ()) {
    qg_timer tm (ctx, &ctx->stats.ms_prep);
    QG_TRY (qg_upload (ctx, ctx->scratch[SC_TILEJOBS], jobs.data (), sizeof (qg_tile_job) * jobs.size ()));
    QG_LAUNCH (qg_sort_tiles_kernel, (unsigned) jobs.size (), 256, 0, ctx->stream,
               ctx->scratch[SC_TILEJOBS].as<qg_tile_job> (), s.d_codes.as<uint16_t> (), 1u << (2 * k), s.d_sorted.as<uint32_t> ());
    QG_TRY (qg_check_launch (ctx, "qg_sort_tiles_kernel"));
  }
  s.sorted_k = k;
  return QG_OK;
}

extern "C" int qg_set_align_model (qg_ctx* ctx, const qg_align_model* m) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !m) return QG_ERR_INVALID;
  if (m->match_k < 1 || m->match_k > 6 || m->gap_k < 0 || m->gap_k > 6) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "model orders K=%d G=%d outside 1..6 / 0..6", m->match_k, m->gap_k);
  qg_model_dev& d = ctx->model;
  d.match_k = m->match_k; d.gap_k = m->gap_k; d.nK = qg_pow4 (m->match_k); d.nG = qg_pow4 (m->gap_k);
  QG_TRY (qg_upload (ctx, d.d_match, m->match, sizeof (double) * 4 * d.nK * QG_NQ1));
  QG_TRY (qg_upload (ctx, d.d_insert, m->insert, sizeof (double) * 4 * QG_NQ1));
  std::vector<double> gap (4 * d.nG);
  for (uint64_t g = 0; g < d.nG; ++g) { gap[g] = m->m2m[g]; gap[d.nG + g] = m->m2i[g]; gap[2 * d.nG + g] = m->m2d[g]; gap[3 * d.nG + g] = m->m2e[g]; }
  QG_TRY (qg_upload (ctx, d.d_gap, gap.data (), sizeof (double) * gap.size ()));
  d.d2d = m->d2d; d.d2m = m->d2m; d.i2i = m->i2i; d.i2m = m->i2m;
  QG_CUDA (ctx, qg_sync (ctx));
  d.set = true;
  return QG_OK;
}

// ---- host arithmetic that the reference also does once per run on the host ---------------------------------
static double qg_log_negbinom (int k, double p, double r) {
  // log(gsl_ran_negative_binomial_pdf(k,p,r)), negbinom.cpp:30-32
  const double f = lgamma (k + r), a = lgamma (r), b = lgamma (k + 1.0);
  return log (exp (f - a - b) * pow (p, r) * pow (1 - p, (double) k));
}

extern "C" int qg_scores_from_params (const qg_params* qp, double* match, double* insert,
                                      double* m2m, double* m2i, double* m2d, double* m2e, double* scal4) {
  if (!qp || !match || !insert || !m2m || !m2i || !m2d || !m2e || !scal4) return QG_ERR_INVALID;
  const uint64_t nK = qg_pow4 (qp->match_k), nG = qg_pow4 (qp->gap_k);
  auto fill = [] (const double* pqr, double* out95) {       // SymQualScores, qmodel.cpp:87-93
    const double lsp = log (pqr[0]);
    for (int k = 0; k < QG_NQUAL; ++k) out95[k] = lsp + qg_log_negbinom (k, pqr[1], pqr[2]);
    out95[QG_NQUAL] = lsp;
  };
  for (int i = 0; i < 4; ++i) {
    fill (qp->insert_pqr + 3 * i, insert + i * QG_NQ1);
    for (uint64_t j = 0; j < nK; ++j) fill (qp->match_pqr + 3 * (i * nK + j), match + (i * nK + j) * QG_NQ1);
  }
  for (uint64_t j = 0; j < nG; ++j) {                        // QuaffScores, qmodel.cpp:312-318
    m2m[j] = log (1 - qp->begin_insert[j]) + log (1 - qp->begin_delete[j]);
    m2i[j] = log (qp->begin_insert[j]);
    m2d[j] = log (1 - qp->begin_insert[j]) + log (qp->begin_delete[j]);
    m2e[j] = log (qp->begin_insert[j]);
  }
  scal4[0] = log (qp->extend_delete); scal4[1] = log (1 - qp->extend_delete);
  scal4[2] = log (qp->extend_insert); scal4[3] = log (1 - qp->extend_insert);
  return QG_OK;
}

extern "C" double qg_null_loglike (double null_emit, const double* null_pqr, const uint8_t* tok, const uint8_t* qual, uint64_t len) {
  // qmodel.cpp:1875-1890: the same addends in the same order (so the same bits); the two per-base terms are functions of
  // (token, quality) only and are tabulated once per call instead of evaluated per base (three lgamma + two pow each)
  double lsym[4], lq[4][QG_NQUAL];
  for (int t = 0; t < 4; ++t) {
    const double* d = null_pqr + 3 * t;
    lsym[t] = log (d[0]);
    if (qual) for (int k = 0; k < QG_NQUAL; ++k) lq[t][k] = qg_log_negbinom (k, d[1], d[2]);
  }
  double ll = len * log (null_emit) + log (1. - null_emit);
  for (uint64_t i = 0; i < len; ++i) {
    ll += lsym[tok[i]];
    if (qual) ll += lq[tok[i]][qual[i]];
  }
  return ll;
}

// ---- envelope stage ------------------------------------------------------------------------------------------
struct qg_env_result {
  std::vector<uint32_t> run_begin;   // [n_pairs+1] into runs
  std::vector<int2> runs;
  std::vector<uint64_t> cu;          // [n_pairs]
  std::vector<uint32_t> ndiag;       // [n_pairs]
};

// which seeding kernel serves this pair list: the shared-memory histogram kernel (k in 5..7, index + ring fit one CTA)
// or the general path (sorted read index + per-diagonal counters in HBM)
static size_t qg_seed_smem_bytes (int k, uint32_t ymax, uint32_t* ring_out) {
  const uint32_t nk = 1u << (2 * k);
  uint32_t need = std::max<uint32_t> (nk, QG_SEED_STEP + ymax + 4), ring = 1;
  while (ring < need) ring <<= 1;
  if (ring_out) *ring_out = ring;
  return (size_t) ring * 4 + (size_t) nk * 8 + (size_t) ((ymax + 2) & ~1u) * 2 + (QG_SEED_STEP / 32 + 2) * 4;
}
// the tile-sorted kernel (qg_seed_tile_kernel): k = 5 or 6, every read's diagonals fit the 32k-counter ring next to one tile
static size_t qg_tseed_smem_bytes (int k, uint32_t ymax) {
  const uint32_t nk = 1u << (2 * k);
  return (size_t) QG_TSEED_RING * 4 + (size_t) (nk + 1) * 8 + (size_t) ((nk + 2 + 3) & ~3u) * 2 + (size_t) ((ymax + 2 + 1) & ~1u) * 2 + (QG_TSEED_ESTEP / 32 + 2) * 4 + 32 * 4;
}
static bool qg_seed_use_tiles (qg_ctx* ctx, int k, uint32_t ymax) {
  if (getenv ("QG_SEED_LEGACY")) return false;                 // tests / profiling: the round-1 kernel
  return (k == 5 || k == 6) && ymax >= (uint32_t) k && ymax - (uint32_t) k + 3 + QG_TILE_POS <= QG_TSEED_RING && qg_tseed_smem_bytes (k, ymax) <= ctx->smem_optin;
}
static bool qg_pair_is_sparse (const qg_dpconfig* cfg, uint32_t xlen, uint32_t ylen) {
  if (!cfg->sparse) return false;
  if (cfg->kmer_threshold >= 0) {                              // diagenv.cpp:23-29
    const uint32_t min_len = 2u * (uint32_t) (cfg->kmer_len + cfg->kmer_threshold);
    if (xlen < min_len || ylen < min_len) return false;
  }
  return true;
}
static bool qg_seed_is_general (qg_ctx* ctx, const qg_dpconfig* cfg, int x_set, size_t n_pairs, const uint32_t* xi, const uint32_t* yi) {
  const qg_seqset& X = ctx->seqs[x_set];
  const qg_seqset& Y = ctx->seqs[QG_READS];
  uint32_t ymax = 0; bool any = false;
  for (size_t p = 0; p < n_pairs; ++p)
    if (xi[p] < X.n && yi[p] < Y.n && qg_pair_is_sparse (cfg, X.len (xi[p]), Y.len (yi[p]))) { any = true; ymax = std::max (ymax, Y.len (yi[p])); }
  if (!any) return false;
  if (getenv ("QG_SEED_GENERAL")) return true;                 // tests: force the general path
  const int k = cfg->kmer_len;
  return k < 5 || k > 7 || ymax > 65000 || qg_seed_smem_bytes (k, ymax, nullptr) > ctx->smem_optin;
}

// per-read sorted (k-mer, start position) index for the general path; keys/vals end up in SC_KEYS1 / SC_VALS1
static int qg_build_read_index (qg_ctx* ctx, int k, const std::vector<qg_index_job>& jobs, uint64_t total) {
  qg_seqset& Y = ctx->seqs[QG_READS];
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_YC64], sizeof (unsigned long long) * (Y.total + 1)));
  QG_LAUNCH (qg_codes64_kernel, (unsigned) ((Y.total + 255) / 256), 256, 0, ctx->stream,
             Y.d_tok.as<uint8_t> (), Y.d_off.as<uint64_t> (), (uint32_t) Y.n, Y.total, k, ctx->scratch[SC_YC64].as<unsigned long long> ());
  QG_TRY (qg_check_launch (ctx, "qg_codes64_kernel"));
  QG_TRY (qg_upload (ctx, ctx->scratch[SC_IDXJOBS], jobs.data (), sizeof (qg_index_job) * jobs.size ()));
  for (int b = SC_KEYS0; b <= SC_KEYS1; ++b) QG_TRY (qg_reserve (ctx, ctx->scratch[b], sizeof (unsigned long long) * (total + 1)));
  for (int b = SC_VALS0; b <= SC_VALS1; ++b) QG_TRY (qg_reserve (ctx, ctx->scratch[b], sizeof (uint32_t) * (total + 1)));
  uint32_t nmax = 0;
  std::vector<uint32_t> segoff (2 * jobs.size ());
  for (size_t t = 0; t < jobs.size (); ++t) { nmax = std::max (nmax, jobs[t].n); segoff[t] = (uint32_t) jobs[t].idx_off; segoff[jobs.size () + t] = (uint32_t) (jobs[t].idx_off + jobs[t].n); }
  QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGOFF], segoff.data (), sizeof (uint32_t) * segoff.size ()));
  dim3 grid ((nmax + 255) / 256, (unsigned) jobs.size ());
  QG_LAUNCH (qg_index_fill_kernel, grid, 256, 0, ctx->stream, ctx->scratch[SC_IDXJOBS].as<qg_index_job> (),
             ctx->scratch[SC_YC64].as<unsigned long long> (), ctx->scratch[SC_KEYS0].as<unsigned long long> (), ctx->scratch[SC_VALS0].as<uint32_t> ());
  QG_TRY (qg_check_launch (ctx, "qg_index_fill_kernel"));
#ifdef QG_EMU
  QG_CUDA (ctx, qg_sync (ctx));
  {
    const unsigned long long* k0 = ctx->scratch[SC_KEYS0].as<unsigned long long> (); const uint32_t* v0 = ctx->scratch[SC_VALS0].as<uint32_t> ();
    unsigned long long* k1 = ctx->scratch[SC_KEYS1].as<unsigned long long> (); uint32_t* v1 = ctx->scratch[SC_VALS1].as<uint32_t> ();
    for (const qg_index_job& jb : jobs) {
      std::vector<uint32_t> perm (jb.n);
      for (uint32_t j = 0; j < jb.n; ++j) perm[j] = j;
      std::stable_sort (perm.begin (), perm.end (), [&] (uint32_t a, uint32_t b) { return k0[jb.idx_off + a] < k0[jb.idx_off + b]; });
      for (uint32_t j = 0; j < jb.n; ++j) { k1[jb.idx_off + j] = k0[jb.idx_off + perm[j]]; v1[jb.idx_off + j] = v0[jb.idx_off + perm[j]]; }
    }
  }
#else
  {
    const uint32_t* segb = ctx->scratch[SC_SEGOFF].as<uint32_t> ();
    const uint32_t* sege = segb + jobs.size ();
    size_t tmp = 0;
    QG_CUDA (ctx, cub::DeviceSegmentedRadixSort::SortPairs (nullptr, tmp, ctx->scratch[SC_KEYS0].as<unsigned long long> (), ctx->scratch[SC_KEYS1].as<unsigned long long> (),
                  ctx->scratch[SC_VALS0].as<uint32_t> (), ctx->scratch[SC_VALS1].as<uint32_t> (), (int) total, (int) jobs.size (), segb, sege, 0, 2 * k, ctx->stream));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_SORTTMP], tmp + 16));
    QG_CUDA (ctx, cub::DeviceSegmentedRadixSort::SortPairs (ctx->scratch[SC_SORTTMP].p, tmp, ctx->scratch[SC_KEYS0].as<unsigned long long> (), ctx->scratch[SC_KEYS1].as<unsigned long long> (),
                  ctx->scratch[SC_VALS0].as<uint32_t> (), ctx->scratch[SC_VALS1].as<uint32_t> (), (int) total, (int) jobs.size (), segb, sege, 0, 2 * k, ctx->stream));
  }
#endif
  return QG_OK;
}

static int qg_envelope_stage_cap (qg_ctx* ctx, const qg_dpconfig* cfg, uint64_t cell_size, int x_set,
                                  size_t n_pairs, const uint32_t* xi, const uint32_t* yi, qg_env_result& out, uint32_t run_cap, bool* overflow) {
  *overflow = false;
  const qg_seqset& X = ctx->seqs[x_set];
  const qg_seqset& Y = ctx->seqs[QG_READS];
  const int k = cfg->kmer_len;
  std::vector<qg_pair_desc> pd (n_pairs);
  std::vector<qg_seed_item> items;
  bool any_sparse = false;
  // diagonals per work item: large enough that the per-item index build is amortised, small enough to fill the GPU
  uint64_t total_diags = 0;
  for (size_t p = 0; p < n_pairs; ++p) if (xi[p] < X.n && yi[p] < Y.n) total_diags += (uint64_t) X.len (xi[p]) + Y.len (yi[p]);
  int64_t chunk = (int64_t) (total_diags / (uint64_t) (16 * std::max (ctx->sm_count, 1)));
  const int64_t chunk_min = (int64_t) qg_env_size ("QG_SEED_CHUNK", QG_SEED_CHUNK);      // lowered by the tests only: several items per pair on short inputs
  chunk = std::max<int64_t> (chunk_min, std::min<int64_t> (chunk, 8 * chunk_min));
  chunk = std::max<int64_t> (QG_SEED_STEP, (chunk / QG_SEED_STEP) * QG_SEED_STEP);
  const bool memory_mode = cfg->kmer_threshold < 0;
  if (cfg->sparse && (k < 1 || k > 32)) QG_FAIL (ctx, QG_ERR_INVALID, "-kmatch %d: k-mer length must be 1..32", k);
  const bool general = qg_seed_is_general (ctx, cfg, x_set, n_pairs, xi, yi);
  const int64_t gen_chunk = 32768;                               // reference positions per work item on the general path
  uint32_t ymax_sparse = 0;
  for (size_t p = 0; p < n_pairs; ++p)
    if (xi[p] < X.n && yi[p] < Y.n && qg_pair_is_sparse (cfg, X.len (xi[p]), Y.len (yi[p]))) ymax_sparse = std::max (ymax_sparse, Y.len (yi[p]));
  const bool tiles = !general && cfg->sparse && qg_seed_use_tiles (ctx, k, ymax_sparse);
  if (tiles) chunk = std::max<int64_t> (QG_TILE_POS, (chunk / QG_TILE_POS) * QG_TILE_POS);   // item boundaries on tile boundaries
  std::vector<qg_index_job> idx_jobs; std::map<uint32_t, uint64_t> idx_of_read; uint64_t idx_total = 0;
  std::vector<uint32_t> mem_pairs;
  uint64_t count_total = 0, bits_total = 0, hist_total = 0;
  uint32_t ymax = 0;
  uint64_t run_total = 0;
  for (size_t p = 0; p < n_pairs; ++p) {
    if (xi[p] >= X.n || yi[p] >= Y.n) QG_FAIL (ctx, QG_ERR_INVALID, "pair %zu: sequence index out of range", p);
    qg_pair_desc& d = pd[p];
    d.xseq = xi[p]; d.yseq = yi[p]; d.xlen = X.len (xi[p]); d.ylen = Y.len (yi[p]);
    d.xoff = X.off[xi[p]]; d.yoff = Y.off[yi[p]];
    if (d.xlen == 0 || d.ylen == 0) QG_FAIL (ctx, QG_ERR_INVALID, "pair %zu: empty sequence", p);
    const bool full = !qg_pair_is_sparse (cfg, d.xlen, d.ylen);
    d.full = full ? 1 : 0;
    d.idx_off = 0;
    d.item_begin = (uint32_t) items.size ();
    d.run_cap = 0; d.pad_ = 0; d.count_off = 0; d.bits_off = 0; d.hist_off = 0;
    uint64_t runs_here = 1;
    if (!full) {
      if (d.xlen < (uint32_t) k || d.ylen < (uint32_t) k)
        QG_FAIL (ctx, QG_ERR_PRECONDITION, "pair %zu: sequence shorter than k=%d (the reference's KmerIndex underflows here, fastseq.cpp:247)", p, k);
      any_sparse = true;
      ymax = std::max (ymax, d.ylen);
      if (general) {
        for (int64_t b = 0; b <= (int64_t) d.xlen - k; b += gen_chunk) {           // items over reference positions
          qg_seed_item it; it.pair = (uint32_t) p; it.d_begin = (int32_t) b; it.d_end = (int32_t) std::min<int64_t> (b + gen_chunk, (int64_t) d.xlen - k + 1);
          items.push_back (it);
        }
        auto f = idx_of_read.find (d.yseq);
        if (f == idx_of_read.end ()) {
          qg_index_job jb; jb.yoff = d.yoff; jb.idx_off = idx_total; jb.n = d.ylen - (uint32_t) k + 1; jb.pad_ = 0;
          idx_jobs.push_back (jb);
          f = idx_of_read.insert (std::make_pair (d.yseq, idx_total)).first;
          idx_total += jb.n;
          if (idx_total > 0x7FFFFFF0ull) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "read index of one seeding batch exceeds 2^31 k-mers; split the pair list");
        }
        d.idx_off = f->second;
      } else {
        const int64_t dmin = -((int64_t) d.ylen - k), dmax = (int64_t) d.xlen - k;   // diagonals that can receive hits
        if (tiles) {
          // [dmin, chunk), [chunk, 2 chunk), ...: every item but the first starts on a tile boundary of the reference
          for (int64_t b = dmin; b <= dmax; ) {
            const int64_t base = b < 0 ? 0 : b, e = base - base % chunk + chunk;
            qg_seed_item it; it.pair = (uint32_t) p; it.d_begin = (int32_t) b; it.d_end = (int32_t) std::min<int64_t> (e, dmax + 1);
            items.push_back (it);
            b = e;
          }
        } else
        for (int64_t b = dmin; b <= dmax; b += chunk) {
          qg_seed_item it; it.pair = (uint32_t) p; it.d_begin = (int32_t) b; it.d_end = (int32_t) std::min<int64_t> (b + chunk, dmax + 1);
          items.push_back (it);
        }
      }
      runs_here = (uint64_t) (items.size () - d.item_begin) * run_cap + 1;
      if (memory_mode || general) {
        d.full = 2;
        mem_pairs.push_back ((uint32_t) p);
        d.count_off = count_total; count_total += (uint64_t) d.xlen + d.ylen;
        d.bits_off = bits_total; bits_total += 3ull * (((uint64_t) d.xlen + d.ylen + 3 + 31) / 32);
        d.hist_off = hist_total; hist_total += (uint64_t) d.ylen + 2;
        runs_here = ((uint64_t) d.xlen + d.ylen) / (2ull * ((unsigned) cfg->band_size / 2) + 2) + 4;
        d.run_cap = (uint32_t) runs_here;
      }
    }
    d.item_end = (uint32_t) items.size ();
    d.run_out = (uint32_t) run_total;
    run_total += runs_here;
    if (run_total > 0xFFFFFFF0ull) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "too many seeding work items in one call; split the pair list");
  }

  qg_dbuf &dPD = ctx->scratch[SC_PAIRDESC], &dIT = ctx->scratch[SC_ITEMS], &dIR = ctx->scratch[SC_ITEMRUNS], &dIN = ctx->scratch[SC_ITEMNRUNS];
  qg_dbuf &dPR = ctx->scratch[SC_PAIRRUNS], &dPI = ctx->scratch[SC_PAIRINFO], &dPC = ctx->scratch[SC_PAIRCU], &dFL = ctx->scratch[SC_FLAGS];
  QG_TRY (qg_upload (ctx, dPD, pd.data (), sizeof (qg_pair_desc) * n_pairs));
  QG_TRY (qg_upload (ctx, dIT, items.data (), sizeof (qg_seed_item) * items.size ()));
  QG_TRY (qg_reserve (ctx, dIR, sizeof (int2) * ((general ? 0 : items.size () * run_cap) + 1)));
  QG_TRY (qg_reserve (ctx, dIN, sizeof (uint32_t) * (items.size () + 1)));
  QG_TRY (qg_reserve (ctx, dPR, sizeof (int2) * (run_total + 1)));
  QG_TRY (qg_reserve (ctx, dPI, sizeof (uint2) * (n_pairs + 1)));
  QG_TRY (qg_reserve (ctx, dPC, sizeof (unsigned long long) * (n_pairs + 1)));
  QG_TRY (qg_reserve (ctx, dFL, 64));
  QG_CUDA (ctx, cudaMemsetAsync (dFL.p, 0, 64, ctx->stream));

  if ((memory_mode || general) && !mem_pairs.empty ()) {
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_STORE], sizeof (uint32_t) * (count_total + 1)));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ROWACC], sizeof (uint32_t) * (hist_total + 1)));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_MISC1], sizeof (uint32_t) * (bits_total + 1)));
    QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], mem_pairs.data (), sizeof (uint32_t) * mem_pairs.size ()));
    QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_STORE].p, 0, sizeof (uint32_t) * (count_total + 1), ctx->stream));
  }
  if (!items.empty () && general) {
    qg_seqset& XS = ctx->seqs[x_set];
    qg_timer tm (ctx, &ctx->stats.ms_seed);
    QG_TRY (qg_build_read_index (ctx, k, idx_jobs, idx_total));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_XC64], sizeof (unsigned long long) * (XS.total + 1)));
    QG_LAUNCH (qg_codes64_kernel, (unsigned) ((XS.total + 255) / 256), 256, 0, ctx->stream,
               XS.d_tok.as<uint8_t> (), XS.d_off.as<uint64_t> (), (uint32_t) XS.n, XS.total, k, ctx->scratch[SC_XC64].as<unsigned long long> ());
    QG_TRY (qg_check_launch (ctx, "qg_codes64_kernel"));
    QG_LAUNCH (qg_seed_general_kernel, (unsigned) items.size (), 256, 0, ctx->stream,
               dIT.as<qg_seed_item> (), dPD.as<qg_pair_desc> (), ctx->scratch[SC_XC64].as<unsigned long long> (),
               ctx->scratch[SC_KEYS1].as<unsigned long long> (), ctx->scratch[SC_VALS1].as<uint32_t> (), k,
               ctx->scratch[SC_STORE].as<uint32_t> (), (unsigned long long*) ((char*) dFL.p + 8));
    QG_TRY (qg_check_launch (ctx, "qg_seed_general_kernel"));
  }
  if (!items.empty ()) {
    if (tiles) {
      QG_TRY (qg_ensure_sorted (ctx, x_set, k));
      QG_TRY (qg_ensure_codes (ctx, QG_READS, k));
      const size_t smem = qg_tseed_smem_bytes (k, ymax);
      QG_CUDA (ctx, cudaFuncSetAttribute (qg_seed_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_optin));
      QG_CUDA (ctx, cudaFuncSetAttribute (qg_seed_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_optin));
      qg_timer tm (ctx, &ctx->stats.ms_seed);
      auto kfn = memory_mode ? qg_seed_tile_kernel<true> : qg_seed_tile_kernel<false>;
      QG_LAUNCH (kfn, (unsigned) items.size (), QG_TSEED_THREADS, smem, ctx->stream,
                 dIT.as<qg_seed_item> (), dPD.as<qg_pair_desc> (), ctx->seqs[x_set].d_sorted.as<uint32_t> (), ctx->seqs[QG_READS].d_codes.as<uint16_t> (),
                 k, cfg->kmer_threshold, (int) ((unsigned) cfg->band_size / 2), ymax, run_cap,
                 dIR.as<int2> (), dIN.as<uint32_t> (), (unsigned long long*) ((char*) dFL.p + 8), (uint32_t*) dFL.p,
                 memory_mode ? ctx->scratch[SC_STORE].as<uint32_t> () : (uint32_t*) nullptr);
      QG_TRY (qg_check_launch (ctx, "qg_seed_tile_kernel"));
    } else if (!general) {
    QG_TRY (qg_ensure_codes (ctx, x_set, k));
    QG_TRY (qg_ensure_codes (ctx, QG_READS, k));
    uint32_t ring = 1;
    const size_t smem = qg_seed_smem_bytes (k, ymax, &ring);
    if (smem > ctx->smem_optin)
      QG_FAIL (ctx, QG_ERR_CUDA, "internal: seeding kernel selection (%zu B of shared memory needed, limit %zu)", smem, ctx->smem_optin);
    // the attribute belongs to the function, not to this context: always the device maximum, so that contexts running
    // on other host threads with other read lengths never shrink it under a launch
    QG_CUDA (ctx, cudaFuncSetAttribute (qg_seed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_optin));
    QG_CUDA (ctx, cudaFuncSetAttribute (qg_seed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_optin));
    {
      qg_timer tm (ctx, &ctx->stats.ms_seed);
      auto kfn = memory_mode ? qg_seed_kernel<true> : qg_seed_kernel<false>;
      QG_LAUNCH (kfn, (unsigned) items.size (), QG_SEED_THREADS, smem, ctx->stream,
                 dIT.as<qg_seed_item> (), dPD.as<qg_pair_desc> (), ctx->seqs[x_set].d_codes.as<uint16_t> (), ctx->seqs[QG_READS].d_codes.as<uint16_t> (),
                 k, cfg->kmer_threshold, (int) ((unsigned) cfg->band_size / 2), ring, ymax, run_cap,
                 dIR.as<int2> (), dIN.as<uint32_t> (), (unsigned long long*) ((char*) dFL.p + 8), (uint32_t*) dFL.p,
                 memory_mode ? ctx->scratch[SC_STORE].as<uint32_t> () : (uint32_t*) nullptr);
      QG_TRY (qg_check_launch (ctx, "qg_seed_kernel"));
    }
    }
    if (memory_mode || general) {
      qg_timer tm (ctx, &ctx->stats.ms_envelope);
      QG_LAUNCH (qg_memtier_kernel, (unsigned) mem_pairs.size (), 256, 0, ctx->stream,
                 dPD.as<qg_pair_desc> (), ctx->scratch[SC_MISC0].as<uint32_t> (), ctx->scratch[SC_STORE].as<uint32_t> (),
                 ctx->scratch[SC_ROWACC].as<uint32_t> (), ctx->scratch[SC_MISC1].as<uint32_t> (), k, (int) ((unsigned) cfg->band_size / 2),
                 memory_mode ? -1 : cfg->kmer_threshold, (unsigned long long) cell_size, (unsigned long long) cfg->max_size,
                 dPR.as<int2> (), dPI.as<uint2> (), (uint32_t*) dFL.p);
      QG_TRY (qg_check_launch (ctx, "qg_memtier_kernel"));
    }
  }
  {
    qg_timer tm (ctx, &ctx->stats.ms_envelope);
    QG_LAUNCH (qg_envelope_finalize_kernel, (unsigned) ((n_pairs + 127) / 128), 128, 0, ctx->stream,
               dPD.as<qg_pair_desc> (), (uint32_t) n_pairs, dIR.as<int2> (), dIN.as<uint32_t> (), run_cap,
               dPR.as<int2> (), dPI.as<uint2> (), dPC.as<unsigned long long> ());
    QG_TRY (qg_check_launch (ctx, "qg_envelope_finalize_kernel"));
  }
  // small read-back: flags, per-pair run counts, runs
  uint64_t flags[2] = {0, 0};
  std::vector<uint2> info (n_pairs);
  std::vector<unsigned long long> cu (n_pairs);
  std::vector<int2> pr (run_total + 1);
  {
    qg_timer tm (ctx, &ctx->stats.ms_d2h);
    QG_TRY (qg_fetch (ctx, flags, dFL.p, 16));
    QG_TRY (qg_fetch (ctx, info.data (), dPI.p, sizeof (uint2) * n_pairs));
    QG_TRY (qg_fetch (ctx, cu.data (), dPC.p, sizeof (unsigned long long) * n_pairs));
    QG_TRY (qg_fetch (ctx, pr.data (), dPR.p, sizeof (int2) * run_total));
    QG_TRY (qg_fetch_wait (ctx));
  }
  if ((uint32_t) flags[0] == 2) QG_FAIL (ctx, QG_ERR_CUDA, "internal: memory-guided envelope produced more runs than the bound");
  if ((uint32_t) flags[0]) { *overflow = true; return QG_OK; }
  ctx->stats.kmer_hits += flags[1];
  out.run_begin.resize (n_pairs + 1);
  out.runs.clear ();
  out.cu.assign (cu.begin (), cu.end ());
  out.ndiag.resize (n_pairs);
  for (size_t p = 0; p < n_pairs; ++p) {
    out.run_begin[p] = (uint32_t) out.runs.size ();
    for (uint32_t r = 0; r < info[p].x; ++r) out.runs.push_back (pr[pd[p].run_out + r]);
    out.ndiag[p] = info[p].y;
  }
  out.run_begin[n_pairs] = (uint32_t) out.runs.size ();
  ctx->stats.n_pairs += n_pairs;
  return QG_OK;
}

// run buffers are sized for the common case (a handful of runs per 192k-diagonal chunk) and grown on demand up
// to the most runs a chunk can hold, chunk / (band + 2) + 2
static int qg_envelope_stage_retry (qg_ctx* ctx, const qg_dpconfig* cfg, uint64_t cell_size, int x_set,
                                    size_t n_pairs, const uint32_t* xi, const uint32_t* yi, qg_env_result& out) {
  const uint32_t half = (uint32_t) cfg->band_size / 2;
  const uint32_t cap_max = (8 * QG_SEED_CHUNK + 65536) / (2 * half + 2) + 3;      // the tile kernel's first item also owns the read's negative diagonals
  uint32_t cap = (uint32_t) qg_env_size ("QG_RUN_CAP", 16);
  while (true) {
    bool overflow = false;
    QG_TRY (qg_envelope_stage_cap (ctx, cfg, cell_size, x_set, n_pairs, xi, yi, out, cap, &overflow));
    if (!overflow) return QG_OK;
    if (cap >= cap_max) QG_FAIL (ctx, QG_ERR_CUDA, "internal: run buffer overflow at the theoretical maximum of %u runs per chunk", cap);
    cap = std::min<uint64_t> ((uint64_t) cap * 16, cap_max);
  }
}

// The memory-guided mode and the general seeding path keep one 32-bit counter per diagonal per pair in HBM: the pair
// list is cut into sub-batches whose counters fit the budget.
// The E-step runs Forward over every pair, then Forward + Backward over the gated ones (qmodel.cpp:2247-2262 builds one
// envelope per pair and uses it for both matrices): the first pass leaves its envelopes here, the second picks its pairs
// out of them instead of seeding again.  Per thread = per calling context (one host thread drives a context at a time).
struct qg_env_reuse { qg_env_result* keep = nullptr; const qg_env_result* src = nullptr; const std::vector<size_t>* pick = nullptr; };
static thread_local qg_env_reuse g_env_reuse;

static int qg_envelope_stage (qg_ctx* ctx, const qg_dpconfig* cfg, uint64_t cell_size, int x_set,
                              size_t n_pairs, const uint32_t* xi, const uint32_t* yi, qg_env_result& out) {
  if (g_env_reuse.src && g_env_reuse.pick && g_env_reuse.pick->size () == n_pairs) {
    const qg_env_result& S = *g_env_reuse.src;
    out = qg_env_result ();
    out.run_begin.push_back (0);
    for (size_t p : *g_env_reuse.pick) {
      out.runs.insert (out.runs.end (), S.runs.begin () + S.run_begin[p], S.runs.begin () + S.run_begin[p + 1]);
      out.run_begin.push_back ((uint32_t) out.runs.size ());
      out.cu.push_back (S.cu[p]); out.ndiag.push_back (S.ndiag[p]);
    }
    return QG_OK;
  }
  const bool counters = cfg->sparse && (cfg->kmer_threshold < 0 || qg_seed_is_general (ctx, cfg, x_set, n_pairs, xi, yi));
  if (!counters) return qg_envelope_stage_retry (ctx, cfg, cell_size, x_set, n_pairs, xi, yi, out);
  const qg_seqset& X = ctx->seqs[x_set];
  const qg_seqset& Y = ctx->seqs[QG_READS];
  size_t freeb = 0, totb = 0;
  QG_CUDA (ctx, cudaMemGetInfo (&freeb, &totb));
  const uint64_t budget = (uint64_t) qg_env_size ("QG_COUNT_BUDGET_MB", std::min<size_t> (freeb / 3, (size_t) 16 << 30) >> 20) << 20;
  out = qg_env_result ();
  out.run_begin.push_back (0);
  size_t p0 = 0;
  while (p0 < n_pairs) {
    size_t p1 = p0; uint64_t bytes = 0, kmers = 0;
    while (p1 < n_pairs) {
      if (xi[p1] >= X.n || yi[p1] >= Y.n) QG_FAIL (ctx, QG_ERR_INVALID, "pair %zu: sequence index out of range", p1);
      const uint64_t need = ((uint64_t) X.len (xi[p1]) + Y.len (yi[p1])) * 5 + 64;     // counters + three bitmaps + count histogram
      if (p1 > p0 && (bytes + need > budget || kmers + Y.len (yi[p1]) > 0x40000000ull)) break;
      bytes += need; kmers += Y.len (yi[p1]); ++p1;
    }
    qg_env_result part;
    QG_TRY (qg_envelope_stage_retry (ctx, cfg, cell_size, x_set, p1 - p0, xi + p0, yi + p0, part));
    for (size_t p = 0; p < p1 - p0; ++p) {
      for (uint32_t r = part.run_begin[p]; r < part.run_begin[p + 1]; ++r) out.runs.push_back (part.runs[r]);
      out.run_begin.push_back ((uint32_t) out.runs.size ());
      out.cu.push_back (part.cu[p]); out.ndiag.push_back (part.ndiag[p]);
    }
    p0 = p1;
  }
  return QG_OK;
}

extern "C" int qg_envelopes (qg_ctx* ctx, const qg_dpconfig* cfg, uint64_t cell_size, int x_set,
                             size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                             int32_t** diags_out, uint64_t* diag_offsets, uint64_t* cell_updates) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !xi || !yi || !diags_out || !diag_offsets) return QG_ERR_INVALID;
  if (x_set != QG_REFS && x_set != QG_READS) QG_FAIL (ctx, QG_ERR_INVALID, "x_set must be QG_REFS or QG_READS");
  qg_env_result er;
  QG_TRY (qg_envelope_stage (ctx, cfg, cell_size, x_set, n_pairs, xi, yi, er));
  uint64_t total = 0;
  for (size_t p = 0; p < n_pairs; ++p) { diag_offsets[p] = total; total += er.ndiag[p]; }
  diag_offsets[n_pairs] = total;
  int32_t* d = (int32_t*) malloc (sizeof (int32_t) * (total + 1));
  if (!d) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
  uint64_t n = 0;
  for (size_t p = 0; p < n_pairs; ++p)
    for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r)
      for (int v = er.runs[r].x; v <= er.runs[r].y; ++v) d[n++] = v;
  *diags_out = d;
  if (cell_updates) for (size_t p = 0; p < n_pairs; ++p) cell_updates[p] = er.cu[p];
  return QG_OK;
}

// ---- DP staging shared by Viterbi / Forward / Backward ----------------------------------------------------------
struct qg_dp_plan {
  std::vector<qg_segment> segs;      // grouped by launch class, pairs ascending inside a class? no: see seg_of_pair
  std::vector<qg_pair_dp> pairs;     // per pair: its segments are contiguous in `segs`
  struct launch { int R; int nw; uint32_t begin, count; };
  std::vector<launch> launches;      // each covers segs_sorted[begin, begin+count)
  std::vector<uint32_t> order;       // launch order -> index into segs
  std::vector<qg_segment> segs_sorted;
  uint64_t trace_words = 0, store_doubles = 0, aux_slots = 0, rp_rows = 0, acc_rows = 0;
  std::vector<qg_rp_job> rp_jobs;
};

static int qg_pick_R (uint32_t width, int* R, int* nw) {
  if (width <= 256) { int r = (int) ((width + 31) / 32); if (r < 2) r = 2; *R = r; *nw = 1; return QG_OK; }
  *R = 8; *nw = (int) ((width + 255) / 256);
  return (*nw <= QG_MAX_NW) ? QG_OK : QG_ERR_UNSUPPORTED;
}

// pairs [p0, p1) of the call; trace / store sizing according to `mode` (0 Viterbi, 1 Forward, 2 Forward+store)
static int qg_build_plan (qg_ctx* ctx, const qg_env_result& er, size_t p0, size_t p1, const uint32_t* xi, const uint32_t* yi,
                          int x_set, int mode, qg_dp_plan& plan, bool diag_forward = false) {
  const qg_seqset& X = ctx->seqs[x_set];
  const qg_seqset& Y = ctx->seqs[QG_READS];
  plan = qg_dp_plan ();
  std::map<uint32_t, uint64_t> rp_of_read;
  for (size_t p = p0; p < p1; ++p) {
    qg_pair_dp pp;
    pp.seg_begin = (uint32_t) plan.segs.size ();
    pp.xlen = X.len (xi[p]); pp.ylen = Y.len (yi[p]);
    auto it = rp_of_read.find (yi[p]);
    if (it == rp_of_read.end ()) {
      qg_rp_job jb; jb.yseq = yi[p]; jb.ylen = pp.ylen; jb.yoff = Y.off[yi[p]]; jb.rp_off = plan.rp_rows;
      plan.rp_jobs.push_back (jb);
      it = rp_of_read.insert (std::make_pair (yi[p], plan.rp_rows)).first;
      plan.rp_rows += (uint64_t) pp.ylen + 2;
    }
    int dmin = 0, dmax = 0; bool first = true;
    for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r) {
      qg_segment sg;
      memset (&sg, 0, sizeof (sg));
      sg.pair = (uint32_t) (p - p0);
      sg.dlo = er.runs[r].x; sg.width = (uint32_t) (er.runs[r].y - er.runs[r].x + 1);
      sg.xlen = pp.xlen; sg.ylen = pp.ylen; sg.xseq = xi[p]; sg.yseq = yi[p];
      int R, nw;
      if (qg_pick_R (sg.width, &R, &nw) != QG_OK)
        QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "pair %zu: a run of %u consecutive diagonals exceeds the %d this build fills with one CTA", p, sg.width, 256 * QG_MAX_NW);
      sg.R = (uint32_t) R; sg.nwarps = (uint32_t) nw;
      sg.rp_off = it->second;
      const uint64_t lanes = 32ull * nw;
      const bool narrow = mode == 0 && sg.width <= 4 && !getenv ("QG_VIT_GENERIC");    // one thread per run (qg_vit_narrow_kernel)
      if (narrow) { sg.R = sg.width; sg.nwarps = 0; sg.trace_off = plan.trace_words; plan.trace_words += (((uint64_t) pp.ylen + 4) / 4) * 4; }
      else if (diag_forward && (mode == 1 || mode == 2) && sg.width == 1) { sg.R = 1; sg.nwarps = 0; }   // isolated diagonal: one thread (qg_forward_prob_diag_kernel), closed-form Backward
      else if (mode == 0 && nw == 1 && R <= 4 && !getenv ("QG_VIT_GENERIC")) {
        // qg_vit_kernel<2..4>: R nibbles per lane fit a 16-bit word -- half the pointer bytes of the band (4 bit per slot)
        sg.half = 1; sg.trace_off = plan.trace_words; plan.trace_words += (((uint64_t) pp.ylen + lanes + 1) * lanes + 1) / 2;
      }
      else if (mode == 0 || mode == 3) { sg.trace_off = plan.trace_words; plan.trace_words += ((uint64_t) pp.ylen + lanes + 1) * lanes; }
      if (mode == 3) { sg.acc_off = plan.acc_rows; plan.acc_rows += (uint64_t) pp.ylen + 2; }
      if (mode == 2) { if (sg.nwarps != 0) { sg.store_off = plan.store_doubles; plan.store_doubles += ((uint64_t) pp.ylen + lanes + 1) * 3 * lanes * R; }
                       sg.acc_off = plan.acc_rows; plan.acc_rows += (uint64_t) pp.ylen + 2; }
      sg.seg_id = plan.segs.size ();
      if (mode == 0) sg.aux_off = plan.segs.size ();           // Viterbi: index of the segment in pair order
      else { sg.aux_off = plan.aux_slots; plan.aux_slots += lanes * R; }
      plan.segs.push_back (sg);
      if (first) { dmin = er.runs[r].x; dmax = er.runs[r].y; first = false; } else { dmin = std::min (dmin, er.runs[r].x); dmax = std::max (dmax, er.runs[r].y); }
    }
    pp.seg_end = (uint32_t) plan.segs.size ();
    const uint64_t xspan = std::min<uint64_t> (pp.xlen, (uint64_t) pp.ylen + (uint64_t) (dmax - dmin) + 1);
    pp.path_cap = (uint32_t) (pp.ylen + xspan + 1);
    pp.path_off = 0; pp.want_path = 0;
    plan.pairs.push_back (pp);
  }
  // launch classes: (nw, R)
  std::map<std::pair<int,int>, std::vector<uint32_t> > cls;
  for (uint32_t s = 0; s < plan.segs.size (); ++s) cls[std::make_pair ((int) plan.segs[s].nwarps, (int) plan.segs[s].R)].push_back (s);
  for (auto& kv : cls) {
    qg_dp_plan::launch L; L.nw = kv.first.first; L.R = kv.first.second; L.begin = (uint32_t) plan.order.size (); L.count = (uint32_t) kv.second.size ();
    for (uint32_t s : kv.second) plan.order.push_back (s);
    plan.launches.push_back (L);
  }
  return QG_OK;
}

// Launch classes run concurrently on side streams, but every CTA is one long dependent chain and the biggest class alone nearly
// fills the resident-CTA slots: a small class launched after it waits for a slot and then adds a whole chain to the stage's
// time.  Small classes first: they take their few slots at once and the big class fills the rest (Viterbi fill of 1536 reads:
// 15.2 -> 14.1 ms).
template<class LaunchVec>
static std::vector<size_t> qg_launch_order (const LaunchVec& launches) {
  std::vector<size_t> order (launches.size ());
  std::iota (order.begin (), order.end (), (size_t) 0);
  if (!getenv ("QG_LAUNCH_PLAN_ORDER"))
    std::stable_sort (order.begin (), order.end (), [&] (size_t u, size_t v) { return launches[u].count < launches[v].count; });
  return order;
}

template<int MODE>
static int qg_launch_fill (qg_ctx* ctx, const qg_dp_plan& plan, qg_fill_args a, const qg_segment* d_segs_launch_order) {
  QG_TRY (qg_fork (ctx));
  int kcls = 0;
  for (size_t li : qg_launch_order (plan.launches)) {
    const auto& L = plan.launches[li];
    cudaStream_t st; QG_TRY (qg_side (ctx, kcls++, &st));
    a.segs = d_segs_launch_order + L.begin;
    if (L.nw == 0) {                                        // narrow Viterbi segments, one thread each
      qg_vit_args va; va.segs = a.segs; va.xpacked = a.xpacked; va.xpoff = a.xpoff; va.rps = ctx->scratch[SC_RPS].as<double2> ();
      va.i2i = a.i2i; va.i2m = a.i2m; va.d2d = a.d2d; va.d2m = a.d2m; va.local = a.local; va.trace = a.trace; va.endvals = a.endvals;
      const unsigned grid = (L.count + 63) / 64;
      switch (L.R) {
        case 1: QG_LAUNCH (qg_vit_narrow_kernel<1>, grid, 64, 0, st, va, L.count); break;
        case 2: QG_LAUNCH (qg_vit_narrow_kernel<2>, grid, 64, 0, st, va, L.count); break;
        case 3: QG_LAUNCH (qg_vit_narrow_kernel<3>, grid, 64, 0, st, va, L.count); break;
        case 4: QG_LAUNCH (qg_vit_narrow_kernel<4>, grid, 64, 0, st, va, L.count); break;
        default: QG_FAIL (ctx, QG_ERR_INVALID, "internal: narrow segment of width %d", L.R);
      }
      QG_TRY (qg_check_launch (ctx, "qg_vit_narrow_kernel"));
      continue;
    }
    const bool multi = L.nw > 1;
    const unsigned block = 32u * L.nw;
#define QG_CASE(RR) case RR: \
      if (multi) { auto kfn = qg_fill_kernel<8, MODE, true>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } \
      else if (MODE == 0 && !getenv ("QG_VIT_GENERIC")) { \
        qg_vit_args va; va.segs = a.segs; va.xpacked = a.xpacked; va.xpoff = a.xpoff; va.rps = ctx->scratch[SC_RPS].as<double2> (); \
        va.i2i = a.i2i; va.i2m = a.i2m; va.d2d = a.d2d; va.d2m = a.d2m; va.local = a.local; va.trace = a.trace; va.endvals = a.endvals; \
        auto kfn = qg_vit_kernel<RR>; QG_LAUNCH (kfn, L.count, block, 0, st, va); } \
      else { auto kfn = qg_fill_kernel<RR, MODE, false>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } break;
    switch (L.R) {
      QG_CASE (2) QG_CASE (3) QG_CASE (4) QG_CASE (5) QG_CASE (6) QG_CASE (7) QG_CASE (8)
      default: QG_FAIL (ctx, QG_ERR_INVALID, "internal: unsupported R=%d", L.R);
    }
#undef QG_CASE
    QG_TRY (qg_check_launch (ctx, "qg_fill_kernel"));
  }
  QG_TRY (qg_join (ctx, kcls));
  return QG_OK;
}

static int qg_stage_rowparams (qg_ctx* ctx, const qg_dp_plan& plan) {
  const qg_seqset& Y = ctx->seqs[QG_READS];
  const qg_model_dev& m = ctx->model;
  QG_TRY (qg_upload (ctx, ctx->scratch[SC_RPJOBS], plan.rp_jobs.data (), sizeof (qg_rp_job) * plan.rp_jobs.size ()));
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_RP], sizeof (qg_rowp) * (plan.rp_rows + 1)));
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_RPS], sizeof (qg_rowp) * (plan.rp_rows + 1)));
  if (!plan.rp_jobs.empty ()) {
    QG_LAUNCH (qg_rowparams_kernel, (unsigned) plan.rp_jobs.size (), 256, 0, ctx->stream,
               ctx->scratch[SC_RPJOBS].as<qg_rp_job> (), Y.d_tok.as<uint8_t> (), Y.has_qual ? Y.d_qual.as<uint8_t> () : (const uint8_t*) nullptr,
               m.d_match.as<double> (), m.d_insert.as<double> (), m.d_gap.as<double> (), m.match_k, m.gap_k, ctx->scratch[SC_RP].as<qg_rowp> (),
               ctx->scratch[SC_RPS].as<double2> ());
    QG_TRY (qg_check_launch (ctx, "qg_rowparams_kernel"));
  }
  return QG_OK;
}

static qg_fill_args qg_base_args (qg_ctx* ctx, const qg_dpconfig* cfg, int x_set) {
  qg_fill_args a;
  memset (&a, 0, sizeof (a));
  a.xpacked = ctx->seqs[x_set].d_packed.as<uint64_t> ();
  a.xpoff = ctx->seqs[x_set].d_poff.as<uint64_t> ();
  a.rp = ctx->scratch[SC_RP].as<qg_rowp> ();
  a.lse = ctx->d_lse.as<double> ();
  a.i2i = ctx->model.i2i; a.i2m = ctx->model.i2m; a.d2d = ctx->model.d2d; a.d2m = ctx->model.d2m;
  a.local = cfg->local;
  return a;
}

static int qg_check_ready (qg_ctx* ctx, const qg_dpconfig* cfg) {
  if (!ctx->model.set) QG_FAIL (ctx, QG_ERR_STATE, "no align model: call qg_set_align_model first");
  if (!ctx->seqs[QG_REFS].n || !ctx->seqs[QG_READS].n) QG_FAIL (ctx, QG_ERR_STATE, "sequence sets not uploaded: call qg_set_seqs for QG_REFS and QG_READS");
  if (cfg->band_size < 0) QG_FAIL (ctx, QG_ERR_INVALID, "negative band size");
  return QG_OK;
}

static uint64_t qg_plan_cells (const qg_env_result& er, size_t p0, size_t p1) {
  uint64_t cu = 0;
  for (size_t p = p0; p < p1; ++p) cu += er.cu[p];
  return cu;
}

// ---- wide pairs: runs of more than 256 * QG_MAX_NW diagonals (qg_tile.cuh) -----------------------------------------
static bool qg_pair_is_wide (const qg_env_result& er, size_t p) {
  const uint32_t min_wide = (uint32_t) qg_env_size ("QG_WIDE_MIN_DIAGS", 256u * QG_MAX_NW);   // lowered by the tests only
  for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r)
    if ((uint32_t) (er.runs[r].y - er.runs[r].x + 1) > min_wide) return true;
  return false;
}

// a run of more than 256 diagonals needs several warps: the probability-space Forward / Backward kernels are single-warp, such
// pairs take the log-space kernels.  The choice is a function of the pair alone (not of its batch-mates): a call that holds
// both kinds is split and each part served by its kernels.
static bool qg_pair_is_multiwarp (const qg_env_result& er, size_t p) {
  for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r)
    if ((uint32_t) (er.runs[r].y - er.runs[r].x + 1) > 256u) return true;
  return false;
}

// mode 0: Viterbi (+ traceback of the pairs flagged in want); mode 1: Forward log-likelihood.  idx: the wide pairs'
// positions in the caller's pair list; outputs are written at those positions.
// mode 2: Forward with every cell kept, Backward + counts (qg_tile_backward_kernel); fb carries its inputs and outputs.
struct qg_wide_fb { const double* weights = nullptr; double* back = nullptr; double* counts_sum = nullptr; double* counts_per_pair = nullptr; uint64_t nC = 0; };

static int qg_wide_run (qg_ctx* ctx, const qg_dpconfig* cfg, const qg_env_result& er, const std::vector<size_t>& idx,
                        const uint32_t* xi, const uint32_t* yi, int mode, const uint8_t* want,
                        double* score, uint32_t* x_start, uint32_t* x_end, std::vector<std::vector<uint8_t> >* paths,
                        const qg_wide_fb* fb = nullptr) {
  const qg_seqset& X = ctx->seqs[QG_REFS];
  const qg_seqset& Y = ctx->seqs[QG_READS];
  size_t freeb = 0, totb = 0;
  QG_CUDA (ctx, cudaMemGetInfo (&freeb, &totb));
  const uint64_t budget = (uint64_t) qg_env_size ("QG_TRACE_BUDGET_MB", std::min<size_t> (freeb / 2, (size_t) 48 << 30) >> 20) << 20;
  size_t q0 = 0;
  while (q0 < idx.size ()) {
    // ---- plan a batch of pairs under the memory budget
    std::vector<qg_wseg> segs; std::vector<qg_tile> tiles; std::vector<qg_wpair> wp; std::vector<long long> tables;
    std::vector<qg_tile> btiles;                             // mode 2: tiles of the reversed matrices
    qg_dp_plan rpplan; std::map<uint32_t, uint64_t> rp_of_read;
    uint64_t end_d = 0, bytes = 0, scratch_bytes = 0, acc_rows = 0;
    const uint64_t tile_bytes = (mode == 0 ? QG_TILE_WORDS * 4ull : 0) + (QG_TRH + QG_TCW) * 24ull + 2 * 16 + sizeof (qg_tile)
                              + (mode == 2 ? QG_TILE_CELLS3 * 8 + (QG_TRH + QG_TCW) * 24ull + sizeof (qg_tile) : 0);
    size_t q1 = q0;
    while (q1 < idx.size ()) {
      const size_t p = idx[q1];
      const uint32_t xlen = X.len (xi[p]), ylen = Y.len (yi[p]);
      const uint32_t nCB = (xlen + QG_TCW - 1) / QG_TCW, nRB = (ylen + QG_TRH - 1) / QG_TRH;
      // tiles a run touches in row block b: columns [jLo + dlo, jHi + dhi] clipped to [1, xLen]
      auto col_blocks = [&] (int dlo, int dhi, uint32_t b, int64_t* aLo, int64_t* aHi) {
        const int64_t jLo = (int64_t) b * QG_TRH + 1, jHi = std::min<int64_t> ((int64_t) (b + 1) * QG_TRH, ylen);
        const int64_t iLo = std::max<int64_t> (1, jLo + dlo), iHi = std::min<int64_t> (xlen, jHi + dhi);
        if (iLo > iHi) { *aLo = 0; *aHi = -1; return; }
        *aLo = (iLo - 1) / QG_TCW; *aHi = (iHi - 1) / QG_TCW;
      };
      uint64_t ntile = 0;
      for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r)
        for (uint32_t b = 0; b < nRB; ++b) { int64_t aLo, aHi; col_blocks (er.runs[r].x, er.runs[r].y, b, &aLo, &aHi); ntile += (uint64_t) (aHi - aLo + 1); }
      const uint64_t nruns = er.run_begin[p + 1] - er.run_begin[p];
      const uint64_t need = ntile * tile_bytes + nruns * ((uint64_t) nCB * nRB * 8 + (uint64_t) (xlen + 1) * 8 + (mode == 2 ? ((uint64_t) ylen + 2) * 64 : 0)) + 2ull * (xlen + ylen)
                            + (mode == 2 && fb ? fb->nC * 8 : 0);
      if (q1 > q0 && (bytes + need > budget || tiles.size () + ntile > 0x7FFFFFF0ull)) break;
      if (need > budget || ntile > 0x7FFFFFF0ull) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "pair %zu needs %llu MB of device memory for its %llu tiles (budget %llu MB)", p,
                                  (unsigned long long) (need >> 20), (unsigned long long) ntile, (unsigned long long) (budget >> 20));
      bytes += need;
      qg_wpair w; memset (&w, 0, sizeof (w));
      w.seg_begin = (uint32_t) segs.size (); w.tile_begin = (uint32_t) tiles.size (); w.xlen = xlen; w.ylen = ylen;
      auto it = rp_of_read.find (yi[p]);
      if (it == rp_of_read.end ()) {
        qg_rp_job jb; jb.yseq = yi[p]; jb.ylen = ylen; jb.yoff = Y.off[yi[p]]; jb.rp_off = rpplan.rp_rows;
        rpplan.rp_jobs.push_back (jb);
        it = rp_of_read.insert (std::make_pair (yi[p], rpplan.rp_rows)).first;
        rpplan.rp_rows += (uint64_t) ylen + 2;
      }
      int dmin = 0, dmax = 0; bool first = true;
      for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r) {
        qg_wseg ws; memset (&ws, 0, sizeof (ws));
        ws.pair = (uint32_t) wp.size (); ws.xseq = xi[p]; ws.xlen = xlen; ws.ylen = ylen; ws.dlo = er.runs[r].x; ws.dhi = er.runs[r].y;
        ws.nCB = nCB; ws.nRB = nRB; ws.rp_off = it->second;
        ws.table_off = tables.size (); tables.resize (tables.size () + (size_t) nCB * nRB, -1);
        ws.end_off = end_d; end_d += (uint64_t) xlen + 1;
        ws.acc_off = acc_rows; if (mode == 2) acc_rows += (uint64_t) ylen + 2;
        if (mode == 2) {
          // the reversed matrix of this run: band [xLen - yLen - dhi, xLen - yLen - dlo], its own block table (host only)
          const int rlo = ((int) xlen - (int) ylen) - ws.dhi, rhi = ((int) xlen - (int) ylen) - ws.dlo;
          std::vector<long long> rtab ((size_t) nCB * nRB, -1);
          for (uint32_t b = 0; b < nRB; ++b) {
            int64_t aLo, aHi; col_blocks (rlo, rhi, b, &aLo, &aHi);
            for (int64_t a = aLo; a <= aHi; ++a) {
              qg_tile t; t.seg = (uint32_t) segs.size (); t.a = (uint32_t) a; t.b = b; t.id = (uint32_t) btiles.size (); t.pad_ = 0;
              t.left = a > 0 ? (int32_t) rtab[(size_t) (a - 1) * nRB + b] : -1;
              t.up = b > 0 ? (int32_t) rtab[(size_t) a * nRB + (b - 1)] : -1;
              t.diag = (a > 0 && b > 0) ? (int32_t) rtab[(size_t) (a - 1) * nRB + (b - 1)] : -1;
              rtab[(size_t) a * nRB + b] = (long long) t.id;
              btiles.push_back (t);
            }
          }
        }
        long long* tab = tables.data () + ws.table_off;
        for (uint32_t b = 0; b < nRB; ++b) {
          int64_t aLo, aHi; col_blocks (ws.dlo, ws.dhi, b, &aLo, &aHi);
          for (int64_t a = aLo; a <= aHi; ++a) {
            qg_tile t; t.seg = (uint32_t) segs.size (); t.a = (uint32_t) a; t.b = b; t.id = (uint32_t) tiles.size (); t.pad_ = 0;
            t.left = a > 0 ? (int32_t) tab[(size_t) (a - 1) * nRB + b] : -1;
            t.up = b > 0 ? (int32_t) tab[(size_t) a * nRB + (b - 1)] : -1;
            t.diag = (a > 0 && b > 0) ? (int32_t) tab[(size_t) (a - 1) * nRB + (b - 1)] : -1;
            tab[(size_t) a * nRB + b] = (long long) t.id;
            tiles.push_back (t);
          }
        }
        segs.push_back (ws);
        if (first) { dmin = ws.dlo; dmax = ws.dhi; first = false; } else { dmin = std::min (dmin, ws.dlo); dmax = std::max (dmax, ws.dhi); }
      }
      w.seg_end = (uint32_t) segs.size (); w.tile_end = (uint32_t) tiles.size ();
      const uint64_t xspan = std::min<uint64_t> (xlen, (uint64_t) ylen + (uint64_t) (dmax - dmin) + 1);
      w.path_cap = (uint32_t) (ylen + xspan + 1);
      w.want_path = (mode == 0 && want && want[p]) ? 1 : 0;
      w.path_off = scratch_bytes;
      if (w.want_path) scratch_bytes += w.path_cap;
      wp.push_back (w);
      ++q1;
    }
    const uint64_t trace_words = mode == 0 ? (uint64_t) tiles.size () * QG_TILE_WORDS : 0;
    const uint64_t ntl = std::max (tiles.size (), btiles.size ());
    const uint64_t col_d = ntl * QG_TRH * 3, row_d = ntl * QG_TCW * 3;
    const size_t np = wp.size ();
    std::vector<qg_tile> bsorted (btiles);
    std::stable_sort (bsorted.begin (), bsorted.end (), [] (const qg_tile& u, const qg_tile& v) { return u.a + u.b < v.a + v.b; });
    // tiles in wavefront order
    std::vector<qg_tile> sorted (tiles);
    std::stable_sort (sorted.begin (), sorted.end (), [] (const qg_tile& u, const qg_tile& v) { return u.a + u.b < v.a + v.b; });
    {
      qg_timer tm (ctx, &ctx->stats.ms_prep);
      QG_TRY (qg_stage_rowparams (ctx, rpplan));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGS], segs.data (), sizeof (qg_wseg) * segs.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_ITEMS], sorted.data (), sizeof (qg_tile) * sorted.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], wp.data (), sizeof (qg_wpair) * np));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], tables.data (), sizeof (long long) * tables.size ()));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_TRACE], sizeof (uint32_t) * (trace_words + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_STORE], sizeof (double) * (col_d + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ROWACC], sizeof (double) * (row_d + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDVALS], sizeof (double) * (std::max<uint64_t> (2 * tiles.size (), end_d) + 2)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT0], sizeof (double) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT1], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT2], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT3], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHSCR], scratch_bytes + 16));
      if (mode == 2) {
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_TRACE], sizeof (double) * ((uint64_t) tiles.size () * QG_TILE_CELLS3 + 1)));      // the Forward cells
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_ZE], bsorted.data (), sizeof (qg_tile) * bsorted.size ()));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ZM], sizeof (qg_rowrec) * (acc_rows + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_MISC1], sizeof (double) * 12 * (segs.size () + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHSCR], sizeof (double) * fb->nC * np + 16));
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_ZM].p, 0, sizeof (qg_rowrec) * (acc_rows + 1), ctx->stream));
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_MISC1].p, 0, sizeof (double) * 12 * (segs.size () + 1), ctx->stream));
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_PATHSCR].p, 0, sizeof (double) * fb->nC * np, ctx->stream));
        ctx->stats.fwd_store_bytes += (uint64_t) tiles.size () * QG_TILE_CELLS3 * 8;
      }
    }
    uint64_t cu = 0;
    for (size_t q = q0; q < q1; ++q) cu += er.cu[idx[q]];
    if (mode == 2) cu *= 2;
    ctx->stats.cell_updates += cu;
    ctx->stats.n_segments += segs.size ();
    ctx->stats.trace_bytes += trace_words * 4;
    qg_tile_args a;
    memset (&a, 0, sizeof (a));
    a.segs = ctx->scratch[SC_SEGS].as<qg_wseg> (); a.tiles = ctx->scratch[SC_ITEMS].as<qg_tile> ();
    a.xpacked = X.d_packed.as<uint64_t> (); a.xpoff = X.d_poff.as<uint64_t> ();
    a.rp = ctx->scratch[SC_RP].as<qg_rowp> (); a.lse = ctx->d_lse.as<double> ();
    a.i2i = ctx->model.i2i; a.i2m = ctx->model.i2m; a.d2d = ctx->model.d2d; a.d2m = ctx->model.d2m; a.local = cfg->local;
    a.trace = ctx->scratch[SC_TRACE].as<uint32_t> (); a.colbuf = ctx->scratch[SC_STORE].as<double> (); a.rowbuf = ctx->scratch[SC_ROWACC].as<double> ();
    a.tile_best = ctx->scratch[SC_ENDVALS].as<double> (); a.endvals = ctx->scratch[SC_ENDVALS].as<double> ();
    if (mode == 2) {
      a.fstore = ctx->scratch[SC_TRACE].as<double> (); a.btiles = ctx->scratch[SC_ZE].as<qg_tile> ();
      a.tables = ctx->scratch[SC_MISC0].as<long long> (); a.pair_z = ctx->scratch[SC_OUT0].as<double> ();
      a.rowacc = ctx->scratch[SC_ZM].as<double> (); a.seg_scal = ctx->scratch[SC_MISC1].as<double> ();
    }
    {
      qg_timer tm (ctx, mode == 0 ? &ctx->stats.ms_viterbi : &ctx->stats.ms_forward);
      size_t t0 = 0;
      while (t0 < sorted.size ()) {                          // one launch per tile anti-diagonal
        size_t t1 = t0; const uint32_t w = sorted[t0].a + sorted[t0].b;
        while (t1 < sorted.size () && sorted[t1].a + sorted[t1].b == w) ++t1;
        if (mode == 0) { QG_LAUNCH (qg_tile_kernel<0>, (unsigned) (t1 - t0), 32, 0, ctx->stream, a, (uint32_t) t0); }
        else { QG_LAUNCH (qg_tile_kernel<1>, (unsigned) (t1 - t0), 32, 0, ctx->stream, a, (uint32_t) t0); }
        QG_TRY (qg_check_launch (ctx, "qg_tile_kernel"));
        t0 = t1;
      }
      if (mode == 0) {
        QG_LAUNCH (qg_wide_score_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_wpair> (), (uint32_t) np, ctx->scratch[SC_ENDVALS].as<double> (),
                   ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_OUT2].as<uint32_t> ());
        QG_TRY (qg_check_launch (ctx, "qg_wide_score_kernel"));
      } else {
        QG_LAUNCH (qg_wide_forward_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_wpair> (), (uint32_t) np, ctx->scratch[SC_SEGS].as<qg_wseg> (),
                   ctx->scratch[SC_ENDVALS].as<double> (), ctx->d_lse.as<double> (), ctx->scratch[SC_OUT0].as<double> ());
        QG_TRY (qg_check_launch (ctx, "qg_wide_forward_finalize_kernel"));
      }
    }
    if (mode == 2) {
      // Backward over the reversed matrices, one launch per tile anti-diagonal; then the counts as for banded pairs
      const qg_seqset& Yr = ctx->seqs[QG_READS];
      const qg_model_dev& m = ctx->model;
      {
        qg_timer tm (ctx, &ctx->stats.ms_backward);
        size_t t0 = 0;
        while (t0 < bsorted.size ()) {
          size_t t1 = t0; const uint32_t w = bsorted[t0].a + bsorted[t0].b;
          while (t1 < bsorted.size () && bsorted[t1].a + bsorted[t1].b == w) ++t1;
          QG_LAUNCH (qg_tile_backward_kernel, (unsigned) (t1 - t0), 32, 0, ctx->stream, a, (uint32_t) t0);
          QG_TRY (qg_check_launch (ctx, "qg_tile_backward_kernel"));
          t0 = t1;
        }
        QG_LAUNCH (qg_wide_backward_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_wpair> (), (uint32_t) np, ctx->scratch[SC_SEGS].as<qg_wseg> (),
                   ctx->scratch[SC_ENDVALS].as<double> (), ctx->d_lse.as<double> (), ctx->scratch[SC_OUT1].as<double> ());
        QG_TRY (qg_check_launch (ctx, "qg_wide_backward_finalize_kernel"));
        // the scatter / reduce kernels of the banded path, over run and pair records in their layout
        std::vector<qg_segment> fs (segs.size ()); std::vector<qg_pair_dp> fp (np);
        for (size_t t = 0; t < segs.size (); ++t) {
          qg_segment g; memset (&g, 0, sizeof (g));
          g.pair = segs[t].pair; g.dlo = segs[t].dlo; g.width = (uint32_t) (segs[t].dhi - segs[t].dlo + 1); g.xlen = segs[t].xlen; g.ylen = segs[t].ylen;
          g.xseq = segs[t].xseq; g.yseq = yi[idx[q0 + segs[t].pair]]; g.acc_off = segs[t].acc_off; g.seg_id = t;
          fs[t] = g;
        }
        for (size_t q = 0; q < np; ++q) { qg_pair_dp d; memset (&d, 0, sizeof (d)); d.seg_begin = wp[q].seg_begin; d.seg_end = wp[q].seg_end; d.xlen = wp[q].xlen; d.ylen = wp[q].ylen; fp[q] = d; }
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_KEYS0], fs.data (), sizeof (qg_segment) * fs.size ()));
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_KEYS1], fp.data (), sizeof (qg_pair_dp) * np));
        QG_LAUNCH (qg_counts_scatter_kernel, (unsigned) np, 256, 0, ctx->stream,
                   ctx->scratch[SC_KEYS1].as<qg_pair_dp> (), ctx->scratch[SC_KEYS0].as<qg_segment> (),
                   Yr.d_tok.as<uint8_t> (), Yr.d_qual.as<uint8_t> (), Yr.d_off.as<uint64_t> (),
                   (const qg_rowrec*) ctx->scratch[SC_ZM].p, ctx->scratch[SC_MISC1].as<double> (),
                   m.match_k, m.gap_k, fb->nC, ctx->scratch[SC_PATHSCR].as<double> ());
        QG_TRY (qg_check_launch (ctx, "qg_counts_scatter_kernel"));
        std::vector<double> wq (np, 1.0);
        if (fb->weights) for (size_t q = 0; q < np; ++q) wq[q] = fb->weights[idx[q0 + q]];
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_VALS0], wq.data (), sizeof (double) * np));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_VALS1], sizeof (double) * (fb->nC + 1)));
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_VALS1].p, 0, sizeof (double) * fb->nC, ctx->stream));
        QG_LAUNCH (qg_counts_reduce_kernel, (unsigned) ((fb->nC + 127) / 128), 128, 0, ctx->stream,
                   ctx->scratch[SC_PATHSCR].as<double> (), ctx->scratch[SC_VALS0].as<double> (), (uint32_t) np, fb->nC, ctx->scratch[SC_VALS1].as<double> ());
        QG_TRY (qg_check_launch (ctx, "qg_counts_reduce_kernel"));
      }
      std::vector<double> part (fb->nC), bk (np);
      QG_TRY (qg_download (ctx, part.data (), ctx->scratch[SC_VALS1].p, sizeof (double) * fb->nC));
      if (fb->counts_sum) for (uint64_t k = 0; k < fb->nC; ++k) fb->counts_sum[k] += part[k];
      QG_TRY (qg_download (ctx, bk.data (), ctx->scratch[SC_OUT1].p, sizeof (double) * np));
      if (fb->back) for (size_t q = 0; q < np; ++q) fb->back[idx[q0 + q]] = bk[q];
      if (fb->counts_per_pair) {
        std::vector<double> pp (fb->nC * np);
        QG_TRY (qg_download (ctx, pp.data (), ctx->scratch[SC_PATHSCR].p, sizeof (double) * fb->nC * np));
        for (size_t q = 0; q < np; ++q) memcpy (fb->counts_per_pair + (uint64_t) idx[q0 + q] * fb->nC, pp.data () + q * fb->nC, sizeof (double) * fb->nC);
      }
    }
    std::vector<double> sc (np);
    QG_TRY (qg_download (ctx, sc.data (), ctx->scratch[SC_OUT0].p, sizeof (double) * np));
    if (score) for (size_t q = 0; q < np; ++q) score[idx[q0 + q]] = sc[q];
    if (mode == 0 && x_start) {
      std::vector<uint32_t> plen (np), xs (np), xe (np);
      uint32_t flag = 0;
      {
        qg_timer tm (ctx, &ctx->stats.ms_traceback);
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_FLAGS].p, 0, 64, ctx->stream));
        QG_LAUNCH (qg_wide_traceback_kernel, (unsigned) ((np + 31) / 32), 32, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_wpair> (), (uint32_t) np, ctx->scratch[SC_SEGS].as<qg_wseg> (),
                   ctx->scratch[SC_MISC0].as<long long> (), ctx->scratch[SC_TRACE].as<uint32_t> (), ctx->scratch[SC_OUT0].as<double> (),
                   ctx->scratch[SC_OUT2].as<uint32_t> (), ctx->scratch[SC_OUT1].as<uint32_t> (), ctx->scratch[SC_PATHSCR].as<uint8_t> (),
                   ctx->scratch[SC_OUT3].as<uint32_t> (), (uint32_t*) ctx->scratch[SC_FLAGS].p);
        QG_TRY (qg_check_launch (ctx, "qg_wide_traceback_kernel"));
      }
      QG_TRY (qg_fetch (ctx, xs.data (), ctx->scratch[SC_OUT1].p, sizeof (uint32_t) * np));
      QG_TRY (qg_fetch (ctx, xe.data (), ctx->scratch[SC_OUT2].p, sizeof (uint32_t) * np));
      QG_TRY (qg_fetch (ctx, plen.data (), ctx->scratch[SC_OUT3].p, sizeof (uint32_t) * np));
      QG_TRY (qg_fetch (ctx, &flag, ctx->scratch[SC_FLAGS].p, sizeof (uint32_t)));
      QG_TRY (qg_fetch_wait (ctx));
      if (flag) QG_FAIL (ctx, QG_ERR_CUDA, "internal: tiled traceback left the envelope (code %u)", flag);
      std::vector<uint64_t> goff (np + 1, 0);
      for (size_t q = 0; q < np; ++q) goff[q + 1] = goff[q] + plen[q];
      std::vector<uint8_t> flat (goff[np] + 1);
      if (goff[np]) {
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC1], goff.data (), sizeof (uint64_t) * (np + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHOUT], goff[np] + 16));
        QG_LAUNCH (qg_wide_path_gather_kernel, (unsigned) np, 128, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_wpair> (), (uint32_t) np, ctx->scratch[SC_PATHSCR].as<uint8_t> (),
                   ctx->scratch[SC_OUT3].as<uint32_t> (), ctx->scratch[SC_MISC1].as<uint64_t> (), ctx->scratch[SC_PATHOUT].as<uint8_t> ());
        QG_TRY (qg_check_launch (ctx, "qg_wide_path_gather_kernel"));
        QG_TRY (qg_download (ctx, flat.data (), ctx->scratch[SC_PATHOUT].p, goff[np]));
      }
      for (size_t q = 0; q < np; ++q) {
        const size_t p = idx[q0 + q];
        const bool w = wp[q].want_path != 0;
        x_start[p] = w ? xs[q] : 0; x_end[p] = w ? xe[q] : 0;
        if (paths) (*paths)[p].assign (flat.begin () + goff[q], flat.begin () + goff[q + 1]);
      }
    }
    q0 = q1;
  }
  return QG_OK;
}

// ---- Viterbi -------------------------------------------------------------------------------------------------------
// group > 0: pairs come in consecutive groups of `group` (one read against every reference); the traceback then
// runs only for the best-scoring pair of each group, earliest index on ties (qmodel.cpp:2773-2775).
static int qg_viterbi_impl (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                            const uint8_t* want_path, size_t group, double* score, uint32_t* x_start, uint32_t* x_end,
                            uint8_t** path_out, uint64_t* path_offsets, const qg_env_result* er_in = nullptr) {
  QG_TRY (qg_check_ready (ctx, cfg));
  const bool paths = path_out && path_offsets && x_start && x_end;
  if (path_out) *path_out = nullptr;
  qg_env_result er_local;
  if (!er_in) QG_TRY (qg_envelope_stage (ctx, cfg, 24, QG_REFS, n_pairs, xi, yi, er_local));
  const qg_env_result& er = er_in ? *er_in : er_local;
  if (!er_in) {
    // pairs with runs too wide for one CTA take the tiled path (qg_tile.cuh); the rest continue below
    std::vector<size_t> wide, narrow;
    for (size_t p = 0; p < n_pairs; ++p) (qg_pair_is_wide (er, p) ? wide : narrow).push_back (p);
    if (!wide.empty ()) {
      // scores of every pair first, then the traceback selection (per group or per want_path), then both paths
      std::vector<uint8_t> want (n_pairs, paths ? 1 : 0);
      if (paths && !group && want_path) for (size_t p = 0; p < n_pairs; ++p) want[p] = want_path[p] ? 1 : 0;
      std::vector<std::vector<uint8_t> > per_pair (n_pairs);
      std::vector<uint32_t> xs (n_pairs, 0), xe (n_pairs, 0);
      QG_TRY (qg_wide_run (ctx, cfg, er, wide, xi, yi, 0, want.data (), score, paths ? xs.data () : nullptr, paths ? xe.data () : nullptr, paths ? &per_pair : nullptr));
      if (!narrow.empty ()) {
        qg_env_result sub; std::vector<uint32_t> sxi, syi; std::vector<uint8_t> swant;
        sub.run_begin.push_back (0);
        for (size_t p : narrow) {
          for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r) sub.runs.push_back (er.runs[r]);
          sub.run_begin.push_back ((uint32_t) sub.runs.size ()); sub.cu.push_back (er.cu[p]); sub.ndiag.push_back (er.ndiag[p]);
          sxi.push_back (xi[p]); syi.push_back (yi[p]); swant.push_back (want[p]);
        }
        std::vector<double> ssc (narrow.size ()); std::vector<uint32_t> sxs (narrow.size ()), sxe (narrow.size ()); std::vector<uint64_t> soff (narrow.size () + 1);
        qg_hostbuf sp_;
        uint8_t*& sp = sp_.p;
        QG_TRY (qg_viterbi_impl (ctx, cfg, narrow.size (), sxi.data (), syi.data (), swant.data (), 0, ssc.data (), paths ? sxs.data () : nullptr,
                                 paths ? sxe.data () : nullptr, paths ? &sp : nullptr, paths ? soff.data () : nullptr, &sub));
        for (size_t q = 0; q < narrow.size (); ++q) {
          const size_t p = narrow[q];
          score[p] = ssc[q];
          if (paths) { xs[p] = sxs[q]; xe[p] = sxe[q]; per_pair[p].assign (sp + soff[q], sp + soff[q + 1]); }
        }
      }
      if (paths) {
        // group mode: keep the best pair of each group only
        if (group) for (size_t g0 = 0; g0 < n_pairs; g0 += group) {
          size_t best = g0; bool any = false;
          for (size_t q = g0; q < std::min (n_pairs, g0 + group); ++q) if (score[q] > -INFINITY && (!any || score[q] > score[best])) { best = q; any = true; }
          for (size_t q = g0; q < std::min (n_pairs, g0 + group); ++q) if (!any || q != best) { per_pair[q].clear (); xs[q] = 0; xe[q] = 0; }
        }
        uint64_t total = 0;
        for (size_t p = 0; p < n_pairs; ++p) { path_offsets[p] = total; total += per_pair[p].size (); x_start[p] = xs[p]; x_end[p] = xe[p]; }
        path_offsets[n_pairs] = total;
        uint8_t* buf = (uint8_t*) malloc (total + 1);
        if (!buf) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
        for (size_t p = 0; p < n_pairs; ++p) if (!per_pair[p].empty ()) memcpy (buf + path_offsets[p], per_pair[p].data (), per_pair[p].size ());
        *path_out = buf;
      }
      return QG_OK;
    }
  }

  // sub-batches bounded by pointer memory
  size_t freeb = 0, totb = 0;
  QG_CUDA (ctx, cudaMemGetInfo (&freeb, &totb));
  const uint64_t budget = (uint64_t) qg_env_size ("QG_TRACE_BUDGET_MB", std::min<size_t> (freeb / 2, (size_t) 32 << 30) >> 20) << 20;
  qg_hostbuf all_paths_;                                     // grows by realloc: no zero fill, handed to the caller at the end, freed on any error return
  uint8_t*& all_paths = all_paths_.p;
  std::vector<uint64_t> offs (n_pairs + 1, 0);
  uint64_t path_total = 0;
  const size_t step = group ? group : 1;
  size_t p0 = 0;
  while (p0 < n_pairs) {
    size_t p1 = p0; uint64_t words = 0;
    while (p1 < n_pairs) {
      uint64_t w = 0;
      const size_t pe = std::min (n_pairs, p1 + step);
      for (size_t q = p1; q < pe; ++q)
        for (uint32_t r = er.run_begin[q]; r < er.run_begin[q + 1]; ++r) {
          int R, nw; const uint32_t width = (uint32_t) (er.runs[r].y - er.runs[r].x + 1);
          if (qg_pick_R (width, &R, &nw) != QG_OK) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "pair %zu: a run of %u consecutive diagonals exceeds the %d this build fills with one CTA", q, width, 256 * QG_MAX_NW);
          w += ((uint64_t) ctx->seqs[QG_READS].len (yi[q]) + 32ull * nw + 1) * 32ull * nw;
        }
      if (p1 > p0 && (words + w) * 4 > budget) break;
      words += w; p1 = pe;
    }
    qg_htrace ("envelopes");
    qg_dp_plan plan;
    QG_TRY (qg_build_plan (ctx, er, p0, p1, xi, yi, QG_REFS, 0, plan));
    qg_htrace ("plan");
    const size_t np = p1 - p0;
    for (uint32_t s : plan.order) plan.segs_sorted.push_back (plan.segs[s]);
    {
      qg_timer tm (ctx, &ctx->stats.ms_prep);
      QG_TRY (qg_stage_rowparams (ctx, plan));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGS], plan.segs_sorted.data (), sizeof (qg_segment) * plan.segs_sorted.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], plan.segs.data (), sizeof (qg_segment) * plan.segs.size ()));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_TRACE], sizeof (uint32_t) * (plan.trace_words + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDVALS], sizeof (double) * 2 * (plan.segs.size () + 1)));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], plan.pairs.data (), sizeof (qg_pair_dp) * np));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT0], sizeof (double) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT1], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT2], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT3], sizeof (uint32_t) * (np + 1)));
    }
    ctx->stats.trace_bytes += plan.trace_words * 4;
    ctx->stats.n_segments += plan.segs.size ();
    ctx->stats.cell_updates += qg_plan_cells (er, p0, p1);
    qg_htrace ("prep");
    {
      qg_timer tm (ctx, &ctx->stats.ms_viterbi);
      qg_fill_args a = qg_base_args (ctx, cfg, QG_REFS);
      a.trace = ctx->scratch[SC_TRACE].as<uint32_t> ();
      a.endvals = ctx->scratch[SC_ENDVALS].as<double> ();       // {score, i} per segment, indexed by seg.aux_off (pair order)
      QG_TRY (qg_launch_fill<0> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
      QG_LAUNCH (qg_pair_score_kernel, (unsigned) ((np + 127) / 128), 128, 0, ctx->stream,
                 ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_ENDVALS].as<double> (),
                 ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_OUT2].as<uint32_t> ());
      QG_TRY (qg_check_launch (ctx, "qg_pair_score_kernel"));
    }
    {
      qg_timer tm (ctx, &ctx->stats.ms_d2h);
      QG_TRY (qg_download (ctx, score + p0, ctx->scratch[SC_OUT0].p, sizeof (double) * np));
    }
    qg_htrace ("fill+scores");
    if (paths) {
      // which pairs get a traceback
      uint64_t scratch_bytes = 0;
      for (size_t p = 0; p < np; ++p) plan.pairs[p].want_path = 0;
      if (group) {
        for (size_t g0 = 0; g0 < np; g0 += group) {
          size_t best = g0; bool any = false;
          for (size_t q = g0; q < std::min (np, g0 + group); ++q) {
            const double sc = score[p0 + q];
            if (sc > -INFINITY && (!any || sc > score[p0 + best])) { best = q; any = true; }
          }
          if (any) plan.pairs[best].want_path = 1;
        }
      } else {
        for (size_t p = 0; p < np; ++p) plan.pairs[p].want_path = (!want_path || want_path[p0 + p]) ? 1 : 0;
      }
      for (size_t p = 0; p < np; ++p) {
        plan.pairs[p].path_off = scratch_bytes;
        if (plan.pairs[p].want_path) scratch_bytes += plan.pairs[p].path_cap;
      }
      std::vector<uint32_t> plen (np), xs (np), xe (np);
      uint32_t flag = 0;
      {
        qg_timer tm (ctx, &ctx->stats.ms_traceback);
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], plan.pairs.data (), sizeof (qg_pair_dp) * np));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHSCR], scratch_bytes + 16));
        QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_FLAGS].p, 0, 64, ctx->stream));
        QG_LAUNCH (qg_traceback_warp_kernel, (unsigned) ((np + QG_TB_WARPS - 1) / QG_TB_WARPS), 32 * QG_TB_WARPS, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                   ctx->scratch[SC_TRACE].as<uint32_t> (), ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_OUT2].as<uint32_t> (),
                   ctx->scratch[SC_OUT1].as<uint32_t> (), ctx->scratch[SC_PATHSCR].as<uint8_t> (), ctx->scratch[SC_OUT3].as<uint32_t> (),
                   (uint32_t*) ctx->scratch[SC_FLAGS].p);
        QG_TRY (qg_check_launch (ctx, "qg_traceback_kernel"));
      }
      {
        qg_timer tm (ctx, &ctx->stats.ms_d2h);
        QG_TRY (qg_fetch (ctx, xs.data (), ctx->scratch[SC_OUT1].p, sizeof (uint32_t) * np));
        QG_TRY (qg_fetch (ctx, xe.data (), ctx->scratch[SC_OUT2].p, sizeof (uint32_t) * np));
        QG_TRY (qg_fetch (ctx, plen.data (), ctx->scratch[SC_OUT3].p, sizeof (uint32_t) * np));
        QG_TRY (qg_fetch (ctx, &flag, ctx->scratch[SC_FLAGS].p, sizeof (uint32_t)));
        QG_TRY (qg_fetch_wait (ctx));
      }
      if (flag) QG_FAIL (ctx, QG_ERR_CUDA, "internal: traceback left the envelope (code %u)", flag);
      for (size_t p = 0; p < np; ++p) {
        const bool w = plan.pairs[p].want_path != 0;
        x_start[p0 + p] = w ? xs[p] : 0; x_end[p0 + p] = w ? xe[p] : 0;
      }
      std::vector<uint64_t> goff (np + 1, 0);
      for (size_t p = 0; p < np; ++p) goff[p + 1] = goff[p] + plen[p];
      for (size_t p = 0; p < np; ++p) offs[p0 + p] = path_total + goff[p];
      if (goff[np]) {
        qg_timer tm (ctx, &ctx->stats.ms_d2h);
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC1], goff.data (), sizeof (uint64_t) * (np + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHOUT], goff[np] + 16));
        QG_LAUNCH (qg_path_gather_kernel, (unsigned) np, 128, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_PATHSCR].as<uint8_t> (),
                   ctx->scratch[SC_OUT3].as<uint32_t> (), ctx->scratch[SC_MISC1].as<uint64_t> (), ctx->scratch[SC_PATHOUT].as<uint8_t> ());
        QG_TRY (qg_check_launch (ctx, "qg_path_gather_kernel"));
        uint8_t* grown = (uint8_t*) realloc (all_paths, path_total + goff[np] + 1);
        if (!grown) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
        all_paths = grown;
        QG_TRY (qg_download (ctx, all_paths + path_total, ctx->scratch[SC_PATHOUT].p, goff[np]));
      }
      path_total += goff[np];
    }
    p0 = p1;
  }
  if (paths) {
    offs[n_pairs] = path_total;
    memcpy (path_offsets, offs.data (), sizeof (uint64_t) * (n_pairs + 1));
    if (!all_paths) all_paths = (uint8_t*) malloc (1);
    if (!all_paths) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
    *path_out = all_paths_.release ();
  }
  return QG_OK;
}

extern "C" int qg_viterbi (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                           const uint8_t* want_path, double* score, uint32_t* x_start, uint32_t* x_end,
                           uint8_t** path_out, uint64_t* path_offsets) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !xi || !yi || !score) return QG_ERR_INVALID;
  return qg_viterbi_impl (ctx, cfg, n_pairs, xi, yi, want_path, 0, score, x_start, x_end, path_out, path_offsets);
}

// ---- seam A: QuaffAligner::align (qmodel.cpp:2624-2646) = QuaffAlignmentTask::run for every read (qmodel.cpp:2764-2778)
extern "C" int qg_align_reads_range (qg_ctx* ctx, const qg_dpconfig* cfg, size_t first_read, size_t n_reads, const double* null_loglike,
                                     uint32_t* best_ref, double* score, uint32_t* x_start, uint32_t* x_end,
                                     uint8_t** path_out, uint64_t* path_offsets) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !null_loglike || !best_ref || !score || !x_start || !x_end || !path_out || !path_offsets) return QG_ERR_INVALID;
  QG_TRY (qg_check_ready (ctx, cfg));
  const size_t nx = ctx->seqs[QG_REFS].n, ny = n_reads;
  if (first_read + n_reads > ctx->seqs[QG_READS].n) QG_FAIL (ctx, QG_ERR_INVALID, "read range exceeds the READS set");
  const size_t np = nx * ny;
  std::vector<uint32_t> xi (np), yi (np);
  for (size_t y = 0; y < ny; ++y) for (size_t x = 0; x < nx; ++x) { xi[y * nx + x] = (uint32_t) x; yi[y * nx + x] = (uint32_t) (first_read + y); }
  std::vector<double> sc (np);
  std::vector<uint32_t> xs (np), xe (np);
  std::vector<uint64_t> poff (np + 1);
  uint8_t* paths = nullptr;
  qg_htrace ("entry");
  QG_TRY (qg_viterbi_impl (ctx, cfg, np, xi.data (), yi.data (), nullptr, nx, sc.data (), xs.data (), xe.data (), &paths, poff.data ()));
  qg_htrace ("traceback+out");
  // the traced pair of each read is its best one; compact the path list to one entry per read
  uint64_t total = 0;
  for (size_t y = 0; y < ny; ++y) {
    best_ref[y] = 0xFFFFFFFFu; score[y] = -INFINITY; x_start[y] = 0; x_end[y] = 0;
    path_offsets[y] = total;
    size_t best = 0; bool any = false;
    for (size_t x = 0; x < nx; ++x) {
      const double v = sc[y * nx + x];
      if (v > -INFINITY && (!any || v > sc[y * nx + best])) { best = x; any = true; }
    }
    if (any) {
      const size_t p = y * nx + best;
      best_ref[y] = (uint32_t) best;
      score[y] = sc[p] - null_loglike[y];                       // scoreAdjustedAlignment, qmodel.cpp:1648-1654
      x_start[y] = xs[p]; x_end[y] = xe[p];
      const uint64_t len = poff[p + 1] - poff[p];
      memmove (paths + total, paths + poff[p], len);            // total <= poff[p]: compaction moves data towards the front
      total += len;
    }
  }
  path_offsets[ny] = total;
  *path_out = paths;
  qg_htrace ("compact");
  return QG_OK;
}

extern "C" int qg_align_reads (qg_ctx* ctx, const qg_dpconfig* cfg, const double* null_loglike,
                               uint32_t* best_ref, double* score, uint32_t* x_start, uint32_t* x_end,
                               uint8_t** path_out, uint64_t* path_offsets) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx) return QG_ERR_INVALID;
  return qg_align_reads_range (ctx, cfg, 0, ctx->seqs[QG_READS].n, null_loglike, best_ref, score, x_start, x_end, path_out, path_offsets);
}

// ---- probability-space Forward / Backward (qg_prob.cuh): used unless QG_OPT_FB_EXACT or a run needs several warps ----
static bool qg_plan_single_warp (const qg_dp_plan& plan) {
  for (const auto& L : plan.launches) if (L.nw > 1) return false;
  return true;
}

static int qg_stage_rowq (qg_ctx* ctx, const qg_dp_plan& plan) {
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_RQ], sizeof (qg_rowq) * (plan.rp_rows + 1)));
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_RS], sizeof (double) * (plan.rp_rows + 1)));
  QG_TRY (qg_reserve (ctx, ctx->scratch[SC_RPS], sizeof (qg_rowq) * (plan.rp_rows + 1)));     // structure-of-arrays copy (the Viterbi kernels' slot: not live in a Forward / Backward call)
  if (!plan.rp_jobs.empty ()) {
    QG_LAUNCH (qg_rowq_kernel, (unsigned) plan.rp_jobs.size (), 256, 0, ctx->stream,
               ctx->scratch[SC_RPJOBS].as<qg_rp_job> (), ctx->scratch[SC_RP].as<qg_rowp> (), ctx->scratch[SC_RQ].as<qg_rowq> (), ctx->scratch[SC_RS].as<double> (),
               ctx->scratch[SC_RPS].as<double2> ());
    QG_TRY (qg_check_launch (ctx, "qg_rowq_kernel"));
  }
  return QG_OK;
}

static qg_prob_args qg_prob_base_args (qg_ctx* ctx, const qg_dpconfig* cfg, int x_set) {
  qg_prob_args a;
  memset (&a, 0, sizeof (a));
  a.xpacked = ctx->seqs[x_set].d_packed.as<uint64_t> ();
  a.xpoff = ctx->seqs[x_set].d_poff.as<uint64_t> ();
  a.rq = ctx->scratch[SC_RQ].as<qg_rowq> ();
  a.rqs = ctx->scratch[SC_RPS].as<double2> ();
  a.rs = ctx->scratch[SC_RS].as<double> ();
  a.pi2i = exp (ctx->model.i2i); a.pi2m = exp (ctx->model.i2m); a.pd2d = exp (ctx->model.d2d); a.pd2m = exp (ctx->model.d2m);
  a.local = cfg->local;
  return a;
}

template<int BACKWARD>
static int qg_launch_prob (qg_ctx* ctx, const qg_dp_plan& plan, qg_prob_args a, const qg_segment* d_segs_launch_order) {
  QG_TRY (qg_fork (ctx));
  int kcls = 0;
  for (const auto& L : plan.launches) {                   // plan order: small-first (qg_launch_order) measured no gain for these kernels
    cudaStream_t st; QG_TRY (qg_side (ctx, kcls++, &st));
    a.segs = d_segs_launch_order + L.begin;
    if (L.nw == 0) {                                        // isolated diagonals: one thread each (Forward), closed form (Backward)
      if (BACKWARD) { auto kfn = qg_backward_prob_kernel<2>; QG_LAUNCH (kfn, L.count, 32, 0, st, a); }
      else QG_LAUNCH (qg_forward_prob_diag_kernel, (L.count + 63) / 64, 64, 0, st, a, L.count);
      QG_TRY (qg_check_launch (ctx, BACKWARD ? "qg_backward_prob_kernel (diagonal)" : "qg_forward_prob_diag_kernel"));
      continue;
    }
#define QG_CASE(RR) case RR: \
      if (BACKWARD) { auto kfn = qg_backward_prob_kernel<RR>; QG_LAUNCH (kfn, L.count, 32, 0, st, a); } \
      else { auto kfn = qg_forward_prob_kernel<RR>; QG_LAUNCH (kfn, L.count, 32, 0, st, a); } break;
    switch (L.R) {
      QG_CASE (2) QG_CASE (3) QG_CASE (4) QG_CASE (5) QG_CASE (6) QG_CASE (7) QG_CASE (8)
      default: QG_FAIL (ctx, QG_ERR_INVALID, "internal: unsupported R=%d", L.R);
    }
#undef QG_CASE
    QG_TRY (qg_check_launch (ctx, BACKWARD ? "qg_backward_prob_kernel" : "qg_forward_prob_kernel"));
  }
  QG_TRY (qg_join (ctx, kcls));
  return QG_OK;
}

// ---- Forward --------------------------------------------------------------------------------------------------------
extern "C" int qg_forward (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* loglike) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !xi || !yi || !loglike) return QG_ERR_INVALID;
  QG_TRY (qg_check_ready (ctx, cfg));
  qg_env_result er;
  QG_TRY (qg_envelope_stage (ctx, cfg, 48, QG_REFS, n_pairs, xi, yi, er));
  if (g_env_reuse.keep) *g_env_reuse.keep = er;
  {
    std::vector<size_t> wide, narrow;
    for (size_t p = 0; p < n_pairs; ++p) (qg_pair_is_wide (er, p) ? wide : narrow).push_back (p);
    if (!wide.empty ()) {
      QG_TRY (qg_wide_run (ctx, cfg, er, wide, xi, yi, 1, nullptr, loglike, nullptr, nullptr, nullptr));
      if (!narrow.empty ()) {                               // the rest through the ordinary path, with the envelopes already computed
        std::vector<uint32_t> sxi, syi; std::vector<double> sll (narrow.size ());
        for (size_t p : narrow) { sxi.push_back (xi[p]); syi.push_back (yi[p]); }
        const qg_env_reuse saved = g_env_reuse;
        g_env_reuse.keep = nullptr; g_env_reuse.src = &er; g_env_reuse.pick = &narrow;
        const int rc = qg_forward (ctx, cfg, narrow.size (), sxi.data (), syi.data (), sll.data ());
        g_env_reuse = saved;
        QG_TRY (rc);
        for (size_t q = 0; q < narrow.size (); ++q) loglike[narrow[q]] = sll[q];
      }
      return QG_OK;
    }
    if (!ctx->fb_exact) {
      // probability-space kernels for the pairs whose runs fit one warp, log-space kernels for the others (see qg_pair_is_multiwarp)
      std::vector<size_t> part[2];
      for (size_t p = 0; p < n_pairs; ++p) part[qg_pair_is_multiwarp (er, p) ? 1 : 0].push_back (p);
      if (!part[0].empty () && !part[1].empty ()) {
        for (int c = 0; c < 2; ++c) {
          std::vector<uint32_t> sxi, syi; std::vector<double> sll (part[c].size ());
          for (size_t p : part[c]) { sxi.push_back (xi[p]); syi.push_back (yi[p]); }
          const qg_env_reuse saved = g_env_reuse;
          g_env_reuse.keep = nullptr; g_env_reuse.src = &er; g_env_reuse.pick = &part[c];
          const int rc = qg_forward (ctx, cfg, part[c].size (), sxi.data (), syi.data (), sll.data ());
          g_env_reuse = saved;
          QG_TRY (rc);
          for (size_t q = 0; q < part[c].size (); ++q) loglike[part[c][q]] = sll[q];
        }
        return QG_OK;
      }
    }
  }
  qg_dp_plan plan;
  QG_TRY (qg_build_plan (ctx, er, 0, n_pairs, xi, yi, QG_REFS, 1, plan, !ctx->fb_exact && !getenv ("QG_VIT_GENERIC")));
  if (!ctx->fb_exact && !qg_plan_single_warp (plan))        // multi-warp runs use the log-space kernels for the whole call: no thread-per-diagonal class there
    QG_TRY (qg_build_plan (ctx, er, 0, n_pairs, xi, yi, QG_REFS, 1, plan));
  // Forward folds end values per pair in pair order: keep segs in pair order, launch classes through an index
  for (uint32_t s : plan.order) plan.segs_sorted.push_back (plan.segs[s]);
  {
    qg_timer tm (ctx, &ctx->stats.ms_prep);
    QG_TRY (qg_stage_rowparams (ctx, plan));
    QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGS], plan.segs_sorted.data (), sizeof (qg_segment) * plan.segs_sorted.size ()));
    QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], plan.segs.data (), sizeof (qg_segment) * plan.segs.size ()));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDVALS], sizeof (double) * (plan.aux_slots + 1)));
    QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], plan.pairs.data (), sizeof (qg_pair_dp) * n_pairs));
    QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT0], sizeof (double) * (n_pairs + 1)));
  }
  ctx->stats.n_segments += plan.segs.size ();
  ctx->stats.cell_updates += qg_plan_cells (er, 0, n_pairs);
  const bool fast = !ctx->fb_exact && qg_plan_single_warp (plan);
  if (fast) {
    {
      qg_timer tm (ctx, &ctx->stats.ms_prep);
      QG_TRY (qg_stage_rowq (ctx, plan));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDEX], sizeof (int) * (plan.aux_slots + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ZM], sizeof (double) * (n_pairs + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ZE], sizeof (int) * (n_pairs + 1)));
    }
    qg_timer tm (ctx, &ctx->stats.ms_forward);
    qg_prob_args a = qg_prob_base_args (ctx, cfg, QG_REFS);
    a.endvals = ctx->scratch[SC_ENDVALS].as<double> (); a.endex = ctx->scratch[SC_ENDEX].as<int> ();
    QG_TRY (qg_launch_prob<0> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
    QG_LAUNCH (qg_forward_prob_finalize_kernel, (unsigned) ((n_pairs + 63) / 64), 64, 0, ctx->stream,
               ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) n_pairs, ctx->scratch[SC_MISC0].as<qg_segment> (),
               ctx->scratch[SC_ENDVALS].as<double> (), ctx->scratch[SC_ENDEX].as<int> (), ctx->scratch[SC_RS].as<double> (),
               ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_ZM].as<double> (), ctx->scratch[SC_ZE].as<int> ());
    QG_TRY (qg_check_launch (ctx, "qg_forward_prob_finalize_kernel"));
  } else {
    qg_timer tm (ctx, &ctx->stats.ms_forward);
    qg_fill_args a = qg_base_args (ctx, cfg, QG_REFS);
    a.endvals = ctx->scratch[SC_ENDVALS].as<double> ();
    QG_TRY (qg_launch_fill<1> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
    QG_LAUNCH (qg_forward_finalize_kernel, (unsigned) ((n_pairs + 63) / 64), 64, 0, ctx->stream,
               ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) n_pairs, ctx->scratch[SC_MISC0].as<qg_segment> (),
               ctx->scratch[SC_ENDVALS].as<double> (), ctx->d_lse.as<double> (), ctx->scratch[SC_OUT0].as<double> ());
    QG_TRY (qg_check_launch (ctx, "qg_forward_finalize_kernel"));
  }
  {
    qg_timer tm (ctx, &ctx->stats.ms_d2h);
    QG_TRY (qg_download (ctx, loglike, ctx->scratch[SC_OUT0].p, sizeof (double) * n_pairs));
  }
  return QG_OK;
}

// ---- Backward + counts ----------------------------------------------------------------------------------------------
template<int DUMMY>
static int qg_launch_backward (qg_ctx* ctx, const qg_dp_plan& plan, qg_fill_args a, const qg_segment* d_segs_launch_order) {
  QG_TRY (qg_fork (ctx));
  int kcls = 0;
  for (const auto& L : plan.launches) {
    cudaStream_t st; QG_TRY (qg_side (ctx, kcls++, &st));
    a.segs = d_segs_launch_order + L.begin;
    const bool multi = L.nw > 1;
    const unsigned block = 32u * L.nw;
#define QG_CASE(RR) case RR: \
      if (multi) { auto kfn = qg_backward_kernel<8, true>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } \
      else { auto kfn = qg_backward_kernel<RR, false>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } break;
    switch (L.R) {
      QG_CASE (2) QG_CASE (3) QG_CASE (4) QG_CASE (5) QG_CASE (6) QG_CASE (7) QG_CASE (8)
      default: QG_FAIL (ctx, QG_ERR_INVALID, "internal: unsupported R=%d", L.R);
    }
#undef QG_CASE
    QG_TRY (qg_check_launch (ctx, "qg_backward_kernel"));
  }
  QG_TRY (qg_join (ctx, kcls));
  return QG_OK;
}

// The E-step's single pass (qg_estep): pair groups (one per read) that a memory batch must not split, and a callback that
// turns a batch's Forward results into its posterior weights once they are on the host.  QG_RETRY_TWO_PASS: the batch
// contains wide pairs, which this form does not serve -- the caller falls back to Forward first, Backward second.
#define QG_RETRY_TWO_PASS (-1000)
struct qg_weight_hook {
  const std::vector<size_t>* group_first = nullptr;         // [n_groups + 1] pair indices
  std::function<void (size_t p0, size_t p1, const double* fwd, double* w)> fn;
};
static thread_local const qg_weight_hook* g_weight_hook = nullptr;

extern "C" int qg_backward_counts (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                                   const double* weights, double* fwd_loglike, double* back_loglike,
                                   double* counts_sum, double* counts_per_pair) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !xi || !yi) return QG_ERR_INVALID;
  const qg_weight_hook* hook = g_weight_hook;
  g_weight_hook = nullptr;                                  // this call only (the wide/narrow split below re-enters)
  QG_TRY (qg_check_ready (ctx, cfg));
  const qg_seqset& Y = ctx->seqs[QG_READS];
  if (!Y.has_qual) QG_FAIL (ctx, QG_ERR_INVALID, "Forward-Backward requires quality scores (qmodel.cpp:1398)");
  const qg_model_dev& m = ctx->model;
  const uint64_t nC = qg_counts_size (m.match_k, m.gap_k);
  qg_env_result er;
  QG_TRY (qg_envelope_stage (ctx, cfg, 48, QG_REFS, n_pairs, xi, yi, er));
  {
    // pairs with runs too wide for one CTA take the tiled path (qg_tile.cuh: log space, every Forward cell kept); the
    // others go through this function again on their own, with the envelopes already computed
    // (pairs whose runs need several warps but fit one CTA: log-space kernels; the rest of a mixed call keeps the probability-space ones)
    std::vector<size_t> wide, part[2];
    for (size_t p = 0; p < n_pairs; ++p) {
      if (qg_pair_is_wide (er, p)) wide.push_back (p);
      else part[(!ctx->fb_exact && qg_pair_is_multiwarp (er, p)) ? 1 : 0].push_back (p);
    }
    const bool mixed = !part[0].empty () && !part[1].empty ();
    if ((!wide.empty () || mixed) && hook) return QG_RETRY_TWO_PASS;
    if (!wide.empty () || mixed) {
      std::vector<double> sum (nC, 0.0);
      for (int cls = 0; cls < 2; ++cls) {
        const std::vector<size_t>& narrow = part[cls];
        if (narrow.empty ()) continue;
        std::vector<double> csum (nC, 0.0);
        std::vector<uint32_t> nx (narrow.size ()), ny (narrow.size ());
        std::vector<double> nw (narrow.size (), 1.0), nf (narrow.size ()), nb (narrow.size ()), npp;
        for (size_t q = 0; q < narrow.size (); ++q) { nx[q] = xi[narrow[q]]; ny[q] = yi[narrow[q]]; if (weights) nw[q] = weights[narrow[q]]; }
        if (counts_per_pair) npp.resize (nC * narrow.size ());
        const qg_env_reuse saved = g_env_reuse;
        g_env_reuse.keep = nullptr; g_env_reuse.src = &er; g_env_reuse.pick = &narrow;
        const int rc = qg_backward_counts (ctx, cfg, narrow.size (), nx.data (), ny.data (), weights ? nw.data () : nullptr, nf.data (), nb.data (),
                                           csum.data (), counts_per_pair ? npp.data () : nullptr);
        g_env_reuse = saved;
        QG_TRY (rc);
        for (uint64_t k = 0; k < nC; ++k) sum[k] += csum[k];
        for (size_t q = 0; q < narrow.size (); ++q) {
          if (fwd_loglike) fwd_loglike[narrow[q]] = nf[q];
          if (back_loglike) back_loglike[narrow[q]] = nb[q];
          if (counts_per_pair) memcpy (counts_per_pair + (uint64_t) narrow[q] * nC, npp.data () + q * nC, sizeof (double) * nC);
        }
      }
      if (!wide.empty ()) {
        qg_wide_fb fb; fb.weights = weights; fb.back = back_loglike; fb.counts_sum = sum.data (); fb.counts_per_pair = counts_per_pair; fb.nC = nC;
        QG_TRY (qg_wide_run (ctx, cfg, er, wide, xi, yi, 2, nullptr, fwd_loglike, nullptr, nullptr, nullptr, &fb));
      }
      if (counts_sum) memcpy (counts_sum, sum.data (), sizeof (double) * nC);
      QG_CUDA (ctx, qg_sync (ctx));
      return QG_OK;
    }
  }
  size_t freeb = 0, totb = 0;
  QG_CUDA (ctx, cudaMemGetInfo (&freeb, &totb));
  // what this context already holds for the store is available to it again: count it with the free memory
  const size_t avail = freeb + ctx->scratch[SC_STORE].cap;
  const uint64_t budget = (uint64_t) qg_env_size ("QG_STORE_BUDGET_MB", std::min<size_t> (avail / 20 * 13, (size_t) 128 << 30) >> 20) << 20;
  qg_dbuf& dSum = ctx->scratch[SC_OUT3];
  QG_TRY (qg_reserve (ctx, dSum, sizeof (double) * (nC + 1)));
  QG_CUDA (ctx, cudaMemsetAsync (dSum.p, 0, sizeof (double) * nC, ctx->stream));
  std::vector<double> hook_w, hook_f;
  size_t p0 = 0, g0 = 0;                                     // g0: first group of the batch (hook only)
  while (p0 < n_pairs) {
    size_t p1 = p0; uint64_t bytes = 0;
    auto pair_bytes = [&] (size_t p, uint64_t* out) -> int {
      uint64_t b = nC * 8;
      for (uint32_t r = er.run_begin[p]; r < er.run_begin[p + 1]; ++r) {
        int R, nw; const uint32_t width = (uint32_t) (er.runs[r].y - er.runs[r].x + 1);
        if (qg_pick_R (width, &R, &nw) != QG_OK) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "pair %zu: a run of %u consecutive diagonals exceeds the %d this build fills with one CTA", p, width, 256 * QG_MAX_NW);
        b += ((uint64_t) Y.len (yi[p]) + 32ull * nw + 1) * 3 * 32ull * nw * R * 8 + ((uint64_t) Y.len (yi[p]) + 2) * 64;
      }
      *out = b; return QG_OK;
    };
    if (hook) {
      // whole groups (all pairs of a read) per batch: the weights of a read need the Forward results of all its pairs
      const std::vector<size_t>& gf = *hook->group_first;
      size_t g1 = g0;
      while (g1 + 1 < gf.size ()) {
        uint64_t gb = 0;
        for (size_t p = gf[g1]; p < gf[g1 + 1]; ++p) { uint64_t b; QG_TRY (pair_bytes (p, &b)); gb += b; }
        if (g1 > g0 && bytes + gb > budget) break;
        bytes += gb; ++g1;
      }
      p1 = gf[g1]; g0 = g1;
      if (p1 == p0) { if (g1 + 1 >= gf.size ()) break; continue; }      // empty groups only
    } else
    while (p1 < n_pairs) {
      uint64_t b; QG_TRY (pair_bytes (p1, &b));
      if (p1 > p0 && bytes + b > budget) break;
      bytes += b; ++p1;
    }
    const size_t np = p1 - p0;
    qg_dp_plan plan;
    QG_TRY (qg_build_plan (ctx, er, p0, p1, xi, yi, QG_REFS, 2, plan, !ctx->fb_exact && !getenv ("QG_VIT_GENERIC")));
    if (!ctx->fb_exact && !qg_plan_single_warp (plan)) QG_TRY (qg_build_plan (ctx, er, p0, p1, xi, yi, QG_REFS, 2, plan));   // log-space kernels: no diagonal class
    for (uint32_t s : plan.order) plan.segs_sorted.push_back (plan.segs[s]);
    {
      qg_timer tm (ctx, &ctx->stats.ms_prep);
      QG_TRY (qg_stage_rowparams (ctx, plan));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGS], plan.segs_sorted.data (), sizeof (qg_segment) * plan.segs_sorted.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], plan.segs.data (), sizeof (qg_segment) * plan.segs.size ()));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDVALS], sizeof (double) * (plan.aux_slots + 1)));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], plan.pairs.data (), sizeof (qg_pair_dp) * np));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT0], sizeof (double) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT1], sizeof (double) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_STORE], sizeof (double) * (plan.store_doubles + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ROWACC], sizeof (qg_rowrec) * (plan.acc_rows + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_MISC1], sizeof (double) * 12 * (plan.segs.size () + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHSCR], sizeof (double) * nC * np + 16));
      QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_MISC1].p, 0, sizeof (double) * 12 * (plan.segs.size () + 1), ctx->stream));
      QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_PATHSCR].p, 0, sizeof (double) * nC * np, ctx->stream));
      QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_ROWACC].p, 0, sizeof (qg_rowrec) * (plan.acc_rows + 1), ctx->stream));
    }
    ctx->stats.n_segments += plan.segs.size ();
    ctx->stats.cell_updates += 2 * qg_plan_cells (er, p0, p1);
    ctx->stats.fwd_store_bytes += plan.store_doubles * 8;
    const bool fast = !ctx->fb_exact && qg_plan_single_warp (plan);
    if (fast) {
      {
        qg_timer tm (ctx, &ctx->stats.ms_prep);
        QG_TRY (qg_stage_rowq (ctx, plan));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ENDEX], sizeof (int) * (plan.aux_slots + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_STOREEX], sizeof (int) * 32 * (plan.acc_rows + 32 * plan.segs.size () + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ZM], sizeof (double) * (np + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_ZE], sizeof (int) * (np + 1)));
      }
      qg_prob_args a = qg_prob_base_args (ctx, cfg, QG_REFS);
      a.endvals = ctx->scratch[SC_ENDVALS].as<double> (); a.endex = ctx->scratch[SC_ENDEX].as<int> ();
      a.store = ctx->scratch[SC_STORE].as<double> (); a.store_ex = ctx->scratch[SC_STOREEX].as<int> ();
      a.rowacc = ctx->scratch[SC_ROWACC].as<double> ();
      a.pair_zm = ctx->scratch[SC_ZM].as<double> (); a.pair_ze = ctx->scratch[SC_ZE].as<int> ();
      a.seg_scal = ctx->scratch[SC_MISC1].as<double> ();
      a.do_store = 1;
      {
        qg_timer tm (ctx, &ctx->stats.ms_forward);
        QG_TRY (qg_launch_prob<0> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
        QG_LAUNCH (qg_forward_prob_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                   ctx->scratch[SC_ENDVALS].as<double> (), ctx->scratch[SC_ENDEX].as<int> (), ctx->scratch[SC_RS].as<double> (),
                   ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_ZM].as<double> (), ctx->scratch[SC_ZE].as<int> ());
        QG_TRY (qg_check_launch (ctx, "qg_forward_prob_finalize_kernel"));
      }
      {
        qg_timer tm (ctx, &ctx->stats.ms_backward);
        QG_TRY (qg_launch_prob<1> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
        QG_LAUNCH (qg_backward_prob_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                   ctx->scratch[SC_ENDVALS].as<double> (), ctx->scratch[SC_ENDEX].as<int> (), ctx->scratch[SC_RS].as<double> (),
                   ctx->scratch[SC_OUT1].as<double> ());
        QG_TRY (qg_check_launch (ctx, "qg_backward_prob_finalize_kernel"));
      }
    } else {
    qg_fill_args a = qg_base_args (ctx, cfg, QG_REFS);
    a.endvals = ctx->scratch[SC_ENDVALS].as<double> ();
    a.store = ctx->scratch[SC_STORE].as<double> ();
    a.rowacc = ctx->scratch[SC_ROWACC].as<double> ();
    a.pair_z = ctx->scratch[SC_OUT0].as<double> ();
    a.seg_scal = ctx->scratch[SC_MISC1].as<double> ();
    {
      qg_timer tm (ctx, &ctx->stats.ms_forward);
      QG_TRY (qg_launch_fill<2> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
      QG_LAUNCH (qg_forward_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                 ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                 ctx->scratch[SC_ENDVALS].as<double> (), ctx->d_lse.as<double> (), ctx->scratch[SC_OUT0].as<double> ());
      QG_TRY (qg_check_launch (ctx, "qg_forward_finalize_kernel"));
    }
    {
      qg_timer tm (ctx, &ctx->stats.ms_backward);
      QG_TRY (qg_launch_backward<0> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
      QG_LAUNCH (qg_backward_finalize_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                 ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                 ctx->scratch[SC_ENDVALS].as<double> (), ctx->d_lse.as<double> (), ctx->scratch[SC_OUT1].as<double> ());
      QG_TRY (qg_check_launch (ctx, "qg_backward_finalize_kernel"));
    }
    }
    {
      qg_timer tm (ctx, &ctx->stats.ms_backward);
      QG_LAUNCH (qg_counts_scatter_kernel, (unsigned) np, 256, 0, ctx->stream,
                 ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), ctx->scratch[SC_MISC0].as<qg_segment> (),
                 Y.d_tok.as<uint8_t> (), Y.d_qual.as<uint8_t> (), Y.d_off.as<uint64_t> (),
                 (const qg_rowrec*) ctx->scratch[SC_ROWACC].p, ctx->scratch[SC_MISC1].as<double> (),
                 m.match_k, m.gap_k, nC, ctx->scratch[SC_PATHSCR].as<double> ());
      QG_TRY (qg_check_launch (ctx, "qg_counts_scatter_kernel"));
      const double* dW = nullptr;
      if (hook) {
        // the batch's Forward results -> its posterior weights (the host waits here; Backward and the scatter are already queued)
        hook_f.resize (np); hook_w.assign (np, 0.0);
        QG_TRY (qg_download (ctx, hook_f.data (), ctx->scratch[SC_OUT0].p, sizeof (double) * np));
        hook->fn (p0, p1, hook_f.data (), hook_w.data ());
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_OUT2], hook_w.data (), sizeof (double) * np)); dW = ctx->scratch[SC_OUT2].as<double> ();
      } else
      if (weights) { QG_TRY (qg_upload (ctx, ctx->scratch[SC_OUT2], weights + p0, sizeof (double) * np)); dW = ctx->scratch[SC_OUT2].as<double> (); }
      QG_LAUNCH (qg_counts_reduce_kernel, (unsigned) ((nC + 127) / 128), 128, 0, ctx->stream,
                 ctx->scratch[SC_PATHSCR].as<double> (), dW, (uint32_t) np, nC, dSum.as<double> ());
      QG_TRY (qg_check_launch (ctx, "qg_counts_reduce_kernel"));
    }
    {
      qg_timer tm (ctx, &ctx->stats.ms_d2h);
      if (fwd_loglike) QG_TRY (qg_fetch (ctx, fwd_loglike + p0, ctx->scratch[SC_OUT0].p, sizeof (double) * np));
      if (back_loglike) QG_TRY (qg_fetch (ctx, back_loglike + p0, ctx->scratch[SC_OUT1].p, sizeof (double) * np));
      if (counts_per_pair) QG_TRY (qg_fetch (ctx, counts_per_pair + (uint64_t) p0 * nC, ctx->scratch[SC_PATHSCR].p, sizeof (double) * nC * np));
      QG_TRY (qg_fetch_wait (ctx));
    }
    p0 = p1;
  }
  if (counts_sum) { qg_timer tm (ctx, &ctx->stats.ms_d2h); QG_TRY (qg_download (ctx, counts_sum, dSum.p, sizeof (double) * nC)); }
  QG_CUDA (ctx, qg_sync (ctx));
  return QG_OK;
}

// ---- seam C: QuaffTrainer::getCounts (qmodel.cpp:2005-2032) = QuaffCountingTask::run per read (qmodel.cpp:2238-2271) ----
// The DP runs on the device; what stays here is the reference's per-read control flow: the order-dependent
// "F >= yLL - 20" gate, the posterior weights and the pruning of sortOrder.
static double qg_host_lse (const std::vector<double>& tab, double a, double b) {     // logsumexp.cpp:34-59, 84-103
  double mx, diff;
  if (a == b) { mx = a; diff = 0; } else if (a < b) { mx = b; diff = b - a; } else { mx = a; diff = a - b; }
  double u = 0;
  if (!(diff >= 10.0 || std::isnan (diff) || std::isinf (diff))) {
    const int n = (int) (diff / .0001);
    const double dx = diff - (n * .0001);
    const double f0 = tab[n], f1 = tab[n + 1];
    u = f0 + (f1 - f0) * (dx / .0001);
  }
  return mx + u;
}

extern "C" int qg_estep (qg_ctx* ctx, const qg_dpconfig* cfg, int use_null, const double* null_loglike,
                         uint32_t* sort_order, uint32_t* sort_len, double* y_loglike,
                         double* param_counts, double* loglike_sum) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !sort_order || !sort_len || !y_loglike || !param_counts || !loglike_sum) return QG_ERR_INVALID;
  if (use_null && !null_loglike) QG_FAIL (ctx, QG_ERR_INVALID, "qg_estep: use_null without null_loglike");
  QG_TRY (qg_check_ready (ctx, cfg));
  const size_t nx = ctx->seqs[QG_REFS].n, ny = ctx->seqs[QG_READS].n;
  const int mk = ctx->model.match_k, gk = ctx->model.gap_k;
  const uint64_t nK = qg_pow4 (mk), nG = qg_pow4 (gk), nC = qg_counts_size (mk, gk);
  std::vector<double> tab (100002);
  { const int n = ((int) (10 / .0001)) + 1; for (int t = 0; t < n; ++t) { const double x = t * .0001; tab[t] = log (1. + exp (-x)); } tab[n] = 0; }

  // 1. Forward for every (read, ref in sortOrder) pair
  std::vector<uint32_t> xi, yi; std::vector<size_t> first (ny + 1);
  for (size_t y = 0; y < ny; ++y) {
    first[y] = xi.size ();
    if (sort_len[y] > nx) QG_FAIL (ctx, QG_ERR_INVALID, "qg_estep: sort_len[%zu] > number of references", y);
    for (uint32_t s = 0; s < sort_len[y]; ++s) {
      const uint32_t n = sort_order[y * nx + s];
      if (n >= nx) QG_FAIL (ctx, QG_ERR_INVALID, "qg_estep: sort_order entry out of range");
      xi.push_back (n); yi.push_back ((uint32_t) y);
    }
  }
  first[ny] = xi.size ();
  std::vector<double> F (xi.size ());
  std::vector<std::vector<double> > xyLL (ny, std::vector<double> (nx, -INFINITY));
  std::vector<double> qc (nC, 0.0);
  *loglike_sum = 0;
  // the reference's per-read replay (qmodel.cpp:2247-2270) over the pairs [first[y], first[y+1]) of read y: the
  // order-dependent gate "F >= yLL - 20", the running yLL, the posterior weights of the gated pairs and the new sortOrder
  auto replay_read = [&] (size_t y, const double* Fy, double* wy, std::vector<size_t>* gated_out) {
    double yLL = use_null ? null_loglike[y] : -INFINITY;
    std::vector<size_t> gated;
    const size_t n = first[y + 1] - first[y];
    for (size_t t = 0; t < n; ++t) {
      xyLL[y][xi[first[y] + t]] = Fy[t];
      if (Fy[t] >= yLL - 20) gated.push_back (t);                 // MAX_TRAINING_LOG_DELTA, qmodel.cpp:23, 2252
      yLL = qg_host_lse (tab, yLL, Fy[t]);
    }
    if (wy) { for (size_t t = 0; t < n; ++t) wy[t] = 0.0; for (size_t t : gated) wy[t] = exp (Fy[t] - yLL); }
    if (gated_out) for (size_t t : gated) gated_out->push_back (first[y] + t);
    y_loglike[y] = yLL;
    // sortOrder := refs by descending F, cut at the first one below yLL - 20 (qmodel.cpp:2264-2270)
    std::vector<size_t> idx (nx);
    std::iota (idx.begin (), idx.end (), (size_t) 0);
    const std::vector<double>& v = xyLL[y];
    std::sort (idx.begin (), idx.end (), [&] (size_t a, size_t b) { return v[a] < v[b]; });
    uint32_t len = 0;
    for (size_t t = nx; t-- > 0; ) { if (v[idx[t]] < yLL - 20) break; sort_order[y * nx + len++] = (uint32_t) idx[t]; }
    sort_len[y] = len;
  };

  // 1+3 in one pass (default): Forward with the cells kept and Backward for EVERY pair of a memory batch, the posterior
  // weights computed from the batch's Forward results while its Backward runs; pairs that fail the gate get weight 0
  // (they are the wrong-strand pairs: a handful of narrow runs each, so filling them backwards costs little, whereas
  // the two-pass form filled the right-strand pairs forwards twice)
  bool done = false;
  if (!xi.empty () && cfg->sparse && !getenv ("QG_ESTEP_TWO_PASS")) {
    qg_weight_hook hook;
    hook.group_first = &first;
    hook.fn = [&] (size_t p0, size_t p1, const double* fwd, double* w) {
      // a batch is a range of whole reads: first[y] >= p0 for its first read
      for (size_t y = std::lower_bound (first.begin (), first.end (), p0) - first.begin (); y < ny && first[y] < p1; ++y)
        if (first[y + 1] > first[y]) replay_read (y, fwd + (first[y] - p0), w + (first[y] - p0), nullptr);
    };
    g_weight_hook = &hook;
    const int rc = qg_backward_counts (ctx, cfg, xi.size (), xi.data (), yi.data (), nullptr, F.data (), nullptr, qc.data (), nullptr);
    g_weight_hook = nullptr;
    if (rc == QG_OK) {
      done = true;
      for (size_t y = 0; y < ny; ++y) if (first[y] == first[y + 1]) replay_read (y, nullptr, nullptr, nullptr);   // reads without pairs
    } else if (rc != QG_RETRY_TWO_PASS) return rc;
    else std::fill (qc.begin (), qc.end (), 0.0);
  }
  if (!done) {
    // two passes: Forward for every (read, ref in sortOrder) pair, the replay, then Forward + Backward of the gated pairs
    qg_env_result env_all;                                         // the Forward pass's envelopes, reused by the Backward pass
    if (!xi.empty ()) {
      g_env_reuse.keep = &env_all;
      const int rc = qg_forward (ctx, cfg, xi.size (), xi.data (), yi.data (), F.data ());
      g_env_reuse.keep = nullptr;
      QG_TRY (rc);
    }
    std::vector<uint32_t> bxi, byi; std::vector<double> bw; std::vector<size_t> bpick;
    std::vector<double> wtmp;
    for (size_t y = 0; y < ny; ++y) {
      const size_t before = bpick.size ();
      wtmp.assign (first[y + 1] - first[y], 0.0);
      replay_read (y, F.data () + first[y], wtmp.data (), &bpick);
      for (size_t q = before; q < bpick.size (); ++q) { const size_t p = bpick[q]; bxi.push_back (xi[p]); byi.push_back ((uint32_t) y); bw.push_back (wtmp[p - first[y]]); }
    }
    if (!bxi.empty ()) {
      const bool reuse = env_all.run_begin.size () == xi.size () + 1 && !getenv ("QG_ESTEP_RESEED");
      if (reuse) { g_env_reuse.src = &env_all; g_env_reuse.pick = &bpick; }
      const int rc = qg_backward_counts (ctx, cfg, bxi.size (), bxi.data (), byi.data (), bw.data (), nullptr, nullptr, qc.data (), nullptr);
      g_env_reuse.src = nullptr; g_env_reuse.pick = nullptr;
      QG_TRY (rc);
    }
  }
  for (size_t y = 0; y < ny; ++y) *loglike_sum += y_loglike[y];      // accumulate(yLogLike, 0.), qmodel.cpp:2420-2422
  const size_t nEmit = 4 * nK * QG_NQUAL + 4 * QG_NQUAL;
  memcpy (param_counts, qc.data (), sizeof (double) * nEmit);
  const double *m2m = qc.data () + nEmit, *m2i = m2m + nG, *m2d = m2i + nG, *m2e = m2d + nG, *sc = m2e + nG;
  double *bINo = param_counts + nEmit, *bIYes = bINo + nG, *bDNo = bIYes + nG, *bDYes = bDNo + nG, *ext = bDYes + nG;
  for (uint64_t g = 0; g < nG; ++g) { bINo[g] = m2m[g] + m2d[g]; bIYes[g] = m2i[g] + m2e[g]; bDNo[g] = m2m[g]; bDYes[g] = m2d[g]; }
  ext[0] = sc[3]; ext[1] = sc[2]; ext[2] = sc[1]; ext[3] = sc[0];     // extendInsertNo=i2m, Yes=i2i, extendDeleteNo=d2m, Yes=d2d
  return QG_OK;
}

// ---- overlap ---------------------------------------------------------------------------------------------------------
extern "C" int qg_set_overlap_model (qg_ctx* ctx, const qg_overlap_model* m) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !m) return QG_ERR_INVALID;
  // K <= 2: the pair emission is tabulated (16^K * 94^2 doubles per strand); above that it is evaluated per cell from
  // per-position factors (qg_overlap_fill_kernel, a.ea / a.eb).  G: the three [nG][nG] transition tables (24 * 16^G bytes)
  if (m->match_k < 1 || m->match_k > 6 || m->gap_k < 0 || m->gap_k > 5)
    QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "overlap model orders K=%d G=%d outside 1..6 / 0..5", m->match_k, m->gap_k);
  qg_overlap_dev& d = ctx->omodel;
  d.match_k = m->match_k; d.gap_k = m->gap_k; d.nK = qg_pow4 (m->match_k); d.nG = qg_pow4 (m->gap_k);
  QG_TRY (qg_upload (ctx, d.d_match, m->match, sizeof (double) * 4 * d.nK * QG_NQ1));
  QG_TRY (qg_upload (ctx, d.d_insert, m->insert, sizeof (double) * 4 * QG_NQ1));
  for (int r = 0; r < 4; ++r) d.log_ref_base[r] = m->log_ref_base[r];
  // transitions: qoverlap.cpp:22-48 (host arithmetic, once per model)
  const uint64_t nG = d.nG;
  std::vector<double> gapOpen (nG), m2m (nG * nG), m2i (nG * nG), m2d (nG * nG);
  double sumP = 0, sumA = 0;
  for (uint64_t j = 0; j < nG; ++j) {
    const double readInsertProb = m->begin_insert[j];
    const double readDeleteProb = (1 - m->begin_insert[j]) * m->begin_delete[j];
    gapOpen[j] = readInsertProb + readDeleteProb;
    const double pGapIsInsert = readInsertProb / gapOpen[j];
    const double gapAdjacentProb = pGapIsInsert * readInsertProb + (1 - pGapIsInsert) * gapOpen[j] / (1 - m->extend_delete * (1 - gapOpen[j]));
    sumP += pGapIsInsert; sumA += gapAdjacentProb;
  }
  for (uint64_t i = 0; i < nG; ++i)
    for (uint64_t j = 0; j < nG; ++j) {
      m2m[i * nG + j] = log (1 - gapOpen[i]) + log (1 - gapOpen[j]);
      m2i[i * nG + j] = log (gapOpen[i]);
      m2d[i * nG + j] = log (1 - gapOpen[i]) + log (gapOpen[j]);
    }
  const double pGapIsInsert = sumP / nG;
  const double meanGapLength = pGapIsInsert / m->extend_insert + (1 - pGapIsInsert) / m->extend_delete;
  const double gapExtendProb = 1 / meanGapLength;
  const double gapAdjacentProb = sumA / nG;
  d.i2i = d.d2d = log (gapExtendProb);
  d.i2d = d.d2i = log (1 - gapExtendProb) + log (gapAdjacentProb);
  d.i2m = d.d2m = log (1 - gapExtendProb) + log (1 - gapAdjacentProb);
  QG_TRY (qg_upload (ctx, d.d_m2m, m2m.data (), sizeof (double) * nG * nG));
  QG_TRY (qg_upload (ctx, d.d_m2i, m2i.data (), sizeof (double) * nG * nG));
  QG_TRY (qg_upload (ctx, d.d_m2d, m2d.data (), sizeof (double) * nG * nG));
  QG_CUDA (ctx, qg_sync (ctx));
  d.built[0] = d.built[1] = false; d.d_xonly[0].cap = d.d_xonly[1].cap = 0;
  d.set = true;
  return QG_OK;
}

static int qg_overlap_ensure_table (qg_ctx* ctx, int strand, bool with_qual) {
  qg_overlap_dev& d = ctx->omodel;
  // d_none doubles as "table for this strand without qualities"; d_pair with
  qg_dbuf& T = with_qual ? d.d_pair[strand] : d.d_none[strand];
  const bool have = with_qual ? d.built[strand] : (d.d_none[strand].p != nullptr && d.d_xonly[strand].cap == 1);
  if (have) return QG_OK;
  const uint64_t nq = with_qual ? QG_NQUAL : 1;
  const uint64_t total = d.nK * d.nK * nq * nq;
  QG_TRY (qg_reserve (ctx, T, sizeof (double) * (total + 1)));
  {
    qg_timer tm (ctx, &ctx->stats.ms_prep);
    QG_LAUNCH (qg_overlap_pair_table_kernel, (unsigned) ((total + 127) / 128), 128, 0, ctx->stream,
               d.d_match.as<double> (), d.d_insert.as<double> (), ctx->d_lse.as<double> (),
               d.log_ref_base[0], d.log_ref_base[1], d.log_ref_base[2], d.log_ref_base[3], d.match_k, strand, with_qual ? 1 : 0, T.as<double> ());
    QG_TRY (qg_check_launch (ctx, "qg_overlap_pair_table_kernel"));
  }
  if (with_qual) d.built[strand] = true; else d.d_xonly[strand].cap = 1;   // marker only; d_xonly holds no memory
  return QG_OK;
}

template<int DUMMY>
static int qg_launch_overlap_fill (qg_ctx* ctx, const qg_dp_plan& plan, qg_ofill_args a, const qg_segment* d_segs_launch_order) {
  QG_TRY (qg_fork (ctx));
  int kcls = 0;
  for (const auto& L : plan.launches) {
    cudaStream_t st; QG_TRY (qg_side (ctx, kcls++, &st));
    a.segs = d_segs_launch_order + L.begin;
    const bool multi = L.nw > 1;
    const unsigned block = 32u * L.nw;
#define QG_CASE(RR) case RR: \
      if (multi) { auto kfn = qg_overlap_fill_kernel<8, true>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } \
      else { auto kfn = qg_overlap_fill_kernel<RR, false>; QG_LAUNCH (kfn, L.count, block, 0, st, a); } break;
    switch (L.R) {
      QG_CASE (2) QG_CASE (3) QG_CASE (4) QG_CASE (5) QG_CASE (6) QG_CASE (7) QG_CASE (8)
      default: QG_FAIL (ctx, QG_ERR_INVALID, "internal: unsupported R=%d", L.R);
    }
#undef QG_CASE
    QG_TRY (qg_check_launch (ctx, "qg_overlap_fill_kernel"));
  }
  QG_TRY (qg_join (ctx, kcls));
  return QG_OK;
}

extern "C" int qg_overlap_viterbi (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_pairs, const uint32_t* xi, const uint32_t* yi,
                                   const uint8_t* y_complemented, const uint8_t* want_path,
                                   double* score, uint32_t* coords4, uint8_t** path_out, uint64_t* path_offsets) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !xi || !yi || !y_complemented || !score) return QG_ERR_INVALID;
  if (!ctx->omodel.set) QG_FAIL (ctx, QG_ERR_STATE, "no overlap model: call qg_set_overlap_model first");
  const qg_seqset& Y = ctx->seqs[QG_READS];
  if (!Y.n) QG_FAIL (ctx, QG_ERR_STATE, "read set not uploaded");
  const bool paths = path_out && path_offsets && coords4;
  if (path_out) *path_out = nullptr;
  const bool with_qual = Y.has_qual;
  qg_overlap_dev& om = ctx->omodel;
  // emission: table look-up (K <= 2, or no qualities: 16^K entries, each a 94^2-term sum) or evaluated per cell
  const bool fly = with_qual && (om.match_k > 2 || qg_env_size ("QG_OVERLAP_FLY", 0) != 0);
  if (!with_qual && om.match_k > 4)
    QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "overlap of reads without qualities with K=%d: the quality-marginal emission table (16^K entries of 94^2 x 4 log-sum-exp terms) is built for K <= 4", om.match_k);
  bool need[2] = {false, false};
  for (size_t p = 0; p < n_pairs; ++p) need[y_complemented[p] ? 1 : 0] = true;
  if (!fly) for (int s = 0; s < 2; ++s) if (need[s]) QG_TRY (qg_overlap_ensure_table (ctx, s, with_qual));

  qg_env_result er;
  QG_TRY (qg_envelope_stage (ctx, cfg, 24, QG_READS, n_pairs, xi, yi, er));
  size_t freeb = 0, totb = 0;
  QG_CUDA (ctx, cudaMemGetInfo (&freeb, &totb));
  const uint64_t budget = (uint64_t) qg_env_size ("QG_TRACE_BUDGET_MB", std::min<size_t> (freeb / 2, (size_t) 32 << 30) >> 20) << 20;
  std::vector<uint8_t> all_paths;
  std::vector<uint64_t> offs (n_pairs + 1, 0);
  uint64_t path_total = 0;
  size_t p0 = 0;
  while (p0 < n_pairs) {
    size_t p1 = p0; uint64_t words = 0;
    while (p1 < n_pairs) {
      uint64_t w = 0;
      for (uint32_t r = er.run_begin[p1]; r < er.run_begin[p1 + 1]; ++r) {
        int R, nw; const uint32_t width = (uint32_t) (er.runs[r].y - er.runs[r].x + 1);
        if (qg_pick_R (width, &R, &nw) != QG_OK) QG_FAIL (ctx, QG_ERR_UNSUPPORTED, "pair %zu: a run of %u consecutive diagonals exceeds the %d this build fills with one CTA", p1, width, 256 * QG_MAX_NW);
        w += ((uint64_t) Y.len (yi[p1]) + 32ull * nw + 1) * 32ull * nw;
      }
      if (p1 > p0 && (words + w) * 8 > budget) break;
      words += w; ++p1;
    }
    const size_t np = p1 - p0;
    qg_dp_plan plan;
    QG_TRY (qg_build_plan (ctx, er, p0, p1, xi, yi, QG_READS, 3, plan));
    for (uint32_t s : plan.order) plan.segs_sorted.push_back (plan.segs[s]);
    std::vector<qg_opair> op (np);
    uint64_t xa_tot = 0, yb_tot = 0, gx_tot = 0, gy_tot = 0, scratch_bytes = 0;
    for (size_t p = 0; p < np; ++p) {
      qg_opair& o = op[p];
      o.xseq = xi[p0 + p]; o.yseq = yi[p0 + p]; o.xlen = Y.len (o.xseq); o.ylen = Y.len (o.yseq);
      o.xoff = Y.off[o.xseq]; o.yoff = Y.off[o.yseq];
      o.xa_off = xa_tot; xa_tot += o.xlen; o.yb_off = yb_tot; yb_tot += o.ylen;
      o.gx_off = gx_tot; gx_tot += o.xlen + 1; o.gy_off = gy_tot; gy_tot += o.ylen + 1;
      o.y_comp = y_complemented[p0 + p] ? 1 : 0; o.pad_ = 0;
      plan.pairs[p].want_path = paths && (!want_path || want_path[p0 + p]) ? 1 : 0;
      plan.pairs[p].path_off = scratch_bytes;
      if (plan.pairs[p].want_path) scratch_bytes += plan.pairs[p].path_cap;
    }
    qg_dbuf &dOP = ctx->scratch[SC_RPJOBS], &dXA = ctx->scratch[SC_RP], &dYB = ctx->scratch[SC_STORE], &dGX = ctx->scratch[SC_ITEMRUNS], &dGY = ctx->scratch[SC_ITEMNRUNS];
    qg_dbuf &dINS = ctx->scratch[SC_OUT1], &dLR = ctx->scratch[SC_ENDVALS], &dLC = ctx->scratch[SC_ROWACC];
    qg_dbuf &dEA = ctx->scratch[SC_RQ], &dEB = ctx->scratch[SC_RS];
    {
      qg_timer tm (ctx, &ctx->stats.ms_prep);
      QG_TRY (qg_upload (ctx, dOP, op.data (), sizeof (qg_opair) * np));
      QG_TRY (qg_reserve (ctx, dXA, sizeof (uint32_t) * (xa_tot + 1)));
      QG_TRY (qg_reserve (ctx, dYB, sizeof (uint32_t) * (yb_tot + 1)));
      QG_TRY (qg_reserve (ctx, dGX, sizeof (uint32_t) * (gx_tot + 1)));
      QG_TRY (qg_reserve (ctx, dGY, sizeof (uint32_t) * (gy_tot + 1)));
      QG_TRY (qg_reserve (ctx, dINS, sizeof (double) * 2 * (np + 1)));
      if (fly) { QG_TRY (qg_reserve (ctx, dEA, sizeof (double) * 5 * (xa_tot + 1))); QG_TRY (qg_reserve (ctx, dEB, sizeof (double) * 5 * (yb_tot + 1))); }
      QG_TRY (qg_reserve (ctx, dLR, sizeof (double) * (plan.aux_slots + 1)));
      QG_TRY (qg_reserve (ctx, dLC, sizeof (double) * (plan.acc_rows + 1)));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_SEGS], plan.segs_sorted.data (), sizeof (qg_segment) * plan.segs_sorted.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC0], plan.segs.data (), sizeof (qg_segment) * plan.segs.size ()));
      QG_TRY (qg_upload (ctx, ctx->scratch[SC_PAIRDP], plan.pairs.data (), sizeof (qg_pair_dp) * np));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_TRACE], sizeof (unsigned long long) * (plan.trace_words + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT0], sizeof (double) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT2], sizeof (uint32_t) * 4 * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_OUT3], sizeof (uint32_t) * (np + 1)));
      QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHSCR], scratch_bytes + 16));
      QG_LAUNCH (qg_overlap_prep_kernel, (unsigned) np, 256, 0, ctx->stream,
                 dOP.as<qg_opair> (), Y.d_tok.as<uint8_t> (), with_qual ? Y.d_qual.as<uint8_t> () : (const uint8_t*) nullptr,
                 om.match_k, om.gap_k, with_qual ? 1 : 0, dXA.as<uint32_t> (), dYB.as<uint32_t> (), dGX.as<uint32_t> (), dGY.as<uint32_t> (),
                 om.d_insert.as<double> (), dINS.as<double> (),
                 om.d_match.as<double> (), om.log_ref_base[0], om.log_ref_base[1], om.log_ref_base[2], om.log_ref_base[3],
                 fly ? dEA.as<double> () : (double*) nullptr, fly ? dEB.as<double> () : (double*) nullptr);
      QG_TRY (qg_check_launch (ctx, "qg_overlap_prep_kernel"));
      QG_LAUNCH (qg_fill_neginf_kernel, (unsigned) ((plan.acc_rows + 256) / 256), 256, 0, ctx->stream, dLC.as<double> (), plan.acc_rows + 1);
      QG_TRY (qg_check_launch (ctx, "qg_fill_neginf_kernel"));
    }
    ctx->stats.trace_bytes += plan.trace_words * 8;
    ctx->stats.n_segments += plan.segs.size ();
    ctx->stats.cell_updates += qg_plan_cells (er, p0, p1);
    {
      qg_timer tm (ctx, &ctx->stats.ms_overlap);
      qg_ofill_args a;
      memset (&a, 0, sizeof (a));
      a.pairs = dOP.as<qg_opair> ();
      a.xa = dXA.as<uint32_t> (); a.yb = dYB.as<uint32_t> (); a.gx = dGX.as<uint32_t> (); a.gy = dGY.as<uint32_t> ();
      a.table = with_qual ? om.d_pair[0].as<double> () : om.d_none[0].as<double> ();
      a.table1 = with_qual ? om.d_pair[1].as<double> () : om.d_none[1].as<double> ();
      if (fly) { a.table = a.table1 = nullptr; a.ea = dEA.as<double> (); a.eb = dEB.as<double> (); }
      a.m2m = om.d_m2m.as<double> (); a.m2i = om.d_m2i.as<double> (); a.m2d = om.d_m2d.as<double> ();
      a.lse = ctx->d_lse.as<double> ();
      // accessor swaps of qoverlap.h:46-51
      a.effI2M = om.i2i; a.effI2I = om.i2m; a.effI2D = om.i2d; a.effD2M = om.d2i; a.effD2I = om.d2m; a.effD2D = om.d2d;
      a.nG = (int) om.nG;
      a.trace = (unsigned long long*) ctx->scratch[SC_TRACE].p;
      a.lastrow = dLR.as<double> (); a.lastcol = dLC.as<double> ();
      QG_TRY (qg_launch_overlap_fill<0> (ctx, plan, a, ctx->scratch[SC_SEGS].as<qg_segment> ()));
    }
    std::vector<uint32_t> plen (np), co (4 * np);
    uint32_t flag = 0;
    {
      qg_timer tm (ctx, &ctx->stats.ms_traceback);
      QG_CUDA (ctx, cudaMemsetAsync (ctx->scratch[SC_FLAGS].p, 0, 64, ctx->stream));
      QG_LAUNCH (qg_overlap_traceback_kernel, (unsigned) ((np + 63) / 64), 64, 0, ctx->stream,
                 ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_MISC0].as<qg_segment> (),
                 dLR.as<double> (), dLC.as<double> (), dINS.as<double> (), (const unsigned long long*) ctx->scratch[SC_TRACE].p,
                 ctx->scratch[SC_OUT0].as<double> (), ctx->scratch[SC_OUT2].as<uint32_t> (), ctx->scratch[SC_PATHSCR].as<uint8_t> (),
                 ctx->scratch[SC_OUT3].as<uint32_t> (), (uint32_t*) ctx->scratch[SC_FLAGS].p);
      QG_TRY (qg_check_launch (ctx, "qg_overlap_traceback_kernel"));
    }
    {
      qg_timer tm (ctx, &ctx->stats.ms_d2h);
      QG_TRY (qg_fetch (ctx, score + p0, ctx->scratch[SC_OUT0].p, sizeof (double) * np));
      QG_TRY (qg_fetch (ctx, co.data (), ctx->scratch[SC_OUT2].p, sizeof (uint32_t) * 4 * np));
      QG_TRY (qg_fetch (ctx, plen.data (), ctx->scratch[SC_OUT3].p, sizeof (uint32_t) * np));
      QG_TRY (qg_fetch (ctx, &flag, ctx->scratch[SC_FLAGS].p, sizeof (uint32_t)));
      QG_TRY (qg_fetch_wait (ctx));
    }
    if (flag) QG_FAIL (ctx, QG_ERR_CUDA, "internal: overlap traceback left the envelope (code %u)", flag);
    if (coords4) memcpy (coords4 + 4 * p0, co.data (), sizeof (uint32_t) * 4 * np);
    if (paths) {
      std::vector<uint64_t> goff (np + 1, 0);
      for (size_t p = 0; p < np; ++p) goff[p + 1] = goff[p] + plen[p];
      for (size_t p = 0; p < np; ++p) offs[p0 + p] = path_total + goff[p];
      if (goff[np]) {
        qg_timer tm (ctx, &ctx->stats.ms_d2h);
        QG_TRY (qg_upload (ctx, ctx->scratch[SC_MISC1], goff.data (), sizeof (uint64_t) * (np + 1)));
        QG_TRY (qg_reserve (ctx, ctx->scratch[SC_PATHOUT], goff[np] + 16));
        QG_LAUNCH (qg_path_gather_kernel, (unsigned) np, 128, 0, ctx->stream,
                   ctx->scratch[SC_PAIRDP].as<qg_pair_dp> (), (uint32_t) np, ctx->scratch[SC_PATHSCR].as<uint8_t> (),
                   ctx->scratch[SC_OUT3].as<uint32_t> (), ctx->scratch[SC_MISC1].as<uint64_t> (), ctx->scratch[SC_PATHOUT].as<uint8_t> ());
        QG_TRY (qg_check_launch (ctx, "qg_path_gather_kernel"));
        all_paths.resize (path_total + goff[np]);
        QG_TRY (qg_download (ctx, all_paths.data () + path_total, ctx->scratch[SC_PATHOUT].p, goff[np]));
      }
      path_total += goff[np];
    }
    p0 = p1;
  }
  if (paths) {
    offs[n_pairs] = path_total;
    memcpy (path_offsets, offs.data (), sizeof (uint64_t) * (n_pairs + 1));
    uint8_t* buf = (uint8_t*) malloc (path_total + 1);
    if (!buf) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
    if (path_total) memcpy (buf, all_paths.data (), path_total);
    *path_out = buf;
  }
  return QG_OK;
}

// host string assembly, with the reference's squashing of adjacent insertions and deletions (qoverlap.cpp:231-267):
// a gap run with deleted x bases X and inserted y bases Y becomes [X_t over Y_t for t < min] [rest of X over gaps] [gaps over rest of Y]
extern "C" int qg_overlap_rows (const uint8_t* x_tok, const uint8_t* y_tok, const uint32_t* coords4,
                                const uint8_t* path, uint64_t path_len, char** xrow, char** yrow) {
  if (!x_tok || !y_tok || !coords4 || (!path && path_len) || !xrow || !yrow) return QG_ERR_INVALID;
  static const char alph[] = "ACGT";
  std::string xr, yr;
  uint64_t i = coords4[0] ? coords4[0] - 1 : 0, j = coords4[2] ? coords4[2] - 1 : 0;
  uint64_t t = 0;
  while (t < path_len) {
    if (path[t] == QG_OP_MATCH) { xr += alph[x_tok[i++] & 3]; yr += alph[y_tok[j++] & 3]; ++t; continue; }
    uint64_t nd = 0, ni = 0, e = t;
    while (e < path_len && path[e] != QG_OP_MATCH) { if (path[e] == QG_OP_DELETE) ++nd; else ++ni; ++e; }
    const uint64_t sh = nd < ni ? nd : ni;
    for (uint64_t s = 0; s < sh; ++s) { xr += alph[x_tok[i + s] & 3]; yr += alph[y_tok[j + s] & 3]; }
    for (uint64_t s = sh; s < nd; ++s) { xr += alph[x_tok[i + s] & 3]; yr += '-'; }
    for (uint64_t s = sh; s < ni; ++s) { xr += '-'; yr += alph[y_tok[j + s] & 3]; }
    i += nd; j += ni; t = e;
  }
  *xrow = (char*) malloc (xr.size () + 1); *yrow = (char*) malloc (yr.size () + 1);
  if (!*xrow || !*yrow) return QG_ERR_INVALID;
  memcpy (*xrow, xr.c_str (), xr.size () + 1); memcpy (*yrow, yr.c_str (), yr.size () + 1);
  return QG_OK;
}

// ---- seam B: QuaffOverlapAligner::align (qoverlap.cpp:312-334); pair order of QuaffOverlapScheduler (qoverlap.cpp:473-478, 528-547)
extern "C" int qg_overlap_reads (qg_ctx* ctx, const qg_dpconfig* cfg, size_t n_originals, const double* null_loglike,
                                 size_t* n_pairs_out, uint32_t** xi_out, uint32_t** yi_out,
                                 double** score_out, uint32_t** coords4_out, uint8_t** path_out, uint64_t** path_offsets_out) {
  if (ctx) cudaSetDevice (ctx->device);                    // the caller may be a host thread that never selected this context's GPU
  if (!ctx || !cfg || !null_loglike || !n_pairs_out || !xi_out || !yi_out || !score_out || !coords4_out || !path_out || !path_offsets_out) return QG_ERR_INVALID;
  const size_t N = ctx->seqs[QG_READS].n;
  if (n_originals > N) QG_FAIL (ctx, QG_ERR_INVALID, "n_originals exceeds the read set");
  std::vector<uint32_t> xi, yi; std::vector<uint8_t> yc;
  for (size_t nx = 0; nx + 1 < n_originals; ++nx)
    for (size_t ny = nx + 1; ny < N; ++ny) { xi.push_back ((uint32_t) nx); yi.push_back ((uint32_t) ny); yc.push_back (ny >= n_originals ? 1 : 0); }
  const size_t np = xi.size ();
  double* sc = (double*) malloc (sizeof (double) * (np + 1));
  uint32_t* co = (uint32_t*) malloc (sizeof (uint32_t) * 4 * (np + 1));
  uint64_t* po = (uint64_t*) malloc (sizeof (uint64_t) * (np + 2));
  uint32_t* xo = (uint32_t*) malloc (sizeof (uint32_t) * (np + 1));
  uint32_t* yo = (uint32_t*) malloc (sizeof (uint32_t) * (np + 1));
  if (!sc || !co || !po || !xo || !yo) QG_FAIL (ctx, QG_ERR_INVALID, "out of host memory");
  uint8_t* paths = nullptr;
  po[0] = 0;
  if (np) {
    const int rc = qg_overlap_viterbi (ctx, cfg, np, xi.data (), yi.data (), yc.data (), nullptr, sc, co, &paths, po);
    if (rc != QG_OK) { free (sc); free (co); free (po); free (xo); free (yo); return rc; }
    // scoreAdjustedAlignment (qoverlap.cpp:292-302): minus the null log-likelihood of x and of y as stored
    for (size_t p = 0; p < np; ++p) { if (sc[p] > -INFINITY) { sc[p] -= null_loglike[xi[p]]; sc[p] -= null_loglike[yi[p]]; } xo[p] = xi[p]; yo[p] = yi[p]; }
  } else paths = (uint8_t*) malloc (1);
  *n_pairs_out = np; *xi_out = xo; *yi_out = yo; *score_out = sc; *coords4_out = co; *path_out = paths; *path_offsets_out = po;
  return QG_OK;
}

#include "qg_pool.cuh"
