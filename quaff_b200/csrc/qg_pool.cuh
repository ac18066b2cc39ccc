// Several contexts on one or more devices behind ONE handle: the multi-GPU form of the three seams.
// Reads shard over the contexts (the reference's tasks are independent per read: QuaffAlignmentTask, qmodel.cpp:2575-2622;
// QuaffCountingTask, qmodel.cpp:1879-1960); every context holds the reference set, its tile index and the model.
//   align : the read set is cut into chunks; each context's host thread takes the next chunk when it is free (dynamic
//           balance, as the reference's thread pool does, qmodel.cpp:2650-2668), and hands the finished chunk to the
//           caller's callback ON THAT THREAD, so that output formatting runs in parallel with the other contexts' GPU work
//   E-step: contiguous read ranges, one per context; the partial QuaffParamCounts are summed in context order on the host
//           (a single process owns all devices here: the sum of <= 24 509 doubles needs no collective; with one process
//           per GPU the caller all-reduces instead, quaff_b200/dist.py)
#ifndef QG_POOL_CUH
#define QG_POOL_CUH
#include <thread>
#include <atomic>
#include <mutex>

struct qg_pool {
  std::vector<qg_ctx*> ctx;
  std::vector<int> device;
  std::string err;
  int K = 1, G = 0;
};

static thread_local std::string qg_pool_create_error;

extern "C" int qg_device_count (void) {
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
  return n;
}

// The first touch of a device (driver initialisation + primary context) costs seconds on a multi-GPU box.  A host that knows
// early which devices it will use (the CLI does, from its flags) calls this from a side thread while it still parses its
// inputs; the later qg_create / qg_pool_create find the devices ready.
extern "C" int qg_init_devices (const int* devices, int n_devices) {
  if (!devices || n_devices < 1) return QG_ERR_INVALID;
  std::vector<int> rc (n_devices, QG_OK);
  std::vector<std::thread> th;
  for (int d = 0; d < n_devices; ++d)                        // one thread per device: the primary contexts are created concurrently
    th.emplace_back ([&, d] { if (cudaSetDevice (devices[d]) != cudaSuccess || cudaFree (0) != cudaSuccess) { cudaGetLastError (); rc[d] = QG_ERR_NO_DEVICE; } });
  for (auto& t : th) t.join ();
  for (int d = 0; d < n_devices; ++d) if (rc[d] != QG_OK) return rc[d];
  return QG_OK;
}

extern "C" int qg_pool_create (qg_pool** out, const int* devices, int n_devices, int contexts_per_device) {
  if (!out || n_devices < 1 || contexts_per_device < 1 || !devices) return QG_ERR_INVALID;
  *out = nullptr;
  qg_pool* p = new qg_pool;
  // context order: device-major inside every round, so that the first n_devices workers sit on different devices.
  // Created concurrently: the first touch of a device (CUDA context creation) dominates and is independent per device.
  const int n = n_devices * contexts_per_device;
  p->ctx.assign (n, nullptr); p->device.resize (n);
  std::vector<int> rcs (n, QG_OK);
  std::vector<std::string> errs (n);
  std::vector<std::thread> th;
  for (int w = 0; w < n; ++w) {
    p->device[w] = devices[w % n_devices];
    th.emplace_back ([&, w] { rcs[w] = qg_create (&p->ctx[w], p->device[w]); if (rcs[w] != QG_OK) errs[w] = qg_last_error (nullptr); });
  }
  for (auto& t : th) t.join ();
  for (int w = 0; w < n; ++w)
    if (rcs[w] != QG_OK) {
      qg_pool_create_error = errs[w];
      for (qg_ctx* y : p->ctx) if (y) qg_destroy (y);
      delete p;
      return rcs[w];
    }
  *out = p;
  return QG_OK;
}

extern "C" void qg_pool_destroy (qg_pool* p) {
  if (!p) return;
  for (qg_ctx* x : p->ctx) qg_destroy (x);
  delete p;
}

extern "C" const char* qg_pool_last_error (const qg_pool* p) { return p ? p->err.c_str () : qg_pool_create_error.c_str (); }
extern "C" int qg_pool_size (const qg_pool* p) { return p ? (int) p->ctx.size () : 0; }
extern "C" qg_ctx* qg_pool_context (qg_pool* p, int i) { return (p && i >= 0 && i < (int) p->ctx.size ()) ? p->ctx[i] : nullptr; }

// run fn (worker) on one host thread per context; first error wins
template<class F>
static int qg_pool_run (qg_pool* p, F fn) {
  const int n = (int) p->ctx.size ();
  std::vector<int> rc (n, QG_OK);
  std::vector<std::thread> th;
  for (int w = 1; w < n; ++w) th.emplace_back ([&, w] { rc[w] = fn (w); });
  rc[0] = fn (0);
  for (auto& t : th) t.join ();
  for (int w = 0; w < n; ++w)
    if (rc[w] != QG_OK) { p->err = std::string ("context ") + std::to_string (w) + " (device " + std::to_string (p->device[w]) + "): " + qg_last_error (p->ctx[w]); return rc[w]; }
  return QG_OK;
}

extern "C" int qg_pool_set_refs (qg_pool* p, size_t n, const uint8_t* tok, const uint64_t* offsets) {
  if (!p) return QG_ERR_INVALID;
  return qg_pool_run (p, [&] (int w) { return qg_set_seqs (p->ctx[w], QG_REFS, n, tok, nullptr, offsets); });
}

extern "C" int qg_pool_set_align_model (qg_pool* p, const qg_align_model* m) {
  if (!p || !m) return QG_ERR_INVALID;
  p->K = m->match_k; p->G = m->gap_k;
  return qg_pool_run (p, [&] (int w) { return qg_set_align_model (p->ctx[w], m); });
}

extern "C" int qg_pool_set_option (qg_pool* p, int option, int64_t value) {
  if (!p) return QG_ERR_INVALID;
  for (qg_ctx* x : p->ctx) { const int rc = qg_set_option (x, option, value); if (rc != QG_OK) { p->err = qg_last_error (x); return rc; } }
  return QG_OK;
}

extern "C" int qg_pool_align_reads (qg_pool* p, const qg_dpconfig* cfg, size_t n_reads, const uint8_t* tok, const uint8_t* qual,
                                    const uint64_t* offsets, const double* null_loglike, size_t chunk_reads,
                                    qg_chunk_fn on_chunk, void* user) {
  if (!p || !cfg || !offsets || !null_loglike || !on_chunk || (n_reads && !tok)) return QG_ERR_INVALID;
  if (chunk_reads == 0) chunk_reads = 1536;
  const size_t n_chunks = (n_reads + chunk_reads - 1) / chunk_reads;
  std::atomic<size_t> next (0);
  std::atomic<int> failed (0);
  return qg_pool_run (p, [&] (int w) -> int {
    qg_ctx* ctx = p->ctx[w];
    std::vector<uint64_t> off, poff;
    std::vector<uint32_t> best, xs, xe;
    std::vector<double> score;
    for (;;) {
      const size_t c = next.fetch_add (1);
      if (c >= n_chunks || failed.load ()) return QG_OK;
      const size_t r0 = c * chunk_reads, r1 = std::min (n_reads, r0 + chunk_reads), nr = r1 - r0;
      off.resize (nr + 1);
      for (size_t r = 0; r <= nr; ++r) off[r] = offsets[r0 + r] - offsets[r0];
      int rc = qg_set_seqs (ctx, QG_READS, nr, tok + offsets[r0], qual ? qual + offsets[r0] : nullptr, off.data ());
      if (rc != QG_OK) { failed = 1; return rc; }
      best.resize (nr); xs.resize (nr); xe.resize (nr); score.resize (nr); poff.resize (nr + 1);
      uint8_t* path = nullptr;
      rc = qg_align_reads (ctx, cfg, null_loglike + r0, best.data (), score.data (), xs.data (), xe.data (), &path, poff.data ());
      if (rc != QG_OK) { failed = 1; return rc; }
      on_chunk (user, w, r0, nr, best.data (), score.data (), xs.data (), xe.data (), path, poff.data ());
      qg_free (path);
    }
  });
}

extern "C" int qg_pool_estep (qg_pool* p, const qg_dpconfig* cfg, int use_null, size_t n_refs, size_t n_reads, const uint8_t* tok, const uint8_t* qual,
                              const uint64_t* offsets, const double* null_loglike, uint32_t* sort_order, uint32_t* sort_len,
                              double* y_loglike, double* param_counts, double* loglike_sum) {
  if (!p || !cfg || !offsets || !sort_order || !sort_len || !y_loglike || !param_counts || !loglike_sum || (n_reads && (!tok || !qual))) return QG_ERR_INVALID;
  const int n = (int) p->ctx.size ();
  const size_t nc = qg_counts_size (p->K, p->G);
  std::vector<std::vector<double>> part (n, std::vector<double> (nc, 0.0));
  std::vector<double> ll (n, 0.0);
  // contiguous ranges balanced by bases (the DP work of a read is proportional to its length)
  std::vector<size_t> cut (n + 1, n_reads);
  cut[0] = 0;
  { const uint64_t total = offsets[n_reads] - offsets[0]; size_t r = 0;
    for (int w = 1; w < n; ++w) { const uint64_t want = offsets[0] + total * (uint64_t) w / (uint64_t) n; while (r < n_reads && offsets[r] < want) ++r; cut[w] = r; } }
  const int rc = qg_pool_run (p, [&] (int w) -> int {
    const size_t r0 = cut[w], r1 = cut[w + 1], nr = r1 - r0;
    if (!nr) return QG_OK;
    qg_ctx* ctx = p->ctx[w];
    std::vector<uint64_t> off (nr + 1);
    for (size_t r = 0; r <= nr; ++r) off[r] = offsets[r0 + r] - offsets[r0];
    QG_TRY (qg_set_seqs (ctx, QG_READS, nr, tok + offsets[r0], qual + offsets[r0], off.data ()));
    return qg_estep (ctx, cfg, use_null, null_loglike ? null_loglike + r0 : nullptr, sort_order + r0 * n_refs, sort_len + r0, y_loglike + r0, part[w].data (), &ll[w]);
  });
  if (rc != QG_OK) return rc;
  for (size_t k = 0; k < nc; ++k) { double s = 0; for (int w = 0; w < n; ++w) s += part[w][k]; param_counts[k] = s; }
  double s = 0; for (int w = 0; w < n; ++w) s += ll[w];
  *loglike_sum = s;
  return QG_OK;
}

// seam B over the pool ("overlap: partition the pair list; all reads replicated; no collective"): every context holds
// the whole read set and the overlap model; the scheduler's pair list (qoverlap.cpp:528-547) is cut into contiguous ranges
// balanced by cells (x length x y length); outputs are concatenated in pair order, so they equal qg_overlap_reads'.
extern "C" int qg_pool_overlap_reads (qg_pool* p, const qg_dpconfig* cfg, const qg_overlap_model* model,
                                      size_t n_seqs, const uint8_t* tok, const uint8_t* qual, const uint64_t* offsets,
                                      size_t n_originals, const double* null_loglike,
                                      size_t* n_pairs_out, uint32_t** xi_out, uint32_t** yi_out,
                                      double** score_out, uint32_t** coords4_out, uint8_t** path_out, uint64_t** path_offsets_out) {
  if (!p || !cfg || !model || !offsets || !null_loglike || !n_pairs_out || !xi_out || !yi_out || !score_out || !coords4_out || !path_out || !path_offsets_out
      || (n_seqs && !tok)) return QG_ERR_INVALID;
  if (n_originals > n_seqs) { p->err = "n_originals exceeds the read set"; return QG_ERR_INVALID; }
  std::vector<uint32_t> xi, yi; std::vector<uint8_t> yc;
  std::vector<double> work;
  for (size_t nx = 0; nx + 1 < n_originals; ++nx)
    for (size_t ny = nx + 1; ny < n_seqs; ++ny) {
      xi.push_back ((uint32_t) nx); yi.push_back ((uint32_t) ny); yc.push_back (ny >= n_originals ? 1 : 0);
      work.push_back ((double) (offsets[nx + 1] - offsets[nx]) * (double) (offsets[ny + 1] - offsets[ny]));
    }
  const size_t np = xi.size ();
  const int n = (int) p->ctx.size ();
  std::vector<size_t> cut (n + 1, np);
  cut[0] = 0;
  { double total = 0; for (double w : work) total += w;
    double acc = 0; size_t q = 0;
    for (int w = 1; w < n; ++w) { const double want = total * w / n; while (q < np && acc < want) acc += work[q++]; cut[w] = q; } }
  std::vector<std::vector<double> > sc (n); std::vector<std::vector<uint32_t> > co (n); std::vector<std::vector<uint64_t> > po (n);
  std::vector<uint8_t*> pa (n, nullptr);
  const int rc = qg_pool_run (p, [&] (int w) -> int {
    const size_t q0 = cut[w], q1 = cut[w + 1], nq = q1 - q0;
    if (!nq) return QG_OK;
    qg_ctx* ctx = p->ctx[w];
    QG_TRY (qg_set_seqs (ctx, QG_READS, n_seqs, tok, qual, offsets));
    QG_TRY (qg_set_overlap_model (ctx, model));
    sc[w].resize (nq); co[w].resize (4 * nq); po[w].resize (nq + 1);
    return qg_overlap_viterbi (ctx, cfg, nq, xi.data () + q0, yi.data () + q0, yc.data () + q0, nullptr, sc[w].data (), co[w].data (), &pa[w], po[w].data ());
  });
  if (rc != QG_OK) { for (uint8_t* b : pa) free (b); return rc; }
  uint64_t path_total = 0;
  for (int w = 0; w < n; ++w) if (!po[w].empty ()) path_total += po[w].back ();
  double* osc = (double*) malloc (sizeof (double) * (np + 1));
  uint32_t* oco = (uint32_t*) malloc (sizeof (uint32_t) * 4 * (np + 1));
  uint64_t* opo = (uint64_t*) malloc (sizeof (uint64_t) * (np + 2));
  uint32_t* oxi = (uint32_t*) malloc (sizeof (uint32_t) * (np + 1));
  uint32_t* oyi = (uint32_t*) malloc (sizeof (uint32_t) * (np + 1));
  uint8_t* opa = (uint8_t*) malloc (path_total + 1);
  if (!osc || !oco || !opo || !oxi || !oyi || !opa) {
    free (osc); free (oco); free (opo); free (oxi); free (oyi); free (opa); for (uint8_t* b : pa) free (b);
    p->err = "out of host memory"; return QG_ERR_INVALID;
  }
  uint64_t base = 0;
  opo[0] = 0;
  for (int w = 0; w < n; ++w) {
    const size_t q0 = cut[w], nq = cut[w + 1] - q0;
    for (size_t q = 0; q < nq; ++q) {
      const size_t g = q0 + q;
      double v = sc[w][q];
      if (v > -INFINITY) { v -= null_loglike[xi[g]]; v -= null_loglike[yi[g]]; }     // scoreAdjustedAlignment, qoverlap.cpp:292-302
      osc[g] = v; oxi[g] = xi[g]; oyi[g] = yi[g];
      for (int t = 0; t < 4; ++t) oco[4 * g + t] = co[w][4 * q + t];
      opo[g + 1] = base + po[w][q + 1];
    }
    if (nq) { if (po[w].back ()) memcpy (opa + base, pa[w], po[w].back ()); base += po[w].back (); }
    free (pa[w]);
  }
  *n_pairs_out = np; *xi_out = oxi; *yi_out = oyi; *score_out = osc; *coords4_out = oco; *path_out = opa; *path_offsets_out = opo;
  return QG_OK;
}
#endif
