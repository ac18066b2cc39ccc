// Seams A/B/C of the reference (QuaffAligner::align qmodel.cpp:2624, QuaffOverlapAligner::align qoverlap.cpp:312,
// QuaffTrainer::getCounts qmodel.cpp:2005) routed to libquaffgpu.  The reference keeps everything else: flag parsing,
// sequence loading, params / null model, the Alignment type and every output writer.
#include <fstream>
#include <algorithm>
#include <cmath>
#include "qmodel.h"
#include "qoverlap.h"
#include "quaffgpu.h"

int quaffGpuDevice = -1;

bool quaffGpuParseArg (std::deque<std::string>& argvec) {
  if (argvec.size() && argvec[0] == "-gpu") {
    argvec.pop_front();
    quaffGpuDevice = 0;
    if (argvec.size() && !argvec[0].empty() && isdigit (argvec[0][0]) && argvec[0].size() <= 2) {
      quaffGpuDevice = atoi (argvec[0].c_str());
      argvec.pop_front();
    }
    return true;
  }
  return false;
}

namespace {
struct Flat { std::vector<uint8_t> tok, qual; std::vector<uint64_t> off; bool quals; };

Flat flatten (const vguard<FastSeq>& seqs, bool wantQual) {
  Flat f;
  f.quals = wantQual && !seqs.empty();
  for (const auto& s : seqs) f.quals = f.quals && s.hasQual();
  if (wantQual && !f.quals)
    for (const auto& s : seqs)
      Require (!s.hasQual(), "-gpu: either all reads or no reads must carry quality scores (%s)", s.name.c_str());
  f.off.push_back (0);
  for (const auto& s : seqs) {
    for (auto t : s.tokens (dnaAlphabet)) f.tok.push_back ((uint8_t) t);
    if (f.quals) for (auto q : s.qualScores()) f.qual.push_back ((uint8_t) q);
    f.off.push_back (f.tok.size());
  }
  return f;
}

struct Gpu {
  qg_ctx* ctx;
  Gpu () : ctx (NULL) { Require (qg_create (&ctx, quaffGpuDevice) == QG_OK, "-gpu: %s", qg_last_error (NULL)); }
  ~Gpu () { qg_destroy (ctx); }
  void ok (int rc) const { Require (rc == QG_OK, "-gpu: %s", qg_last_error (ctx)); }
};

qg_dpconfig gpuConfig (const QuaffDPConfig& c) {
  qg_dpconfig g;
  g.sparse = c.sparse; g.kmer_len = c.kmerLen; g.kmer_threshold = c.kmerThreshold; g.band_size = c.bandSize;
  g.local = c.local; g.max_size = (uint64_t) c.effectiveMaxSize();
  return g;
}

struct ScoreTables { std::vector<double> match, insert; int K, G; };

ScoreTables tables (const QuaffScores& qs) {
  ScoreTables t;
  t.K = qs.matchContext.kmerLen; t.G = qs.indelContext.kmerLen;
  const Kmer nK = qs.matchContext.numKmers;
  t.match.resize (4 * nK * QG_NQ1); t.insert.resize (4 * QG_NQ1);
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i) {
    for (Kmer j = 0; j < nK; ++j) {
      for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q) t.match[(i * nK + j) * QG_NQ1 + q] = qs.match[i][j].logSymQualProb[q];
      t.match[(i * nK + j) * QG_NQ1 + QG_NQUAL] = qs.match[i][j].logSymProb;
    }
    for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q) t.insert[i * QG_NQ1 + q] = qs.insert[i].logSymQualProb[q];
    t.insert[i * QG_NQ1 + QG_NQUAL] = qs.insert[i].logSymProb;
  }
  return t;
}

void setAlignModel (const Gpu& g, const QuaffParams& params) {
  const QuaffScores qs (params);
  const ScoreTables t = tables (qs);
  qg_align_model m;
  m.match_k = t.K; m.gap_k = t.G; m.match = t.match.data(); m.insert = t.insert.data();
  m.m2m = qs.m2m.data(); m.m2i = qs.m2i.data(); m.m2d = qs.m2d.data(); m.m2e = qs.m2e.data();
  m.d2d = qs.d2d; m.d2m = qs.d2m; m.i2i = qs.i2i; m.i2m = qs.i2m;
  g.ok (qg_set_align_model (g.ctx, &m));
}

// the second half of QuaffViterbiMatrix::alignment(): gapped rows, names and coordinates from the state path
Alignment alignmentFromPath (const FastSeq& x, const FastSeq& y, uint32_t xStart, uint32_t xEnd, const uint8_t* path, uint64_t n,
                             double score, bool local) {
  std::string xRow, yRow, yQual;
  xRow.reserve (n); yRow.reserve (n);
  size_t i = xStart - 1, j = 0;
  for (uint64_t t = 0; t < n; ++t) {
    switch (path[t]) {
    case QG_OP_MATCH:  xRow += x.seq[i++]; yRow += y.seq[j]; if (y.hasQual()) yQual += y.qual[j]; ++j; break;
    case QG_OP_INSERT: xRow += Alignment::gapChar; yRow += y.seq[j]; if (y.hasQual()) yQual += y.qual[j]; ++j; break;
    default:           xRow += x.seq[i++]; yRow += Alignment::gapChar; if (y.hasQual()) yQual += FastSeq::maxQualityChar; break;
    }
  }
  Alignment align (2);
  align.gappedSeq[0].name = "Ref";
  align.gappedSeq[0].comment = local ? "substr(" + x.name + "," + to_string (xStart) + ".." + to_string (xEnd) + ")" : x.name;
  align.gappedSeq[1].name = "Read";
  align.gappedSeq[1].comment = y.name;
  align.gappedSeq[0].seq = xRow;
  align.gappedSeq[1].seq = yRow;
  align.gappedSeq[1].qual = yQual;
  align.gappedSeq[0].source.name = x.name;
  align.gappedSeq[0].source.start = xStart;
  align.gappedSeq[0].source.end = xEnd;
  align.gappedSeq[1].source.name = y.name;
  align.gappedSeq[1].source.start = 1;
  align.gappedSeq[1].source.end = y.length();
  align.gappedSeq[0].source = align.gappedSeq[0].source.compose (x.source);
  align.gappedSeq[1].source = align.gappedSeq[1].source.compose (y.source);
  align.score = score;
  return align;
}
}  // namespace

// ---- seam A -------------------------------------------------------------------------------------------------------
void quaffGpuAlign (QuaffAligner& aligner, std::ostream& out, const vguard<FastSeq>& x, const vguard<FastSeq>& y,
                    const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config) {
  Gpu g;
  const Flat fx = flatten (x, false), fy = flatten (y, true);
  g.ok (qg_set_seqs (g.ctx, QG_REFS, x.size(), fx.tok.data(), NULL, fx.off.data()));
  g.ok (qg_set_seqs (g.ctx, QG_READS, y.size(), fy.tok.data(), fy.quals ? fy.qual.data() : NULL, fy.off.data()));
  setAlignModel (g, params);
  const qg_dpconfig gc = gpuConfig (config);
  std::vector<double> nullLL (y.size());
  for (size_t n = 0; n < y.size(); ++n) nullLL[n] = nullModel.logLikelihood (y[n]);
  aligner.writeAlignmentHeader (out, x, true);
  if (!aligner.printAllAlignments) {
    std::vector<uint32_t> best (y.size()), xs (y.size()), xe (y.size());
    std::vector<double> score (y.size());
    std::vector<uint64_t> off (y.size() + 1);
    uint8_t* path = NULL;
    g.ok (qg_align_reads (g.ctx, &gc, nullLL.data(), best.data(), score.data(), xs.data(), xe.data(), &path, off.data()));
    for (size_t n = 0; n < y.size(); ++n)
      if (best[n] != 0xFFFFFFFFu)
        aligner.writeAlignment (out, alignmentFromPath (x[best[n]], y[n], xs[n], xe[n], path + off[n], off[n+1] - off[n], score[n], config.local));
    qg_free (path);
  } else {
    // -printall: every reference with a finite score, best first, earlier reference first on ties (multiset order)
    const size_t nx = x.size(), np = nx * y.size();
    std::vector<uint32_t> xi (np), yi (np), xs (np), xe (np);
    for (size_t n = 0; n < y.size(); ++n) for (size_t m = 0; m < nx; ++m) { xi[n * nx + m] = (uint32_t) m; yi[n * nx + m] = (uint32_t) n; }
    std::vector<double> score (np);
    std::vector<uint64_t> off (np + 1);
    uint8_t* path = NULL;
    g.ok (qg_viterbi (g.ctx, &gc, np, xi.data(), yi.data(), NULL, score.data(), xs.data(), xe.data(), &path, off.data()));
    for (size_t n = 0; n < y.size(); ++n) {
      std::vector<size_t> order;
      for (size_t m = 0; m < nx; ++m) if (score[n * nx + m] > -numeric_limits<double>::infinity()) order.push_back (n * nx + m);
      std::stable_sort (order.begin(), order.end(), [&] (size_t a, size_t b) { return score[a] > score[b]; });
      for (size_t p : order)
        aligner.writeAlignment (out, alignmentFromPath (x[xi[p]], y[n], xs[p], xe[p], path + off[p], off[p+1] - off[p], score[p] - nullLL[n], config.local));
    }
    qg_free (path);
  }
}

// ---- seam B -------------------------------------------------------------------------------------------------------
void quaffGpuOverlap (QuaffOverlapAligner& aligner, std::ostream& out, const vguard<FastSeq>& seqs, size_t nOriginals,
                      const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config) {
  Gpu g;
  const Flat f = flatten (seqs, true);
  g.ok (qg_set_seqs (g.ctx, QG_READS, seqs.size(), f.tok.data(), f.quals ? f.qual.data() : NULL, f.off.data()));
  const QuaffScores qs (params);
  const ScoreTables t = tables (qs);
  qg_overlap_model om;
  om.match_k = t.K; om.gap_k = t.G; om.match = t.match.data(); om.insert = t.insert.data();
  for (int r = 0; r < 4; ++r) om.log_ref_base[r] = log (params.refBase[r]);
  om.begin_insert = params.beginInsert.data(); om.begin_delete = params.beginDelete.data();
  om.extend_insert = params.extendInsert; om.extend_delete = params.extendDelete;
  g.ok (qg_set_overlap_model (g.ctx, &om));
  const qg_dpconfig gc = gpuConfig (config);
  // stored reverse strands: nullModel.logLikelihood (y.revcomp()) of the reference = the original read's
  std::vector<double> nullLL (seqs.size());
  for (size_t n = 0; n < seqs.size(); ++n)
    nullLL[n] = nullModel.logLikelihood (n >= nOriginals ? seqs[n].revcomp() : seqs[n]);
  size_t np = 0; uint32_t *xi = NULL, *yi = NULL, *co = NULL; double* score = NULL; uint8_t* path = NULL; uint64_t* off = NULL;
  g.ok (qg_overlap_reads (g.ctx, &gc, nOriginals, nullLL.data(), &np, &xi, &yi, &score, &co, &path, &off));
  aligner.writeAlignmentHeader (out, seqs, false);
  for (size_t p = 0; p < np; ++p) {
    if (!(score[p] > -numeric_limits<double>::infinity())) continue;
    const FastSeq& x = seqs[xi[p]]; const FastSeq& y = seqs[yi[p]];
    const uint32_t xStart = co[4*p], xEnd = co[4*p+1], yStart = co[4*p+2], yEnd = co[4*p+3];
    // rows with the reference's squashing of adjacent insertions and deletions (qoverlap.cpp:231-267)
    std::string xRow, yRow, xQual, yQual;
    size_t i = xStart - 1, j = yStart - 1;
    const uint8_t* ops = path + off[p]; const uint64_t n = off[p+1] - off[p];
    uint64_t tt = 0;
    while (tt < n) {
      if (ops[tt] == QG_OP_MATCH) {
        xRow += x.seq[i]; yRow += y.seq[j];
        if (x.hasQual()) xQual += x.qual[i];
        if (y.hasQual()) yQual += y.qual[j];
        ++i; ++j; ++tt; continue;
      }
      uint64_t nd = 0, ni = 0, e = tt;
      while (e < n && ops[e] != QG_OP_MATCH) { if (ops[e] == QG_OP_DELETE) ++nd; else ++ni; ++e; }
      const uint64_t sh = std::min (nd, ni);
      for (uint64_t s = 0; s < sh; ++s) { xRow += x.seq[i+s]; yRow += y.seq[j+s]; if (x.hasQual()) xQual += x.qual[i+s]; if (y.hasQual()) yQual += y.qual[j+s]; }
      for (uint64_t s = sh; s < nd; ++s) { xRow += x.seq[i+s]; yRow += Alignment::gapChar; if (x.hasQual()) xQual += x.qual[i+s]; if (y.hasQual()) yQual += FastSeq::maxQualityChar; }
      for (uint64_t s = sh; s < ni; ++s) { xRow += Alignment::gapChar; yRow += y.seq[j+s]; if (x.hasQual()) xQual += FastSeq::maxQualityChar; if (y.hasQual()) yQual += y.qual[j+s]; }
      i += nd; j += ni; tt = e;
    }
    Alignment align (2);
    align.gappedSeq[0].name = "read_x";
    align.gappedSeq[0].comment = "substr(" + x.name + "," + to_string (xStart) + ".." + to_string (xEnd) + ")";
    align.gappedSeq[1].name = "read_y";
    align.gappedSeq[1].comment = "substr(" + y.name + "," + to_string (yStart) + ".." + to_string (yEnd) + ")";
    align.gappedSeq[0].seq = xRow; align.gappedSeq[1].seq = yRow;
    align.gappedSeq[0].qual = xQual; align.gappedSeq[1].qual = yQual;
    align.gappedSeq[0].source.name = x.name; align.gappedSeq[0].source.start = xStart; align.gappedSeq[0].source.end = xEnd;
    align.gappedSeq[1].source.name = y.name; align.gappedSeq[1].source.start = yStart; align.gappedSeq[1].source.end = yEnd;
    align.gappedSeq[0].source = align.gappedSeq[0].source.compose (x.source);
    align.gappedSeq[1].source = align.gappedSeq[1].source.compose (y.source);
    align.score = score[p];
    aligner.writeAlignment (out, align);
  }
  qg_free (xi); qg_free (yi); qg_free (co); qg_free (score); qg_free (path); qg_free (off);
}

// ---- seam C -------------------------------------------------------------------------------------------------------
QuaffParamCounts quaffGpuGetCounts (QuaffTrainer& trainer, const vguard<FastSeq>& x, const vguard<FastSeq>& y,
                                    const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config,
                                    vguard<vguard<size_t> >& sortOrder, double& logLike) {
  Gpu g;
  const Flat fx = flatten (x, false), fy = flatten (y, true);
  Require (fy.quals, "Forward-Backward algorithm requires quality scores to fit model");
  g.ok (qg_set_seqs (g.ctx, QG_REFS, x.size(), fx.tok.data(), NULL, fx.off.data()));
  g.ok (qg_set_seqs (g.ctx, QG_READS, y.size(), fy.tok.data(), fy.qual.data(), fy.off.data()));
  setAlignModel (g, params);
  if (getenv ("QUAFF_GPU_EXACT")) g.ok (qg_set_option (g.ctx, QG_OPT_FB_EXACT, 1));
  const qg_dpconfig gc = gpuConfig (config);
  const size_t nx = x.size(), ny = y.size();
  std::vector<double> nullLL (ny), yLL (ny);
  for (size_t n = 0; n < ny; ++n) nullLL[n] = nullModel.logLikelihood (y[n]);
  std::vector<uint32_t> so (nx * ny, 0), soLen (ny);
  for (size_t n = 0; n < ny; ++n) {
    soLen[n] = (uint32_t) sortOrder[n].size();
    for (size_t s = 0; s < sortOrder[n].size(); ++s) so[n * nx + s] = (uint32_t) sortOrder[n][s];
  }
  const unsigned int K = params.matchContext.kmerLen, G = params.indelContext.kmerLen;
  std::vector<double> flat (qg_counts_size (K, G));
  g.ok (qg_estep (g.ctx, &gc, trainer.allowNullModel ? 1 : 0, nullLL.data(), so.data(), soLen.data(), yLL.data(), flat.data(), &logLike));
  for (size_t n = 0; n < ny; ++n) sortOrder[n] = vguard<size_t> (so.begin() + n * nx, so.begin() + n * nx + soLen[n]);
  QuaffParamCounts counts (K, G);
  size_t k = 0;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (Kmer j = 0; j < counts.matchContext.numKmers; ++j)
      for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q) counts.match[i][j].qualCount[q] = flat[k++];
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q) counts.insert[i].qualCount[q] = flat[k++];
  const Kmer nG = counts.indelContext.numKmers;
  for (Kmer j = 0; j < nG; ++j) counts.beginInsertNo[j] = flat[k++];
  for (Kmer j = 0; j < nG; ++j) counts.beginInsertYes[j] = flat[k++];
  for (Kmer j = 0; j < nG; ++j) counts.beginDeleteNo[j] = flat[k++];
  for (Kmer j = 0; j < nG; ++j) counts.beginDeleteYes[j] = flat[k++];
  counts.extendInsertNo = flat[k++]; counts.extendInsertYes = flat[k++];
  counts.extendDeleteNo = flat[k++]; counts.extendDeleteYes = flat[k++];
  if (trainer.rawCountsFilename.size()) {
    ofstream outf (trainer.rawCountsFilename);
    counts.writeJson (outf) << endl;
  }
  return counts;
}
