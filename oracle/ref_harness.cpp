// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// C-callable taps into the UNMODIFIED ihh/quaff classes, compiled from the sources where they
// lie under /root/reference by oracle/Makefile into oracle/_ref/libquaffref.so.  It is used
//   * to pin oracle/quaff_oracle.c (the portable CPU restatement) against the real reference,
//   * to generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * optionally as the "reference" CPU baseline of bench.py.
// Every entry point just constructs the reference object named in its comment and copies
// public members out; there is no algorithm in this file.
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>
#include <limits>
#include <chrono>
#include "qmodel.h"
#include "qoverlap.h"
#include "logsumexp.h"

namespace {
char* dupString (const std::string& s) {
  char* p = (char*) malloc (s.size() + 1);
  memcpy (p, s.c_str(), s.size() + 1);
  return p;
}

struct Cfg {
  int sparse, kmerLen, kmerThreshold, bandSize, local;
  uint64_t maxSize;
};

void fillConfig (QuaffDPConfig& config, const Cfg* c) {
  config.sparse = c->sparse != 0;
  config.kmerLen = c->kmerLen;
  config.kmerThreshold = c->kmerThreshold;
  config.bandSize = c->bandSize;
  config.local = c->local != 0;
  config.maxSize = (size_t) c->maxSize;
  config.autoMemSize = false;
  config.threads = 1;
}

// flat QuaffCounts layout shared with include/quaffgpu.h:
//   match[4][4^K][94], insert[4][94], m2m[4^G], m2i[4^G], m2d[4^G], m2e[4^G], d2d, d2m, i2i, i2m
size_t flattenCounts (const QuaffCounts& qc, double* out) {
  size_t n = 0;
  const Kmer nK = qc.matchContext.numKmers, nG = qc.indelContext.numKmers;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (Kmer j = 0; j < nK; ++j)
      for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
	out[n++] = qc.match[i][j].qualCount[q];
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
      out[n++] = qc.insert[i].qualCount[q];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.m2m[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.m2i[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.m2d[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.m2e[j];
  out[n++] = qc.d2d; out[n++] = qc.d2m; out[n++] = qc.i2i; out[n++] = qc.i2m;
  return n;
}

// flat QuaffParamCounts layout:
//   match[4][4^K][94], insert[4][94], beginInsertNo[4^G], beginInsertYes[4^G], beginDeleteNo[4^G],
//   beginDeleteYes[4^G], extendInsertNo, extendInsertYes, extendDeleteNo, extendDeleteYes
size_t flattenParamCounts (const QuaffParamCounts& qc, double* out) {
  size_t n = 0;
  const Kmer nK = qc.matchContext.numKmers, nG = qc.indelContext.numKmers;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (Kmer j = 0; j < nK; ++j)
      for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
	out[n++] = qc.match[i][j].qualCount[q];
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i)
    for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
      out[n++] = qc.insert[i].qualCount[q];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.beginInsertNo[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.beginInsertYes[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.beginDeleteNo[j];
  for (Kmer j = 0; j < nG; ++j) out[n++] = qc.beginDeleteYes[j];
  out[n++] = qc.extendInsertNo; out[n++] = qc.extendInsertYes;
  out[n++] = qc.extendDeleteNo; out[n++] = qc.extendDeleteYes;
  return n;
}

template<class Matrix>
double* dumpCells (const Matrix& m, const DiagonalEnvelope& env, uint64_t* nCells) {
  std::vector<double> v;
  for (SeqIdx j = 1; j <= m.yLen; ++j)
    for (DiagonalEnvelope::iterator pi = env.begin(j); !pi.finished(); ++pi) {
      v.push_back (m.mat(*pi,j));
      v.push_back (m.ins(*pi,j));
      v.push_back (m.del(*pi,j));
    }
  *nCells = v.size() / 3;
  double* out = (double*) malloc (v.size() * sizeof(double) + 8);
  memcpy (out, v.data(), v.size() * sizeof(double));
  return out;
}
}  // namespace

extern "C" {

void qref_free (void* p) { free (p); }

// QuaffParams::readJson (qmodel.cpp:270-273, 210-268)
void* qref_params_from_json (const char* text) {
  QuaffParams* qp = new QuaffParams();
  std::istringstream in (text);
  qp->readJson (in);
  return qp;
}
void qref_params_free (void* p) { delete (QuaffParams*) p; }
int qref_params_orders (void* p, int* matchK, int* gapK) {
  const QuaffParams* qp = (const QuaffParams*) p;
  *matchK = qp->matchContext.kmerLen;
  *gapK = qp->indelContext.kmerLen;
  return 0;
}
// QuaffParams::writeJson
char* qref_params_to_json (void* p) {
  std::ostringstream out;
  ((const QuaffParams*) p)->writeJson (out);
  return dupString (out.str());
}


// the parameter values exactly as the reference's JSON reader (gason) parsed them; gason's
// hand-rolled number parser is not correctly rounded, so these can differ from strtod by an ulp
int qref_params_values (void* p, double* beginInsert, double* beginDelete, double* extend2, double* insertPQR, double* matchPQR, double* refBase) {
  const QuaffParams* qp = (const QuaffParams*) p;
  const Kmer nK = qp->matchContext.numKmers, nG = qp->indelContext.numKmers;
  for (Kmer g = 0; g < nG; ++g) { beginInsert[g] = qp->beginInsert[g]; beginDelete[g] = qp->beginDelete[g]; }
  extend2[0] = qp->extendInsert; extend2[1] = qp->extendDelete;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i) {
    refBase[i] = qp->refBase[i];
    insertPQR[3*i] = qp->insert[i].symProb; insertPQR[3*i+1] = qp->insert[i].qualTrialSuccessProb; insertPQR[3*i+2] = qp->insert[i].qualNumSuccessfulTrials;
    for (Kmer j = 0; j < nK; ++j) {
      const SymQualDist& d = qp->match[i][j];
      matchPQR[3*(i*nK+j)] = d.symProb; matchPQR[3*(i*nK+j)+1] = d.qualTrialSuccessProb; matchPQR[3*(i*nK+j)+2] = d.qualNumSuccessfulTrials;
    }
  }
  return 0;
}
int qref_null_values (void* p, double* nullEmit, double* nullPQR) {
  const QuaffNullParams* np = (const QuaffNullParams*) p;
  *nullEmit = np->nullEmit;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i) {
    nullPQR[3*i] = np->null[i].symProb; nullPQR[3*i+1] = np->null[i].qualTrialSuccessProb; nullPQR[3*i+2] = np->null[i].qualNumSuccessfulTrials;
  }
  return 0;
}


// QuaffNullParams (const vguard<FastSeq>&, double) -- the null model auto-fitted from the reads when -null is
// absent (qmodel.cpp:1811-1843; t/quaff.cpp:419-424); returned through qref_null_values
void* qref_null_fit (void** seqs, int n) {
  vguard<FastSeq> v;
  for (int i = 0; i < n; ++i) v.push_back (*(const FastSeq*) seqs[i]);
  return new QuaffNullParams (v);
}

// QuaffNullParams::readJson (qmodel.cpp:1845-1848)
void* qref_null_from_json (const char* text) {
  QuaffNullParams* np = new QuaffNullParams();
  std::istringstream in (text);
  np->readJson (in);
  return np;
}
void qref_null_free (void* p) { delete (QuaffNullParams*) p; }

void* qref_seq_new (const char* name, const char* seq, const char* qual) {
  FastSeq* fs = new FastSeq();
  fs->name = name;
  fs->seq = seq;
  if (qual) fs->qual = qual;
  return fs;
}
void qref_seq_free (void* p) { delete (FastSeq*) p; }
// FastSeq::revcomp (fastseq.cpp:218-230)
void* qref_seq_revcomp (void* p) { return new FastSeq (((const FastSeq*) p)->revcomp()); }
char* qref_seq_bases (void* p) { return dupString (((const FastSeq*) p)->seq); }
char* qref_seq_quals (void* p) { return dupString (((const FastSeq*) p)->qual); }

// FastSeq::kmers (fastseq.cpp:85-99)
int qref_seq_kmers (void* p, int k, uint64_t* out) {
  const vguard<Kmer> km = ((const FastSeq*) p)->kmers (dnaAlphabet, k);
  for (size_t n = 0; n < km.size(); ++n) out[n] = km[n];
  return (int) km.size();
}

// QuaffNullParams::logLikelihood (qmodel.cpp:1875-1890)
double qref_null_loglike (void* np, void* seq) {
  return ((const QuaffNullParams*) np)->logLikelihood (*(const FastSeq*) seq);
}

// log_sum_exp (logsumexp.cpp:34-59)
double qref_lse (double a, double b) { return log_sum_exp (a, b); }
double qref_lse_unary (double x) { return log_sum_exp_unary (x); }

// QuaffScores (qmodel.cpp:296-325) flattened: match[4][4^K][95] with slot 94 = logSymProb, insert[4][95]
int qref_scores (void* p, double* match, double* insert, double* m2m, double* m2i, double* m2d, double* m2e, double* scal4) {
  const QuaffScores qs (*(const QuaffParams*) p);
  const Kmer nK = qs.matchContext.numKmers, nG = qs.indelContext.numKmers;
  const size_t Q1 = FastSeq::qualScoreRange + 1;
  for (AlphTok i = 0; i < dnaAlphabetSize; ++i) {
    for (Kmer j = 0; j < nK; ++j) {
      for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
	match[(i * nK + j) * Q1 + q] = qs.match[i][j].logSymQualProb[q];
      match[(i * nK + j) * Q1 + FastSeq::qualScoreRange] = qs.match[i][j].logSymProb;
    }
    for (QualScore q = 0; q < FastSeq::qualScoreRange; ++q)
      insert[i * Q1 + q] = qs.insert[i].logSymQualProb[q];
    insert[i * Q1 + FastSeq::qualScoreRange] = qs.insert[i].logSymProb;
  }
  for (Kmer j = 0; j < nG; ++j) { m2m[j] = qs.m2m[j]; m2i[j] = qs.m2i[j]; m2d[j] = qs.m2d[j]; m2e[j] = qs.m2e[j]; }
  scal4[0] = qs.d2d; scal4[1] = qs.d2m; scal4[2] = qs.i2i; scal4[3] = qs.i2m;
  return 0;
}

// KmerIndex (fastseq.cpp:240-256) + QuaffDPConfig::makeEnvelope (qmodel.cpp:1049-1056)
//   -> DiagonalEnvelope::initSparse / initFull (diagenv.cpp:11-106)
int qref_envelope (void* x, void* y, const Cfg* c, uint64_t cellSize, int** diagsOut, uint64_t* totalStorage, uint64_t* cellUpdates) {
  QuaffDPConfig config;
  fillConfig (config, c);
  const FastSeq& xfs = *(const FastSeq*) x;
  const FastSeq& yfs = *(const FastSeq*) y;
  const KmerIndex yKmerIndex (yfs, dnaAlphabet, config.kmerLen);
  const DiagonalEnvelope env = config.makeEnvelope (xfs, yKmerIndex, (size_t) cellSize);
  int* d = (int*) malloc (sizeof(int) * (env.diagonals.size() + 1));
  for (size_t n = 0; n < env.diagonals.size(); ++n) d[n] = env.diagonals[n];
  *diagsOut = d;
  *totalStorage = env.totalStorageSize;
  uint64_t cu = 0;
  for (SeqIdx j = 1; j <= env.yLen; ++j)
    cu += env.endIntersecting(j) - env.beginIntersecting(j);
  *cellUpdates = cu;
  return (int) env.diagonals.size();
}

static void rowsFromAlignment (const Alignment& a, char** xrow, char** yrow) {
  *xrow = dupString (a.gappedSeq[0].seq);
  *yrow = dupString (a.gappedSeq[1].seq);
}

// QuaffViterbiMatrix ctor (qmodel.cpp:1512-1560) + alignment() (qmodel.cpp:1562-1646)
int qref_viterbi (void* x, void* y, void* params, const Cfg* c, double* result,
		  uint32_t* xStart, uint32_t* xEnd, char** xrow, char** yrow,
		  double** cellsOut, uint64_t* nCells, double* seconds) {
  QuaffDPConfig config;
  fillConfig (config, c);
  const FastSeq& xfs = *(const FastSeq*) x;
  const FastSeq& yfs = *(const FastSeq*) y;
  const KmerIndex yKmerIndex (yfs, dnaAlphabet, config.kmerLen);
  const auto t0 = std::chrono::steady_clock::now();
  const DiagonalEnvelope env = config.makeEnvelope (xfs, yKmerIndex, QuaffDPMatrixContainer::cellSize());
  const auto t1 = std::chrono::steady_clock::now();
  const QuaffViterbiMatrix viterbi (env, *(const QuaffParams*) params, config);
  const auto t2 = std::chrono::steady_clock::now();
  *result = viterbi.result;
  *xrow = *yrow = NULL;
  *xStart = *xEnd = 0;
  if (viterbi.resultIsFinite()) {
    const Alignment a = viterbi.alignment();
    rowsFromAlignment (a, xrow, yrow);
    *xStart = a.gappedSeq[0].source.start;
    *xEnd = a.gappedSeq[0].source.end;
  }
  const auto t3 = std::chrono::steady_clock::now();
  if (cellsOut) *cellsOut = dumpCells (viterbi, env, nCells);
  if (seconds) {
    seconds[0] = std::chrono::duration<double> (t1 - t0).count();
    seconds[1] = std::chrono::duration<double> (t2 - t1).count();
    seconds[2] = std::chrono::duration<double> (t3 - t2).count();
  }
  return 0;
}

// QuaffForwardMatrix ctor (qmodel.cpp:1343-1391)
int qref_forward (void* x, void* y, void* params, const Cfg* c, double* result, double** cellsOut, uint64_t* nCells, double* seconds) {
  QuaffDPConfig config;
  fillConfig (config, c);
  const FastSeq& xfs = *(const FastSeq*) x;
  const FastSeq& yfs = *(const FastSeq*) y;
  const KmerIndex yKmerIndex (yfs, dnaAlphabet, config.kmerLen);
  const auto t0 = std::chrono::steady_clock::now();
  const DiagonalEnvelope env = config.makeEnvelope (xfs, yKmerIndex, 2*QuaffDPMatrixContainer::cellSize());
  const auto t1 = std::chrono::steady_clock::now();
  const QuaffForwardMatrix fwd (env, *(const QuaffParams*) params, config);
  const auto t2 = std::chrono::steady_clock::now();
  *result = fwd.result;
  if (cellsOut) *cellsOut = dumpCells (fwd, env, nCells);
  if (seconds) {
    seconds[0] = std::chrono::duration<double> (t1 - t0).count();
    seconds[1] = std::chrono::duration<double> (t2 - t1).count();
  }
  return 0;
}

// QuaffForwardMatrix + QuaffBackwardMatrix ctor (qmodel.cpp:1393-1503)
int qref_backward (void* x, void* y, void* params, const Cfg* c, double* fwdResult, double* backResult, double* countsFlat,
		   double** cellsOut, uint64_t* nCells, double* seconds) {
  QuaffDPConfig config;
  fillConfig (config, c);
  const FastSeq& xfs = *(const FastSeq*) x;
  const FastSeq& yfs = *(const FastSeq*) y;
  const KmerIndex yKmerIndex (yfs, dnaAlphabet, config.kmerLen);
  const DiagonalEnvelope env = config.makeEnvelope (xfs, yKmerIndex, 2*QuaffDPMatrixContainer::cellSize());
  const auto t0 = std::chrono::steady_clock::now();
  const QuaffForwardMatrix fwd (env, *(const QuaffParams*) params, config);
  const auto t1 = std::chrono::steady_clock::now();
  const QuaffBackwardMatrix back (fwd);
  const auto t2 = std::chrono::steady_clock::now();
  *fwdResult = fwd.result;
  *backResult = back.result;
  flattenCounts (back.qc, countsFlat);
  if (cellsOut) *cellsOut = dumpCells (back, env, nCells);
  if (seconds) {
    seconds[0] = std::chrono::duration<double> (t1 - t0).count();
    seconds[1] = std::chrono::duration<double> (t2 - t1).count();
  }
  return 0;
}

// QuaffOverlapViterbiMatrix ctor (qoverlap.cpp:77-160) + alignment() (qoverlap.cpp:162-290)
int qref_overlap (void* x, void* y, void* params, const Cfg* c, int yComplemented, double* result,
		  uint32_t* coords4, char** xrow, char** yrow, double** cellsOut, uint64_t* nCells) {
  QuaffDPConfig config;
  fillConfig (config, c);
  const FastSeq& xfs = *(const FastSeq*) x;
  const FastSeq& yfs = *(const FastSeq*) y;
  const KmerIndex yKmerIndex (yfs, dnaAlphabet, config.kmerLen);
  const DiagonalEnvelope env = config.makeEnvelope (xfs, yKmerIndex, QuaffDPMatrixContainer::cellSize());
  const QuaffOverlapViterbiMatrix viterbi (env, *(const QuaffParams*) params, yComplemented != 0);
  *result = viterbi.result;
  *xrow = *yrow = NULL;
  coords4[0] = coords4[1] = coords4[2] = coords4[3] = 0;
  if (viterbi.resultIsFinite()) {
    const Alignment a = viterbi.alignment();
    rowsFromAlignment (a, xrow, yrow);
    // coordinates before compose() with the sequences' own source intervals are not kept by the
    // reference; recover them from the "substr(name,s..e)" comments it builds (qoverlap.cpp:271-274)
    for (int r = 0; r < 2; ++r) {
      const std::string& cm = a.gappedSeq[r].comment;
      const size_t comma = cm.rfind (','), dots = cm.rfind (".."), close = cm.rfind (')');
      coords4[2*r] = (uint32_t) atol (cm.substr (comma + 1, dots - comma - 1).c_str());
      coords4[2*r+1] = (uint32_t) atol (cm.substr (dots + 2, close - dots - 2).c_str());
    }
  }
  if (cellsOut) *cellsOut = dumpCells (viterbi, env, nCells);
  return 0;
}

// QuaffOverlapScores (qoverlap.cpp:9-75): six transition scalars as stored (NOT through the swapped
// accessors), plus m2m/m2i/m2d [4^G][4^G] and the matchMinusInsert[iK][jK].logSymQualPairProb table
int qref_overlap_scores (void* params, int yComplemented, double* scal6, double* m2m, double* m2i, double* m2d,
			 double* pairTable /* [nK][nK][94][94] or NULL */, double* xOnly /* [nK][nK][94] */, double* yOnly, double* none /* [nK][nK] */) {
  const QuaffOverlapScores qos (*(const QuaffParams*) params, yComplemented != 0);
  const Kmer nK = qos.matchContext.numKmers, nG = qos.indelContext.numKmers;
  const size_t Q = FastSeq::qualScoreRange;
  scal6[0] = qos.i2m; scal6[1] = qos.i2i; scal6[2] = qos.i2d; scal6[3] = qos.d2m; scal6[4] = qos.d2i; scal6[5] = qos.d2d;
  for (Kmer i = 0; i < nG; ++i)
    for (Kmer j = 0; j < nG; ++j) {
      m2m[i*nG+j] = qos.m2m[i][j]; m2i[i*nG+j] = qos.m2i[i][j]; m2d[i*nG+j] = qos.m2d[i][j];
    }
  for (Kmer i = 0; i < nK; ++i)
    for (Kmer j = 0; j < nK; ++j) {
      const SymQualPairScores& s = qos.matchMinusInsert[i][j];
      if (pairTable)
	for (size_t a = 0; a < Q; ++a)
	  for (size_t b = 0; b < Q; ++b)
	    pairTable[((i*nK+j)*Q + a)*Q + b] = s.logSymQualPairProb[a][b];
      for (size_t a = 0; a < Q; ++a) {
	if (xOnly) xOnly[(i*nK+j)*Q + a] = s.logSymPairXQualProb[a];
	if (yOnly) yOnly[(i*nK+j)*Q + a] = s.logSymPairYQualProb[a];
      }
      if (none) none[i*nK+j] = s.logSymPairProb;
    }
  return 0;
}

// QuaffCountingTask::run for every read (qmodel.cpp:2238-2271), then the sums of
// QuaffCountingScheduler::finalCounts / finalLogLike (qmodel.cpp:2416-2422)
// sortOrder: in/out, flattened [ny][nx] with per-read lengths in sortLen (in/out)
int qref_estep (void** xs, int nx, void** ys, int ny, void* params, void* nullp, int useNull, const Cfg* c,
		uint32_t* sortOrder, uint32_t* sortLen, double* yLogLike, double* paramCountsFlat) {
  QuaffDPConfig config;
  fillConfig (config, c);
  vguard<FastSeq> x;
  for (int n = 0; n < nx; ++n) x.push_back (*(const FastSeq*) xs[n]);
  const QuaffParams& qp = *(const QuaffParams*) params;
  const QuaffNullParams& np = *(const QuaffNullParams*) nullp;
  QuaffParamCounts total (qp.matchContext.kmerLen, qp.indelContext.kmerLen);
  for (int m = 0; m < ny; ++m) {
    vguard<size_t> so (sortOrder + (size_t) m * nx, sortOrder + (size_t) m * nx + sortLen[m]);
    double ll = 0;
    QuaffParamCounts yc (qp.matchContext.kmerLen, qp.indelContext.kmerLen);
    QuaffCountingTask task (x, *(const FastSeq*) ys[m], (size_t) m, qp, np, useNull != 0, config, so, ll, yc);
    task.run();
    yLogLike[m] = ll;
    sortLen[m] = (uint32_t) so.size();
    for (size_t n = 0; n < so.size(); ++n) sortOrder[(size_t) m * nx + n] = (uint32_t) so[n];
    total = total + yc;
  }
  flattenParamCounts (total, paramCountsFlat);
  return 0;
}

}  // extern "C"
