// Seams A/B/C of the reference (QuaffAligner::align qmodel.cpp:2624, QuaffOverlapAligner::align qoverlap.cpp:312,
// QuaffTrainer::getCounts qmodel.cpp:2005) routed to libquaffgpu.  The reference keeps everything else: flag parsing,
// sequence loading, params / null model, the Alignment type and every output writer.
#include <time.h>
#include <zlib.h>
#include <ctype.h>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <thread>
#include <future>
#include <memory>
#include "qmodel.h"
#include "qoverlap.h"
#include "quaffgpu.h"

int quaffGpuDevice = -1;
std::vector<int> quaffGpuDevices;                 // -gpu 0,1,2,3 / -gpu all: the reads shard over these devices

namespace {
std::shared_future<void> gpuWarmDone;                      // detached worker: an error exit of the CLI must not trip over a joinable thread
bool gpuWarmStarted = false;
}
void quaffGpuWarmStart () {
  if (gpuWarmStarted || getenv ("QUAFF_GPU_NO_WARM_START")) return;
  gpuWarmStarted = true;
  const std::vector<int> devs = quaffGpuDevices;
  std::shared_ptr<std::promise<void> > done (new std::promise<void>);
  gpuWarmDone = done->get_future().share();
  std::thread ([devs, done] { qg_init_devices (devs.data(), (int) devs.size()); done->set_value(); }).detach();   // errors resurface in qg_create
}
void quaffGpuWarmJoin () { if (gpuWarmStarted && gpuWarmDone.valid()) gpuWarmDone.wait(); }

// -gpu [device | device,device,... | all]
bool quaffGpuParseArg (std::deque<std::string>& argvec) {
  if (argvec.size() && argvec[0] == "-gpu") {
    argvec.pop_front();
    quaffGpuDevices.clear();
    if (argvec.size() && argvec[0] == "all") {
      const int n = qg_device_count();
      for (int d = 0; d < n; ++d) quaffGpuDevices.push_back (d);
      argvec.pop_front();
    } else if (argvec.size() && !argvec[0].empty() && argvec[0].size() <= 48
               && argvec[0].find_first_not_of ("0123456789,") == std::string::npos && isdigit (argvec[0][0]) && isdigit (argvec[0][argvec[0].size() - 1])) {
      std::stringstream ss (argvec[0]);
      std::string tok;
      while (std::getline (ss, tok, ',')) if (!tok.empty()) quaffGpuDevices.push_back (atoi (tok.c_str()));
      argvec.pop_front();
    }
    if (quaffGpuDevices.empty()) quaffGpuDevices.push_back (0);
    quaffGpuDevice = quaffGpuDevices[0];
    // the devices are known before any sequence file has been read: initialise them on a side thread while the
    // reference's loader parses FASTA / FASTQ (joined by the seams before they create their contexts)
    quaffGpuWarmStart();
    return true;
  }
  return false;
}

namespace {
// [0, n) in contiguous ranges over the host's cores (tokenising and the null model are per-read work)
template<class F>
void parallelRanges (size_t n, F fn) {
  size_t nt = std::min<size_t> (std::max (1u, std::thread::hardware_concurrency()), 16);
  if (n < 256 || nt < 2) { fn ((size_t) 0, n); return; }
  std::vector<std::thread> th;
  for (size_t t = 0; t < nt; ++t) { const size_t a = n * t / nt, b = n * (t + 1) / nt; if (a < b) th.emplace_back ([=] { fn (a, b); }); }
  for (auto& x : th) x.join();
}

struct Flat { std::vector<uint8_t> tok, qual; std::vector<uint64_t> off; bool quals; };

// FastSeq::tokens / FastSeq::qualScores (fastseq.cpp:71-83, 101-109) for a whole set into flat arrays, through 256-entry
// tables built from the reference's own per-character functions (one pass, no per-read vectors)
Flat flatten (const vguard<FastSeq>& seqs, bool wantQual) {
  Flat f;
  f.quals = wantQual && !seqs.empty();
  for (const auto& s : seqs) f.quals = f.quals && s.hasQual();
  if (wantQual && !f.quals)
    for (const auto& s : seqs)
      Require (!s.hasQual(), "-gpu: either all reads or no reads must carry quality scores (%s)", s.name.c_str());
  int tokOf[256]; uint8_t qualOf[256];
  for (int c = 0; c < 256; ++c) { tokOf[c] = c ? tokenize ((char) c, dnaAlphabet) : -1; qualOf[c] = (uint8_t) FastSeq::qualScoreForChar ((char) c); }
  f.off.resize (seqs.size() + 1);
  f.off[0] = 0;
  for (size_t n = 0; n < seqs.size(); ++n) f.off[n+1] = f.off[n] + seqs[n].length();
  f.tok.resize (f.off.back());
  if (f.quals) f.qual.resize (f.off.back());
  const std::string* bad = NULL; char badChar = 0;
  std::mutex badMx;
  parallelRanges (seqs.size(), [&] (size_t n0, size_t n1) {
    for (size_t n = n0; n < n1; ++n) {
      const FastSeq& s = seqs[n];
      uint8_t* t = f.tok.data() + f.off[n];
      const size_t len = s.length();
      for (size_t i = 0; i < len; ++i) {
        const int v = tokOf[(unsigned char) s.seq[i]];
        if (v < 0) { std::lock_guard<std::mutex> lock (badMx); if (!bad || &s.name < bad) { bad = &s.name; badChar = s.seq[i]; } return; }
        t[i] = (uint8_t) v;
      }
      if (f.quals) { uint8_t* q = f.qual.data() + f.off[n]; for (size_t i = 0; i < len; ++i) q[i] = qualOf[(unsigned char) s.qual[i]]; }
    }
  });
  if (bad) { cerr << "Unknown symbol " << badChar << " in sequence " << *bad << endl; throw; }      // FastSeq::tokens, fastseq.cpp:76-79
  return f;
}

struct Gpu {
  qg_ctx* ctx;
  Gpu () : ctx (NULL) { quaffGpuWarmJoin(); Require (qg_create (&ctx, quaffGpuDevice) == QG_OK, "-gpu: %s", qg_last_error (NULL)); }
  ~Gpu () { qg_destroy (ctx); }
  void ok (int rc) const { Require (rc == QG_OK, "-gpu: %s", qg_last_error (ctx)); }
};

// QUAFF_GPU_TRACE=1: wall-clock stamps of the seam's host phases on stderr
struct Trace {
  bool on; double t0;
  static double now () { struct timespec ts; clock_gettime (CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  Trace () : on (getenv ("QUAFF_GPU_TRACE") != NULL), t0 (now()) { }
  void mark (const char* what) { if (on) { const double t = now(); fprintf (stderr, "[quaff-gpu] %-28s %8.3f s\n", what, t - t0); t0 = t; } }
};

size_t envSize (const char* name, size_t dflt) { const char* v = getenv (name); return (v && *v) ? (size_t) strtoull (v, NULL, 10) : dflt; }

// every device of -gpu, `perDevice` contexts on each (one host thread + one stream per context, inside the library)
struct GpuPool {
  qg_pool* pool;
  GpuPool (int perDevice) : pool (NULL) {
    if (quaffGpuDevices.empty()) quaffGpuDevices.push_back (quaffGpuDevice);
    quaffGpuWarmJoin();
    Require (qg_pool_create (&pool, quaffGpuDevices.data(), (int) quaffGpuDevices.size(), perDevice) == QG_OK, "-gpu: %s", qg_pool_last_error (NULL));
  }
  ~GpuPool () { qg_pool_destroy (pool); }
  void ok (int rc) const { Require (rc == QG_OK, "-gpu: %s", qg_pool_last_error (pool)); }
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

// QuaffNullParams::logLikelihood (qmodel.cpp:1875-1890) for every sequence of a flattened set: the library's table-driven
// form of the same sum (same addends, same order); the reference evaluates a negative-binomial pdf per base (~1.4 ms per 8 kb read)
std::vector<double> nullLogLikes (const QuaffNullParams& nullModel, const Flat& f, const vguard<FastSeq>& seqs) {
  double pqr[12];
  for (int t = 0; t < 4; ++t) { pqr[3*t] = nullModel.null[t].symProb; pqr[3*t+1] = nullModel.null[t].qualTrialSuccessProb; pqr[3*t+2] = nullModel.null[t].qualNumSuccessfulTrials; }
  std::vector<double> ll (seqs.size());
  parallelRanges (seqs.size(), [&] (size_t n0, size_t n1) {
    for (size_t n = n0; n < n1; ++n)
      ll[n] = qg_null_loglike (nullModel.nullEmit, pqr, f.tok.data() + f.off[n], seqs[n].hasQual() ? f.qual.data() + f.off[n] : NULL, f.off[n+1] - f.off[n]);
  });
  return ll;
}

template<class Setter>
void withAlignModel (const QuaffParams& params, Setter set) {
  const QuaffScores qs (params);
  const ScoreTables t = tables (qs);
  qg_align_model m;
  m.match_k = t.K; m.gap_k = t.G; m.match = t.match.data(); m.insert = t.insert.data();
  m.m2m = qs.m2m.data(); m.m2i = qs.m2i.data(); m.m2d = qs.m2d.data(); m.m2e = qs.m2e.data();
  m.d2d = qs.d2d; m.d2m = qs.d2m; m.i2i = qs.i2i; m.i2m = qs.i2m;
  set (&m);
}
void setAlignModel (const Gpu& g, const QuaffParams& params) { withAlignModel (params, [&] (const qg_align_model* m) { g.ok (qg_set_align_model (g.ctx, m)); }); }
void setAlignModel (const GpuPool& g, const QuaffParams& params) { withAlignModel (params, [&] (const qg_align_model* m) { g.ok (qg_pool_set_align_model (g.pool, m)); }); }

// the second half of QuaffViterbiMatrix::alignment(): gapped rows, names and coordinates from the state path
Alignment alignmentFromPath (const FastSeq& x, const FastSeq& y, uint32_t xStart, uint32_t xEnd, const uint8_t* path, uint64_t n,
                             double score, bool local) {
  // rows are written in place (one pass over the op path, no per-character appends)
  const bool q = y.hasQual();
  std::string xRow (n, Alignment::gapChar), yRow (n, Alignment::gapChar), yQual (q ? n : 0, FastSeq::maxQualityChar);
  const char* xs = x.seq.data(); const char* ys = y.seq.data(); const char* yq = q ? y.qual.data() : NULL;
  size_t i = xStart - 1, j = 0;
  for (uint64_t t = 0; t < n; ++t) {
    const uint8_t op = path[t];
    if (op != QG_OP_INSERT) xRow[t] = xs[i++];
    if (op != QG_OP_DELETE) { yRow[t] = ys[j]; if (q) yQual[t] = yq[j]; ++j; }
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
// -format sam without the gapped rows: the record Alignment::writeSam (qmodel.cpp:608-616) prints for the alignment
// alignmentFromPath would build, straight from the op path.  Coordinates go through the reference's own
// SeqIntervalCoords::compose, in the order alignmentFromPath / Alignment::revcomp / FastSeq::revcomp apply it; the CIGAR is
// the run-length of the ops (cigarString, qmodel.cpp:625-653: letter then count), reversed when the reference row is a
// reverse strand.  The reference's cigarString re-allocates the whole string per run and its revcomp() copies both rows:
// ~0.5 ms per 8 kb read, more than the GPU needs for the alignment itself.
void writeSamRecord (std::ostream& out, const FastSeq& x, const FastSeq& y, uint32_t xStart, uint32_t xEnd, const uint8_t* path, uint64_t n, double score) {
  SeqIntervalCoords s0 = SeqIntervalCoords (x.name, xStart, xEnd, false).compose (x.source);
  SeqIntervalCoords s1 = SeqIntervalCoords (y.name, 1, y.length(), false).compose (y.source);
  const bool flip = s0.rev;
  if (flip) {                                              // revcomp().writeSam (out): both gapped rows reverse-complemented
    s0 = SeqIntervalCoords ("Ref", 1, (SeqIdx) n, true).compose (s0);
    s1 = SeqIntervalCoords ("Read", 1, (SeqIdx) n, true).compose (s1);
  }
  static const char letter[3] = { 'M', 'I', 'D' };         // QG_OP_MATCH, QG_OP_INSERT (gap in the reference row), QG_OP_DELETE
  std::string cigar;
  cigar.reserve (n / 2 + 16);
  char buf[24];
  for (uint64_t t = 0; t < n; ) {
    const uint8_t op = flip ? path[n - 1 - t] : path[t];
    uint64_t e = t + 1;
    while (e < n && (flip ? path[n - 1 - e] : path[e]) == op) ++e;
    cigar += letter[op];
    const int len = snprintf (buf, sizeof (buf), "%llu", (unsigned long long) (e - t));
    cigar.append (buf, len);
    t = e;
  }
  const int flag = s1.rev ? 16 : 0;
  out << s1.name << '\t' << flag << '\t' << s0.name << '\t' << s0.start << "\t0\t" << cigar << "\t*\t0\t0\t*\t*\tAS:i:" << ((int) round (score)) << endl;
}
}  // namespace

// ---- ingest --------------------------------------------------------------------------------------------
// readFastSeqs (fastseq.cpp:148-176) with -gpu: the reference reads through kseq one character at a time from a 4 KB buffer
// (~95 MB/s: at GPU speed the loader is the slowest stage of `quaff align`).  Here the file is inflated / read in 4 MB pieces
// into memory and parsed in place by the same state machine as kseq_read (kseq/kseq.h:169-207) -- header character, name up to
// the first whitespace, comment up to the end of the line, sequence = printable characters up to the next '>', '+' or '@',
// then for FASTQ the rest of the '+' line and quality characters 33..127 until the sequence length is reached (one more
// character is consumed, as kseq does) -- so multi-line records, missing qualities and truncated files give the same FastSeq
// vector.  filepos is what gztell returned to the reference before each record: the end of the last 4 KB block kseq had read.
vguard<FastSeq> quaffGpuReadFastSeqs (const char* filename) {
  vguard<FastSeq> seqs;
  // the whole file in memory: read(2) for plain files, zlib for gzip (magic 1f 8b); no zero-filled growth
  struct Buf { char* p; size_t n, cap; Buf () : p (NULL), n (0), cap (0) { } ~Buf () { free (p); }
               void room (size_t more) { if (n + more > cap) { cap = std::max (cap * 2, n + more); p = (char*) realloc (p, cap); Require (p != NULL, "out of memory reading sequences"); } } } data;
  {
    FILE* raw = fopen (filename, "rb");
    Require (raw != NULL, "Couldn't open %s", filename);
    unsigned char magic[2] = {0, 0};
    const size_t got2 = fread (magic, 1, 2, raw);
    const bool gz = got2 == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (!gz) {
      rewind (raw);
      const size_t piece = 8u << 20;
      for (;;) { data.room (piece); const size_t got = fread (data.p + data.n, 1, piece, raw); if (got == 0) break; data.n += got; }
      fclose (raw);
    } else {
      fclose (raw);
      gzFile fp = gzopen (filename, "r");
      Require (fp != Z_NULL, "Couldn't open %s", filename);
      gzbuffer (fp, 1 << 20);
      const size_t piece = 4u << 20;
      for (;;) { data.room (piece); const int got = gzread (fp, data.p + data.n, (unsigned) piece); if (got <= 0) break; data.n += (size_t) got; }
      gzclose (fp);
    }
  }
  // character classes of kseq_read: separators of the name (isspace), sequence characters (isgraph and not a record marker)
  bool isSpace[256], isSeq[256];
  for (int c = 0; c < 256; ++c) { isSpace[c] = isspace (c) != 0; isSeq[c] = isgraph (c) != 0 && c != '>' && c != '+' && c != '@' && c < 128; }
  const char* d = data.p;
  const size_t n = data.n;
  size_t pos = 0;                                          // next unread byte
  int lastChar = 0;
  const std::string fname (filename);
  for (;;) {
    const size_t blocks = (pos + 4095) / 4096;
    const z_off_t filepos = (z_off_t) std::min<size_t> (n, blocks * 4096);
    // kseq_read
    if (lastChar == 0) {
      while (pos < n && d[pos] != '>' && d[pos] != '@') ++pos;
      if (pos >= n) break;
      lastChar = d[pos++];
    }
    if (pos >= n) break;                                   // ks_getuntil on an exhausted stream: -1
    FastSeq fs;
    int c = 0;
    {
      size_t e = pos;
      while (e < n && !isSpace[(unsigned char) d[e]]) ++e;
      fs.name.assign (d + pos, strnlen (d + pos, e - pos));                // string(ks->name.s): up to the first NUL
      if (e < n) { c = d[e]; pos = e + 1; } else { c = 0; pos = n; }
    }
    if (c != '\n') {
      size_t e = pos;
      while (e < n && d[e] != '\n') ++e;
      if (e > pos) fs.comment.assign (d + pos, strnlen (d + pos, e - pos));
      pos = e < n ? e + 1 : n;
    }
    std::string& sq = fs.seq;
    c = -1;
    while (pos < n) {
      size_t e = pos;                                      // a run of sequence characters (whole lines in practice)
      while (e < n && isSeq[(unsigned char) d[e]]) ++e;
      sq.append (d + pos, e - pos);
      pos = e;
      if (pos >= n) break;
      const int t = (unsigned char) d[pos++];
      if (t == '>' || t == '+' || t == '@') { c = t; break; }       // otherwise a separator (newline, blank, control character): skipped
    }
    if (c == '>' || c == '@') lastChar = c;
    { const size_t z = sq.find ('\0'); if (z != std::string::npos) sq.resize (z); }      // string(ks->seq.s)
    bool haveQual = false;
    if (c == '+') {
      while (pos < n && d[pos] != '\n') ++pos;
      if (pos < n) {
        ++pos;
        std::string ql;
        ql.reserve (sq.size());
        const size_t want = fs.seq.size();                 // kseq counts the raw sequence length; a NUL inside a sequence is not a case worth keeping apart
        while (pos < n && ql.size() < want) {
          size_t e = pos;                                  // a run of quality characters, at most what is still missing
          const size_t lim = std::min (n, pos + (want - ql.size()));
          while (e < lim && (unsigned char) d[e] >= 33 && (unsigned char) d[e] <= 127) ++e;
          ql.append (d + pos, e - pos);
          pos = e;
          if (ql.size() < want && pos < n) ++pos;          // a character outside 33..127 (the newline of a multi-line record): skipped
        }
        if (pos < n && ql.size() >= want) ++pos;           // kseq reads one more character before it notices the quality string is complete
        lastChar = 0;
        haveQual = ql.size() == want;
        if (haveQual) fs.qual.swap (ql);
      } else lastChar = 0;                                  // truncated after '+': kseq returns -2, the record is kept without qualities
    }
    fs.filename = fname;
    fs.filepos = filepos;
    seqs.push_back (FastSeq());
    std::swap (seqs.back(), fs);
  }
  LogThisAt(3, "Read " << plural(seqs.size(),"sequence") << " from " << filename << endl);
  if (seqs.empty())
    Warn ("Couldn't read any sequences from %s", filename);
  return seqs;
}

// ---- seam A -------------------------------------------------------------------------------------------------------
namespace {
// Chunks of reads come back from the pool's worker threads in any order: each worker builds the Alignment objects of its
// chunk and formats them (writeAlignment into a string, as the reference's own tasks do, qmodel.cpp:2796-2812) while the
// other contexts keep the GPUs busy; chunks are then written in read order.
struct AlignSink {
  const QuaffAligner& aligner; std::ostream& out; const vguard<FastSeq>& x; const vguard<FastSeq>& y; bool local; size_t chunk;
  bool toFile;                                             // -savealign: writeAlignment ignores the stream it is given; format serially
  std::mutex mx;
  std::map<size_t, std::string> text;
  std::map<size_t, std::vector<Alignment> > held;
  size_t nextFirst;
  AlignSink (const QuaffAligner& a, std::ostream& o, const vguard<FastSeq>& x, const vguard<FastSeq>& y, bool local, size_t chunk)
    : aligner (a), out (o), x (x), y (y), local (local), chunk (chunk), toFile (a.usingOutputFile()), nextFirst (0) { }
  static void onChunk (void* user, int, size_t first, size_t n, const uint32_t* best, const double* score, const uint32_t* xs, const uint32_t* xe,
                       const uint8_t* paths, const uint64_t* off) {
    AlignSink& s = *(AlignSink*) user;
    std::vector<Alignment> al;
    std::ostringstream os;
    const bool samFast = !s.toFile && s.aligner.format == QuaffAlignmentPrinter::SamAlignment && !getenv ("QUAFF_GPU_GENERIC_WRITER");
    for (size_t r = 0; r < n; ++r)
      if (best[r] != 0xFFFFFFFFu) {
        if (samFast) {
          if (score[r] >= s.aligner.logOddsThreshold)
            writeSamRecord (os, s.x[best[r]], s.y[first + r], xs[r], xe[r], paths + off[r], off[r+1] - off[r], score[r]);
          continue;
        }
        Alignment a = alignmentFromPath (s.x[best[r]], s.y[first + r], xs[r], xe[r], paths + off[r], off[r+1] - off[r], score[r], s.local);
        if (s.toFile) al.push_back (a); else s.aligner.writeAlignment (os, a);
      }
    std::lock_guard<std::mutex> lock (s.mx);
    if (s.toFile) s.held[first].swap (al); else s.text[first] = os.str();
    for (;;) {                                             // flush whatever is now contiguous
      if (s.toFile) {
        auto it = s.held.find (s.nextFirst);
        if (it == s.held.end()) break;
        for (const auto& a : it->second) s.aligner.writeAlignment (s.out, a);
        s.held.erase (it);
      } else {
        auto it = s.text.find (s.nextFirst);
        if (it == s.text.end()) break;
        s.out << it->second;
        s.text.erase (it);
      }
      s.nextFirst += s.chunk;
    }
  }
};
}  // namespace

void quaffGpuAlign (QuaffAligner& aligner, std::ostream& out, const vguard<FastSeq>& x, const vguard<FastSeq>& y,
                    const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config) {
  Trace tr;
  const Flat fx = flatten (x, false), fy = flatten (y, true);
  tr.mark ("tokenise (flatten)");
  const qg_dpconfig gc = gpuConfig (config);
  const std::vector<double> nullLL = nullLogLikes (nullModel, fy, y);
  tr.mark ("null-model log-likelihoods");
  if (!aligner.printAllAlignments) {
    // chunks of reads over every context of every device (QUAFF_GPU_CHUNK reads each, QUAFF_GPU_CONTEXTS contexts per device)
    const size_t chunk = std::max<size_t> (1, envSize ("QUAFF_GPU_CHUNK", 1536));
    const size_t nChunks = (y.size() + chunk - 1) / chunk, nDev = std::max<size_t> (1, quaffGpuDevices.size());
    const size_t perDevice = std::max<size_t> (1, std::min<size_t> (envSize ("QUAFF_GPU_CONTEXTS", 4), (nChunks + nDev - 1) / nDev));
    GpuPool g ((int) perDevice);
    tr.mark ("contexts");
    g.ok (qg_pool_set_refs (g.pool, x.size(), fx.tok.data(), fx.off.data()));
    setAlignModel (g, params);
    tr.mark ("reference set + model");
    aligner.writeAlignmentHeader (out, x, true);
    AlignSink sink (aligner, out, x, y, config.local, chunk);
    g.ok (qg_pool_align_reads (g.pool, &gc, y.size(), fy.tok.data(), fy.quals ? fy.qual.data() : NULL, fy.off.data(), nullLL.data(), chunk,
                               AlignSink::onChunk, &sink));
    tr.mark ("align + format + write");
    return;
  }
  Gpu g;
  g.ok (qg_set_seqs (g.ctx, QG_REFS, x.size(), fx.tok.data(), NULL, fx.off.data()));
  g.ok (qg_set_seqs (g.ctx, QG_READS, y.size(), fy.tok.data(), fy.quals ? fy.qual.data() : NULL, fy.off.data()));
  setAlignModel (g, params);
  aligner.writeAlignmentHeader (out, x, true);
  {
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
  // the scheduler's pair list over every device of -gpu (QUAFF_GPU_OVERLAP_CONTEXTS contexts on each, default 1)
  GpuPool g ((int) std::max<size_t> (1, envSize ("QUAFF_GPU_OVERLAP_CONTEXTS", 1)));
  const Flat f = flatten (seqs, true);
  const QuaffScores qs (params);
  const ScoreTables t = tables (qs);
  qg_overlap_model om;
  om.match_k = t.K; om.gap_k = t.G; om.match = t.match.data(); om.insert = t.insert.data();
  for (int r = 0; r < 4; ++r) om.log_ref_base[r] = log (params.refBase[r]);
  om.begin_insert = params.beginInsert.data(); om.begin_delete = params.beginDelete.data();
  om.extend_insert = params.extendInsert; om.extend_delete = params.extendDelete;
  const qg_dpconfig gc = gpuConfig (config);
  // stored reverse strands: nullModel.logLikelihood (y.revcomp()) of the reference = the original read's
  std::vector<double> nullLL (seqs.size());
  for (size_t n = 0; n < seqs.size(); ++n)
    nullLL[n] = nullModel.logLikelihood (n >= nOriginals ? seqs[n].revcomp() : seqs[n]);
  size_t np = 0; uint32_t *xi = NULL, *yi = NULL, *co = NULL; double* score = NULL; uint8_t* path = NULL; uint64_t* off = NULL;
  g.ok (qg_pool_overlap_reads (g.pool, &gc, &om, seqs.size(), f.tok.data(), f.quals ? f.qual.data() : NULL, f.off.data(), nOriginals, nullLL.data(),
                               &np, &xi, &yi, &score, &co, &path, &off));
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
  // one context per device; the reads split into contiguous ranges, the partial counts are summed by the library
  GpuPool g (1);
  const Flat fx = flatten (x, false), fy = flatten (y, true);
  Require (fy.quals, "Forward-Backward algorithm requires quality scores to fit model");
  g.ok (qg_pool_set_refs (g.pool, x.size(), fx.tok.data(), fx.off.data()));
  setAlignModel (g, params);
  if (getenv ("QUAFF_GPU_EXACT")) g.ok (qg_pool_set_option (g.pool, QG_OPT_FB_EXACT, 1));
  const qg_dpconfig gc = gpuConfig (config);
  const size_t nx = x.size(), ny = y.size();
  std::vector<double> yLL (ny);
  const std::vector<double> nullLL = nullLogLikes (nullModel, fy, y);
  std::vector<uint32_t> so (nx * ny, 0), soLen (ny);
  for (size_t n = 0; n < ny; ++n) {
    soLen[n] = (uint32_t) sortOrder[n].size();
    for (size_t s = 0; s < sortOrder[n].size(); ++s) so[n * nx + s] = (uint32_t) sortOrder[n][s];
  }
  const unsigned int K = params.matchContext.kmerLen, G = params.indelContext.kmerLen;
  std::vector<double> flat (qg_counts_size (K, G));
  g.ok (qg_pool_estep (g.pool, &gc, trainer.allowNullModel ? 1 : 0, nx, ny, fy.tok.data(), fy.qual.data(), fy.off.data(), nullLL.data(),
                       so.data(), soLen.data(), yLL.data(), flat.data(), &logLike));
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
