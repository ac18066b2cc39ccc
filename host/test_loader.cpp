// TEST INFRASTRUCTURE: the -gpu loader (quaffGpuReadFastSeqs) against the reference's readFastSeqs on the files given on the
// command line: every field of every record must be equal.  Built by `make -C host loader-test` from the staged reference objects.
#include <iostream>
#include "fastseq.h"
#include "quaff_gpu_seams.h"
int main (int argc, char** argv) {
  int bad = 0;
  for (int a = 1; a < argc; ++a) {
    quaffGpuDevice = -1;
    const vguard<FastSeq> r = readFastSeqs (argv[a]);
    const vguard<FastSeq> g = quaffGpuReadFastSeqs (argv[a]);
    bool same = r.size() == g.size();
    for (size_t n = 0; same && n < r.size(); ++n)
      same = r[n].name == g[n].name && r[n].comment == g[n].comment && r[n].seq == g[n].seq && r[n].qual == g[n].qual
          && r[n].filename == g[n].filename && r[n].filepos == g[n].filepos;
    std::cout << argv[a] << ": " << r.size() << " records " << (same ? "identical" : "DIFFERENT") << std::endl;
    if (!same) ++bad;
  }
  return bad ? 1 : 0;
}
