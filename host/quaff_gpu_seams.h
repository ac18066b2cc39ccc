// Host side of the drop-in: what a quaff maintainer adds to call libquaffgpu from the reference's three seams
// (INTEGRATION.md).  This header is force-included (-include) into the reference's translation units
// by host/Makefile; the bodies are in quaff_gpu_seams.cpp.  Written for this repository against the reference's public types; the only
// part that mirrors reference statements is the field-by-field fill of its own `Alignment` type (qmodel.cpp:1624-1645,
// qoverlap.cpp:269-289), which any caller of that type has to write.
#ifndef QUAFF_GPU_SEAMS_INCLUDED
#define QUAFF_GPU_SEAMS_INCLUDED
#ifdef __cplusplus
#include <deque>
#include <string>
#include <iostream>

struct QuaffAligner; struct QuaffOverlapAligner; struct QuaffTrainer; struct QuaffParams; struct QuaffNullParams;
struct QuaffDPConfig; struct QuaffParamCounts; struct FastSeq;
template<typename T> class vguard;

extern int quaffGpuDevice;                       // -1 = CPU path (default); >= 0 = CUDA device used by the three seams
bool quaffGpuParseArg (std::deque<std::string>& argvec);     // consumes "-gpu [device | device,device,... | all]"

void quaffGpuWarmStart ();
vguard<FastSeq> quaffGpuReadFastSeqs (const char* filename);   // readFastSeqs (fastseq.cpp:148) at memory speed, same records
void quaffGpuAlign (QuaffAligner& aligner, std::ostream& out, const vguard<FastSeq>& x, const vguard<FastSeq>& y,
                    const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config);
void quaffGpuOverlap (QuaffOverlapAligner& aligner, std::ostream& out, const vguard<FastSeq>& seqs, size_t nOriginals,
                      const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config);
QuaffParamCounts quaffGpuGetCounts (QuaffTrainer& trainer, const vguard<FastSeq>& x, const vguard<FastSeq>& y,
                                    const QuaffParams& params, const QuaffNullParams& nullModel, QuaffDPConfig& config,
                                    vguard<vguard<size_t> >& sortOrder, double& logLike);
#endif
#endif
