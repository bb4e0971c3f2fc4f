// Link-time seam: mems::EliminateOverlaps_v2<MatchList>(ml, eliminate_both) and mems::IdentifyBreakpoints<MatchList> on the GPU.
//
// Both are function templates (LM/ProgressiveAligner.h:396-406, LM/GreedyBreakpointElimination.h:161-226) whose instantiations for
// MatchList are weak symbols of ProgressiveAligner.o; the explicit specializations below -- strong definitions of the same symbols --
// are the ones the linker keeps (the technique of sol_seam.cpp).  A list of two-genome matches goes through mcu_eliminate_overlaps /
// mcu_lcbs (csrc/lcb.cu, results identical to the reference's incl. where starts tie); any other list (three genomes, undefined
// coordinates) takes the reference's own code: the three-argument overload of EliminateOverlaps_v2, which holds the algorithm and is
// not replaced, and for IdentifyBreakpoints its template body instantiated for a forwarding vector type.
// MAUVE_CUDA_LCB_SEAM=0 leaves everything with the reference's code.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "libMems/MatchList.h"
#include "libMems/ProgressiveAligner.h"
#include "libMems/GreedyBreakpointElimination.h"

#include "CudaLcb.h"

namespace mems {

namespace {
unsigned long long g_eo_device = 0, g_eo_reference = 0, g_eo_ties = 0, g_bp_device = 0, g_bp_reference = 0;
struct LcbSeamReport {
	~LcbSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT")) {
			fprintf(stderr, "EliminateOverlaps_v2 seam: %llu lists on the device (%llu tied keys ordered like std::sort), %llu in the reference's code\n",
			        g_eo_device, g_eo_ties, g_eo_reference);
			fprintf(stderr, "IdentifyBreakpoints seam: %llu lists on the device, %llu in the reference's code\n", g_bp_device, g_bp_reference);
		}
	}
};
LcbSeamReport g_lcb_report;
bool seam_on()
{
	static const bool off = getenv("MAUVE_CUDA_LCB_SEAM") && getenv("MAUVE_CUDA_LCB_SEAM")[0] == '0';
	return !off;
}
// a second vector type over the same pointers: lets the reference's template body run under another instantiation
struct MatchPtrVector : public std::vector<Match*> {};
}  // namespace

template <>
void EliminateOverlaps_v2<MatchList>(MatchList& ml, bool eliminate_both)
{
	if (ml.size() < 2) return;
	uint64_t ties = 0;
	if (seam_on() && CudaEliminateOverlaps(ml, eliminate_both, 0, &ties)) {
		++g_eo_device;
		g_eo_ties += ties;
		return;
	}
	++g_eo_reference;
	uint seq_count = ml[0]->SeqCount();
	std::vector<uint> seq_ids(seq_count);
	for (uint i = 0; i < seq_count; ++i) seq_ids[i] = i;
	EliminateOverlaps_v2(ml, seq_ids, eliminate_both);
}

template <>
void IdentifyBreakpoints<MatchList>(MatchList& mlist, std::vector<gnSeqI>& breakpoints)
{
	if (mlist.size() == 0) return;
	if (seam_on() && CudaIdentifyBreakpoints(mlist, breakpoints)) {
		++g_bp_device;
		return;
	}
	++g_bp_reference;
	MatchPtrVector v;
	v.assign(mlist.begin(), mlist.end());
	IdentifyBreakpoints(v, breakpoints);
	for (size_t i = 0; i < v.size(); ++i) mlist[i] = v[i];
}

}  // namespace mems
