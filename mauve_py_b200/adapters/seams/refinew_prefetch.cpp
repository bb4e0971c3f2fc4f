// Link-time seam: the gapped DP of muscle::RefineW's windows, prefetched on the GPU in ONE call.
//
// MuscleInterface::RefineFast(ga, 200|500) (LM/MuscleInterface.cpp:823-894, called per LCB from LM/ProgressiveAligner.cpp:1207-1211)
// cuts the alignment into windows and realigns every window from scratch with MUSCLE() (MU/refinew.cpp:76-190); for two
// genomes that is one progressive step = ONE GlobalAlign per window (MU/progalign.cpp:106 -> MU/aligntwoprofs.cpp:23), and the
// windows do not depend on each other: 19,967 of the 61,773 GlobalAlign calls of the MDS42 run.  Re-stating MUSCLE()'s
// bookkeeping (guide tree orientation, estrings, MakeRootMSA) would be fragile, so nothing of it is touched:
//   * RefineW (this file, under the original name; the reference's body is kept as RefineW_reference in a copy of refinew.o)
//     first extracts the ungapped letters of every window -- what SeqVectFromMSACols (MU/refinew.cpp:49-74) hands to MUSCLE() --
//     and aligns all of them, in both operand orders, in one mcu_nw_batch call; the paths go into a cache keyed by the two
//     letter strings; then it runs the reference's RefineW unchanged;
//   * GlobalAlign (this file, under the original name; original kept as GlobalAlign_reference) answers from that cache when both
//     profiles are single ungapped ACGT sequences whose strings are in it (after the SetTermGaps calls NWSmall would have made,
//     MU/nwsmall.cpp:506-507), and otherwise runs the reference's NWSmall.
// So the reference's control flow is executed as it is and only the origin of the path changes; a cache miss costs nothing but
// the CPU time it always cost.  MAUVE_CUDA_REFINE_SEAM=0 switches the prefetch off.
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>
#include <unordered_map>
#include <vector>

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"

#include "CudaGlobalAlign.h"

namespace muscle {

SCORE GlobalAlign_reference(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
    asm("_ZN6muscle21GlobalAlign_referenceEPKNS_7ProfPosEjS2_jRNS_6PWPathE");
void RefineW_reference(const MSA& msaIn, MSA& msaOut) asm("_ZN6muscle17RefineW_referenceERKNS_3MSAERS0_");

static std::unordered_map<std::string, std::string> g_paths;   // "A|B" -> PWPath edge types, first to last
static unsigned long long g_rw_calls = 0, g_rw_prefetched = 0, g_ga_hits = 0, g_ga_misses = 0;
struct RefineSeamReport {
	~RefineSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
			fprintf(stderr, "RefineW seam: %llu calls, %llu DP problems prefetched on the device, %llu GlobalAlign calls answered from them, %llu not\n",
			        g_rw_calls, g_rw_prefetched, g_ga_hits, g_ga_misses);
	}
};
static RefineSeamReport g_refine_report;

SCORE GlobalAlign(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
{
	if (!g_paths.empty()) {
		std::string a, b;
		if (cuda_detail::ProfileToString(PA, uLengthA, a) && cuda_detail::ProfileToString(PB, uLengthB, b)) {
			std::unordered_map<std::string, std::string>::const_iterator it = g_paths.find(a + '|' + b);
			if (it != g_paths.end()) {
				SetTermGaps(PA, uLengthA);   // the side effect NWSmall has on the caller's profiles (MU/nwsmall.cpp:506-507)
				SetTermGaps(PB, uLengthB);
				Path.Clear();
				unsigned ua = 0, ub = 0;
				const std::string& e = it->second;
				for (size_t j = 0; j < e.size(); ++j) {
					if (e[j] != 'I') ++ua;
					if (e[j] != 'D') ++ub;
					Path.AppendEdge(e[j], ua, ub);
				}
				++g_ga_hits;
				return 0;   // NWSmall's own return value is literally 0 (MU/nwsmall.cpp:669)
			}
		}
		++g_ga_misses;
	}
	return GlobalAlign_reference(PA, uLengthA, PB, uLengthB, Path);
}

// ungapped upper-case letters of row `uSeqIndex` in columns [uColFrom, uColTo]; false when a letter is not A, C, G or T
static bool WindowLetters(const MSA& msa, unsigned uSeqIndex, unsigned uColFrom, unsigned uColTo, std::string& s)
{
	s.clear();
	for (unsigned uColIndex = uColFrom; uColIndex <= uColTo; ++uColIndex) {
		char c = msa.GetChar(uSeqIndex, uColIndex);
		if (IsGapChar(c)) continue;
		if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
		if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
		s += c;
	}
	return true;
}

void RefineW(const MSA& msaIn, MSA& msaOut)
{
	static const bool off = getenv("MAUVE_CUDA_REFINE_SEAM") && getenv("MAUVE_CUDA_REFINE_SEAM")[0] == '0';
	++g_rw_calls;
	g_paths.clear();
	const unsigned uColCount = msaIn.GetColCount();
	if (!off && msaIn.GetSeqCount() == 2 && uColCount > 0 && g_uRefineWindow.get() > 0) {
		// the window bounds of MU/refinew.cpp:91-113 (g_uWindowTo == 0 means "to the last window")
		const unsigned uWindowCount = (uColCount + g_uRefineWindow.get() - 1) / g_uRefineWindow.get();
		const unsigned uWindowTo = 0 == g_uWindowTo.get() ? uWindowCount - 1 : g_uWindowTo.get();
		std::vector<std::string> keys;
		std::string a, b, s0, s1;
		std::vector<uint64_t> a_off(1, 0), b_off(1, 0), p_off(1, 0);
		for (unsigned uWindowIndex = g_uWindowFrom.get(); uWindowIndex <= uWindowTo; ++uWindowIndex) {
			const unsigned uColFrom = g_uWindowOffset.get() + uWindowIndex * g_uRefineWindow.get();
			if (uColFrom >= uColCount) break;
			unsigned uColTo = uColFrom + g_uRefineWindow.get() - 1;
			if (uColTo >= uColCount) uColTo = uColCount - 1;
			if (!WindowLetters(msaIn, 0, uColFrom, uColTo, s0) || !WindowLetters(msaIn, 1, uColFrom, uColTo, s1)) continue;
			if (s0.empty() || s1.empty()) continue;   // MUSCLE() is not called for a window with one empty row (MU/refinew.cpp:141-142)
			for (int order = 0; order < 2; ++order) {   // the guide tree decides which sequence is operand A: both orders are prepared
				const std::string& x = order ? s1 : s0;
				const std::string& y = order ? s0 : s1;
				const std::string key = x + '|' + y;
				if (g_paths.find(key) != g_paths.end()) continue;
				g_paths[key] = std::string();
				keys.push_back(key);
				a += x;
				b += y;
				a_off.push_back(a.size());
				b_off.push_back(b.size());
				p_off.push_back(p_off.back() + x.size() + y.size());
			}
		}
		const size_t m = keys.size();
		if (m) {
			std::vector<char> path(p_off.back());
			std::vector<uint32_t> plen(m);
			std::vector<int64_t> score(m);
			const int rc = mcu_nw_batch(m, a.data(), &a_off[0], b.data(), &b_off[0], &p_off[0], &path[0], &plen[0], &score[0], NULL);
			if (rc != MCU_OK) {   // RefineFast's callers would carry on with a half-refined alignment: stop, like MUSCLE's Quit()
				fprintf(stderr, "\n*** FATAL: RefineW prefetch: %s\n", mcu_last_error());
				exit(3);
			}
			for (size_t k = 0; k < m; ++k) g_paths[keys[k]].assign(&path[p_off[k]], plen[k]);
			g_rw_prefetched += m;
		}
	}
	RefineW_reference(msaIn, msaOut);
	g_paths.clear();
}

}  // namespace muscle
