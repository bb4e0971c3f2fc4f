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
// the CPU time it always cost.  MAUVE_CUDA_REFINE_SEAM=0 switches the prefetch off; MAUVE_CUDA_WILD=1 also prefetches windows with
// DNA wildcard letters (N, X, ...) through mcu_nw_batch_wild, the reference's float arithmetic on the device.
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
		bool wa = false, wb = false;
		if (cuda_detail::ProfileToStringWild(PA, uLengthA, a, &wa) && cuda_detail::ProfileToStringWild(PB, uLengthB, b, &wb)) {
			std::unordered_map<std::string, std::string>::const_iterator it = g_paths.find(a + '|' + b);
			if (it != g_paths.end()) {
				SetTermGaps(PA, uLengthA);   // the side effect NWSmall has on the caller's profiles (MU/nwsmall.cpp:506-507)
				SetTermGaps(PB, uLengthB);
				Path.Clear();
				cuda_detail::EdgesToPath(it->second.data(), (uint32_t)it->second.size(), Path);
				++g_ga_hits;
				return 0;   // NWSmall's own return value is literally 0 (MU/nwsmall.cpp:669)
			}
		}
		++g_ga_misses;
	}
	return GlobalAlign_reference(PA, uLengthA, PB, uLengthB, Path);
}

// ungapped upper-case letters of row `uSeqIndex` in columns [uColFrom, uColTo].  Wildcards are written as the two profile classes
// they fall into ('X' -> 'X', every other DNA wildcard -> 'N': what cuda_detail::ProfileToStringWild reads back from a profile);
// false when a byte is neither a letter nor a wildcard, or when a wildcard is met and `wild` is off.
static bool WindowLetters(const MSA& msa, unsigned uSeqIndex, unsigned uColFrom, unsigned uColTo, std::string& s, bool wild, bool* has_wildcard)
{
	s.clear();
	for (unsigned uColIndex = uColFrom; uColIndex <= uColTo; ++uColIndex) {
		char c = msa.GetChar(uSeqIndex, uColIndex);
		if (IsGapChar(c)) continue;
		if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
		if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
			if (!wild) return false;
			switch (c) {
			case 'X': break;
			case 'M': case 'R': case 'W': case 'S': case 'Y': case 'K': case 'V': case 'H': case 'D': case 'B': case 'N': c = 'N'; break;
			default: return false;
			}
			*has_wildcard = true;
		}
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
		// group 0: all letters A/C/G/T (mcu_nw_batch); group 1 (MAUVE_CUDA_WILD=1): windows with wildcard letters (mcu_nw_batch_wild)
		static const bool wild = !(getenv("MAUVE_CUDA_WILD") && getenv("MAUVE_CUDA_WILD")[0] == '0');
		struct Group {
			std::vector<std::string> keys;
			std::string a, b;
			std::vector<uint64_t> a_off, b_off, p_off;
			Group() : a_off(1, 0), b_off(1, 0), p_off(1, 0) {}
		} grp[2];
		std::string s0, s1;
		for (unsigned uWindowIndex = g_uWindowFrom.get(); uWindowIndex <= uWindowTo; ++uWindowIndex) {
			const unsigned uColFrom = g_uWindowOffset.get() + uWindowIndex * g_uRefineWindow.get();
			if (uColFrom >= uColCount) break;
			unsigned uColTo = uColFrom + g_uRefineWindow.get() - 1;
			if (uColTo >= uColCount) uColTo = uColCount - 1;
			bool has_wildcard = false;
			if (!WindowLetters(msaIn, 0, uColFrom, uColTo, s0, wild, &has_wildcard) || !WindowLetters(msaIn, 1, uColFrom, uColTo, s1, wild, &has_wildcard)) continue;
			if (s0.empty() || s1.empty()) continue;   // MUSCLE() is not called for a window with one empty row (MU/refinew.cpp:141-142)
			Group& g = grp[has_wildcard ? 1 : 0];
			for (int order = 0; order < 2; ++order) {   // the guide tree decides which sequence is operand A: both orders are prepared
				const std::string& x = order ? s1 : s0;
				const std::string& y = order ? s0 : s1;
				const std::string key = x + '|' + y;
				if (g_paths.find(key) != g_paths.end()) continue;
				g_paths[key] = std::string();
				g.keys.push_back(key);
				g.a += x;
				g.b += y;
				g.a_off.push_back(g.a.size());
				g.b_off.push_back(g.b.size());
				g.p_off.push_back(g.p_off.back() + x.size() + y.size());
			}
		}
		for (int k = 0; k < 2; ++k) {
			Group& g = grp[k];
			const size_t m = g.keys.size();
			if (!m) continue;
			std::vector<char> path(g.p_off.back());
			std::vector<uint32_t> plen(m);
			std::vector<int64_t> score(m);
			std::vector<float> fscore(m);
			const int rc = k == 0 ? mcu_nw_batch(m, g.a.data(), &g.a_off[0], g.b.data(), &g.b_off[0], &g.p_off[0], &path[0], &plen[0], &score[0], NULL)
			                      : mcu_nw_batch_wild(m, g.a.data(), &g.a_off[0], g.b.data(), &g.b_off[0], &g.p_off[0], &path[0], &plen[0], &fscore[0], NULL);
			if (rc != MCU_OK) {   // RefineFast's callers would carry on with a half-refined alignment: stop, like MUSCLE's Quit()
				fprintf(stderr, "\n*** FATAL: RefineW prefetch: %s\n", mcu_last_error());
				exit(3);
			}
			for (size_t j = 0; j < m; ++j) g_paths[g.keys[j]].assign(&path[g.p_off[j]], plen[j]);
			g_rw_prefetched += m;
		}
	}
	RefineW_reference(msaIn, msaOut);
	g_paths.clear();
}

}  // namespace muscle
