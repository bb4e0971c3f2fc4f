// CudaGlobalAlignBatch -- batch replacement for muscle::GlobalAlign (MU/glbalign.cpp:69-81 -> NWSmall,
// MU/nwsmall.cpp:500-670, + BitTraceBack MU/bittraceback.cpp:138-) over the inter-anchor ranges of one window.
//
// AnchoredProfileProfile (MU/anchoredpp.cpp:501-549) calls ProfileProfile -> AlignTwoProfs -> GlobalAlign once per
// range; the batch seam collects the (ProfPos*, length) pairs of all ranges and submits them in one call.  For the
// two-genome path both profiles hold one ungapped ACGT sequence, the case the integer-exact device kernel covers
// (SURVEY.md 8a-13); with `wildcards` set, ranges with N / X columns go to the float kernel (mcu_nw_batch_wild); a range that is of
// neither form is reported through `handled[i] == false` and the caller runs the reference GlobalAlign for it.  Paths are the reference's PWPath edge lists (MU/pwpath.h:46-51).
#ifndef CUDA_GLOBAL_ALIGN_H_
#define CUDA_GLOBAL_ALIGN_H_

#include <stdexcept>
#include <string>
#include <vector>

#ifndef MUSCLE_LONG_VERSION   /* muscle.h has no include guard of its own */
#include "libMUSCLE/muscle.h"
#endif
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"
#include "libMUSCLE/alpha.h"
#include "libMUSCLE/params.h"
#include "mauve_cuda.h"

namespace muscle {

struct CudaDPRange {
	const ProfPos* PA; unsigned uLengthA;
	const ProfPos* PB; unsigned uLengthB;
};

namespace cuda_detail {

// the single letter of a one-sequence profile column, or 0 when the column is not of that form
inline char SingleLetter(const ProfPos& pp)
{
	if (pp.m_bAllGaps) return 0;
	const unsigned u = pp.m_uSortOrder[0];
	if (u >= 4 || pp.m_fcCounts[u] != 1.0f) return 0;   // NX_A, NX_C, NX_G, NX_T = 0..3 (MU/alpha.h:49-52)
	return "ACGT"[u];
}

inline bool ProfileToString(const ProfPos* P, unsigned n, std::string& out)
{
	out.resize(n);
	for (unsigned i = 0; i < n; ++i) {
		const char c = SingleLetter(P[i]);
		if (!c) return false;
		out[i] = c;
	}
	return n > 0;
}

// the letter of a one-sequence profile column that may be a DNA wildcard: A C G T, 'X' (MSA::GetFractionalWeightedCounts gives it G and A
// halves, MU/msa2.cpp:71-79 with NX_X == AX_R) or 'N' for every other wildcard (1/20 for each letter); 0 for any other column
inline char SingleLetterOrWildcard(const ProfPos& pp)
{
	const char c = SingleLetter(pp);
	if (c || pp.m_bAllGaps) return c;
	const FCOUNT* f = pp.m_fcCounts;
	const FCOUNT n = (FCOUNT)1.0 / 20;
	if (f[0] == n && f[1] == n && f[2] == n && f[3] == n) return 'N';
	if (f[0] == (FCOUNT)0.5 && f[2] == (FCOUNT)0.5 && f[1] == 0 && f[3] == 0) return 'X';
	return 0;
}

// false when a column is neither; *has_wildcard tells which device entry takes the string
inline bool ProfileToStringWild(const ProfPos* P, unsigned n, std::string& out, bool* has_wildcard)
{
	out.resize(n);
	*has_wildcard = false;
	for (unsigned i = 0; i < n; ++i) {
		const char c = SingleLetterOrWildcard(P[i]);
		if (!c) return false;
		if (c == 'N' || c == 'X') *has_wildcard = true;
		out[i] = c;
	}
	return n > 0;
}

// The device kernels have the reference's DEFAULT DNA scoring built in: substitution NUC_SP (MU/nucmx.cpp:8-25 with the +60 centre),
// gap open = close = -400 / 2 per column, terminal gaps free, gap extend 0 (MU/params.cpp:296-313; libMems calls MUSCLE with these:
// LM/MuscleInterface.cpp:1086-1106).  --muscle-args or a caller of CallMuscleFast can change them; a range whose profiles were built
// with anything else must stay with the reference's ProfileProfile.  Checked on what the profile columns themselves carry.
inline bool DefaultScoring(const ProfPos* P, unsigned n)
{
	static const SCORE nuc[4][4] = {{151, -54, 29, -63}, {-54, 160, -65, 29}, {29, -65, 160, -54}, {-63, 29, -54, 151}};
	if (g_scoreGapExtend.get() != 0 || g_PPScore.get() != PPSCORE_SPN) return false;
	for (unsigned i = 0; i < n; ++i) {
		const ProfPos& pp = P[i];
		if (pp.m_scoreGapOpen != (SCORE)-200 && !(i == 0 && pp.m_scoreGapOpen == 0)) return false;
		if (pp.m_scoreGapClose != (SCORE)-200 && !(i + 1 == n && pp.m_scoreGapClose == 0)) return false;
		const unsigned u = pp.m_uSortOrder[0];
		if (u < 4 && pp.m_fcCounts[u] == 1.0f)   // a plain letter: its AAScores row is the substitution matrix row
			for (unsigned j = 0; j < 4; ++j)
				if (pp.m_AAScores[j] != nuc[u][j]) return false;
	}
	return true;
}

// edge string -> PWPath (PWEdge prefix lengths count the letters consumed including this edge)
inline void EdgesToPath(const char* e, uint32_t len, PWPath& P)
{
	unsigned ua = 0, ub = 0;
	for (uint32_t j = 0; j < len; ++j) {
		if (e[j] != 'I') ++ua;
		if (e[j] != 'D') ++ub;
		P.AppendEdge(e[j], ua, ub);
	}
}

// The same check on the GLOBALS the profile columns are derived from (ProfileFromMSA, MU/profilefrommsa.cpp:283-297: a column of a
// one-sequence alignment without gaps gets gap open = close = g_scoreGapOpen / 2 and the substitution matrix row of its letter), for
// callers that hold the sequences themselves and never build the profiles.
inline bool DefaultScoringGlobals()
{
	static const SCORE nuc[4][4] = {{151, -54, 29, -63}, {-54, 160, -65, 29}, {29, -65, 160, -54}, {-63, 29, -54, 151}};
	if (g_scoreGapExtend.get() != 0 || g_PPScore.get() != PPSCORE_SPN || g_scoreGapOpen.get() != (SCORE)-400) return false;
	if (g_Alpha.get() != ALPHA_DNA) return false;
	for (unsigned i = 0; i < 4; ++i)
		for (unsigned j = 0; j < 4; ++j)
			if ((*g_ptrScoreMatrix.get())[i][j] != nuc[i][j]) return false;
	return true;
}

// the device letter of a character of a one-sequence DNA alignment row (MU/alpha.cpp:123-141, case folded as CharToLetterEx does):
// A C G T, 'X', 'N' for the other wildcards (same classes as SingleLetterOrWildcard); 0 for anything else
inline char DeviceLetterOfChar(char ch)
{
	switch (ch) {
	case 'A': case 'a': return 'A';
	case 'C': case 'c': return 'C';
	case 'G': case 'g': return 'G';
	case 'T': case 't': return 'T';
	case 'X': case 'x': return 'X';
	case 'M': case 'R': case 'W': case 'S': case 'Y': case 'K': case 'V': case 'H': case 'D': case 'B': case 'N':
	case 'm': case 'r': case 'w': case 's': case 'y': case 'k': case 'v': case 'h': case 'd': case 'b': case 'n': return 'N';
	default: return 0;
	}
}

// row 0 of a one-sequence alignment without gap columns as a device string; false when a character is outside the DNA alphabet
inline bool MsaRowToString(const MSA& msa, std::string& out, bool* has_wildcard)
{
	const unsigned n = msa.GetColCount();
	out.resize(n);
	*has_wildcard = false;
	for (unsigned i = 0; i < n; ++i) {
		const char c = DeviceLetterOfChar(msa.GetChar(0, i));
		if (!c) return false;
		if (c == 'N' || c == 'X') *has_wildcard = true;
		out[i] = c;
	}
	return n > 0;
}

}  // namespace cuda_detail

// what the two device entries take: the ranges whose columns are all A/C/G/T (integer wavefront kernels, mcu_nw_batch) and the ranges
// with N / X columns (float wavefront kernel with the reference's arithmetic, mcu_nw_batch_wild)
struct CudaDPGroups {
	struct Group {
		std::string a, b;
		std::vector<uint64_t> a_off, b_off, p_off;
		std::vector<size_t> index;
		Group() : a_off(1, 0), b_off(1, 0), p_off(1, 0) {}
	} g[2];
	void Add(size_t index, const std::string& sa, const std::string& sb, bool wildcard)
	{
		Group& x = g[wildcard ? 1 : 0];
		x.a += sa; x.b += sb;
		x.a_off.push_back(x.a.size()); x.b_off.push_back(x.b.size()); x.p_off.push_back(x.p_off.back() + sa.size() + sb.size());
		x.index.push_back(index);
	}
	void Run(PWPath* paths, std::vector<bool>& handled, std::vector<long long>* scores)
	{
		for (int k = 0; k < 2; ++k) {
			const size_t m = g[k].index.size();
			if (!m) continue;
			std::vector<char> path(g[k].p_off.back());
			std::vector<uint32_t> plen(m);
			std::vector<int64_t> score(m);
			std::vector<float> fscore(m);
			const int rc = k == 0
				? mcu_nw_batch(m, g[k].a.data(), &g[k].a_off[0], g[k].b.data(), &g[k].b_off[0], &g[k].p_off[0], &path[0], &plen[0], &score[0], NULL)
				: mcu_nw_batch_wild(m, g[k].a.data(), &g[k].a_off[0], g[k].b.data(), &g[k].b_off[0], &g[k].p_off[0], &path[0], &plen[0], &fscore[0], NULL);
			if (rc != MCU_OK) throw std::runtime_error(std::string("CudaGlobalAlignBatch: ") + mcu_last_error());
			for (size_t j = 0; j < m; ++j) {
				cuda_detail::EdgesToPath(&path[g[k].p_off[j]], plen[j], paths[g[k].index[j]]);
				handled[g[k].index[j]] = true;
				if (scores) (*scores)[g[k].index[j]] = k == 0 ? score[j] : (long long)fscore[j];
			}
		}
	}
};

// Aligns every range on the device; paths (ranges.size() caller-owned objects: PWPath is not copyable) is filled for
// handled[i] == true.
inline void CudaGlobalAlignBatch(const std::vector<CudaDPRange>& ranges, PWPath* paths, std::vector<bool>& handled,
                                 std::vector<long long>* scores = NULL, bool wildcards = true)
{
	const size_t n = ranges.size();
	for (size_t i = 0; i < n; ++i) paths[i].Clear();
	handled.assign(n, false);
	if (scores) scores->assign(n, 0);
	CudaDPGroups groups;
	std::string sa, sb;
	for (size_t i = 0; i < n; ++i) {
		bool wa = false, wb = false;
		if (!cuda_detail::ProfileToStringWild(ranges[i].PA, ranges[i].uLengthA, sa, &wa) || !cuda_detail::ProfileToStringWild(ranges[i].PB, ranges[i].uLengthB, sb, &wb))
			continue;
		if ((wa || wb) && !wildcards) continue;
		if (!cuda_detail::DefaultScoring(ranges[i].PA, ranges[i].uLengthA) || !cuda_detail::DefaultScoring(ranges[i].PB, ranges[i].uLengthB)) continue;
		groups.Add(i, sa, sb, wa || wb);
	}
	groups.Run(paths, handled, scores);
}

// The same for pairs of ONE-SEQUENCE alignments without gap columns (what AnchoredProfileProfile's ranges are for two genomes): the
// letters are read from the alignment rows and the profiles (ProfileFromMSA: counts, SortCounts, score rows per column -- 1.4 s of the
// MDS42 run's host time) are never built.  A pair is left to the caller (handled[i] == false) when a row holds a character outside the
// DNA alphabet or MUSCLE's scoring globals are not the defaults the kernels have built in.
inline void CudaGlobalAlignBatchRows(const std::vector<std::pair<const MSA*, const MSA*> >& pairs, PWPath* paths, std::vector<bool>& handled,
                                     bool wildcards = true)
{
	const size_t n = pairs.size();
	for (size_t i = 0; i < n; ++i) paths[i].Clear();
	handled.assign(n, false);
	if (!cuda_detail::DefaultScoringGlobals()) return;
	CudaDPGroups groups;
	std::string sa, sb;
	for (size_t i = 0; i < n; ++i) {
		bool wa = false, wb = false;
		if (pairs[i].first->GetSeqCount() != 1 || pairs[i].second->GetSeqCount() != 1) continue;
		if (!cuda_detail::MsaRowToString(*pairs[i].first, sa, &wa) || !cuda_detail::MsaRowToString(*pairs[i].second, sb, &wb)) continue;
		if ((wa || wb) && !wildcards) continue;
		groups.Add(i, sa, sb, wa || wb);
	}
	groups.Run(paths, handled, NULL);
}

}  // namespace muscle

#endif
