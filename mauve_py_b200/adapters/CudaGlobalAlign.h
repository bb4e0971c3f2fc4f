// CudaGlobalAlignBatch -- batch replacement for muscle::GlobalAlign (MU/glbalign.cpp:69-81 -> NWSmall,
// MU/nwsmall.cpp:500-670, + BitTraceBack MU/bittraceback.cpp:138-) over the inter-anchor ranges of one window.
//
// AnchoredProfileProfile (MU/anchoredpp.cpp:501-549) calls ProfileProfile -> AlignTwoProfs -> GlobalAlign once per
// range; the batch seam collects the (ProfPos*, length) pairs of all ranges and submits them in one call.  For the
// two-genome path both profiles hold one ungapped ACGT sequence, the case the integer-exact device kernel covers
// (SURVEY.md 8a-13); a range that is not of that form is reported through `handled[i] == false` and the caller runs
// the reference GlobalAlign for it.  Paths are the reference's PWPath edge lists (MU/pwpath.h:46-51).
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

}  // namespace cuda_detail

// Aligns every range on the device; paths (ranges.size() caller-owned objects: PWPath is not copyable) is filled for
// handled[i] == true.
inline void CudaGlobalAlignBatch(const std::vector<CudaDPRange>& ranges, PWPath* paths, std::vector<bool>& handled,
                                 std::vector<long long>* scores = NULL)
{
	const size_t n = ranges.size();
	for (size_t i = 0; i < n; ++i) paths[i].Clear();
	handled.assign(n, false);
	if (scores) scores->assign(n, 0);
	std::string a, b, sa, sb;
	std::vector<uint64_t> a_off(1, 0), b_off(1, 0), p_off(1, 0);
	std::vector<size_t> index;
	for (size_t i = 0; i < n; ++i) {
		if (!cuda_detail::ProfileToString(ranges[i].PA, ranges[i].uLengthA, sa) || !cuda_detail::ProfileToString(ranges[i].PB, ranges[i].uLengthB, sb))
			continue;
		a += sa; b += sb;
		a_off.push_back(a.size()); b_off.push_back(b.size()); p_off.push_back(p_off.back() + sa.size() + sb.size());
		index.push_back(i);
	}
	const size_t m = index.size();
	if (!m) return;
	std::vector<char> path(p_off.back());
	std::vector<uint32_t> plen(m);
	std::vector<int64_t> score(m);
	const int rc = mcu_nw_batch(m, a.data(), &a_off[0], b.data(), &b_off[0], &p_off[0], &path[0], &plen[0], &score[0], NULL);
	if (rc != MCU_OK) throw std::runtime_error(std::string("CudaGlobalAlignBatch: ") + mcu_last_error());
	for (size_t k = 0; k < m; ++k) {
		PWPath& P = paths[index[k]];
		unsigned ua = 0, ub = 0;
		const char* e = &path[p_off[k]];
		for (uint32_t j = 0; j < plen[k]; ++j) {   // PWEdge prefix lengths count the letters consumed including this edge
			if (e[j] != 'I') ++ua;
			if (e[j] != 'D') ++ub;
			P.AppendEdge(e[j], ua, ub);
		}
		handled[index[k]] = true;
		if (scores) (*scores)[index[k]] = score[k];
	}
}

}  // namespace muscle

#endif
