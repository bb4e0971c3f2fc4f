// CudaFindAnchorColsPP -- drop-in for muscle::FindAnchorColsPP (MU/anchoredpp.cpp:354-409): the anchor columns of an
// alignment window, i.e. which ranges AnchoredProfileProfile (:443-552) hands to the gapped DP.
// Same arguments, same result: AnchorCols[] / *ptruAnchorColCount as the reference fills them (per-column scores and their smoothed
// form are the reference's floats: LetterObjScoreXP :256-329, WindowSmooth MU/anchors.cpp:9-47, FindBestColsComboPP :335-351,
// MergeBestCols MU/anchors.cpp:137-186 all run on the device, mcu_anchor_cols_batch).  Like the reference it leaves
// g_uSmoothWindowLength = 21 and g_uAnchorSpacing = 96 behind (:370-371).  The settings are read from MUSCLE's globals at the time of
// the call; returns false (nothing touched) when the alphabet is not a four-letter one -- the caller keeps the reference's function.
#ifndef CUDA_ANCHOR_COLS_H_
#define CUDA_ANCHOR_COLS_H_

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef MUSCLE_LONG_VERSION   /* muscle.h has no include guard of its own */
#include "libMUSCLE/muscle.h"
#endif
#include "libMUSCLE/msa.h"
#include "libMUSCLE/alpha.h"
#include "libMUSCLE/params.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/objscore.h"
#include "mauve_cuda.h"

namespace cuda_cols_detail {
// MUSCLE's globals as the column scoring reads them
inline void CudaAnchorParamsFromGlobals(mcu_anchor_params& p)
{
	using namespace muscle;
	memset(&p, 0, sizeof p);
	for (unsigned a = 0; a < 4; ++a)
		for (unsigned b = 0; b < 4; ++b) p.subst[a][b] = (*g_ptrScoreMatrix.get())[a][b];
	p.gap_open = g_scoreGapOpen.get();
	p.gap_extend = g_scoreGapExtend.get();
	p.term_gap = TermGapScore(true);
	p.smooth_ceil = g_dSmoothScoreCeil.get();
	p.min_best_col = g_dMinBestColScore.get();
	p.min_smooth = g_dMinSmoothScore.get();
	p.smooth_window = g_uSmoothWindowLength.get();
	p.anchor_spacing = g_uAnchorSpacing.get();
	for (unsigned c = 0; c < 256; ++c) {
		const unsigned l = CharToLetterEx((char)c);
		p.letter_of_char[c] = IsGapChar((char)c) ? (uint8_t)MCU_AC_GAP : (uint8_t)(l > 254 ? 254 : l);
	}
}
}  // namespace cuda_cols_detail

inline bool CudaFindAnchorColsPP(const muscle::MSA& msa1, const muscle::MSA& msa2, unsigned AnchorCols[], unsigned* ptruAnchorColCount)
{
	using namespace muscle;
	if (g_AlphaSize.get() != 4) return false;
	const unsigned uColCount = msa1.GetColCount();
	if (uColCount != msa2.GetColCount()) {   // :358-362
		*ptruAnchorColCount = 0;
		return true;
	}
	g_uSmoothWindowLength.get() = 21;   // :370-371
	g_uAnchorSpacing.get() = 96;
	const uint32_t n1 = msa1.GetSeqCount(), n2 = msa2.GetSeqCount();
	if (uColCount == 0 || n1 == 0 || n2 == 0) return false;
	mcu_anchor_params p;
	cuda_cols_detail::CudaAnchorParamsFromGlobals(p);
	std::vector<char> rows((size_t)(n1 + n2) * uColCount);
	std::vector<float> weights(n1 + n2);
	for (uint32_t r = 0; r < n1; ++r) {
		memcpy(&rows[(size_t)r * uColCount], msa1.GetSeqBuffer(r), uColCount);
		weights[r] = msa1.GetSeqWeight(r);
	}
	for (uint32_t r = 0; r < n2; ++r) {
		memcpy(&rows[(size_t)(n1 + r) * uColCount], msa2.GetSeqBuffer(r), uColCount);
		weights[n1 + r] = msa2.GetSeqWeight(r);
	}
	const uint64_t row_off[2] = {0, (uint64_t)rows.size()}, col_off[2] = {0, uColCount};
	const uint32_t ncol = uColCount;
	std::vector<uint32_t> cols(uColCount);
	uint32_t count = 0;
	const int rc = mcu_anchor_cols_batch(1, &rows[0], row_off, &ncol, &n1, &n2, &weights[0], &p, col_off, &cols[0], &count, NULL, NULL, NULL);
	if (rc != MCU_OK) throw std::runtime_error(std::string("CudaFindAnchorColsPP: ") + mcu_last_error());
	for (uint32_t i = 0; i < count; ++i) AnchorCols[i] = cols[i];
	*ptruAnchorColCount = count;
	return true;
}

#endif
