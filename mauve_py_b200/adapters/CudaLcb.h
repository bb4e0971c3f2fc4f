// CudaEliminateOverlaps / CudaIdentifyBreakpoints -- drop-ins for the step after the match list when the list holds matches of TWO
// genomes:
//   mems::EliminateOverlaps_v2(ml, eliminate_both)  LM/ProgressiveAligner.h:396-406 (-> :300-394), followed by
//   ml.LengthFilter(min_length)                     LM/MatchList.h:680-692         as pairwiseAnchorSearch does (LM/ProgressiveAligner.cpp:656-660)
//   mems::IdentifyBreakpoints(ml, breakpoints)      LM/GreedyBreakpointElimination.h:161-226
// Same arguments, same effect on the list: afterwards it holds the matches the reference's functions would leave, in their order
// (the matches are rebuilt from the device's rows: a Match of a two-genome list is its two starts and its length).
// CudaTwoGenomeRows says whether a list is in the form the device entry covers; callers keep the reference's function otherwise.
#ifndef CUDA_LCB_H_
#define CUDA_LCB_H_

#include <stdexcept>
#include <string>
#include <vector>

#include "libMems/MatchList.h"
#include "mauve_cuda.h"

namespace cuda_detail {
// rows of a list whose matches are all defined in exactly two genomes, forward in the first; false otherwise
inline bool CudaTwoGenomeRows(const mems::MatchList& ml, std::vector<mcu_match>& rows)
{
	rows.clear();
	rows.reserve(ml.size());
	for (size_t i = 0; i < ml.size(); ++i) {
		const mems::Match* m = ml[i];
		if (m == NULL || m->SeqCount() != 2 || m->Start(0) <= 0 || m->Start(1) == 0 || m->Length() == 0) return false;
		mcu_match r;
		r.len = (int64_t)m->Length();
		r.start0 = m->Start(0);
		r.start1 = m->Start(1);
		rows.push_back(r);
	}
	return true;
}

inline void CudaRowsToList(const mcu_match* rows, size_t n, mems::MatchList& ml)
{
	for (size_t i = 0; i < ml.size(); ++i)
		if (ml[i]) ml[i]->Free();
	ml.clear();
	mems::Match proto(2);
	for (size_t i = 0; i < n; ++i) {
		mems::Match* m = proto.Copy();
		m->SetStart(0, rows[i].start0);
		m->SetStart(1, rows[i].start1);
		m->SetLength((gnSeqI)rows[i].len);
		ml.push_back(m);
	}
}
}  // namespace cuda_detail

// returns false (list untouched) when the list is not a two-genome list
inline bool CudaEliminateOverlaps(mems::MatchList& ml, bool eliminate_both, gnSeqI min_length = 0, uint64_t* ties = NULL)
{
	if (ml.size() < 2 && min_length == 0) return true;   // the reference returns at once
	std::vector<mcu_match> rows;
	if (!cuda_detail::CudaTwoGenomeRows(ml, rows)) return false;
	std::vector<mcu_match> out(rows.size() ? rows.size() : 1);
	uint64_t n_out = 0, t = 0;
	const int rc = mcu_eliminate_overlaps(rows.empty() ? NULL : &rows[0], rows.size(), eliminate_both ? 1 : 0, min_length, &out[0], &n_out, &t);
	if (rc != MCU_OK) throw std::runtime_error(std::string("CudaEliminateOverlaps: ") + mcu_last_error());
	if (ties) *ties = t;
	cuda_detail::CudaRowsToList(&out[0], n_out, ml);
	return true;
}

// the list is ordered on genome 0 (as the reference leaves it) and `breakpoints` filled; false (nothing touched) for other lists
inline bool CudaIdentifyBreakpoints(mems::MatchList& ml, std::vector<gnSeqI>& breakpoints, uint64_t* ties = NULL)
{
	if (ml.size() == 0) return true;
	std::vector<mcu_match> rows;
	if (!cuda_detail::CudaTwoGenomeRows(ml, rows)) return false;
	std::vector<mcu_match> sorted(rows.size());
	std::vector<uint64_t> bp(rows.size() + 1);
	uint64_t n_bp = 0, t = 0;
	const int rc = mcu_lcbs(&rows[0], rows.size(), &sorted[0], &bp[0], &n_bp, &t);
	if (rc != MCU_OK) throw std::runtime_error(std::string("CudaIdentifyBreakpoints: ") + mcu_last_error());
	if (ties) *ties = t;
	cuda_detail::CudaRowsToList(&sorted[0], sorted.size(), ml);
	breakpoints.assign(bp.begin(), bp.begin() + n_bp);
	return true;
}

#endif
