// CudaSeedOccurrenceList / CudaPairwiseAnchorScores -- drop-ins for the consumers of the sorted mer list after matching
// (SURVEY.md 8f-2), in the reference's own language, compiled against the reference's headers and linked with
// libmauve_cuda.so (include/mauve_cuda.h).
//
//   CudaSeedOccurrenceList::construct(sml)   replaces SeedOccurrenceList::construct (LM/SeedOccurrenceList.h:22-78): the
//       multiplicities and their smoothing are computed on the GPU (mcu_sol_build); the result is written to the same
//       temporary file and mapped through the base class's members, so getFrequency() (:81-84) and the destructor are the
//       inherited ones and every caller -- ProgressiveAligner::align (LM/ProgressiveAligner.cpp:3908-3912),
//       GetPairwiseAnchorScore -- sees the same object.  Only the 2-bit sequence of the list matters to the result, so the
//       bases are taken from the list itself (SortedMerList::GetBSequence), whatever SML class the caller holds.
//   CudaPairwiseAnchorScores(LCB_list, seq_table, scoring, sol_1, sol_2, scores)   replaces the loop
//       `for lcbI: lcb_scores[lcbI] = GetPairwiseAnchorScore(LCB_list[lcbI], ...)` (LM/ProgressiveAligner.cpp:3421-3422 and
//       the per-pair call at :1825) with ONE device call for all LCBs of the genome pair (mcu_anchor_scores).
//   CudaGetPairwiseAnchorScore(lcb, ...)     the single-LCB form with the reference's signature.
// Matches must be ungapped two-genome matches (mems::Match); penalize_gaps is not supported (the aligner never sets it).
// There is no CPU fallback: a failing device call throws.
#ifndef CUDA_SEED_OCCURRENCE_LIST_H_
#define CUDA_SEED_OCCURRENCE_LIST_H_

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "libMems/SeedOccurrenceList.h"
#include "libMems/SubstitutionMatrix.h"
#include "libMems/GreedyBreakpointElimination.h"
#include "mauve_cuda.h"

namespace mems {

namespace cuda_detail {

// getFrequency() values of every position of the list's sequence, computed on the device
template <typename SMLType>
inline void SolFrequencies(SMLType& sml, std::vector<SeedOccurrenceList::frequency_type>& count)
{
	const size_t total_len = sml.Length();
	std::string bases(total_len, 'A');
	if (total_len) {   // 2-bit codes -> letters (the SML build on the device packs them back to the same codes)
		const gnSeqI words = (total_len * 2) / 32 + (((total_len * 2) % 32) ? 1 : 0);
		std::vector<uint32> buf(words + 4, 0);
		uint32* packed = &buf[1];   // GetBSequence masks dest[-1] when the sequence is shorter than one word (LM/SortedMerList.cpp:369-372)
		sml.GetBSequence(packed, total_len, 0);
		static const char letters[4] = {'A', 'C', 'G', 'T'};
		for (gnSeqI i = 0; i < total_len; ++i) bases[i] = letters[(packed[i >> 4] >> (30 - 2 * (i & 15))) & 3];
	}
	count.assign(total_len ? total_len : 1, 1.0f);
	if (mcu_sol_build(bases.data(), total_len, sml.Seed(), &count[0]) != MCU_OK) {
		std::cerr << "CudaSeedOccurrenceList::construct: " << mcu_last_error() << std::endl;
		throw "CudaSeedOccurrenceList::construct failed";
	}
}

}  // namespace cuda_detail

class CudaSeedOccurrenceList : public SeedOccurrenceList
{
public:
	CudaSeedOccurrenceList() : total_len(0) {}

	template <typename SMLType>
	void construct(SMLType& sml)
	{
		total_len = sml.Length();
		std::vector<frequency_type> count;
		cuda_detail::SolFrequencies(sml, count);
		// as the reference from here on (:67-77): temporary file, memory mapped
		tmpfile = CreateTempFileName("sol");
		{
			std::ofstream tfout;
			tfout.open(tmpfile.c_str(), std::ios::binary);
			tfout.write((const char*)&count[0], total_len * sizeof(frequency_type));
			tfout.close();
		}
		data.close();
		data.open(tmpfile);
	}

	const frequency_type* frequencies() const { return (const frequency_type*)data.data(); }
	size_t length() const { return total_len; }

private:
	size_t total_len;
};

// all LCBs of one genome pair in one device call; scores[lcbI] is what GetPairwiseAnchorScore(LCB_list[lcbI], ...) returns
template <class MatchVector>
void CudaPairwiseAnchorScores(std::vector<MatchVector>& LCB_list, std::vector<genome::gnSequence*>& seq_table,
                              const PairwiseScoringScheme& subst_scoring, CudaSeedOccurrenceList& sol_1, CudaSeedOccurrenceList& sol_2,
                              std::vector<double>& scores)
{
	std::vector<mcu_match> rows;
	std::vector<uint64_t> off(1, 0);
	for (size_t l = 0; l < LCB_list.size(); ++l) {
		for (typename MatchVector::iterator it = LCB_list[l].begin(); it != LCB_list[l].end(); ++it) {
			mcu_match r;
			r.len = (int64_t)(*it)->Length(0);
			r.start0 = (*it)->Start(0);
			r.start1 = (*it)->Start(1);
			if ((*it)->SeqCount() != 2 || (*it)->Length(1) != (*it)->Length(0) || (*it)->AlignmentLength() != (*it)->Length(0))
				throw "CudaPairwiseAnchorScores: ungapped two-genome matches only";
			rows.push_back(r);
		}
		off.push_back(rows.size());
	}
	// explicit lengths: gnRAWSequence::ToString() with default arguments drops the last two bases (LM/gnRAWSequence.h:157-161)
	const std::string s0 = seq_table[0]->ToString(seq_table[0]->length(), 1), s1 = seq_table[1]->ToString(seq_table[1]->length(), 1);
	int32_t matrix[16];
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) matrix[4 * i + j] = subst_scoring.matrix[i][j];
	scores.assign(LCB_list.size(), 0.0);
	if (LCB_list.empty()) return;
	if (sol_1.length() != s0.size() || sol_2.length() != s1.size()) throw "CudaPairwiseAnchorScores: seed occurrence lists of other sequences";
	const int rc = mcu_anchor_scores(s0.data(), s0.size(), s1.data(), s1.size(), 0, sol_1.frequencies(), sol_2.frequencies(),
	                                 rows.empty() ? NULL : &rows[0], rows.size(), &off[0], LCB_list.size(), matrix, penalize_repeats ? 1 : 0,
	                                 &scores[0], NULL);
	if (rc != MCU_OK) {
		std::cerr << "CudaPairwiseAnchorScores: " << mcu_last_error() << std::endl;
		throw "CudaPairwiseAnchorScores failed";
	}
}

template <class MatchVector>
double CudaGetPairwiseAnchorScore(MatchVector& lcb, std::vector<genome::gnSequence*>& seq_table, const PairwiseScoringScheme& subst_scoring,
                                  CudaSeedOccurrenceList& sol_1, CudaSeedOccurrenceList& sol_2, bool penalize_gaps = false)
{
	if (penalize_gaps) throw "CudaGetPairwiseAnchorScore: penalize_gaps is not supported";
	std::vector<MatchVector> one(1, lcb);
	std::vector<double> scores;
	CudaPairwiseAnchorScores(one, seq_table, subst_scoring, sol_1, sol_2, scores);
	return scores[0];
}

}  // namespace mems

#endif
