// CudaPairwiseMatchFinder / CudaMemHash -- drop-ins for mems::PairwiseMatchFinder and mems::MemHash whose
// FindMatches() runs seed-match enumeration + extension on the GPU for two genomes.
//
// Replaces MemHash::FindMatches (LM/MemHash.cpp:109-127): AddSequence x2, MatchFinder::FindMatchSeeds /
// SearchRange (LM/MatchFinder.cpp:137-340), EnumerateMatches (LM/PairwiseMatchFinder.cpp:37-71 or, for MemHash
// with repeat_tolerance 0 / enumeration_tolerance 1, LM/MemHash.cpp:139-162), HashMatch / SetDirection
// (:167-203), AddHashEntry (:209-251), ExtendMatch (LM/MatchFinder.h:218-374) and GetMatchList
// (LM/MemHash.h:183-203).  The match rows come back in the reference's list order and are appended to the
// MatchList as slot-allocated Match copies, exactly what GetMatchList leaves there; the caller frees them as today
// (LM/MatchList.h:475-479).  More than two genomes, or tolerances other than the MUM settings, take the inherited
// reference path.  Instances: `PairwiseMatchFinder pmf` (MA/progressiveMauve.cpp:500), `gap_mh`
// (LM/Aligner.h:198, LM/ProgressiveAligner.cpp:651).
#ifndef CUDA_MATCH_FINDER_H_
#define CUDA_MATCH_FINDER_H_

#include <iostream>
#include <string>

#include "libMems/MemHash.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/MatchList.h"
#include "mauve_cuda.h"

namespace mems {

namespace cuda_detail {

// true when the list was produced on the device
inline bool FindMatchesTwoGenomes(MatchList& ml, int rule, uint64& mem_count, uint64& collision_count)
{
	const std::string s0 = ml.seq_table[0]->ToString(), s1 = ml.seq_table[1]->ToString();
	mcu_match* rows = NULL;
	uint64_t n = 0, stats[8];
	const int rc = mcu_find_mums(s0.data(), s0.size(), s1.data(), s1.size(), ml.sml_table[0]->Seed(), rule, &rows, &n, stats);
	if (rc == MCU_EGAP) throw "ERROR: gap character encountered in input sequence";
	if (rc != MCU_OK) {
		std::cerr << "Cuda match finder: " << mcu_last_error() << std::endl;
		Throw_gnEx(InvalidData());   // mems::InvalidData, LM/MatchFinder.h:123
	}
	mem_count = stats[1];
	collision_count = stats[2];
	ml.clear();  // GetMatchList starts with mem_list.clear()
	Match mm(2);
	for (uint64_t i = 0; i < n; ++i) {
		Match* m = mm.Copy();
		m->SetStart(0, rows[i].start0);
		m->SetStart(1, rows[i].start1);
		m->SetLength(rows[i].len);
		ml.push_back(m);
	}
	mcu_free(rows);
	return true;
}

}  // namespace cuda_detail

class CudaPairwiseMatchFinder : public PairwiseMatchFinder
{
public:
	virtual void FindMatches(MatchList& ml)
	{
		if (ml.seq_table.size() != 2 || ml.sml_table.size() != 2) { PairwiseMatchFinder::FindMatches(ml); return; }
		cuda_detail::FindMatchesTwoGenomes(ml, MCU_RULE_PAIRWISE, m_mem_count, m_collision_count);
	}
};

class CudaMemHash : public MemHash
{
public:
	virtual void FindMatches(MatchList& ml)
	{
		// the MUM settings of gap_mh (LM/ProgressiveAligner.cpp:649-651): only then is "unique in every genome" the rule
		if (ml.seq_table.size() != 2 || ml.sml_table.size() != 2 || m_repeat_tolerance != 0 || m_enumeration_tolerance != 1) {
			MemHash::FindMatches(ml);
			return;
		}
		cuda_detail::FindMatchesTwoGenomes(ml, MCU_RULE_MEMHASH, m_mem_count, m_collision_count);
	}
};

}  // namespace mems

#endif
