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
#include <utility>
#include <vector>

#include "libMems/MemHash.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/MatchList.h"
#include "libMems/SeedMasks.h"
#include "mauve_cuda.h"

namespace mems {

namespace cuda_detail {

// true when the list was produced on the device
inline bool FindMatchesTwoGenomes(MatchList& ml, int rule, uint64& mem_count, uint64& collision_count)
{
	// explicit lengths: progressiveMauve holds gnRAWSequence objects, whose ToString() with default arguments drops the last two
	// bases (LM/gnRAWSequence.h:157-161)
	const std::string s0 = ml.seq_table[0]->ToString(ml.seq_table[0]->length(), 1), s1 = ml.seq_table[1]->ToString(ml.seq_table[1]->length(), 1);
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

// One round of recursive anchoring in one device call: the matches MemHash (MUM settings) finds in every gap pair, i.e.
// what pairwiseAnchorSearch (LM/ProgressiveAligner.cpp:590-679) has in gap_list after gap_mh.FindMatches (:651) and before
// EliminateOverlaps_v2 / LengthFilter (:656-660), which stay with the caller together with the coordinate shift (:662-671).
// The seed of a gap is the reference's: getSeed(getDefaultSeedWeight((len0 + len1) / 2), 0), no search below
// MIN_DNA_SEED_WEIGHT (:617-634).  (--seed-family, three seeds per gap, is not batched: use CudaMemHash per gap.)
// out[i]: slot-allocated Match copies for gap i, in GetMatchList order, coordinates local to the gap sequences.
inline void CudaGapSearchBatch(const std::vector<std::pair<std::string, std::string> >& gaps, std::vector<std::vector<Match*> >& out)
{
	const size_t n = gaps.size();
	out.assign(n, std::vector<Match*>());
	std::string c0, c1;
	std::vector<uint64_t> o0(1, 0), o1(1, 0), seeds(n, 0), out_off(n + 1, 0);
	for (size_t i = 0; i < n; ++i) {
		c0 += gaps[i].first;
		c1 += gaps[i].second;
		o0.push_back(c0.size());
		o1.push_back(c1.size());
		const uint w = getDefaultSeedWeight((gaps[i].first.size() + gaps[i].second.size()) / 2);
		seeds[i] = w < MIN_DNA_SEED_WEIGHT ? 0 : (uint64_t)getSeed(w, 0);
	}
	mcu_match* rows = NULL;
	const int rc = mcu_find_mums_batch(n, c0.data(), &o0[0], c1.data(), &o1[0], n ? &seeds[0] : NULL, MCU_RULE_MEMHASH, &rows, &out_off[0], NULL);
	if (rc == MCU_EGAP) throw "ERROR: gap character encountered in input sequence";
	if (rc != MCU_OK) {
		std::cerr << "CudaGapSearchBatch: " << mcu_last_error() << std::endl;
		Throw_gnEx(InvalidData());
	}
	Match mm(2);
	for (size_t i = 0; i < n; ++i)
		for (uint64_t r = out_off[i]; r < out_off[i + 1]; ++r) {
			Match* m = mm.Copy();
			m->SetStart(0, rows[r].start0);
			m->SetStart(1, rows[r].start1);
			m->SetLength(rows[r].len);
			out[i].push_back(m);
		}
	mcu_free(rows);
}

}  // namespace mems

#endif
