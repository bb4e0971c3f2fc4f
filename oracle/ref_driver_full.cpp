// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI driver around the UNMODIFIED reference sources for the rows next to the hot path, which need
// the whole of libMems (compiled in place by oracle/Makefile.ref into oracle/_ref/libmauve_ref_full.so):
//
//   ref_sol_build      -> mems::SeedOccurrenceList::construct            (LM/SeedOccurrenceList.h:22-78)
//   ref_anchor_scores  -> mems::GetPairwiseAnchorScore                   (LM/GreedyBreakpointElimination.h:403-476)
//   ref_anchor_cols    -> muscle::FindAnchorColsPP                       (MU/anchoredpp.cpp:354-409)
//
// Only the glue below is ours; every algorithmic step runs reference code.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "libGenome/gnSequence.h"
#include "libMems/DNAMemorySML.h"
#include "libMems/Match.h"
#include "libMems/MatchList.h"
#include "libMems/SeedOccurrenceList.h"
#include "libMems/SubstitutionMatrix.h"
#include "libMems/GreedyBreakpointElimination.h"
using namespace std;  // LM/Scoring.h names std types unqualified
#include "libMems/Scoring.h"

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/params.h"
#include "libMUSCLE/alpha.h"

using namespace std;
using namespace genome;
using namespace mems;

namespace muscle {
void FindAnchorColsPP(const MSA& msa1, const MSA& msa2, unsigned AnchorCols[], unsigned* ptruAnchorColCount);
SCORE LetterObjScoreXP(const MSA& msa1, const MSA& msa2, SCORE MatchScore[]);
}

extern "C" {

struct ref_match { int64_t len; int64_t start0; int64_t start1; };

// freq_out: n floats.  Returns n or -1.
long long ref_sol_build(const char* seq, uint64_t n, uint64_t seed, float* freq_out)
{
	try {
		gnSequence s(string(seq, n));
		DNAMemorySML sml;
		sml.Create(s, seed);
		SeedOccurrenceList sol;
		sol.construct(sml);
		for (uint64_t i = 0; i < n; ++i) freq_out[i] = sol.getFrequency(i);
		return (long long)n;
	} catch (...) { return -1; }
}

// lcb_off: n_lcb + 1 offsets into the rows.  Builds both seed occurrence lists with `seed`, then one
// GetPairwiseAnchorScore call per LCB, exactly as CreatePairwiseBPDistance does (LM/ProgressiveAligner.cpp:3420-3422).
int ref_anchor_scores(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, const ref_match* rows,
                      uint64_t n_rows, const uint64_t* lcb_off, uint64_t n_lcb, int penalize, double* lcb_score_out)
{
	try {
		vector<gnSequence*> seq_table;
		seq_table.push_back(new gnSequence(string(seq0, n0)));
		seq_table.push_back(new gnSequence(string(seq1, n1)));
		DNAMemorySML sml0, sml1;
		sml0.Create(*seq_table[0], seed);
		sml1.Create(*seq_table[1], seed);
		SeedOccurrenceList sol0, sol1;
		sol0.construct(sml0);
		sol1.construct(sml1);
		const bool saved = penalize_repeats;
		penalize_repeats = penalize != 0;
		PairwiseScoringScheme pss;
		for (uint64_t l = 0; l < n_lcb; ++l) {
			vector<Match*> lcb;
			for (uint64_t k = lcb_off[l]; k < lcb_off[l + 1] && k < n_rows; ++k) {
				Match mm(2);
				Match* m = mm.Copy();
				m->SetStart(0, rows[k].start0);
				m->SetStart(1, rows[k].start1);
				m->SetLength(rows[k].len);
				lcb.push_back(m);
			}
			lcb_score_out[l] = GetPairwiseAnchorScore(lcb, seq_table, pss, sol0, sol1);
			for (size_t i = 0; i < lcb.size(); ++i) lcb[i]->Free();
		}
		penalize_repeats = saved;
		delete seq_table[0];
		delete seq_table[1];
		return 0;
	} catch (...) { return -1; }
}

}  // extern "C"
