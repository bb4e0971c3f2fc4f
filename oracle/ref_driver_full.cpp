// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI driver around the UNMODIFIED reference sources for the rows next to the hot path, which need
// the whole of libMems (compiled in place by oracle/Makefile.ref into oracle/_ref/libmauve_ref_full.so):
//
//   ref_sol_build      -> mems::SeedOccurrenceList::construct            (LM/SeedOccurrenceList.h:22-78)
//   ref_anchor_scores  -> mems::GetPairwiseAnchorScore                   (LM/GreedyBreakpointElimination.h:403-476)
//   ref_anchor_cols    -> muscle::FindAnchorColsPP                       (MU/anchoredpp.cpp:354-409)
//   ref_eliminate_overlaps -> mems::EliminateOverlaps_v2 + LengthFilter  (LM/ProgressiveAligner.h:300-406, LM/MatchList.h:680-692)
//   ref_lcbs           -> mems::IdentifyBreakpoints + ComputeLCBs_v2     (LM/GreedyBreakpointElimination.h:161-250)
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
#include "libMems/ProgressiveAligner.h"
using namespace std;  // LM/Scoring.h names std types unqualified
#include "libMems/Scoring.h"

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/params.h"
#include "libMUSCLE/alpha.h"
#include "libMUSCLE/profile.h"

using namespace std;
using namespace genome;
using namespace mems;

namespace muscle {
void FindAnchorColsPP(const MSA& msa1, const MSA& msa2, unsigned AnchorCols[], unsigned* ptruAnchorColCount);
SCORE LetterObjScoreXP(const MSA& msa1, const MSA& msa2, SCORE MatchScore[]);
void PrepareMSAforScoring(MSA& msa);
SCORE TermGapScore(bool Gap);
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

static void rows_to_list(const ref_match* rows, uint64_t n, MatchList& ml)
{
	Match mm(2);
	for (uint64_t k = 0; k < n; ++k) {
		Match* m = mm.Copy();
		m->SetStart(0, rows[k].start0);
		m->SetStart(1, rows[k].start1);
		m->SetLength(rows[k].len);
		ml.push_back(m);
	}
}

// EliminateOverlaps_v2(ml, eliminate_both) followed (min_length > 0) by ml.LengthFilter(min_length): the sequence of
// pairwiseAnchorSearch (LM/ProgressiveAligner.cpp:656-660) and, with eliminate_both, of the pairwise LCB set-up (:3408-3410).
// out: what the list holds afterwards, in its order.  Returns the count or -1.
long long ref_eliminate_overlaps(const ref_match* rows, uint64_t n, int eliminate_both, uint64_t min_length, ref_match* out)
{
	try {
		MatchList ml;
		rows_to_list(rows, n, ml);
		EliminateOverlaps_v2(ml, eliminate_both != 0);
		if (min_length) ml.LengthFilter(min_length);
		for (size_t i = 0; i < ml.size(); ++i) {
			out[i].len = (int64_t)ml[i]->Length();
			out[i].start0 = ml[i]->Start(0);
			out[i].start1 = ml[i]->Start(1);
			ml[i]->Free();
		}
		return (long long)ml.size();
	} catch (...) { return -1; }
}

// IdentifyBreakpoints + ComputeLCBs_v2 on the list: sorted_out = the list afterwards (sorted on genome 0), bp_out = the breakpoints
// (index of the last match of every LCB).  Returns their number or -1.
long long ref_lcbs(const ref_match* rows, uint64_t n, ref_match* sorted_out, uint64_t* bp_out)
{
	try {
		MatchList ml;
		rows_to_list(rows, n, ml);
		vector<gnSeqI> breakpoints;
		IdentifyBreakpoints(ml, breakpoints);
		vector<MatchList> lcbs;
		ComputeLCBs_v2(ml, breakpoints, lcbs);
		size_t k = 0;
		for (size_t l = 0; l < lcbs.size(); ++l) k += lcbs[l].size();
		if (k != ml.size()) return -2;
		for (size_t i = 0; i < ml.size(); ++i) {
			sorted_out[i].len = (int64_t)ml[i]->Length();
			sorted_out[i].start0 = ml[i]->Start(0);
			sorted_out[i].start1 = ml[i]->Start(1);
		}
		for (size_t i = 0; i < breakpoints.size(); ++i) bp_out[i] = breakpoints[i];
		for (size_t i = 0; i < ml.size(); ++i) ml[i]->Free();
		return (long long)breakpoints.size();
	} catch (...) { return -1; }
}

// ---- anchor columns of a window ------------------------------------------------------------------------------------------------
// Global settings exactly as MuscleInterface::ProfileAlignFast (LM/MuscleInterface.cpp:1086-1106).
static void ref_muscle_globals(unsigned nseq)
{
	using namespace muscle;
	g_SeqType.get() = SEQTYPE_DNA;
	g_uMaxIters.get() = 1;
	g_bStable.get() = true;
	g_bQuiet.get() = true;
	g_SeqWeight1.get() = SEQWEIGHT_ClustalW;
	SetMaxIters(g_uMaxIters.get());
	SetSeqWeightMethod(g_SeqWeight1.get());
	MSA::SetIdCount(nseq);
	SetAlpha(ALPHA_DNA);
	SetPPScore(PPSCORE_SPN);
}

static void ref_msa_from_rows(muscle::MSA& msa, const char* rows, unsigned nrows, unsigned ncol, unsigned first_id)
{
	msa.SetSize(nrows, ncol);
	for (unsigned r = 0; r < nrows; ++r) {
		char name[32];
		snprintf(name, sizeof name, "seq%u", first_id + r);
		msa.SetSeqName(r, name);
		msa.SetSeqId(r, first_id + r);
		for (unsigned c = 0; c < ncol; ++c) msa.SetChar(r, c, rows[(size_t)r * ncol + c]);
	}
}

// rows: (n1 + n2) rows of ncol characters ('-' = gap), the first alignment's rows first.  prepare != 0: the rows' weights come from
// PrepareMSAforScoring, as in AnchoredProfileProfile (MU/anchoredpp.cpp:454-455), and are written to weights; prepare == 0: weights
// are given.  cols_out: FindAnchorColsPP's columns; score_out / smooth_out (optional): LetterObjScoreXP's per-column scores and
// WindowSmooth of them with the settings FindAnchorColsPP uses.  fixed_rows_out (optional): the characters after MSA::FixAlpha.
// Returns the number of anchor columns or -1.
long long ref_anchor_cols(const char* rows, unsigned n1, unsigned n2, unsigned ncol, float* weights, int prepare, unsigned* cols_out,
                          float* score_out, float* smooth_out, char* fixed_rows_out)
{
	using namespace muscle;
	try {
		ref_muscle_globals(n1 + n2);
		MSA msa1, msa2;
		ref_msa_from_rows(msa1, rows, n1, ncol, 0);
		ref_msa_from_rows(msa2, rows + (size_t)n1 * ncol, n2, ncol, n1);
		msa1.FixAlpha();
		msa2.FixAlpha();
		SetPPScore(PPSCORE_SPN);
		if (prepare) {
			PrepareMSAforScoring(msa1);
			PrepareMSAforScoring(msa2);
			for (unsigned r = 0; r < n1; ++r) weights[r] = msa1.GetSeqWeight(r);
			for (unsigned r = 0; r < n2; ++r) weights[n1 + r] = msa2.GetSeqWeight(r);
		} else {
			for (unsigned r = 0; r < n1; ++r) msa1.SetSeqWeight(r, weights[r]);
			for (unsigned r = 0; r < n2; ++r) msa2.SetSeqWeight(r, weights[n1 + r]);
		}
		if (fixed_rows_out) {
			for (unsigned r = 0; r < n1; ++r)
				for (unsigned c = 0; c < ncol; ++c) fixed_rows_out[(size_t)r * ncol + c] = msa1.GetChar(r, c);
			for (unsigned r = 0; r < n2; ++r)
				for (unsigned c = 0; c < ncol; ++c) fixed_rows_out[(size_t)(n1 + r) * ncol + c] = msa2.GetChar(r, c);
		}
		unsigned count = 0;
		vector<unsigned> cols(ncol + 1);
		FindAnchorColsPP(msa1, msa2, &cols[0], &count);
		for (unsigned i = 0; i < count; ++i) cols_out[i] = cols[i];
		if (score_out || smooth_out) {
			vector<SCORE> score(ncol + 1), smooth(ncol + 1);
			LetterObjScoreXP(msa1, msa2, &score[0]);
			WindowSmooth(&score[0], ncol, g_uSmoothWindowLength.get(), &smooth[0], g_dSmoothScoreCeil.get());
			for (unsigned c = 0; c < ncol; ++c) {
				if (score_out) score_out[c] = score[c];
				if (smooth_out) smooth_out[c] = smooth[c];
			}
		}
		return (long long)count;
	} catch (...) { return -1; }
}

// the settings the column scoring reads, after the set-up above (for the oracle's / the product's default parameters):
// out[0..15] the score matrix on A C G T, [16] g_scoreGapOpen, [17] g_scoreGapExtend, [18] TermGapScore(true), [19] g_dSmoothScoreCeil,
// [20] g_dMinBestColScore, [21] g_dMinSmoothScore, [22] g_uSmoothWindowLength, [23] g_uAnchorSpacing (both as FindAnchorColsPP sets
// them), [24] g_AlphaSize; letters_out[256]: CharToLetterEx per character (255 where IsGapChar)
void ref_anchor_settings(float* out, unsigned char* letters_out)
{
	using namespace muscle;
	ref_muscle_globals(2);
	for (int a = 0; a < 4; ++a)
		for (int b = 0; b < 4; ++b) out[a * 4 + b] = (*g_ptrScoreMatrix.get())[a][b];
	out[16] = g_scoreGapOpen.get();
	out[17] = g_scoreGapExtend.get();
	out[18] = TermGapScore(true);
	out[19] = g_dSmoothScoreCeil.get();
	out[20] = g_dMinBestColScore.get();
	out[21] = g_dMinSmoothScore.get();
	out[22] = 21;
	out[23] = 96;
	out[24] = (float)g_AlphaSize.get();
	for (int c = 0; c < 256; ++c) {
		unsigned l = CharToLetterEx((char)c);
		letters_out[c] = IsGapChar((char)c) ? 255 : (unsigned char)(l > 254 ? 254 : l);
	}
}

}  // extern "C"
