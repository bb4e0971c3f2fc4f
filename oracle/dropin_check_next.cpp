// TEST INFRASTRUCTURE ONLY.
//
// Drop-in check for the rows next to the hot path, like oracle/dropin_check.cpp but linked with ALL of
// libMems (these callers live in ProgressiveAligner / GreedyBreakpointElimination): the reference's own classes next to the
// adapters of mauve_py_b200/adapters/, same inputs through both, in one process.
//
//   sol    <fasta> <weight> <rank>          SeedOccurrenceList::construct vs CudaSeedOccurrenceList::construct: every getFrequency()
//   scores <a.fa> <b.fa> <weight> <rank>    the reference's own flow of CreatePairwiseBPDistance (LM/ProgressiveAligner.cpp:3395-3422):
//                                           PairwiseMatchFinder -> EliminateOverlaps_v2 -> MultiplicityFilter(2) -> IdentifyBreakpoints
//                                           -> ComputeLCBs_v2, then GetPairwiseAnchorScore per LCB vs ONE CudaPairwiseAnchorScores call
//
// Prints one "key value" line per measurement and "RESULT identical|DIFFERENT"; exit code 0 only when identical.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "libGenome/gnSequence.h"
#include "libMems/DNAMemorySML.h"
#include "libMems/MatchList.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/SeedMasks.h"
#include "libMems/ProgressiveAligner.h"
#include "libMems/GreedyBreakpointElimination.h"
using namespace std;  // LM/Scoring.h names std types unqualified
#include "libMems/Scoring.h"

#include "CudaSeedOccurrenceList.h"

using namespace genome;
using namespace mems;

static double now_s() { return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }

static uint64 pick_seed(const vector<gnSequence*>& seqs, int weight, int rank)
{
	if (weight == 0) {
		gnSeqI total = 0;
		for (size_t i = 0; i < seqs.size(); ++i) total += seqs[i]->length();
		weight = getDefaultSeedWeight(total / seqs.size());
	}
	return (uint64)getSeed(weight, rank);
}

static gnSequence* load(const char* path)
{
	gnSequence* s = new gnSequence();
	s->LoadSource(path);
	return s;
}

static int cmd_sol(int argc, char** argv)
{
	if (argc < 5) return 2;
	vector<gnSequence*> seqs(1, load(argv[2]));
	const uint64 seed = pick_seed(seqs, atoi(argv[3]), atoi(argv[4]));
	DNAMemorySML sml;
	sml.Create(*seqs[0], seed);
	double t0 = now_s();
	SeedOccurrenceList ref;
	ref.construct(sml);
	double t1 = now_s();
	CudaSeedOccurrenceList dev;
	dev.construct(sml);
	double t2 = now_s();
	const gnSeqI n = sml.Length();
	gnSeqI differing = 0, not_one = 0;
	for (gnSeqI i = 0; i < n; ++i) {
		const float a = ref.getFrequency(i), b = dev.getFrequency(i);
		if (memcmp(&a, &b, sizeof a)) ++differing;
		if (a != 1.0f) ++not_one;
	}
	cout << "length " << n << "\nseed 0x" << hex << seed << dec << "\nnot_one " << not_one << "\ndiffering " << differing << "\nreference_s " << t1 - t0
	     << "\ncuda_s " << t2 - t1 << "\nRESULT " << (differing == 0 ? "identical" : "DIFFERENT") << endl;
	return differing == 0 ? 0 : 1;
}

static int cmd_scores(int argc, char** argv)
{
	if (argc < 6) return 2;
	MatchList ml;
	ml.seq_table.push_back(load(argv[2]));
	ml.seq_table.push_back(load(argv[3]));
	ml.seq_filename.push_back(argv[2]);
	ml.seq_filename.push_back(argv[3]);
	const uint64 seed = pick_seed(ml.seq_table, atoi(argv[4]), atoi(argv[5]));
	for (int i = 0; i < 2; ++i) {
		DNAMemorySML* sml = new DNAMemorySML();
		sml->Create(*ml.seq_table[i], seed);
		ml.sml_table.push_back(sml);
	}
	PairwiseMatchFinder pmf;
	pmf.FindMatches(ml);
	pmf.Clear();
	const size_t n_matches = ml.size();
	// LM/ProgressiveAligner.cpp:3408-3418
	EliminateOverlaps_v2(ml, true);
	ml.MultiplicityFilter(2);
	vector<MatchList> LCB_list;
	vector<gnSeqI> breakpoints;
	IdentifyBreakpoints(ml, breakpoints);
	ComputeLCBs_v2(ml, breakpoints, LCB_list);
	size_t rows = 0, reverse = 0;
	for (size_t l = 0; l < LCB_list.size(); ++l)
		for (size_t k = 0; k < LCB_list[l].size(); ++k) { ++rows; if (LCB_list[l][k]->Start(1) < 0) ++reverse; }
	PairwiseScoringScheme pss;
	int rc_all = 0;
	for (int pen = 0; pen < 2; ++pen) {
		penalize_repeats = pen != 0;
		double t0 = now_s();
		SeedOccurrenceList r0, r1;
		r0.construct(*ml.sml_table[0]);
		r1.construct(*ml.sml_table[1]);
		vector<double> ref_scores(LCB_list.size());
		for (size_t l = 0; l < LCB_list.size(); ++l) ref_scores[l] = GetPairwiseAnchorScore(LCB_list[l], ml.seq_table, pss, r0, r1);
		double t1 = now_s();
		CudaSeedOccurrenceList c0, c1;
		c0.construct(*ml.sml_table[0]);
		c1.construct(*ml.sml_table[1]);
		vector<double> dev_scores;
		CudaPairwiseAnchorScores(LCB_list, ml.seq_table, pss, c0, c1, dev_scores);
		double t2 = now_s();
		size_t differing = 0;
		double total = 0;
		for (size_t l = 0; l < LCB_list.size(); ++l) { if (ref_scores[l] != dev_scores[l]) ++differing; total += ref_scores[l]; }
		// the single-LCB form, on the first LCB
		if (!LCB_list.empty() && CudaGetPairwiseAnchorScore(LCB_list[0], ml.seq_table, pss, c0, c1) != ref_scores[0]) ++differing;
		cout << (pen ? "penalized_" : "") << "differing " << differing << "\n" << (pen ? "penalized_" : "") << "total " << (long long)total << "\n"
		     << (pen ? "penalized_" : "") << "reference_s " << t1 - t0 << "\n" << (pen ? "penalized_" : "") << "cuda_s " << t2 - t1 << endl;
		if (differing) rc_all = 1;
	}
	penalize_repeats = false;
	cout << "matches " << n_matches << "\nlcbs " << LCB_list.size() << "\nlcb_rows " << rows << "\nreverse_rows " << reverse << "\nRESULT "
	     << (rc_all == 0 ? "identical" : "DIFFERENT") << endl;
	return rc_all;
}

int main(int argc, char** argv)
{
	if (argc < 2) { cerr << "usage: dropin_check_next sol|scores ..." << endl; return 2; }
	if (mcu_init(0) != MCU_OK) { cerr << "no CUDA device: " << mcu_last_error() << endl; return 3; }
	try {
		const string c = argv[1];
		if (c == "sol") return cmd_sol(argc, argv);
		if (c == "scores") return cmd_scores(argc, argv);
	} catch (const char* msg) {
		cerr << "exception: " << msg << endl;
		return 4;
	} catch (std::exception& e) {
		cerr << "exception: " << e.what() << endl;
		return 4;
	}
	return 2;
}
