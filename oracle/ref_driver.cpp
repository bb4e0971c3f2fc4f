// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled in place from
// /root/reference by oracle/Makefile.ref into oracle/_ref/libmauve_ref.so).
// It exposes the four subsystems of the anchoring hot path exactly as the
// reference executes them so that the C restatement (oracle/mauve_oracle.c)
// and the CUDA path can be checked against the real thing:
//
//   ref_sml_build    -> mems::DNAMemorySML::Create + Read       (LM/MemorySML.cpp:45-82)
//   ref_find_mums    -> PairwiseMatchFinder / MemHash FindMatches (LM/MemHash.cpp:109-127)
//   ref_nw_align     -> muscle::ProfileProfile -> NWSmall         (MU/profile.cpp:68, MU/nwsmall.cpp:500)
//   ref_hmm_run      -> run() + Forward/Backward posteriors       (LM/HomologyHMM/homologymain.cc:24)
//
// Only the glue below is ours; every algorithmic step runs reference code.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <sstream>
#include <iostream>

#include "libGenome/gnSequence.h"
#include "libMems/DNAMemorySML.h"
#include "libMems/MatchList.h"
#include "libMems/MemHash.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/SeedMasks.h"

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"
#include "libMUSCLE/params.h"
#include "libMUSCLE/alpha.h"
#include "libMUSCLE/tree.h"

#include "homology.h"
#include "parameters.h"
#include "dptables.h"

using namespace std;
using namespace genome;
using namespace mems;

extern "C" {

// ---- seeds ---------------------------------------------------------------
uint64_t ref_get_seed(int weight, int rank) { return (uint64_t)getSeed(weight, rank); }
unsigned ref_default_seed_weight(uint64_t avg_len) { return getDefaultSeedWeight((gnSeqI)avg_len); }
int ref_seed_length(uint64_t seed) { return getSeedLength((int64)seed); }
int ref_seed_weight(uint64_t seed) { return getSeedWeight((int64)seed); }

// ---- SML -----------------------------------------------------------------
// pos_out / mer_out must hold n-L+1 entries. Returns SML length, or -1 on error.
long long ref_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* mer_out)
{
	try {
		gnSequence s(string(seq, n));
		DNAMemorySML sml;
		sml.Create(s, seed);
		gnSeqI len = sml.SMLLength();
		vector<bmer> v;
		sml.Read(v, len, 0);
		for (size_t i = 0; i < v.size(); ++i) {
			if (pos_out) pos_out[i] = v[i].position;
			if (mer_out) mer_out[i] = v[i].mer;
		}
		return (long long)v.size();
	} catch (...) { return -1; }
}

// packed sequence words as the reference stores them (binary_seq_len words)
long long ref_pack(const char* seq, uint64_t n, uint32_t* out, uint64_t out_words)
{
	try {
		gnSequence s(string(seq, n));
		DNAMemorySML sml;
		sml.Create(s, (uint64)getSeed(5, 0));
		uint64_t words = (n * 2) / 32 + (((n * 2) % 32) ? 1 : 0);
		if (out_words < words) return -1;
		// GetBSequence copies whole characters; fetch word-aligned
		sml.GetBSequence(out, n, 0);
		return (long long)words;
	} catch (...) { return -1; }
}

// FindMer for every query (LM/SortedMerList.cpp:170-179, bsearch :380-394) and GetMer / GetSeedMer / GetDnaSeedMer at every probe
// position (:321-342, :726-769) on the list the reference builds for `seq`.  found_out / rank_out: nq entries; the three mer arrays: np.
long long ref_sml_probe(const char* seq, uint64_t n, uint64_t seed, const uint64_t* queries, uint64_t nq, uint8_t* found_out, uint64_t* rank_out,
                        const uint64_t* positions, uint64_t np, uint64_t* getmer_out, uint64_t* seedmer_out, uint64_t* dnaseedmer_out)
{
	try {
		gnSequence s(string(seq, n));
		DNAMemorySML sml;
		sml.Create(s, seed);
		for (uint64_t i = 0; i < nq; ++i) {
			gnSeqI r = 0;
			found_out[i] = sml.FindMer(queries[i], r) ? 1 : 0;
			rank_out[i] = (uint64_t)r;
		}
		for (uint64_t i = 0; i < np; ++i) {
			getmer_out[i] = sml.GetMer((gnSeqI)positions[i]);
			seedmer_out[i] = sml.GetSeedMer((gnSeqI)positions[i]);
			dnaseedmer_out[i] = sml.GetDnaSeedMer((gnSeqI)positions[i]);
		}
		return (long long)sml.SMLLength();
	} catch (...) { return -1; }
}

// ---- MUMs ----------------------------------------------------------------
struct ref_match { int64_t len; int64_t start0; int64_t start1; };

// rule 0: PairwiseMatchFinder (progressiveMauve.cpp:500-503)
// rule 1: MemHash with repeat_tolerance 0 / enumeration_tolerance 1 (gap_mh, ProgressiveAligner.cpp:651)
// Returns number of matches; *out is malloc'd (free with ref_free). stats[0]=collisions, stats[1]=mem count.
long long ref_find_mums(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule,
                        ref_match** out, uint64_t* stats)
{
	try {
		MatchList ml;
		ml.seq_table.push_back(new gnSequence(string(seq0, n0)));
		ml.seq_table.push_back(new gnSequence(string(seq1, n1)));
		ml.seq_filename.push_back("a");
		ml.seq_filename.push_back("b");
		for (int i = 0; i < 2; ++i) {
			DNAMemorySML* sml = new DNAMemorySML();
			sml->Create(*ml.seq_table[i], seed);
			ml.sml_table.push_back(sml);
		}
		uint64_t coll = 0, cnt = 0;
		if (rule == 0) {
			PairwiseMatchFinder pmf;
			pmf.FindMatches(ml);
			coll = pmf.MemCollisionCount(); cnt = pmf.MemCount();
			pmf.Clear();
		} else {
			MemHash mh;
			mh.SetRepeatTolerance(0);
			mh.SetEnumerationTolerance(1);
			mh.FindMatches(ml);
			coll = mh.MemCollisionCount(); cnt = mh.MemCount();
			mh.Clear();
		}
		if (stats) { stats[0] = coll; stats[1] = cnt; }
		size_t m = ml.size();
		ref_match* r = (ref_match*)malloc(sizeof(ref_match) * (m ? m : 1));
		for (size_t i = 0; i < m; ++i) {
			r[i].len = (int64_t)ml[i]->Length();
			r[i].start0 = ml[i]->Start(0);
			r[i].start1 = ml[i]->Start(1);
			ml[i]->Free();
		}
		for (int i = 0; i < 2; ++i) { delete ml.sml_table[i]; delete ml.seq_table[i]; }
		ml.sml_table.clear(); ml.seq_table.clear();
		*out = r;
		return (long long)m;
	} catch (...) { return -1; }
}

void ref_free(void* p) { free(p); }

// ---- match list file format (LM/MatchList.h:526-662), as --mums writes and --match-input reads ----
// WriteList of n rows -> malloc'd NUL-terminated text (free with ref_free); the match-id column is the pointer value.
char* ref_write_list(const ref_match* rows, uint64_t n, const char* name0, const char* name1, uint64_t len0, uint64_t len1)
{
	try {
		MatchList ml;
		ml.seq_filename.push_back(name0);
		ml.seq_filename.push_back(name1);
		ml.seq_table.push_back(new gnSequence(string((size_t)len0, 'A')));
		ml.seq_table.push_back(new gnSequence(string((size_t)len1, 'A')));
		Match mm(2);
		for (uint64_t i = 0; i < n; ++i) {
			Match* m = mm.Copy();
			m->SetStart(0, rows[i].start0);
			m->SetStart(1, rows[i].start1);
			m->SetLength(rows[i].len);
			ml.push_back(m);
		}
		stringstream ss;
		WriteList(ml, ss);
		for (size_t i = 0; i < ml.size(); ++i) ml[i]->Free();
		delete ml.seq_table[0]; delete ml.seq_table[1];
		ml.seq_table.clear();
		string t = ss.str();
		char* out = (char*)malloc(t.size() + 1);
		memcpy(out, t.c_str(), t.size() + 1);
		return out;
	} catch (...) { return NULL; }
}

// ReadList of a text -> rows (malloc'd, free with ref_free); returns the row count or -1 when the reference rejects the text
long long ref_read_list(const char* text, ref_match** out)
{
	try {
		MatchList ml;
		stringstream ss(text);
		ReadList(ml, ss);
		size_t m = ml.size();
		ref_match* r = (ref_match*)malloc(sizeof(ref_match) * (m ? m : 1));
		for (size_t i = 0; i < m; ++i) {
			r[i].len = (int64_t)ml[i]->Length();
			r[i].start0 = ml[i]->Start(0);
			r[i].start1 = ml[i]->Start(1);
			ml[i]->Free();
		}
		*out = r;
		return (long long)m;
	} catch (...) { return -1; }
}

// ---- DP ------------------------------------------------------------------
// Global settings exactly as MuscleInterface::ProfileAlignFast (LM/MuscleInterface.cpp:1086-1106).
static void ref_dp_globals(unsigned nseq)
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

static void ref_msa_from_string(muscle::MSA& msa, const char* s, unsigned n, unsigned id)
{
	msa.SetSize(1, n);
	stringstream ss; ss << "seq" << id;
	msa.SetSeqName(0, ss.str().c_str());
	msa.SetSeqId(0, id);
	for (unsigned i = 0; i < n; ++i) msa.SetChar(0, i, s[i]);
}

// Aligns two ungapped DNA strings through ProfileProfile (-> GlobalAlign -> NWSmall -> BitTraceBack).
// path_out receives the edge types ('M','D','I') from first to last edge, must hold la+lb bytes.
// Returns the path length or -1.
long long ref_nw_align(const char* a, unsigned la, const char* b, unsigned lb, char* path_out)
{
	using namespace muscle;
	try {
		ref_dp_globals(2);
		MSA msa1, msa2;
		ref_msa_from_string(msa1, a, la, 0);
		ref_msa_from_string(msa2, b, lb, 1);
		msa1.FixAlpha();
		msa2.FixAlpha();
		SetPPScore(PPSCORE_SPN);
		// ProfileProfile = ProfileFromMSALocal x2 -> AlignTwoProfs -> GlobalAlign -> NWSmall -> BitTraceBack
		// -> AlignTwoMSAsGivenPath (MU/profile.cpp:68-93).  The path is read back from the two output rows.
		MSA msaOut;
		ProfileProfile(msa1, msa2, msaOut);
		unsigned n = msaOut.GetColCount();
		if (msaOut.GetSeqCount() != 2) return -1;
		unsigned ia = 0, ib = 1; // AlignTwoMSAsGivenPath emits msa1 rows then msa2 rows
		for (unsigned i = 0; i < n; ++i) {
			bool ga = msaOut.IsGap(ia, i), gb = msaOut.IsGap(ib, i);
			path_out[i] = (!ga && !gb) ? 'M' : (!ga ? 'D' : 'I');
		}
		return (long long)n;
	} catch (...) { return -1; }
}

// ---- HMM -----------------------------------------------------------------
// params_out (20 doubles): iStartHomologous, iGoHomologous, iGoUnrelated, iGoStopFromUnrelated,
// iGoStopFromHomologous, aEmitHomologous[8], aEmitUnrelated[8]  (homology.h:169-177 order)
void ref_hmm_params(double gc, double go_homologous, double go_unrelated, double pct_id, double* params_out)
{
	Params p = getAdaptedHoxdMatrixParameters(gc);
	if (go_homologous > 0) p.iGoHomologous = go_homologous;
	if (go_unrelated > 0) p.iGoUnrelated = go_unrelated;
	if (pct_id > 0) adaptToPercentIdentity(p, pct_id);
	params_out[0] = p.iStartHomologous; params_out[1] = p.iGoHomologous; params_out[2] = p.iGoUnrelated;
	params_out[3] = p.iGoStopFromUnrelated; params_out[4] = p.iGoStopFromHomologous;
	for (int i = 0; i < 8; ++i) { params_out[5 + i] = p.aEmitHomologous[i]; params_out[13 + i] = p.aEmitUnrelated[i]; }
}

static Params ref_params_from(const double* a)
{
	Params p;
	p.iStartHomologous = a[0]; p.iGoHomologous = a[1]; p.iGoUnrelated = a[2];
	p.iGoStopFromUnrelated = a[3]; p.iGoStopFromHomologous = a[4];
	for (int i = 0; i < 8; ++i) { p.aEmitHomologous[i] = a[5 + i]; p.aEmitUnrelated[i] = a[13 + i]; }
	return p;
}

// sym: string over '1'..'8'. pred_out: 'H'/'N' per column via the reference run().
// post_out (optional): posterior of "homologous" per column computed exactly as homologymain.cc:48.
int ref_hmm_run(const char* sym, uint64_t len, const double* params, char* pred_out, double* post_out)
{
	try {
		Params p = ref_params_from(params);
		string s(sym, len), pred;
		run(s, pred, p);
		memcpy(pred_out, pred.data(), len);
		if (post_out) {
			char* aSeq = new char[len];
			memcpy(aSeq, sym, len);
			HomologyDPTable *pFW, *pBW;
			HomologyBaumWelch bw;
			bfloat fw = Forward(&pFW, p, aSeq, (int)len);
			bfloat bwp = Backward(bw, pFW, &pBW, p, aSeq, (int)len);
			(void)bwp;
			for (uint64_t i = 0; i < len; ++i) {
				double post = pFW->getProb("homologous", (int)i + 1) * pBW->getProb("homologous", (int)i + 1) / fw;
				post_out[i] = post;
			}
			delete[] aSeq; delete pFW; delete pBW;
		}
		return 0;
	} catch (...) { return -1; }
}

} // extern "C"
