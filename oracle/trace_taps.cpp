// TEST INFRASTRUCTURE ONLY.
//
// Link-time taps on the three seams of the hot path inside the UNMODIFIED reference binary, for minting fixtures from the REAL
// pipeline and for knowing what the callers really ask for.  The object files that define the tapped functions are copied with
// that one symbol renamed (objcopy, oracle/Makefile.ref); this file supplies the functions under their original names, forwards
// to the renamed originals and records inputs and results.  No source of the reference is touched, results are unchanged.
//
//   muscle::GlobalAlign (MU/glbalign.cpp:69-81)      $MAUVE_DP_TRACE : per call "A B path" (letter strings when every column is one
//                                                    ungapped ACGT letter, the case mcu_nw_batch covers; else "- lenA lenB")
//   mems::MemHash::FindMatches (LM/MemHash.cpp:109)  $MAUVE_MH_TRACE : per call "@ class repeat_tol enum_tol nseq seed len0 len1 nmatch",
//                                                    the sequences (two genomes, each <= 2 Mbp) and the rows "len start0 start1"
//   run() (LM/HomologyHMM/homologymain.cc:24)        $MAUVE_HMM_TRACE: per call "columns sequence prediction" + the 21 parameters
#include <cstdio>
#include <cstdlib>
#include <string>
#include <typeinfo>

#include "libGenome/gnSequence.h"
#include "libMems/MemHash.h"
#include "libMems/MatchList.h"
#include "libMUSCLE/muscle.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"
#include "homology.h"

static FILE* tap_file(const char* env, FILE*& f, bool& tried)
{
	if (!tried) {
		tried = true;
		const char* p = getenv(env);
		if (p) f = fopen(p, "w");
	}
	return f;
}

// ---- MemHash::FindMatches ----
namespace mems {
void MemHash_FindMatches_reference(MemHash* self, MatchList& ml)
    asm("_ZN4mems7MemHash21FindMatches_referenceERNS_16GenericMatchListIPNS_22UngappedLocalAlignmentINS_19HybridAbstractMatchILj2ESaIxESaIjEEEEEEE");

void MemHash::FindMatches(MatchList& ml)
{
	static FILE* f = NULL;
	static bool tried = false;
	tap_file("MAUVE_MH_TRACE", f, tried);
	std::string s0, s1;
	const size_t nseq = ml.seq_table.size();
	unsigned long long seed = nseq && ml.sml_table.size() ? ml.sml_table[0]->Seed() : 0;
	unsigned long long l0 = nseq > 0 ? ml.seq_table[0]->length() : 0, l1 = nseq > 1 ? ml.seq_table[1]->length() : 0;
	const bool keep = f && nseq == 2 && l0 <= 2000000 && l1 <= 2000000;
	if (keep) { s0 = ml.seq_table[0]->ToString(); s1 = ml.seq_table[1]->ToString(); }
	const unsigned rt = m_repeat_tolerance, et = m_enumeration_tolerance;
	MemHash_FindMatches_reference(this, ml);
	if (f) {
		fprintf(f, "@ %s %u %u %zu %llu %llu %llu %zu\n", typeid(*this).name(), rt, et, nseq, seed, l0, l1, ml.size());
		if (keep) {
			fprintf(f, "%s\n%s\n", s0.c_str(), s1.c_str());
			for (size_t i = 0; i < ml.size(); ++i) fprintf(f, "%llu %lld %lld\n", (unsigned long long)ml[i]->Length(), (long long)ml[i]->Start(0), (long long)ml[i]->Start(1));
		}
		fflush(f);
	}
}
}  // namespace mems

// ---- run() ----
void run_reference(std::string& sequence, std::string& prediction, const Params& params) asm("_Z13run_referenceRNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES5_RK6Params");

void run(std::string& sequence, std::string& prediction, const Params& params)
{
	static FILE* f = NULL;
	static bool tried = false;
	tap_file("MAUVE_HMM_TRACE", f, tried);
	run_reference(sequence, prediction, params);
	if (f) {
		fprintf(f, "%zu %s %s", sequence.size(), sequence.c_str(), prediction.c_str());
		fprintf(f, " %.17g %.17g %.17g %.17g %.17g", params.iStartHomologous, params.iGoHomologous, params.iGoUnrelated, params.iGoStopFromUnrelated, params.iGoStopFromHomologous);
		for (int i = 0; i < 8; ++i) fprintf(f, " %.17g", params.aEmitHomologous[i]);
		for (int i = 0; i < 8; ++i) fprintf(f, " %.17g", params.aEmitUnrelated[i]);
		fprintf(f, "\n");
		fflush(f);
	}
}

namespace muscle {

SCORE GlobalAlign_reference(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
    asm("_ZN6muscle21GlobalAlign_referenceEPKNS_7ProfPosEjS2_jRNS_6PWPathE");

static bool letters(const ProfPos* P, unsigned n, std::string& out)
{
	out.resize(n);
	for (unsigned i = 0; i < n; ++i) {
		const unsigned u = P[i].m_uSortOrder[0];
		if (P[i].m_bAllGaps || u >= 4 || P[i].m_fcCounts[u] != 1.0f) return false;
		out[i] = "ACGT"[u];
	}
	return n > 0;
}

SCORE GlobalAlign(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
{
	static FILE* f = NULL;
	static bool tried = false;
	if (!tried) {
		tried = true;
		const char* p = getenv("MAUVE_DP_TRACE");
		if (p) f = fopen(p, "w");
	}
	std::string a, b;
	const bool ok = f && letters(PA, uLengthA, a) && letters(PB, uLengthB, b);   // read before SetTermGaps touches the profiles
	const SCORE s = GlobalAlign_reference(PA, uLengthA, PB, uLengthB, Path);
	if (f) {
		if (ok) {
			std::string path(Path.GetEdgeCount(), '?');
			for (unsigned e = 0; e < Path.GetEdgeCount(); ++e) path[e] = Path.GetEdge(e).cType;
			fprintf(f, "%s %s %s\n", a.c_str(), b.c_str(), path.c_str());
		} else
			fprintf(f, "- %u %u\n", uLengthA, uLengthB);
		fflush(f);
	}
	return s;
}

}  // namespace muscle
