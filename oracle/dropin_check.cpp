// TEST INFRASTRUCTURE ONLY.
//
// Drop-in check: the reference's own classes (compiled in place by oracle/Makefile.ref) next to the C++ adapters of
// mauve_py_b200/adapters/, which forward to libmauve_cuda.so through the C ABI.  Each sub-command runs the SAME inputs
// through both and compares the results the reference's callers would see:
//
//   sml  <fasta> <weight> <rank>            DNAMemorySML vs CudaDNAMemorySML: Read() over the whole list
//   mums <a.fa> <b.fa> <weight> <rank> [memhash]   PairwiseMatchFinder / MemHash vs the Cuda* finders: MatchList rows, in order
//   gaps <pairs> <seed>                     the per-gap loop of pairwiseAnchorSearch (DNAMemorySML x2 + MemHash) vs one CudaGapSearchBatch call
//   dp   <regions> <seed>                   muscle::GlobalAlign on ProfileFromMSA profiles vs CudaGlobalAlignBatch: PWPath edges
//   hmm  <columns> <seed>                   run() vs run_cuda(): prediction strings
//
// Prints one "key value" line per measurement and "RESULT identical|DIFFERENT"; exit code 0 only when identical.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

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

#include "CudaDNAMemorySML.h"
#include "CudaMatchFinder.h"
#include "CudaGlobalAlign.h"
#include "CudaHomologyHMM.h"

using namespace std;
using namespace genome;
using namespace mems;

static double now_s()
{
	return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

struct Lcg {
	uint64_t s;
	explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ULL + 3037000493ULL) {}
	uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
	double unit() { return next() / 2147483648.0; }
};

static uint64 pick_seed(const vector<gnSequence*>& seqs, int weight, int rank)
{
	if (weight == 0) {
		gnSeqI total = 0;
		for (size_t i = 0; i < seqs.size(); ++i) total += seqs[i]->length();
		weight = getDefaultSeedWeight(total / seqs.size());   // LM/MatchList.h:269, LM/SeedMasks.h:389
	}
	return (uint64)getSeed(weight, rank);
}

// mers equal at every rank, positions equal as multisets inside every equal-mer run
static bool same_sml(SortedMerList& x, SortedMerList& y)
{
	if (x.SMLLength() != y.SMLLength()) return false;
	vector<bmer> vx, vy;
	x.Read(vx, x.SMLLength(), 0);
	y.Read(vy, y.SMLLength(), 0);
	if (vx.size() != vy.size()) return false;
	size_t i = 0;
	while (i < vx.size()) {
		size_t j = i;
		while (j < vx.size() && vx[j].mer == vx[i].mer) ++j;
		vector<gnSeqI> px, py;
		for (size_t k = i; k < j; ++k) {
			if (vy[k].mer != vx[k].mer) return false;
			px.push_back(vx[k].position);
			py.push_back(vy[k].position);
		}
		sort(px.begin(), px.end());
		sort(py.begin(), py.end());
		if (px != py) return false;
		i = j;
	}
	return true;
}

static int cmd_sml(int argc, char** argv)
{
	if (argc < 5) return 2;
	MatchList ml;
	ml.seq_filename.push_back(argv[2]);
	LoadSequences(ml, NULL);
	const uint64 seed = pick_seed(ml.seq_table, atoi(argv[3]), atoi(argv[4]));
	DNAMemorySML ref;
	CudaDNAMemorySML cu;
	double t0 = now_s();
	ref.Create(*ml.seq_table[0], seed);
	double t1 = now_s();
	cu.Create(*ml.seq_table[0], seed);
	double t2 = now_s();
	cu.Create(*ml.seq_table[0], seed);  // second call: device buffers are cached
	double t3 = now_s();
	const bool ok = same_sml(ref, cu) && cu[0].mer == ref[0].mer && cu.Seed() == ref.Seed() && cu.SeedLength() == ref.SeedLength();
	cout << "seed 0x" << hex << seed << dec << "\nsml_length " << ref.SMLLength() << "\nreference_s " << (t1 - t0) << "\ncuda_first_s " << (t2 - t1)
	     << "\ncuda_s " << (t3 - t2) << "\nRESULT " << (ok ? "identical" : "DIFFERENT") << endl;
	return ok ? 0 : 1;
}

template <class Finder, class Sml>
static double find_with(MatchList& ml, const vector<gnSequence*>& seqs, uint64 seed, bool memhash_settings, uint64& mems, uint64& collisions)
{
	ml.seq_table = seqs;
	for (size_t i = 0; i < seqs.size(); ++i) {
		ml.seq_filename.push_back("seq");
		Sml* sml = new Sml();
		sml->Create(*seqs[i], seed);
		ml.sml_table.push_back(sml);
	}
	Finder mf;
	if (memhash_settings) { mf.SetRepeatTolerance(0); mf.SetEnumerationTolerance(1); }   // LM/ProgressiveAligner.cpp:649-650
	const double t0 = now_s();
	mf.FindMatches(ml);
	const double t1 = now_s();
	mems = mf.MemCount();
	collisions = mf.MemCollisionCount();
	mf.Clear();
	return t1 - t0;
}

static int cmd_mums(int argc, char** argv)
{
	if (argc < 6) return 2;
	const bool memhash = argc > 6 && string(argv[6]) == "memhash";
	MatchList loader;
	loader.seq_filename.push_back(argv[2]);
	loader.seq_filename.push_back(argv[3]);
	LoadSequences(loader, NULL);
	const uint64 seed = pick_seed(loader.seq_table, atoi(argv[4]), atoi(argv[5]));
	MatchList ref, cu;
	uint64 rm = 0, rc = 0, cm = 0, cc = 0;
	double tr, tc;
	if (memhash) {
		tr = find_with<MemHash, DNAMemorySML>(ref, loader.seq_table, seed, true, rm, rc);
		tc = find_with<CudaMemHash, CudaDNAMemorySML>(cu, loader.seq_table, seed, true, cm, cc);
	} else {
		tr = find_with<PairwiseMatchFinder, DNAMemorySML>(ref, loader.seq_table, seed, false, rm, rc);
		tc = find_with<CudaPairwiseMatchFinder, CudaDNAMemorySML>(cu, loader.seq_table, seed, false, cm, cc);
	}
	bool ok = ref.size() == cu.size() && rm == cm;
	uint64_t sumlen = 0, nrev = 0;
	for (size_t i = 0; ok && i < ref.size(); ++i) {
		ok = ref[i]->Length() == cu[i]->Length() && ref[i]->Start(0) == cu[i]->Start(0) && ref[i]->Start(1) == cu[i]->Start(1) &&
		     cu[i]->SeqCount() == 2 && cu[i]->Multiplicity() == 2;
		sumlen += ref[i]->Length();
		nrev += ref[i]->Start(1) < 0;
	}
	ostringstream a, b;   // the text form the CLI writes with --mums (LM/MatchList.h:617-662), minus the pointer column
	cout << "seed 0x" << hex << seed << dec << "\nmatches_reference " << ref.size() << "\nmatches_cuda " << cu.size() << "\nsum_len " << sumlen
	     << "\nreverse " << nrev << "\nmem_count " << rm << " " << cm << "\ncollisions " << rc << " " << cc << "\nreference_find_s " << tr
	     << "\ncuda_find_s " << tc << "\nRESULT " << (ok ? "identical" : "DIFFERENT") << endl;
	return ok ? 0 : 1;
}

// ---- gap search: per-gap MemHash (the reference's loop) vs one CudaGapSearchBatch call ----
static int cmd_gaps(int argc, char** argv)
{
	if (argc < 4) return 2;
	const int n = atoi(argv[2]);
	Lcg rng(atoi(argv[3]));
	vector<pair<string, string> > gaps(n);
	for (int k = 0; k < n; ++k) {
		const unsigned la = 20 + (unsigned)(exp(rng.unit() * log(400.0)) * 20.0);   // 40 bp .. 8 kbp
		string a, b;
		for (unsigned i = 0; i < la; ++i) a += "ACGT"[rng.next() & 3];
		for (unsigned i = 0; i < la; ++i) {
			const double u = rng.unit();
			if (u < 0.005) continue;
			if (u < 0.01) b += "ACGT"[rng.next() & 3];
			b += u < 0.04 ? "ACGT"[rng.next() & 3] : a[i];
		}
		if (k % 4 == 0) {   // reverse complement
			string r(b.rbegin(), b.rend());
			for (size_t i = 0; i < r.size(); ++i) r[i] = r[i] == 'A' ? 'T' : r[i] == 'C' ? 'G' : r[i] == 'G' ? 'C' : 'A';
			b = r;
		}
		gaps[k] = make_pair(a, b);
	}
	double t0 = now_s();
	vector<vector<Match*> > cu;
	CudaGapSearchBatch(gaps, cu);
	double t1 = now_s();
	bool ok = true;
	size_t total = 0;
	MemHash gap_mh;
	gap_mh.SetRepeatTolerance(0);
	gap_mh.SetEnumerationTolerance(1);
	for (int k = 0; k < n; ++k) {   // LM/ProgressiveAligner.cpp:609-651
		MatchList gap_list;
		gap_list.seq_table.push_back(new gnSequence(gaps[k].first));
		gap_list.seq_table.push_back(new gnSequence(gaps[k].second));
		gap_list.sml_table.push_back(new DNAMemorySML());
		gap_list.sml_table.push_back(new DNAMemorySML());
		const gnSeqI avg_len = (gap_list.seq_table[0]->length() + gap_list.seq_table[1]->length()) / 2;
		const uint w = getDefaultSeedWeight(avg_len);
		gap_mh.Clear();
		if (w >= MIN_DNA_SEED_WEIGHT) {
			const uint64 seed = getSeed(w, 0);
			for (uint s = 0; s < 2; ++s) gap_list.sml_table[s]->Create(*gap_list.seq_table[s], seed);
			gap_mh.ClearSequences();
			gap_mh.FindMatches(gap_list);
		}
		if (gap_list.size() != cu[k].size()) ok = false;
		for (size_t i = 0; ok && i < gap_list.size(); ++i)
			ok = gap_list[i]->Length() == cu[k][i]->Length() && gap_list[i]->Start(0) == cu[k][i]->Start(0) && gap_list[i]->Start(1) == cu[k][i]->Start(1);
		total += gap_list.size();
		for (size_t i = 0; i < gap_list.size(); ++i) gap_list[i]->Free();
		for (size_t i = 0; i < cu[k].size(); ++i) cu[k][i]->Free();
		for (uint s = 0; s < 2; ++s) { delete gap_list.seq_table[s]; delete gap_list.sml_table[s]; }
	}
	double t2 = now_s();
	cout << "gaps " << n << "\nmatches " << total << "\ncuda_s " << (t1 - t0) << "\nreference_s " << (t2 - t1) << "\nRESULT " << (ok ? "identical" : "DIFFERENT") << endl;
	return ok ? 0 : 1;
}

// ---- DP: globals exactly as MuscleInterface::ProfileAlignFast (LM/MuscleInterface.cpp:1086-1106) ----
static void dp_globals()
{
	using namespace muscle;
	g_SeqType.get() = SEQTYPE_DNA;
	g_uMaxIters.get() = 1;
	g_bStable.get() = true;
	g_bQuiet.get() = true;
	g_SeqWeight1.get() = SEQWEIGHT_ClustalW;
	SetMaxIters(g_uMaxIters.get());
	SetSeqWeightMethod(g_SeqWeight1.get());
	MSA::SetIdCount(2);
	SetAlpha(ALPHA_DNA);
	SetPPScore(PPSCORE_SPN);
}

namespace muscle { bool TreeNeededForWeighting(SEQWEIGHT s); }   // MU/profile.cpp:10 (not in a header)

static muscle::ProfPos* profile_of(const string& s, unsigned id)
{
	using namespace muscle;
	MSA msa;
	msa.SetSize(1, (unsigned)s.size());
	msa.SetSeqName(0, id ? "b" : "a");
	msa.SetSeqId(0, 0);   // ProfileFromMSALocal renumbers ids per profile (MU/profile.cpp:25-26)
	for (unsigned i = 0; i < s.size(); ++i) msa.SetChar(0, i, s[i]);
	msa.FixAlpha();
	Tree tree;   // as ProfileFromMSALocal (MU/profile.cpp:22-34): sequence weights come from the (one-leaf) tree
	if (TreeNeededForWeighting(g_SeqWeight2.get())) {
		TreeFromMSA(msa, tree, g_Cluster2.get(), g_Distance2.get(), g_Root1.get());
		SetMuscleTree(tree);
	}
	return ProfileFromMSA(msa);
}

static int cmd_dp(int argc, char** argv)
{
	using namespace muscle;
	if (argc < 4) return 2;
	const int n = atoi(argv[2]);
	Lcg rng(atoi(argv[3]));
	dp_globals();
	vector<string> A(n), B(n);
	for (int k = 0; k < n; ++k) {
		const unsigned la = 1 + rng.next() % (k % 7 == 0 ? 1500 : 300);
		for (unsigned i = 0; i < la; ++i) A[k] += "ACGT"[rng.next() & 3];
		for (unsigned i = 0; i < la; ++i) {   // B = A with substitutions and indels
			const double u = rng.unit();
			if (u < 0.02) continue;
			if (u < 0.04) B[k] += "ACGT"[rng.next() & 3];
			B[k] += u < 0.12 ? "ACGT"[rng.next() & 3] : A[k][i];
		}
		if (B[k].empty()) B[k] = "A";
	}
	vector<ProfPos*> PA(n), PB(n);
	vector<CudaDPRange> ranges(n);
	for (int k = 0; k < n; ++k) {
		PA[k] = profile_of(A[k], 0);
		PB[k] = profile_of(B[k], 1);
		CudaDPRange r = {PA[k], (unsigned)A[k].size(), PB[k], (unsigned)B[k].size()};
		ranges[k] = r;
	}
	PWPath* cu = new PWPath[n];
	vector<bool> handled;
	double t0 = now_s();
	CudaGlobalAlignBatch(ranges, cu, handled);
	double t1 = now_s();
	bool ok = true;
	double cells = 0;
	for (int k = 0; k < n; ++k) {
		PWPath ref;
		GlobalAlign(PA[k], (unsigned)A[k].size(), PB[k], (unsigned)B[k].size(), ref);   // mutates the terminal gap scores, as in the reference
		cells += (double)A[k].size() * B[k].size();
		if (!handled[k] || ref.GetEdgeCount() != cu[k].GetEdgeCount()) { ok = false; continue; }
		for (unsigned e = 0; e < ref.GetEdgeCount(); ++e)
			if (!ref.GetEdge(e).Equal(cu[k].GetEdge(e))) { ok = false; break; }
	}
	double t2 = now_s();
	delete[] cu;
	cout << "regions " << n << "\ncells " << cells << "\ncuda_s " << (t1 - t0) << "\nreference_s " << (t2 - t1) << "\nRESULT " << (ok ? "identical" : "DIFFERENT") << endl;
	return ok ? 0 : 1;
}

static int cmd_hmm(int argc, char** argv)
{
	if (argc < 4) return 2;
	const int len = atoi(argv[2]);
	Lcg rng(atoi(argv[3]));
	string s(len, '1');
	bool homologous = true;
	for (int i = 0; i < len; ++i) {   // blocks that look homologous (mostly matches) and unrelated (mostly mismatches / gaps)
		if (rng.next() % 400 == 0) homologous = !homologous;
		const double u = rng.unit();
		if (homologous) s[i] = u < 0.45 ? '1' : u < 0.9 ? '2' : (char)('3' + rng.next() % 6);
		else s[i] = u < 0.2 ? (char)('1' + rng.next() % 2) : (char)('3' + rng.next() % 6);
	}
	Params p = getAdaptedHoxdMatrixParameters(0.5);
	p.iGoHomologous = 1e-5;    // CLI defaults, MA/progressiveMauve.cpp:321-323
	p.iGoUnrelated = 1e-9;
	adaptToPercentIdentity(p, 0.7);
	string ref, sc = s;
	double t0 = now_s();
	run(s, ref, p);
	double t1 = now_s();
	vector<string> in(1, sc), out;
	vector<vector<double> > post;
	run_cuda_batch(in, out, p, &post);
	double t2 = now_s();
	const string& cu = out[0];
	// the prediction is posterior >= 0.9 (LM/HomologyHMM/homologymain.cc:50); posteriors agree to 1e-5 relative (north_star), so a
	// column whose posterior sits within that distance of the threshold may legitimately fall on either side
	size_t diff = 0, at_threshold = 0;
	vector<size_t> flips;
	for (size_t i = 0; i < ref.size() && i < cu.size(); ++i) {
		if (ref[i] == cu[i]) continue;
		if (fabs(post[0][i] - 0.9) <= 0.9 * 1e-5) ++at_threshold; else ++diff;
		flips.push_back(i);
	}
	double max_rel = 0;
	{   // the reference's own posteriors (LM/HomologyHMM/homologymain.cc:36-48) against the device's
		vector<char> aSeq(s.begin(), s.end());
		HomologyDPTable *pFW, *pBW;
		HomologyBaumWelch bw;
		bfloat fw = Forward(&pFW, p, &aSeq[0], len);
		Backward(bw, pFW, &pBW, p, &aSeq[0], len);
		for (int i = 0; i < len; ++i) {
			const double rp = pFW->getProb("homologous", i + 1) * pBW->getProb("homologous", i + 1) / fw;
			if (rp > 1e-6) max_rel = max(max_rel, fabs(post[0][i] - rp) / rp);
		}
		for (size_t k = 0; k < flips.size() && k < 8; ++k) {
			const size_t i = flips[k];
			const double rp = pFW->getProb("homologous", (int)i + 1) * pBW->getProb("homologous", (int)i + 1) / fw;
			cout.precision(12);
			cout << "flip " << i << " reference " << rp << " cuda " << post[0][i] << endl;
		}
		delete pFW; delete pBW;
	}
	cout << "max_rel_err " << max_rel << endl;
	const bool ok = ref.size() == cu.size() && diff == 0;
	cout << "columns " << len << "\nhomologous_columns " << count(ref.begin(), ref.end(), 'H') << "\ndiffering_columns " << diff << "\nthreshold_columns " << at_threshold << "\nreference_s " << (t1 - t0)
	     << "\ncuda_s " << (t2 - t1) << "\nRESULT " << (ok ? "identical" : "DIFFERENT") << endl;
	return ok ? 0 : 1;
}

int main(int argc, char** argv)
{
	if (argc < 2) { cerr << "usage: dropin_check sml|mums|dp|hmm ..." << endl; return 2; }
	if (mcu_init(0) != MCU_OK) { cerr << "no CUDA device: " << mcu_last_error() << endl; return 3; }
	try {
		const string c = argv[1];
		if (c == "sml") return cmd_sml(argc, argv);
		if (c == "mums") return cmd_mums(argc, argv);
		if (c == "gaps") return cmd_gaps(argc, argv);
		if (c == "dp") return cmd_dp(argc, argv);
		if (c == "hmm") return cmd_hmm(argc, argv);
	} catch (const char* msg) {
		cerr << "exception: " << msg << endl;
		return 4;
	} catch (std::exception& e) {
		cerr << "exception: " << e.what() << endl;
		return 4;
	}
	return 2;
}
