// Link-time seam: muscle::AnchoredProfileProfile with the gapped DP of all its ranges on the GPU in ONE call.
//
// This is the reference-side binding for the gapped DP: the reference's MU/anchoredpp.cpp is compiled unchanged,
// its one symbol `muscle::AnchoredProfileProfile` is renamed in a COPY of the object file (objcopy, oracle/Makefile.ref) and this
// file supplies the function under the original name.  Everything it calls is the reference's own code (PrepareMSAforScoring,
// FindAnchorColsPP, ColsToRanges, MSAFromColRange, StripGapColumns, ProfileFromMSA, AlignTwoMSAsGivenPath, MSAAppend), in the
// reference's order; the one change is that the per-range chain ProfileProfile -> AlignTwoProfs -> GlobalAlign -> NWSmall
// (MU/profile.cpp:68-93, MU/aligntwoprofs.cpp:23, MU/glbalign.cpp:69-81, MU/nwsmall.cpp:500-670) is split: profiles of ALL ranges
// first, one CudaGlobalAlignBatch call (mcu_nw_batch), then AlignTwoMSAsGivenPath per range.  The profile AlignTwoProfsGivenPath
// builds after every DP (`ProfOut`, MU/profile.cpp:86,92) is deleted unused by the reference and is not built here.
// A range the integer kernel does not cover (a profile column that is not one ungapped ACGT letter, an empty side) takes the
// reference's ProfileProfile unchanged (with MAUVE_CUDA_WILD=1 ranges with wildcard columns go to the float kernel,
// mcu_nw_batch_wild, first).  MAUVE_CUDA_DP_SEAM=0 sends every call to the reference's function.
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <vector>

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/tree.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"
#include "libMUSCLE/refine.h"

#include "CudaGlobalAlign.h"
#include "CudaAnchorCols.h"

namespace muscle {

// the reference's function under its link-time name (objcopy --redefine-sym on a copy of anchoredpp.o)
void AnchoredProfileProfile_reference(MSA& msa1, MSA& msa2, MSA& msaOut) asm("_ZN6muscle32AnchoredProfileProfile_referenceERNS_3MSAES1_S1_");

void PrepareMSAforScoring(MSA& msa);
void FindAnchorColsPP(const MSA& msa1, const MSA& msa2, unsigned AnchorCols[], unsigned* ptruAnchorColCount);
void StripGapColumns(MSA& msa);
bool TreeNeededForWeighting(SEQWEIGHT s);

// static ProfileFromMSALocal of MU/profile.cpp:22-33
static ProfPos* ProfileOf(MSA& msa, Tree& tree)
{
	const unsigned uSeqCount = msa.GetSeqCount();
	for (unsigned uSeqIndex = 0; uSeqIndex < uSeqCount; ++uSeqIndex) msa.SetSeqId(uSeqIndex, uSeqIndex);
	if (TreeNeededForWeighting(g_SeqWeight2.get())) {
		TreeFromMSA(msa, tree, g_Cluster2.get(), g_Distance2.get(), g_Root1.get());
		SetMuscleTree(tree);
	}
	return ProfileFromMSA(msa);
}

static unsigned long long g_seam_calls = 0, g_seam_ranges = 0, g_seam_device = 0, g_seam_cols_device = 0, g_seam_cols = 0;
struct SeamReport {
	~SeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
		{
			fprintf(stderr, "AnchoredProfileProfile seam: %llu calls, %llu ranges, %llu aligned on the device\n", g_seam_calls, g_seam_ranges, g_seam_device);
			fprintf(stderr, "FindAnchorColsPP seam: %llu windows on the device, %llu anchor columns\n", g_seam_cols_device, g_seam_cols);
		}
	}
};
static SeamReport g_seam_report;

void AnchoredProfileProfile(MSA& msa1, MSA& msa2, MSA& msaOut)
{
	static const bool off = getenv("MAUVE_CUDA_DP_SEAM") && getenv("MAUVE_CUDA_DP_SEAM")[0] == '0';
	if (off) { AnchoredProfileProfile_reference(msa1, msa2, msaOut); return; }
	++g_seam_calls;

	// MU/anchoredpp.cpp:446-464
	const unsigned uColCountIn = msa1.GetColCount();
	const unsigned uSeqCountIn = msa1.GetSeqCount() + msa2.GetSeqCount();
	unsigned* AnchorCols = new unsigned[uColCountIn];
	unsigned uAnchorColCount;
	PrepareMSAforScoring(msa1);
	PrepareMSAforScoring(msa2);
	// the anchor columns of the window on the device (mcu_anchor_cols_batch; MAUVE_CUDA_COLS_SEAM=0: the reference's FindAnchorColsPP)
	static const bool cols_off = getenv("MAUVE_CUDA_COLS_SEAM") && getenv("MAUVE_CUDA_COLS_SEAM")[0] == '0';
	bool cols_done = false;
	if (!cols_off) {
		try {
			cols_done = CudaFindAnchorColsPP(msa1, msa2, AnchorCols, &uAnchorColCount);
		} catch (std::exception& e) {
			fprintf(stderr, "\n*** FATAL: %s\n", e.what());   // see below: the caller would swallow the exception
			exit(3);
		}
		if (cols_done) { ++g_seam_cols_device; g_seam_cols += uAnchorColCount; }
	}
	if (!cols_done) FindAnchorColsPP(msa1, msa2, AnchorCols, &uAnchorColCount);
	const unsigned uRangeCount = uAnchorColCount + 1;
	Range* Ranges = new Range[uRangeCount];
	ColsToRanges(AnchorCols, uAnchorColCount, uColCountIn, Ranges);
	ListVertSavings(uColCountIn, uAnchorColCount, Ranges, uRangeCount);
	delete[] AnchorCols;

	// :483-502
	msaOut.SetSize(uSeqCountIn, 0);
	for (unsigned uSeqIndex = 0; uSeqIndex < uSeqCountIn; ++uSeqIndex) {
		const char* ptrName;
		if (uSeqIndex < msa1.GetSeqCount()) {
			msa1.SetSeqId(uSeqIndex, uSeqIndex);
			ptrName = msa1.GetSeqName(uSeqIndex);
		} else {
			msa2.SetSeqId(uSeqIndex - msa1.GetSeqCount(), uSeqIndex);
			ptrName = msa2.GetSeqName(uSeqIndex - msa1.GetSeqCount());
		}
		msaOut.SetSeqName(uSeqIndex, ptrName);
		msaOut.SetSeqId(uSeqIndex, uSeqIndex);
	}

	// :504-549, phase 1: the sub-alignments of every non-empty range, in range order.  For two genomes both are one-sequence alignments
	// without gap columns: their letters go to the device as they are -- no profile is built for them (MAUVE_CUDA_DP_PROFILES=1 takes
	// the round-1 route through ProfileFromMSA and the per-column check of the scoring, for A/B runs)
	static const bool via_profiles = getenv("MAUVE_CUDA_DP_PROFILES") && getenv("MAUVE_CUDA_DP_PROFILES")[0] == '1';
	struct Job {
		MSA* m1; MSA* m2; ProfPos* P1; ProfPos* P2; bool device;
	};
	std::vector<Job> jobs;
	std::vector<CudaDPRange> dp;
	std::vector<std::pair<const MSA*, const MSA*> > rows;
	std::vector<size_t> dp_job;
	for (unsigned uRangeIndex = 0; uRangeIndex < uRangeCount; ++uRangeIndex) {
		const Range& r = Ranges[uRangeIndex];
		const unsigned uFromColIndex = r.m_uBestColLeft;
		const unsigned uRangeColCount = r.m_uBestColRight - uFromColIndex;
		if (0 == uRangeColCount) continue;
		Job j;
		j.m1 = new MSA();
		j.m2 = new MSA();
		j.P1 = j.P2 = 0;
		j.device = false;
		MSAFromColRange(msa1, uFromColIndex, uRangeColCount, *j.m1);
		MSAFromColRange(msa2, uFromColIndex, uRangeColCount, *j.m2);
		StripGapColumns(*j.m1);
		StripGapColumns(*j.m2);
		const unsigned l1 = j.m1->GetColCount(), l2 = j.m2->GetColCount();
		if (l1 > 0 && l2 > 0 && j.m1->GetSeqCount() == 1 && j.m2->GetSeqCount() == 1) {   // the two-genome form the kernels cover
			if (via_profiles) {
				Tree tree1, tree2;
				j.P1 = ProfileOf(*j.m1, tree1);
				j.P2 = ProfileOf(*j.m2, tree2);
				CudaDPRange d;
				d.PA = j.P1; d.uLengthA = l1; d.PB = j.P2; d.uLengthB = l2;
				dp.push_back(d);
			} else
				rows.push_back(std::make_pair((const MSA*)j.m1, (const MSA*)j.m2));
			dp_job.push_back(jobs.size());
		}
		jobs.push_back(j);
	}
	g_seam_ranges += jobs.size();

	// phase 2: every DP of this window in one device call
	const size_t n_dp = dp_job.size();
	PWPath* paths = n_dp == 0 ? 0 : new PWPath[n_dp];
	std::vector<bool> handled;
	if (n_dp) {
		try {
			// ranges with N / X columns go to mcu_nw_batch_wild (MAUVE_CUDA_WILD=0: to the reference's NWSmall, for A/B runs)
			static const bool wild = !(getenv("MAUVE_CUDA_WILD") && getenv("MAUVE_CUDA_WILD")[0] == '0');
			if (via_profiles) CudaGlobalAlignBatch(dp, paths, handled, NULL, wild);
			else CudaGlobalAlignBatchRows(rows, paths, handled, wild);
		} catch (std::exception& e) {
			// MuscleInterface::ProfileAlignFast swallows every exception (LM/MuscleInterface.cpp:1155-1159) and the aligner would go on
			// without this window: a device failure must stop the run, the way MUSCLE's own Quit() does
			fprintf(stderr, "\n*** FATAL: %s\n", e.what());
			exit(3);
		}
	}
	std::vector<long> path_of(jobs.size(), -1);
	for (size_t k = 0; k < n_dp; ++k)
		if (handled[k]) { jobs[dp_job[k]].device = true; path_of[dp_job[k]] = (long)k; ++g_seam_device; }

	// phase 3: output blocks in range order
	for (size_t i = 0; i < jobs.size(); ++i) {
		Job& j = jobs[i];
		MSA msaRangeOut;
		if (j.device) AlignTwoMSAsGivenPath(paths[path_of[i]], *j.m1, *j.m2, msaRangeOut);
		else ProfileProfile(*j.m1, *j.m2, msaRangeOut);   // the reference's chain, unchanged
		for (unsigned uSeqIndex = 0; uSeqIndex < uSeqCountIn; ++uSeqIndex) msaRangeOut.SetSeqId(uSeqIndex, uSeqIndex);
		MSAAppend(msaOut, msaRangeOut);
		delete[] j.P1;
		delete[] j.P2;
		delete j.m1;
		delete j.m2;
	}
	delete[] paths;
	delete[] Ranges;
}

}  // namespace muscle
