// Link-time seam: mems::SeedOccurrenceList::construct<SortedMerList> on the GPU.
//
// ProgressiveAligner::align builds one seed occurrence list per genome (LM/ProgressiveAligner.cpp:3908-3912) through the member
// template SeedOccurrenceList::construct (LM/SeedOccurrenceList.h:22-78); its instantiation for SortedMerList is a weak symbol of
// ProgressiveAligner.o, so the explicit specialization below -- a strong definition of the same symbol -- is the one the linker
// keeps.  It fills the object exactly as the reference does (temporary file of float32, memory mapped: getFrequency and the
// destructor are untouched) with the values mcu_sol_build computes (csrc/sol.cu).  1.5 s of the 8 s of host work that remain
// for the MDS42 pair once the DP and the match finder are on the device (gprof: two calls, one GetSeedMer per rank each).
// Opt-in with MAUVE_CUDA_SOL_SEAM=1 until the kernels have run on a GPU; otherwise the reference's own template body runs,
// instantiated for a forwarding view of the list.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "libMems/SortedMerList.h"
#include "libMems/SeedOccurrenceList.h"

#include "CudaSeedOccurrenceList.h"

namespace mems {

namespace {
// what construct<> needs from a sorted mer list, forwarded: lets the reference's template body run under another type name
struct SmlView {
	SortedMerList& s;
	explicit SmlView(SortedMerList& sml) : s(sml) {}
	gnSeqI Length() const { return s.Length(); }
	gnSeqI SMLLength() const { return s.SMLLength(); }
	uint64 GetSeedMask() const { return s.GetSeedMask(); }
	uint SeedLength() const { return s.SeedLength(); }
	bmer operator[](gnSeqI i) { return s[i]; }
};
unsigned long long g_sol_device = 0, g_sol_reference = 0;
struct SolSeamReport {
	~SolSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
			fprintf(stderr, "SeedOccurrenceList::construct seam: %llu lists on the device, %llu in the reference's code\n", g_sol_device, g_sol_reference);
	}
};
SolSeamReport g_sol_report;
}  // namespace

template <>
void SeedOccurrenceList::construct<SortedMerList>(SortedMerList& sml)
{
	static const bool on = getenv("MAUVE_CUDA_SOL_SEAM") && getenv("MAUVE_CUDA_SOL_SEAM")[0] == '1';
	if (!on) {
		++g_sol_reference;
		SmlView view(sml);
		construct(view);
		return;
	}
	++g_sol_device;
	std::vector<frequency_type> count;
	cuda_detail::SolFrequencies(sml, count);
	// LM/SeedOccurrenceList.h:67-77
	tmpfile = CreateTempFileName("sol");
	{
		std::ofstream tfout;
		tfout.open(tmpfile.c_str(), std::ios::binary);
		tfout.write((const char*)&count[0], sml.Length() * sizeof(frequency_type));
		tfout.close();
	}
	data.close();
	data.open(tmpfile);
}

}  // namespace mems
