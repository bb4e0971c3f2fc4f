// Link-time seam: the pairwise homology HMM's run() on the GPU.
//
// findHssHomologyHMM (LM/Islands.h:124-194) encodes the columns of a pairwise alignment as a string over '1'..'8' and calls
// void run(std::string& sequence, std::string& prediction, const Params&) (LM/HomologyHMM/homologymain.cc:24-62) once per pair of
// genomes per backbone pass: ONE string as long as the alignment (3,983,034 columns for the MDS42 pair).  The reference's definition
// keeps a renamed symbol (objcopy on a copy of homologymain.o, oracle/Makefile.ref); this definition takes its name and sends the
// string through mcu_hmm_batch (csrc/hmm.cu: Forward and Backward as two warp chains in the FP32 form of the bfloat recurrence,
// posteriors bit-identical to the reference's).  MAUVE_CUDA_HMM_SEAM=0 in the environment leaves the call with the reference's code.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "CudaHomologyHMM.h"

void run_reference(std::string& sequence, std::string& prediction, const Params& params) asm("_Z13run_referenceRNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES5_RK6Params");

namespace {
unsigned long long g_hmm_device = 0, g_hmm_reference = 0, g_hmm_columns = 0;
struct HmmSeamReport {
	~HmmSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
			fprintf(stderr, "run() seam (HomologyHMM): %llu strings (%llu columns) on the device, %llu in the reference's code\n", g_hmm_device, g_hmm_columns,
			        g_hmm_reference);
	}
};
HmmSeamReport g_hmm_report;
}  // namespace

void run(std::string& sequence, std::string& prediction, const Params& params)
{
	static const bool off = getenv("MAUVE_CUDA_HMM_SEAM") && getenv("MAUVE_CUDA_HMM_SEAM")[0] == '0';
	if (off || sequence.empty()) {   // the reference's run() on an empty string is its own business (it indexes column 0)
		++g_hmm_reference;
		run_reference(sequence, prediction, params);
		return;
	}
	++g_hmm_device;
	g_hmm_columns += sequence.size();
	run_cuda(sequence, prediction, params);
}
