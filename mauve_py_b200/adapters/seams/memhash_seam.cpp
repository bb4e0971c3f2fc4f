// Link-time seam: mems::MemHash::FindMatches with seed-match enumeration + extension on the GPU for two genomes.
//
// The reference-side binding for the match finder when the call sites cannot be edited: LM/MemHash.cpp is compiled unchanged,
// the definition of MemHash::FindMatches is weakened in a COPY of the object file and given a second name (objcopy,
// oracle/Makefile.ref), and this file supplies the function under its original name -- virtual calls land here because the
// vtable refers to that name.  What goes to the device (through the same code as the CudaPairwiseMatchFinder / CudaMemHash
// adapters, CudaMatchFinder.h -> mcu_find_mums):
//   * `PairwiseMatchFinder pmf` of progressiveMauve (MA/progressiveMauve.cpp:500-503): the initial anchoring of two genomes;
//   * plain MemHash objects with the MUM tolerances 0 / 1 (gap_mh of recursive anchoring, LM/ProgressiveAligner.cpp:643-651;
//     MAUVE_CUDA_GAP_SEAM=0 keeps them in the reference's code).  The caller must take its matches from the list FindMatches fills, which
//     is what pairwiseAnchorSearch does unless --seed-family is given (then it reads the hash table afterwards, :652-654; a process
//     started with --seed-family keeps its gap searches in the reference's code whatever the variable says).
// Everything else (more genomes, other tolerances, other subclasses, sequences shorter than the seed) runs the reference's code.
// MAUVE_CUDA_MH_SEAM=0 switches the seam off.  A device failure throws, as the adapters do; nothing on this path catches it.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <typeinfo>

#include "libGenome/gnSequence.h"
#include "libMems/MemHash.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/MatchList.h"

#include "CudaMatchFinder.h"

namespace mems {

void MemHash_FindMatches_reference(MemHash* self, MatchList& ml)
    asm("_ZN4mems7MemHash21FindMatches_referenceERNS_16GenericMatchListIPNS_22UngappedLocalAlignmentINS_19HybridAbstractMatchILj2ESaIxESaIjEEEEEEE");

// true when the process was started with --seed-family: pairwiseAnchorSearch then calls FindMatches three times per gap and reads the
// matches from the hash table afterwards (LM/ProgressiveAligner.cpp:643-654), which the device path does not fill
static bool started_with_seed_family()
{
	FILE* f = fopen("/proc/self/cmdline", "rb");
	if (!f) return false;
	std::string all;
	char buf[4096];
	size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
	fclose(f);
	for (size_t i = 0; i < all.size();) {   // NUL-separated arguments
		const std::string arg(all.c_str() + i);
		if (arg == "--seed-family" || arg.compare(0, 14, "--seed-family=") == 0) return true;
		i += arg.size() + 1;
	}
	return false;
}

static unsigned long long g_mh_device = 0, g_mh_reference = 0;
struct MemHashSeamReport {
	~MemHashSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
			fprintf(stderr, "MemHash::FindMatches seam: %llu calls on the device, %llu in the reference's code\n", g_mh_device, g_mh_reference);
	}
};
static MemHashSeamReport g_mh_report;

void MemHash::FindMatches(MatchList& ml)
{
	static const bool off = getenv("MAUVE_CUDA_MH_SEAM") && getenv("MAUVE_CUDA_MH_SEAM")[0] == '0';
	static const bool gaps = !(getenv("MAUVE_CUDA_GAP_SEAM") && getenv("MAUVE_CUDA_GAP_SEAM")[0] == '0') && !started_with_seed_family();
	int rule = -1;
	if (!off && ml.seq_table.size() == 2 && ml.sml_table.size() == 2 && ml.sml_table[0]->Seed() == ml.sml_table[1]->Seed()) {
		const gnSeqI L = ml.sml_table[0]->SeedLength();
		if (ml.seq_table[0]->length() >= L && ml.seq_table[1]->length() >= L && L > 0) {
			if (typeid(*this) == typeid(PairwiseMatchFinder)) rule = MCU_RULE_PAIRWISE;
			else if (gaps && typeid(*this) == typeid(MemHash) && m_repeat_tolerance == 0 && m_enumeration_tolerance == 1) rule = MCU_RULE_MEMHASH;
		}
	}
	if (rule < 0) {
		++g_mh_reference;
		MemHash_FindMatches_reference(this, ml);
		return;
	}
	++g_mh_device;
	cuda_detail::FindMatchesTwoGenomes(ml, rule, m_mem_count, m_collision_count);
}

}  // namespace mems
