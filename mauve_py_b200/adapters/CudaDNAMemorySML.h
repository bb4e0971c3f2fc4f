// CudaDNAMemorySML -- drop-in for mems::DNAMemorySML whose Create() builds the sorted mer list on the GPU.
//
// Host side of the boundary in the reference's own language (C++), compiled against the reference's headers
// (libMems/DNAMemorySML.h) and linked with libmauve_cuda.so (include/mauve_cuda.h).  Replaces
// MemorySML::Create (LM/MemorySML.cpp:45-60): FillDnaSeedSML / FillSML + std::sort(bmer_lessthan) + position copy.
// Everything else (Read, operator[], FindMer, GetSeedMer, header, packed `sequence`) is inherited unchanged, so the
// callers -- MatchList::CreateMemorySMLs (LM/MatchList.h:451), pairwiseAnchorSearch (LM/ProgressiveAligner.cpp:613),
// SearchLCBGaps (LM/Aligner.cpp:806), MatchFinder::AddSequence -- see the same object.
// Tie order inside equal-mer runs is position-ascending; the reference's std::sort leaves it unspecified and no
// consumer depends on it.  There is no CPU fallback: a failing device call throws, as
// SortedMerList::Create does (SMLCreateError, LM/SortedMerList.h:287-292).
#ifndef CUDA_DNA_MEMORY_SML_H_
#define CUDA_DNA_MEMORY_SML_H_

#include <iostream>
#include <string>

#include "libMems/DNAMemorySML.h"
#include "mauve_cuda.h"

namespace mems {

class CudaDNAMemorySML : public DNAMemorySML
{
public:
	CudaDNAMemorySML(const uint8* table = SortedMerList::BasicDNATable(), const uint32 alpha_bits = DNA_ALPHA_BITS)
		: DNAMemorySML(table, alpha_bits) {}
	CudaDNAMemorySML* Clone() const
	{   // as DNAMemorySML::Clone (LM/DNAMemorySML.cpp:30-34): the reference declares a copy constructor it never defines
		CudaDNAMemorySML* c = new CudaDNAMemorySML();
		c->DNAMemorySML::operator=(*this);
		return c;
	}
private:
	CudaDNAMemorySML(const CudaDNAMemorySML&);
public:

	virtual void Create(const genome::gnSequence& seq, const uint64 seed)
	{
		SortedMerList::Create(seq, seed);  // header, masks and the 2-bit `sequence`: unchanged host code (LM/SortedMerList.cpp:786-824)
		// explicit length: gnRAWSequence::ToString() with default arguments drops the last two bases (LM/gnRAWSequence.h:157-161)
		const std::string bases = seq.ToString(seq.length(), 1);
		positions.assign(SMLLength(), 0);
		uint64_t n = 0;
		const int rc = mcu_sml_build(bases.data(), bases.size(), seed, positions.empty() ? NULL : &positions[0], NULL, NULL, &n);
		if (rc == MCU_EGAP) throw "ERROR: gap character encountered in input sequence";  // LM/SortedMerList.cpp:436
		if (rc != MCU_OK || n != positions.size()) {
			std::cerr << "CudaDNAMemorySML::Create: " << mcu_last_error() << std::endl;
			Throw_gnEx(SMLCreateError());
		}
	}
};

}  // namespace mems

#endif
