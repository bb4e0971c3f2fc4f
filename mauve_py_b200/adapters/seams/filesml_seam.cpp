// Link-time seam: mems::FileSML::Create with the sorted mer list built on the GPU.
//
// progressiveMauve creates one `<fasta>.sslist` per genome through DNAFileSML (MatchList::LoadSMLs, LM/MatchList.h:296-330 ->
// FileSML::Create, LM/FileSML.cpp:401-459): header + 2-bit sequence from SortedMerList::Create, then FillDnaSeedSML / FillSML and
// std::sort(bmer_lessthan) over 16-byte records, then the positions go to disk.  The definition of FileSML::Create is weakened in
// a COPY of FileSML.o and given a second name (objcopy, oracle/Makefile.ref); this file supplies it under the original name: the
// same sequence of file operations, with the fill + sort replaced by mcu_sml_build (the CudaDNAMemorySML path).  The file that
// results is the reference's format (what mauve_py_b200.libmems.write_sslist writes from Python), so everything that reads it --
// FileSML::Read, operator[], LoadFile of a later run -- is unchanged.  Ties inside equal-mer runs are position-ascending
// (std::sort leaves them unspecified and no consumer depends on them).
// Only DNA lists (DNAFileSML) take the device path; MAUVE_CUDA_SML_SEAM=0 switches the seam off.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "libGenome/gnSequence.h"
#include "libMems/FileSML.h"
#include "libMems/DNAFileSML.h"
#include "mauve_cuda.h"

namespace mems {

void FileSML_Create_reference(FileSML* self, const genome::gnSequence& seq, const uint64 seed) asm("_ZN4mems7FileSML16Create_referenceERKN6genome10gnSequenceEy");

static unsigned long long g_fsml_device = 0, g_fsml_reference = 0;
struct FileSmlSeamReport {
	~FileSmlSeamReport()
	{
		if (getenv("MAUVE_CUDA_SEAM_REPORT"))
			fprintf(stderr, "FileSML::Create seam: %llu lists on the device, %llu in the reference's code\n", g_fsml_device, g_fsml_reference);
	}
};
static FileSmlSeamReport g_fsml_report;

void FileSML::Create(const genome::gnSequence& seq, const uint64 seed)
{
	static const bool off = getenv("MAUVE_CUDA_SML_SEAM") && getenv("MAUVE_CUDA_SML_SEAM")[0] == '0';
	if (off || dynamic_cast<DNAFileSML*>(this) == NULL) {
		++g_fsml_reference;
		FileSML_Create_reference(this, seq, seed);
		return;
	}
	++g_fsml_device;
	OpenForWriting(true);                 // LM/FileSML.cpp:405
	SortedMerList::Create(seq, seed);     // :408: header, masks, the 2-bit `sequence`

	// :410-431 on the device: positions in sorted order
	const std::string bases = seq.ToString(seq.length(), 1);   // explicit length: gnRAWSequence::ToString() drops two bases
	std::vector<smlSeqI_t> positions(SMLLength());
	uint64_t n = 0;
	const int rc = mcu_sml_build(bases.data(), bases.size(), seed, positions.empty() ? NULL : &positions[0], NULL, NULL, &n);
	if (rc == MCU_EGAP) throw "ERROR: gap character encountered in input sequence";   // LM/SortedMerList.cpp:436
	if (rc != MCU_OK || n != positions.size()) {
		std::cerr << "FileSML::Create (device): " << mcu_last_error() << std::endl;
		Throw_gnEx(SMLCreateError());
	}

	// :433-458, unchanged
	sarfile.write((char*)&header, sizeof(struct SMLHeader));
	if (!sarfile.good()) {
		sarfile.clear();
		Throw_gnExMsg(genome::IOStreamFailed(), "Error writing sorted mer list header to disk.\n");
	}
	sarfile.write((char*)sequence, binary_seq_len * sizeof(uint32));
	sarray_start_offset = sarfile.tellg();
	if (!positions.empty()) sarfile.write((char*)&positions[0], positions.size() * sizeof(smlSeqI_t));
	sarfile.flush();
	if (!sarfile.good()) {
		sarfile.clear();
		Throw_gnExMsg(genome::IOStreamFailed(), "Error writing sorted mer list to disk.\n");
	}
	sarfile.close();
	sarfile.open(filename.c_str(), std::ios::binary | std::ios::in);
	if (!sarfile.is_open()) Throw_gnExMsg(genome::FileNotOpened(), "FileSML::Create: Error opening sorted mer list file.\n");
	sardata.open(filename);
}

}  // namespace mems
