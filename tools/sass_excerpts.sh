#!/bin/bash
# profiles/r02_sass_excerpts.txt: the Blackwell-specific instructions the kernels rest on, from the in-tree objects (no GPU needed)
cd "$(dirname "$0")/.."
python -c "import mauve_py_b200._build as b; b.build_library()" || exit 1
fn() { cuobjdump -sass build/obj/$1.o | awk -v pat="$2" '/Function : /{f = ($0 ~ pat)} f'; }
{
echo "# SASS excerpts (round 2; cuobjdump -sass of the in-tree objects, sm_100a; regenerate with tools/sass_excerpts.sh)"
echo
echo "## bk_group3_kernel<12, 1, 5> (bucket.cu): bulk asynchronous copies (cp.async.bulk -> UBLKCP.S.G) completing on an mbarrier"
echo "## (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT), the hash join's shared-memory compare-and-swap"
fn bucket "bk_group3_kernelILi12ELi1ELi5E" | grep -E "UBLKCP|SYNCS|ATOMS.CAS" | awk '{$1=""; print}' | cut -c1-100
echo
echo "## hmm_exact_chain_warp_kernel (hmm.cu): three columns of the FP32 regime chain (D = 0): 4 FMUL + 4 FFMA + 2 FADD per column,"
echo "## coefficient loads two columns ahead, no DMUL / F2F.F64 in the chain"
fn hmm "hmm_exact_chain_warp_kernel" | grep -E "^\s+/\*[0-9a-f]+\*/" | awk '{$1=""; print}' | cut -c1-70 | awk '/LDS.128/{c++} c>=7' | head -60
echo "   DMUL / F2F instructions in the whole kernel (the hmm_exact_step fallback and the start / stop products): $(fn hmm "hmm_exact_chain_warp_kernel" | grep -cE "DMUL|F2F")"
echo
echo "## nw_forward_kernel (dp.cu): fused add+max of the cell update"
fn dp "nw_forward_kernel" | grep -E "VIADDMNMX" | head -6 | awk '{$1=""; print}' | cut -c1-100
echo "   count of VIADDMNMX in nw_forward_kernel: $(fn dp "nw_forward_kernel" | grep -c VIADDMNMX)"
} > profiles/r02_sass_excerpts.txt
wc -l profiles/r02_sass_excerpts.txt
