// run_cuda -- drop-in for the pairwise homology HMM's `run()` (LM/HomologyHMM/homologymain.cc:24-62: Forward,
// Backward, posterior of the homologous state >= 0.9 -> 'H' else 'N'), called from findHssHomologyHMM
// (LM/Islands.h:161) with the column string over '1'..'8' built by the encoder at LM/Islands.h:113-155.
// Signature and semantics are those of `void run(std::string&, std::string&, const Params&)`
// (LM/HomologyHMM/homology.h:47); several LCBs can be submitted at once with run_cuda_batch.
#ifndef CUDA_HOMOLOGY_HMM_H_
#define CUDA_HOMOLOGY_HMM_H_

#include <stdexcept>
#include <string>
#include <vector>

#include "homology.h"
#include "mauve_cuda.h"

namespace cuda_detail {
inline void ParamsToArray(const Params& p, double* v)  // order of LM/HomologyHMM/homology.h:169-177
{
	v[0] = p.iStartHomologous; v[1] = p.iGoHomologous; v[2] = p.iGoUnrelated; v[3] = p.iGoStopFromUnrelated; v[4] = p.iGoStopFromHomologous;
	for (int i = 0; i < 8; ++i) { v[5 + i] = p.aEmitHomologous[i]; v[13 + i] = p.aEmitUnrelated[i]; }
}
}  // namespace cuda_detail

inline void run_cuda_batch(const std::vector<std::string>& sequences, std::vector<std::string>& predictions, const Params& p,
                           std::vector<std::vector<double> >* posteriors = NULL)
{
	double v[21];
	cuda_detail::ParamsToArray(p, v);
	std::string all;
	std::vector<uint64_t> off(1, 0);
	for (size_t i = 0; i < sequences.size(); ++i) { all += sequences[i]; off.push_back(all.size()); }
	std::string pred(all.size(), 'N');
	std::vector<double> post(posteriors ? all.size() : 0);
	const int rc = mcu_hmm_batch(sequences.size(), all.data(), &off[0], v, all.empty() ? NULL : &pred[0], posteriors && !post.empty() ? &post[0] : NULL, NULL);
	if (rc != MCU_OK) throw std::runtime_error(std::string("run_cuda: ") + mcu_last_error());
	predictions.resize(sequences.size());
	if (posteriors) posteriors->resize(sequences.size());
	for (size_t i = 0; i < sequences.size(); ++i) {
		predictions[i] = pred.substr(off[i], off[i + 1] - off[i]);
		if (posteriors) (*posteriors)[i].assign(post.begin() + off[i], post.begin() + off[i + 1]);
	}
}

inline void run_cuda(std::string& sequence, std::string& prediction, const Params& p)
{
	std::vector<std::string> in(1, sequence), out;
	run_cuda_batch(in, out, p);
	prediction = out[0];
}

#endif
