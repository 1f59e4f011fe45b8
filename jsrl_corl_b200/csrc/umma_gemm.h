// tcgen05 (UMMA) TF32 grouped GEMM used by IQL_MATH_TF32_TCGEN05.
#pragma once
#include "engine.h"

namespace iql {
// true when the phase (mode 0 NT / 1 NN / 2 TN) can run on the tcgen05 kernel
// for this batch size and hidden width (256x256x32 tiles must divide the problem).
bool umma_phase_supported(int mode, int batch, int hidden);
// Encode the two TMA tensor maps (A, B) of every problem into h_maps_out
// (2 * nprob CUtensorMap, 128 B each).  Returns 0 on success.
int umma_encode_maps(int mode, const GemmProb* h_probs, int nprob, void* h_maps_out);
void launch_umma_gemm(int mode, const GemmProb* probs, const void* maps, int epi, int nprob, int maxM, int maxN,
                      const StepCtx& ctx, cudaStream_t st);
// bias gradients of a wgrad phase: dbias[m] = sum_k A[k][m]
void launch_colsum(const GemmProb* probs, int nprob, int maxM, cudaStream_t st);
}  // namespace iql
