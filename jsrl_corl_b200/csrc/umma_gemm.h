// tcgen05 (UMMA) TF32 grouped GEMM used by IQL_MATH_TF32_TCGEN05.
#pragma once
#include "engine.h"

namespace iql {
// true when the phase (mode 0 NT / 1 NN / 2 TN) can run on the tcgen05 kernel
// for this batch size and hidden width (tiles must divide the problem).
bool umma_phase_supported(int mode, int batch, int hidden);
void launch_umma_gemm(int mode, const GemmProb* probs, int nprob, int maxM, int maxN, int K, const StepCtx& ctx,
                      cudaStream_t st);
}  // namespace iql
