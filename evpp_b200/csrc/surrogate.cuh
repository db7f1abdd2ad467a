// Planner surrogate g_phi: fit (Adam, MSE) and batched queries (see surrogate.cu).
#pragma once
#include "common.cuh"

namespace evpp {

constexpr int kSurrogateParams = 192 + 64 + 8 * (4096 + 64) + 64 + 1;   // packed state_dict of the reference's MLP: 33601
constexpr int kSurrogateMaxN = 128;                                     // training points per fit (the planners use 100)

size_t surrogate_workspace_floats(int n);   // activations kept for the backward pass
int launch_surrogate_fit(float* params, float* m, float* v, const float* x, const float* target, int n, int epochs, float lr,
                         float* workspace, float* loss_out, cudaStream_t st);
int launch_surrogate_query(const float* params, const float* x, int M, float* out, cudaStream_t st);

}  // namespace evpp
