// gemm_tc.cuh — interface of the tcgen05 tensor-core engines for the policy MLP (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace ppo {

// allocate the tensor-core operand copies for `mode` (PPO_GEMM_TF32X3_TC / PPO_GEMM_BF16_TC);
// fails (no fallback) when a layer shape is not supported by the tcgen05 kernels.
int tc_prepare(ppo_policy* p, int mode);
// rebuild the operand copies from p->params (after policy_write / every Adam step)
int tc_refresh_weights(ppo_policy* p);
// hidden layer l forward: Y = leakyrelu(X W_l + b_l)
int tc_linear_fwd(ppo_policy* p, int l, const float* X, float* Y, int64_t M);
// hidden layer l backward: dW_l, db_l and (if dX != nullptr) dX = (dY W_l^T) .* leakyrelu'(X)
int tc_linear_bwd(ppo_policy* p, int l, const float* X, const float* dY, float* dX, float* dW, float* db, int64_t M);
void tc_destroy(ppo_policy* p);

}  // namespace ppo
