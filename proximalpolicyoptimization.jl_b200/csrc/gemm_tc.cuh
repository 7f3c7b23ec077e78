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

// stand-alone entry points on device pointers (ppo_dense_op / ppo_bench_kernel)
int tc_test_fwd(ppo_ctx* ctx, const float* X, const float* X_lo, const float* WT, const float* WT_lo, const float* bias,
                float* Y, float* Y_lo, int64_t M, int K, int N, int act, float slope);
int tc_test_dgrad(ppo_ctx* ctx, const float* dY, const float* dY_lo, const float* W, const float* W_lo, const float* gate,
                  float* dX, float* dX_lo, int64_t M, int K, int N, float slope);
int tc_test_wgrad(ppo_ctx* ctx, const float* X, const float* X_lo, const float* dY, const float* dY_lo, float* dW,
                  float* partial, size_t partial_bytes, int64_t M, int K, int N);
void tc_set_passes(int n);
int tc_test_split_lo(ppo_ctx* ctx, const float* x, float* lo, int64_t n);
int tc_test_weight_prep(ppo_ctx* ctx, const float* W, float* W_lo, float* WT, float* WT_lo, int K, int N);
int tc_test_colsum(ppo_ctx* ctx, const float* dY, int64_t M, int N, float* partial, float* db);
size_t tc_test_wgrad_partial_bytes(ppo_ctx* ctx, int64_t M, int K, int N);

}  // namespace ppo
