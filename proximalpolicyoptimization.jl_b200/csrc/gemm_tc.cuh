// gemm_tc.cuh — interface of the tcgen05 tensor-core engine for the policy MLP (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace ppo {

// allocate the tensor-core operand copies for `mode` (PPO_GEMM_TF32X3_TC); fails loudly (no fallback)
// when a hidden layer's shape is not supported by the tcgen05 kernels.
int tc_prepare(ppo_policy* p, int mode);
// rebuild the hi/lo operand copies from p->params (after policy_write / every Adam step)
int tc_refresh_weights(ppo_policy* p);
// hidden layer l forward: Y = leakyrelu(X W_l + b_l).  In this mode Y (= p->act[l+1]) receives the tf32-rounded
// hi part and the engine keeps the lo part; tc_act_lo(p, l) returns it (exact activation = hi + lo).
int tc_linear_fwd(ppo_policy* p, int l, const float* X, float* Y, int64_t M);
const float* tc_act_lo(ppo_policy* p, int l);
// hidden layer l backward: dW_l and (if dX != nullptr) dX = (dY W_l^T) .* leakyrelu'(X) as a hi/lo pair, plus
// db_below = colsum(dX) = the bias gradient of layer l-1 (fused into the dgrad epilogue).  db_l itself comes from
// the kernel that produced dY (head_bwd or the dgrad of layer l+1).
int tc_linear_bwd(ppo_policy* p, int l, const float* X, const float* dY, float* dX, float* dW, float* db_below, int64_t M);
float* tc_dact_lo(ppo_policy* p, const float* dact);
void tc_destroy(ppo_policy* p);

// stand-alone entry points on device pointers (ppo_dense_op / ppo_bench_kernel)
int tc_test_fwd(ppo_ctx* ctx, const float* X_hi, const float* X_lo, const float* WT_hi, const float* WT_lo, const float* bias,
                float* Y_hi, float* Y_lo, int64_t M, int K, int N, int act, float slope);
int tc_test_dgrad(ppo_ctx* ctx, const float* dY_hi, const float* dY_lo, const float* W_hi, const float* W_lo, const float* gate,
                  float* dX_hi, float* dX_lo, int64_t M, int K, int N, float slope, float* colsum_scratch, float* colsum_out);
int tc_test_wgrad(ppo_ctx* ctx, const float* X_hi, const float* X_lo, const float* dY_hi, const float* dY_lo, float* dW,
                  float* db, float* partial, size_t partial_bytes, int64_t M, int K, int N);
int tc_test_split(ppo_ctx* ctx, const float* x, float* hi, float* lo, int64_t n);
int tc_test_weight_prep(ppo_ctx* ctx, const float* W, float* W_hi, float* W_lo, float* WT_hi, float* WT_lo, int K, int N);
size_t tc_test_partial_bytes(ppo_ctx* ctx, int64_t M, int K, int N);

}  // namespace ppo
