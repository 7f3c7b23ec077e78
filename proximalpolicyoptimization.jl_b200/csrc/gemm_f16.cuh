// gemm_f16.cuh — interface of the fp16-split tcgen05 engine (PPO_GEMM_F16X3_TC, gemm_f16.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ppo {

// allocate the engine state for the policy; fails loudly (no fallback) when a layer shape is outside the contract
int f16_prepare(ppo_policy* p);
// weight statistics (abs-max, max column / row abs-sums) -> scales -> fp16 hi/lo copies of W and W^T
int f16_refresh_weights(ppo_policy* p);
// the same preceded by the Adam step, all in one launch; xv: peer-memory gradient exchange (dp_p2p.cu) or nullptr;
// d_step: device minibatch counter to advance, or nullptr
int f16_adam_refresh(ppo_policy* p, ppo_opt* opt, const P2PView* xv, int* d_step);
// whole-MLP forward: X fp32 [M][dims[0]] -> p->act[L] (fp32 logits); hidden activations stay fp16 hi/lo pairs.
// mask (optional): the minibatch's action mask [M * apa]; with p->compact_tokens the MLP then runs only on the tokens
// that have at least one unmasked action (the logits of the others never reach the loss: softmax(-Inf) = 0)
// feat_bound (optional): device word holding the bit pattern of an upper bound of |X| (the rollout buffer's abs-max)
int f16_forward(ppo_policy* p, const float* X, int64_t M, const float* mask, const unsigned* feat_bound = nullptr);
// where the loss kernel leaves max |dlogits| for the backward pass (launch_loss's dl_absmax)
unsigned* f16_dlogits_stat(ppo_policy* p);
// whole-MLP backward from p->dlogits -> p->grads
// dl_stat_ready: f16_dlogits_stat() already holds max |dlogits| of this minibatch (no abs-max pass, plan inside the head kernel)
int f16_backward(ppo_policy* p, int64_t M, bool dl_stat_ready = false);
// the leakyrelu' gates of hidden activation l (1..L-1) as the backward pass of the last minibatch applies them
// (tokens the compacted MLP skipped: PPO_GATE_SKIPPED)
int f16_read_gates(ppo_policy* p, int l, int64_t M, uint8_t* d_out);
int f16_active_tokens(ppo_policy* p, int64_t* out);
void f16_destroy(ppo_policy* p);

// ---- stand-alone entry points on device pointers (ppo_dense_op / ppo_bench_kernel) ----
// sc = device {scale, 1/scale}; st = one device word of scratch (zero on entry, zero on exit)
int f16_test_operand(ppo_ctx* ctx, const float* x, __half* hi, __half* lo, int64_t n, float* sc, unsigned* st);
int f16_test_weight(ppo_ctx* ctx, const float* W, __half* W_hi, __half* W_lo, __half* WT_hi, __half* WT_lo, int K, int N,
                    float* sc, unsigned* st);
int f16_test_set_scale(ppo_ctx* ctx, float* sc, float bound);
int f16_test_fwd(ppo_ctx* ctx, const __half* X_hi, const __half* X_lo, const __half* WT_hi, const __half* WT_lo,
                 const float* bias, __half* Y_hi, __half* Y_lo, uint32_t* Y_sign, int64_t M, int K, int N, int act, float slope,
                 const float* sc_x, const float* sc_w, const float* sc_y);
int f16_test_dgrad(ppo_ctx* ctx, const __half* dY_hi, const __half* dY_lo, const __half* W_hi, const __half* W_lo,
                   const uint32_t* gate, __half* dX_hi, __half* dX_lo, int64_t M, int K, int N, float slope,
                   float* colsum_scratch, float* colsum_out, const float* sc_dy, const float* sc_w, const float* sc_dx);
int f16_test_wgrad(ppo_ctx* ctx, const __half* X_hi, const __half* X_lo, const __half* dY_hi, const __half* dY_lo, float* dW,
                   float* partial, size_t partial_bytes, int64_t M, int K, int N, const float* sc_x, const float* sc_dy);
int f16_test_join(ppo_ctx* ctx, const __half* hi, const __half* lo, int64_t n, const float* sc, float* out);
int f16_test_signbits(ppo_ctx* ctx, const float* x, int64_t M, int N, uint32_t* out);   // N % 32 == 0; out: f16_test_sign_words(M, N)
size_t f16_test_sign_words(int64_t M, int N);
size_t f16_test_partial_bytes(ppo_ctx* ctx, int64_t M, int K, int N);
int f16_test_head_fwd(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* W, const float* bias, float* logits,
                      int64_t M, int K, int N, const float* sc_h);
int f16_test_head_bwd(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* dlogits, const float* W,
                      __half* dH_hi, __half* dH_lo, float* dW, float* db, float* db_below, int64_t M, int K, int N,
                      float slope, float* partial, size_t partial_bytes, const float* sc_h, const float* sc_dh);
size_t f16_test_head_partial_bytes(int64_t M, int K, int N);

}  // namespace ppo
