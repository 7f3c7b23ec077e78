// common.cuh — internal types shared by the kernels and the C ABI (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ppo_b200.h"

namespace ppo {

void set_error(const char* fmt, ...);

#define PPO_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ppo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PPO_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define PPO_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            ppo::set_error(__VA_ARGS__);   \
            return PPO_ERR_INVALID;        \
        }                                  \
    } while (0)

#define PPO_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != PPO_OK) return _s; \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

}  // namespace ppo

// ---------------------------------------------------------------------------------------------
// handles (opaque in the public header)
// ---------------------------------------------------------------------------------------------

struct ppo_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // pinned staging for host<->device copies of small results
    double* h_pinned = nullptr;      // PINNED_DOUBLES doubles
    // scratch for the scan's decoupled look-back and misc reductions
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // L2 flush buffer for benches
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    // NCCL (loaded with dlopen on first use)
    int* d_step = nullptr;           // device-side minibatch counter (CUDA-graph replay of the epoch loop)
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // data parallelism: the gradient all-reduce runs layer by layer on a second stream while the backward pass of the
    // layers below is still computing (fork: ev_fork recorded on `stream`; join: `stream` waits for ev_join)
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool comm_pending = false;
    bool p2p_grads = false;          // a policy of this context exchanges gradients over peer memory (dp_p2p.cu)
    // Destruction order: garbage-collected bindings (Julia finalizers at exit) may destroy the context before the
    // buffers / policies / optimisers that live on it.  ppo_ctx_destroy only marks the context dead while children are
    // alive; the last child's destroy call then releases it.
    int children = 0;
    bool dead = false;
    int64_t pinned_doubles = 0;      // capacity of h_pinned (grown on demand)
};

// minibatch staging area (output of the K4 gather; input of the MLP / loss)
struct ppo_batch {
    int64_t cap = 0;   // rows allocated
    float* feat = nullptr;     // [cap][nhe][nf]
    float* mask = nullptr;     // [cap][A]
    int* action = nullptr;     // [cap] 0-based
    float* old_prob = nullptr; // [cap]
    float* adv = nullptr;      // [cap]
    const unsigned* feat_bound = nullptr;   // device word: bit pattern of an upper bound of |feat| (the buffer's abs-max), or nullptr
};

struct ppo_buf {
    ppo_ctx* ctx = nullptr;
    int64_t cap = 0, n = 0;
    int nf = 0, nhe = 0, apa = 0, A = 0;
    float* feat = nullptr;       // [cap][nhe][nf]
    float* mask = nullptr;       // [cap][A]
    int* action = nullptr;       // [cap], 0-based on the device
    float* old_prob = nullptr;   // [cap]
    float* reward = nullptr;     // [cap]; rewards, then returns (compute_state_value! swaps it with reward_alt)
    float* reward_alt = nullptr; // [cap]; output array of the next returns scan
    float* reward_saved = nullptr; // optional snapshot of the raw rewards
    int64_t saved_n = 0;
    uint8_t* terminal = nullptr; // [cap]
    int* perm = nullptr;         // [cap], 0-based on the device
    int64_t perm_len = 0;        // 0 = no permutation set
    // advantage normalisation (extension): device scalars {mean, inv_std}
    int normalize = 0;
    double norm_eps = 1e-8;
    float* d_norm = nullptr;     // 2 floats: mean, 1/(std+eps)
    unsigned* d_feat_absmax = nullptr;   // bit pattern of max |feat| over everything appended since the last clear
    double* d_tile_stats = nullptr; // per scan tile {sum, sumsq}
    int64_t n_tiles_stats = 0;
    bool stats_valid = false;
    bool returns_valid = false;     // reward[] holds returns (compute_returns ran since the last append)
    ppo_batch batch;             // gather destination
};

struct ppo_policy {
    ppo_ctx* ctx = nullptr;
    int L = 0;                       // number of Dense layers
    std::vector<int> dims;           // L+1
    std::vector<int64_t> w_off, b_off;  // offsets into the flat parameter vector
    int64_t P = 0;
    float slope = 0.01f;
    int gemm_mode = PPO_GEMM_FP32_SIMT;
    int compact_tokens = 1;          // fp16-split engine: run the MLP only on tokens with an unmasked action (ppo_policy_set_token_compaction)
    float* params = nullptr;  // flat, Flux.params order: W1[in][out], b1, W2, b2, ...
    float* grads = nullptr;   // same layout
    // workspace for M = rows*nhe tokens
    int64_t ws_tokens = 0;
    std::vector<float*> act;  // act[l], l = 1..L-1: hidden activations [M][dims[l]]; act[L] = logits [M][apa]
    float* dact[2] = {nullptr, nullptr};  // ping-pong activation gradients [M][Hmax]
    float* dlogits = nullptr;             // [M][apa]
    float* partial = nullptr;             // split-K partials for wgrad
    size_t partial_bytes = 0;
    double* d_loss_partials = nullptr;    // loss kernel block partials
    int64_t loss_partials_cap = 0;
    double* d_loss_hist = nullptr;        // [2 * hist_cap] (ppo sum-mean, entropy) per minibatch
    int64_t hist_cap = 0;
    // tensor-core operand copies (tcgen05 modes), maintained by the Adam kernel / policy_write
    void* tc = nullptr;
    void* f16 = nullptr;              // fp16-split engine state (gemm_f16.cu)
    void* dp = nullptr;               // peer-memory gradient exchange (dp_p2p.cu)
    // own minibatch staging for the host-array entry points
    ppo_batch hbatch;
};

struct ppo_opt {
    ppo_ctx* ctx = nullptr;          // kept separately: the optimiser may outlive its policy handle
    ppo_policy* policy = nullptr;
    double eta = 1e-3, beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    float* m = nullptr;
    float* v = nullptr;
    double* d_bp = nullptr;  // device {beta1^t, beta2^t}, advanced by the update kernel
};

// ---------------------------------------------------------------------------------------------
// kernel launchers (one per .cu file); all enqueue on ctx->stream and bump ctx->launches
// ---------------------------------------------------------------------------------------------
namespace ppo {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_STATS_PER_TILE = SCAN_THREADS / 32;   // one {sum, sumsq} pair per warp

// scan.cu (K1, K2)
size_t scan_scratch_bytes(int64_t n);
int launch_returns_scan(ppo_ctx* ctx, const float* reward_in, float* returns_out, const uint8_t* terminal,
                        int64_t n, double discount, int discount_is_f32, double* tile_stats, void* scratch);
// tile_stats == nullptr in launch_returns_scan skips the K2 statistics (advantage normalisation is an extension, off by
// default); launch_returns_stats produces the identical partials from the finished returns
int launch_returns_stats(ppo_ctx* ctx, const float* returns, int64_t n, double* tile_stats);
int launch_norm_finalize(ppo_ctx* ctx, const double* tile_stats, int64_t n_tiles, int64_t n,
                         double eps, float* d_norm);

// shuffle.cu (K3)
int launch_feistel_permutation(ppo_ctx* ctx, int* perm0, int64_t n, uint64_t seed);
int launch_perm_from_host(ppo_ctx* ctx, const int64_t* d_perm1, int* perm0, int64_t n, int64_t limit,
                          int* d_bad);
int launch_perm_to_i64(ppo_ctx* ctx, const int* perm0, int64_t* d_perm1, int64_t n);

// gather.cu (K4)
struct GatherArgs {
    const float* feat; const float* mask; const int* action; const float* old_prob; const float* ret;
    const int* index;          // 0-based row indices (perm + start), length count
    int64_t count; int feat_elems; int mask_elems;
    float* feat_out; float* mask_out; int* action_out; float* prob_out; float* adv_out;
    const float* norm;         // {mean, inv_std} or nullptr
    const int* step;           // optional device scalar: index += (*step) * step_stride (CUDA-graph replay)
    int64_t step_stride;
};
int launch_gather(ppo_ctx* ctx, const GatherArgs& a, int variant);
int launch_permute_inplace_u8(ppo_ctx* ctx, const uint8_t* src, uint8_t* dst, const int* idx, int64_t n);
int launch_convert_actions_in(ppo_ctx* ctx, const int64_t* a1, int* a0, int64_t n, int A, int* d_bad);
int launch_convert_actions_out(ppo_ctx* ctx, const int* a0, int64_t* a1, int64_t n);
int launch_linear_index_in(ppo_ctx* ctx, const int64_t* lin1, int* a0, int64_t n, int A, int* d_bad);
int launch_i64_to_f32(ppo_ctx* ctx, const int64_t* src, float* dst, int64_t n);
int launch_narrow_to_f32(ppo_ctx* ctx, const void* src, int elem_bytes /* 1: int8, 2: int16 */, float* dst, int64_t n,
                         unsigned* absmax_out = nullptr /* optional: atomicMax of the bit pattern of max |value| */);
int launch_normalize_bool(ppo_ctx* ctx, uint8_t* t, int64_t n);
int launch_step_advance(ppo_ctx* ctx, int* d_step);
int launch_mask_from_bits(ppo_ctx* ctx, const uint64_t* bits, float* mask, int64_t n);
int launch_absmax_f32(ppo_ctx* ctx, const float* x, int64_t n, unsigned* out);

// loss.cu (K6)
int64_t loss_num_blocks(int64_t nb, int A);
int launch_loss(ppo_ctx* ctx, const float* logits, const float* mask, const int* action,
                const float* old_prob, const float* adv, int64_t nb, int A, double epsilon,
                double entropy_weight, double inv_nb_global, float* dlogits, double* partials,
                double* loss_out2 /* {ppoloss, entropyloss unweighted} */, float* probs_out,
                const int* step = nullptr /* optional device scalar: write to loss_out2 + 2 * (*step) */,
                unsigned* dl_absmax = nullptr /* optional: atomicMax of the bit pattern of max |dlogits| */);

int launch_sample_actions(ppo_ctx* ctx, const float* probs, int64_t nb, int A, uint64_t seed, int64_t* action1,
                          float* prob_out);

// gemm_simt.cu (K5/K7, fp32 reference path) + skinny last layer
int launch_linear_fwd_simt(ppo_ctx* ctx, const float* X, const float* W, const float* bias, float* Y,
                           int64_t M, int K, int N, bool act, float slope);
int launch_linear_dgrad_simt(ppo_ctx* ctx, const float* dY, const float* W, const float* Hprev,
                             float* dX, int64_t M, int K, int N, float slope);
int launch_linear_wgrad_simt(ppo_ctx* ctx, const float* X, const float* dY, float* dW, float* db,
                             int64_t M, int K, int N, float* partial, size_t partial_bytes);
size_t wgrad_partial_bytes(int64_t M, int K, int N);
// Hlo: optional lo companion of H (tensor-core mode keeps activations as tf32 hi + lo pairs)
int launch_head_fwd(ppo_ctx* ctx, const float* H, const float* Hlo, const float* W, const float* bias,
                    float* logits, int64_t M, int K, int N);
// dHlo: write dH as a tf32 hi/lo pair; db_below: also emit colsum(dH) = bias gradient of the layer below
int launch_head_bwd(ppo_ctx* ctx, const float* H, const float* Hlo, const float* dlogits, const float* W,
                    float* dH, float* dHlo, float* dW, float* db, float* db_below, int64_t M, int K, int N,
                    float slope, float* partial, size_t partial_bytes, bool need_dH);

// adam.cu (K8)
int launch_adam(ppo_ctx* ctx, float* x, float* m, float* v, const float* g, int64_t n, double eta,
                double b1, double b2, double eps, double* d_bp, float grad_scale);

// nccl_dl.cpp
int nccl_unique_id(void* id128);
int nccl_init(ppo_ctx* ctx, int nranks, int rank, const void* id128);
int nccl_destroy(ppo_ctx* ctx);
int nccl_allreduce_f32(ppo_ctx* ctx, float* d_buf, int64_t n);
int nccl_allreduce_f32_on(ppo_ctx* ctx, float* d_buf, int64_t n, cudaStream_t stream);
// overlapped gradient all-reduce (no-ops without a communicator): grads_ready enqueues the all-reduce of a finished
// slice of the gradient vector on the communication stream, ordered after everything enqueued on ctx->stream so far;
// grads_join makes ctx->stream wait for all of them (call before the optimiser).  Returns true from dp_overlap() when the
// engine should use them instead of one all-reduce after the whole backward pass.
bool dp_overlap(ppo_ctx* ctx);
int grads_ready(ppo_ctx* ctx, float* d_slice, int64_t n);
int grads_join(ppo_ctx* ctx);
int nccl_allreduce_f64(ppo_ctx* ctx, double* d_buf, int64_t n);

// dp_p2p.cu: gradient all-reduce over NVLink peer memory fused into the Adam step (one process per GPU, CUDA IPC)
// what a kernel needs to read the peers' published gradients (dp_p2p.cu)
struct P2PView {
    const float* const* peer_xchg;   // [nranks] exchange buffers (own entry = local)
    int64_t Ppad;                    // floats per half of the double buffer
    const unsigned* flags;           // [nranks] this rank's flag words (epoch + 1 once a peer has published)
    unsigned* state;                 // [0] epoch, [1] publish block counter, [2] error flag
    int nranks;
};
bool p2p_active(const ppo_policy* p);
// publish p->grads to the peers (the first half of p2p_reduce_and_step) and describe where to read them
int p2p_publish(ppo_policy* p, P2PView* view);
int p2p_export(ppo_policy* p, void* handle64);
int p2p_connect(ppo_policy* p, int nranks, int rank, const void* handles);
int p2p_reduce_and_step(ppo_policy* p, ppo_opt* opt);
int p2p_check(ppo_policy* p);
int p2p_wait_stats(ppo_policy* p, int64_t* total_ns, int64_t* waits, int reset);
void p2p_destroy(ppo_policy* p);

// l2 flush helper
int flush_l2(ppo_ctx* ctx);

}  // namespace ppo
