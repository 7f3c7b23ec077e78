// abi.cu — the extern "C" boundary of libppo_b200.so (include/ppo_b200.h) and the host-side
// orchestration of the PPO update: buffer management, returns scan, permutation, gather,
// MLP forward/backward, fused loss, Adam, optional NCCL gradient all-reduce.
//
// Mirrors, call for call, the reference's src/rollout_buffer.jl, src/collect_rollouts.jl:26-42
// and src/train.jl:35-158 (each entry point cites its counterpart in the public header).
// There is deliberately no CPU fallback anywhere in this file: every numeric result comes from a
// kernel launched on ctx->stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <new>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "gemm_f16.cuh"

namespace ppo {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

int use(ppo_ctx* ctx) {
    PPO_REQUIRE(ctx != nullptr, "null context");
    PPO_CUDA(cudaSetDevice(ctx->device));
    return PPO_OK;
}

int ensure_scratch(ppo_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return PPO_OK;
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_scratch) PPO_CUDA(cudaFree(ctx->d_scratch));
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = (size_t)round_up((int64_t)bytes, 1 << 20);
    PPO_CUDA(cudaMalloc(&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
    return PPO_OK;
}

template <typename T>
int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    PPO_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return PPO_OK;
}

template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

void free_batch(ppo_batch& b) {
    dev_free(b.feat); dev_free(b.mask); dev_free(b.action); dev_free(b.old_prob); dev_free(b.adv);
    b.cap = 0;
}

int ensure_batch(ppo_ctx* ctx, ppo_batch& b, int64_t rows, int feat_elems, int A) {
    if (rows <= b.cap) return PPO_OK;
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    free_batch(b);
    PPO_TRY(dev_alloc(&b.feat, (size_t)rows * feat_elems));
    PPO_TRY(dev_alloc(&b.mask, (size_t)rows * A));
    PPO_TRY(dev_alloc(&b.action, (size_t)rows));
    PPO_TRY(dev_alloc(&b.old_prob, (size_t)rows));
    PPO_TRY(dev_alloc(&b.adv, (size_t)rows));
    b.cap = rows;
    return PPO_OK;
}

int h2d(ppo_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return PPO_OK;
    PPO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PPO_OK;
}
int d2h(ppo_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return PPO_OK;
    PPO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PPO_OK;
}

int check_bad_flag(ppo_ctx* ctx, int* d_bad, const char* what) {
    int bad = 0;
    PPO_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    PPO_REQUIRE(bad == 0, "%s", what);
    return PPO_OK;
}

// ---- policy workspace -----------------------------------------------------------------------
void free_workspace(ppo_policy* p) {
    for (auto& a : p->act) dev_free(a);
    dev_free(p->dact[0]); dev_free(p->dact[1]); dev_free(p->dlogits); dev_free(p->partial);
    p->ws_tokens = 0;
    p->partial_bytes = 0;
}

int ensure_workspace(ppo_policy* p, int64_t tokens) {
    if (tokens <= p->ws_tokens) return PPO_OK;
    ppo_ctx* ctx = p->ctx;
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    free_workspace(p);
    const int L = p->L;
    int hmax = 1;
    for (int l = 1; l < L; ++l) hmax = std::max(hmax, p->dims[l]);
    p->act.assign(L + 1, nullptr);
    for (int l = 1; l <= L; ++l) PPO_TRY(dev_alloc(&p->act[l], (size_t)tokens * p->dims[l]));
    // the logits start out as zeros: a compacted forward pass leaves the logits of skipped (fully masked) tokens untouched,
    // and the loss adds their -Inf mask to whatever finite value is there
    PPO_CUDA(cudaMemsetAsync(p->act[L], 0, (size_t)tokens * p->dims[L] * sizeof(float), ctx->stream));
    PPO_TRY(dev_alloc(&p->dact[0], (size_t)tokens * hmax));
    PPO_TRY(dev_alloc(&p->dact[1], (size_t)tokens * hmax));
    PPO_TRY(dev_alloc(&p->dlogits, (size_t)tokens * p->dims[L]));
    size_t pb = 0;
    for (int l = 0; l < L; ++l) pb = std::max(pb, wgrad_partial_bytes(tokens, p->dims[l], p->dims[l + 1]));
    PPO_CUDA(cudaMalloc((void**)&p->partial, pb));
    p->partial_bytes = pb;
    p->ws_tokens = tokens;
    return PPO_OK;
}

int ensure_loss_buffers(ppo_policy* p, int64_t nb, int A, int64_t hist) {
    int64_t blocks = loss_num_blocks(nb, A);
    if (blocks > p->loss_partials_cap) {
        PPO_CUDA(cudaStreamSynchronize(p->ctx->stream));
        dev_free(p->d_loss_partials);
        PPO_TRY(dev_alloc(&p->d_loss_partials, (size_t)blocks * 2));
        p->loss_partials_cap = blocks;
    }
    if (hist > p->hist_cap) {
        PPO_CUDA(cudaStreamSynchronize(p->ctx->stream));
        dev_free(p->d_loss_hist);
        PPO_TRY(dev_alloc(&p->d_loss_hist, (size_t)hist * 2));
        p->hist_cap = hist;
    }
    return PPO_OK;
}

// rebuild the tensor-core engines' operand copies of the weights (after policy_write / every Adam step)
int refresh_engine_weights(ppo_policy* p) {
    if (p->gemm_mode == PPO_GEMM_F16X3_TC) return f16_refresh_weights(p);
    if (p->gemm_mode != PPO_GEMM_FP32_SIMT) return tc_refresh_weights(p);
    return PPO_OK;
}

// forward through all Dense layers: act[l] for l = 1..L  (act[L] = logits, linear).  mask: the action mask the logits
// will be added to (lets the fp16-split engine skip fully masked tokens), or nullptr
int policy_forward(ppo_policy* p, const float* X, int64_t M, const float* mask, const unsigned* feat_bound = nullptr) {
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    if (p->gemm_mode == PPO_GEMM_F16X3_TC) return f16_forward(p, X, M, mask, feat_bound);
    const float* in = X;
    for (int l = 0; l < L; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        const float* W = p->params + p->w_off[l];
        const float* b = p->params + p->b_off[l];
        const bool last = (l == L - 1);
        if (last) {
            const float* in_lo = (p->gemm_mode != PPO_GEMM_FP32_SIMT && l > 0) ? tc_act_lo(p, l) : nullptr;
            PPO_TRY(launch_head_fwd(ctx, in, in_lo, W, b, p->act[l + 1], M, K, N));
        } else if (p->gemm_mode == PPO_GEMM_FP32_SIMT) {
            PPO_TRY(launch_linear_fwd_simt(ctx, in, W, b, p->act[l + 1], M, K, N, true, p->slope));
        } else {
            PPO_TRY(tc_linear_fwd(p, l, in, p->act[l + 1], M));
        }
        in = p->act[l + 1];
    }
    return PPO_OK;
}

// backward from p->dlogits: fills p->grads (Flux.params order)
int policy_backward(ppo_policy* p, const float* X, int64_t M, bool dl_stat_ready = false) {
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    if (p->gemm_mode == PPO_GEMM_F16X3_TC) return f16_backward(p, M, dl_stat_ready);
    const bool tc = p->gemm_mode != PPO_GEMM_FP32_SIMT;
    const float* delta = p->dlogits;
    int pp = 0;
    for (int l = L - 1; l >= 0; --l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        const float* W = p->params + p->w_off[l];
        float* dW = p->grads + p->w_off[l];
        float* db = p->grads + p->b_off[l];
        float* db_below = (l > 0) ? p->grads + p->b_off[l - 1] : nullptr;
        const float* in = (l == 0) ? X : p->act[l];
        float* dX = (l == 0) ? nullptr : p->dact[pp];
        if (l == L - 1) {
            // the head: in tensor-core mode it reads the hi/lo activation pair, writes dX as a hi/lo pair and
            // emits the bias gradient of the layer below (colsum of dX) in the same pass
            const float* in_lo = (tc && l > 0) ? tc_act_lo(p, l) : nullptr;
            float* dX_lo = (tc && dX != nullptr) ? tc_dact_lo(p, dX) : nullptr;
            PPO_TRY(launch_head_bwd(ctx, in, in_lo, delta, W, dX, dX_lo, dW, db, tc ? db_below : nullptr, M, K, N,
                                    p->slope, p->partial, p->partial_bytes, l > 0));
        } else if (!tc) {
            PPO_TRY(launch_linear_wgrad_simt(ctx, in, delta, dW, db, M, K, N, p->partial, p->partial_bytes));
            if (l > 0) PPO_TRY(launch_linear_dgrad_simt(ctx, delta, W, in, dX, M, K, N, p->slope));
        } else {
            PPO_TRY(tc_linear_bwd(p, l, in, delta, dX, dW, db_below, M));
        }
        delta = dX;
        pp ^= 1;
    }
    return PPO_OK;
}

// One minibatch of the update on device-resident batch arrays.  Writes {ppoloss, entropyloss}
// (unweighted) to p->d_loss_hist[2*slot..].
// advance_step: the device minibatch counter, to be advanced once the minibatch is done (CUDA-graph replay); the
// fp16-split engine's fused optimiser kernel does it, otherwise a one-thread kernel is appended here
int step_core(ppo_policy* p, ppo_opt* opt, const ppo_batch& bt, int64_t nb, int nhe, double epsilon,
              double entropy_weight, double inv_nb_global, int64_t slot, const int* d_step = nullptr,
              int* advance_step = nullptr) {
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    const int apa = p->dims[L];
    const int A = nhe * apa;
    const int64_t M = nb * nhe;
    PPO_TRY(ensure_workspace(p, M));
    PPO_TRY(ensure_loss_buffers(p, nb, A, slot + 1));
    PPO_TRY(policy_forward(p, bt.feat, M, bt.mask, bt.feat_bound));
    // (fp16-split engine: the loss kernel also leaves max |dlogits| where the backward pass plans its scales from)
    unsigned* dl_stat = p->gemm_mode == PPO_GEMM_F16X3_TC ? f16_dlogits_stat(p) : nullptr;
    PPO_TRY(launch_loss(ctx, p->act[L], bt.mask, bt.action, bt.old_prob, bt.adv, nb, A, epsilon, entropy_weight,
                        inv_nb_global, p->dlogits, p->d_loss_partials, p->d_loss_hist + (d_step ? 0 : 2 * slot), nullptr,
                        d_step, dl_stat));
    PPO_TRY(policy_backward(p, bt.feat, M, dl_stat != nullptr));
    // fp16-split engine: optimiser step, weight statistics, scales, operand copies and the minibatch counter in one launch
    const bool fused = opt != nullptr && p->gemm_mode == PPO_GEMM_F16X3_TC;
    if (ctx->nccl_comm != nullptr && ctx->nranks > 1 && p2p_active(p)) {
        // gradient all-reduce over NVLink peer memory, fused into the Adam kernel (dp_p2p.cu)
        if (fused) {
            P2PView xv{};
            PPO_TRY(p2p_publish(p, &xv));
            return f16_adam_refresh(p, opt, &xv, advance_step);
        }
        PPO_TRY(p2p_reduce_and_step(p, opt));
        if (opt != nullptr) PPO_TRY(refresh_engine_weights(p));
        if (advance_step != nullptr) PPO_TRY(launch_step_advance(ctx, advance_step));
        return PPO_OK;
    }
    if (ctx->nccl_comm != nullptr && ctx->nranks > 1) {
        // the fp16-split engine all-reduces every layer's gradient as soon as it is complete (overlapped with the rest of
        // the backward pass); the other engines reduce the whole flat vector here
        if (p->gemm_mode == PPO_GEMM_F16X3_TC && dp_overlap(ctx)) PPO_TRY(grads_join(ctx));
        else PPO_TRY(nccl_allreduce_f32(ctx, p->grads, p->P));
    }
    if (fused) return f16_adam_refresh(p, opt, nullptr, advance_step);
    if (opt != nullptr) {
        PPO_TRY(launch_adam(ctx, p->params, opt->m, opt->v, p->grads, p->P, opt->eta, opt->beta1, opt->beta2,
                            opt->eps, opt->d_bp, 1.0f));
        PPO_TRY(refresh_engine_weights(p));
    }
    if (advance_step != nullptr) PPO_TRY(launch_step_advance(ctx, advance_step));
    return PPO_OK;
}

int gather_into(ppo_buf* buf, ppo_batch& bt, const int* d_index, int64_t count, int variant,
                const int* d_step = nullptr, int64_t step_stride = 0) {
    GatherArgs a{};
    a.step = d_step; a.step_stride = step_stride;
    a.feat = buf->feat; a.mask = buf->mask; a.action = buf->action; a.old_prob = buf->old_prob; a.ret = buf->reward;
    a.index = d_index; a.count = count; a.feat_elems = buf->nf * buf->nhe; a.mask_elems = buf->A;
    a.feat_out = bt.feat; a.mask_out = bt.mask; a.action_out = bt.action; a.prob_out = bt.old_prob;
    a.adv_out = bt.adv;
    a.norm = buf->normalize ? buf->d_norm : nullptr;
    bt.feat_bound = buf->d_feat_absmax;       // a bound of every row this gather can deliver
    return launch_gather(buf->ctx, a, variant);
}

int read_batch_to_host(ppo_ctx* ctx, const ppo_batch& bt, int64_t count, int feat_elems, int A, float* feat_out,
                       float* mask_out, int64_t* action_out, float* prob_out, float* ret_out) {
    if (feat_out) PPO_TRY(d2h(ctx, feat_out, bt.feat, (size_t)count * feat_elems * 4));
    if (mask_out) PPO_TRY(d2h(ctx, mask_out, bt.mask, (size_t)count * A * 4));
    if (prob_out) PPO_TRY(d2h(ctx, prob_out, bt.old_prob, (size_t)count * 4));
    if (ret_out) PPO_TRY(d2h(ctx, ret_out, bt.adv, (size_t)count * 4));
    if (action_out) {
        PPO_TRY(ensure_scratch(ctx, (size_t)count * 8));
        PPO_TRY(launch_convert_actions_out(ctx, bt.action, (int64_t*)ctx->d_scratch, count));
        PPO_TRY(d2h(ctx, action_out, ctx->d_scratch, (size_t)count * 8));
    }
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

// BufferRollouts grows without bound (push!, src/rollout_buffer.jl:24-38); the device buffer's capacity is an initial
// reservation that grows geometrically: new arrays, device-to-device copies of the n live transitions, old arrays freed
int grow_buffer(ppo_buf* buf, int64_t need) {
    ppo_ctx* ctx = buf->ctx;
    int64_t cap = std::max<int64_t>(need, buf->cap + buf->cap / 2 + 1024);
    PPO_REQUIRE(cap < ((int64_t)1 << 31), "append: %lld transitions exceed the buffer's index range", (long long)need);
    const int64_t fe = (int64_t)buf->nf * buf->nhe;
    const int64_t cap16 = round_up(cap, SCAN_TILE), old16 = round_up(buf->cap, SCAN_TILE);
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    auto move = [&](auto*& arr, size_t new_count, size_t live) -> int {
        using T = std::remove_reference_t<decltype(*arr)>;
        T* fresh = nullptr;
        PPO_TRY(dev_alloc(&fresh, new_count));
        if (arr != nullptr && live > 0)
            PPO_CUDA(cudaMemcpyAsync(fresh, arr, live * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        PPO_CUDA(cudaStreamSynchronize(ctx->stream));
        dev_free(arr);
        arr = fresh;
        return PPO_OK;
    };
    const size_t n = (size_t)buf->n;
    PPO_TRY(move(buf->feat, (size_t)cap * fe, n * fe));
    PPO_TRY(move(buf->mask, (size_t)cap * buf->A, n * buf->A));
    PPO_TRY(move(buf->action, (size_t)cap, n));
    PPO_TRY(move(buf->old_prob, (size_t)cap, n));
    PPO_TRY(move(buf->reward, (size_t)cap16, std::min<size_t>(n, (size_t)old16)));
    PPO_TRY(move(buf->reward_alt, (size_t)cap16, 0));
    PPO_TRY(move(buf->terminal, (size_t)cap16, n));
    PPO_TRY(move(buf->perm, (size_t)cap, (size_t)buf->perm_len));
    PPO_TRY(move(buf->d_tile_stats, (size_t)2 * SCAN_STATS_PER_TILE * ceil_div(cap, SCAN_TILE), 0));
    if (buf->reward_saved) PPO_TRY(move(buf->reward_saved, (size_t)cap, (size_t)buf->saved_n));
    buf->stats_valid = false;
    buf->cap = cap;
    return PPO_OK;
}

// feat_bytes: element size of the host features: 4 = Float32 (copied as is), 8 = Int64, 1 / 2 = Int8 / Int16 (staged in
// scratch, widened to Float32 on the device; exact)
// mask_bits != nullptr: the action mask as one bit per action (1 = allowed) instead of Float32 0 / -Inf
int append_common(ppo_buf* buf, int64_t n, const void* feat, int feat_bytes, const float* mask, const int64_t* action,
                  const float* old_prob, const float* reward, const uint8_t* terminal, const uint64_t* mask_bits = nullptr) {
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(n >= 0, "append: n < 0");
    if (n == 0) return PPO_OK;
    PPO_REQUIRE(feat && (mask || mask_bits) && action && old_prob && reward && terminal, "append: null input");
    if (buf->n + n > buf->cap) PPO_TRY(grow_buffer(buf, buf->n + n));
    const int64_t fe = (int64_t)buf->nf * buf->nhe;
    const int64_t off = buf->n;
    const size_t act_bytes = (size_t)round_up(n * 8, 64);
    size_t scratch = 64 + act_bytes;
    const size_t feat_scratch = feat_bytes != 4 ? (size_t)round_up(n * fe * feat_bytes, 64) : 0;
    const size_t bit_words = mask_bits != nullptr ? (size_t)ceil_div(n * buf->A, 64) : 0;
    scratch += feat_scratch + bit_words * 8;
    PPO_TRY(ensure_scratch(ctx, scratch));
    int* d_bad = (int*)ctx->d_scratch;
    int64_t* d_act = (int64_t*)((char*)ctx->d_scratch + 64);
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    if (feat_bytes != 4) {
        void* d_f = (char*)ctx->d_scratch + 64 + act_bytes;
        PPO_TRY(h2d(ctx, d_f, feat, (size_t)n * fe * feat_bytes));
        // (the buffer keeps max |feature| of everything it holds: the fp16-split engine's bound for any minibatch)
        if (feat_bytes == 8) {
            PPO_TRY(launch_i64_to_f32(ctx, (const int64_t*)d_f, buf->feat + off * fe, n * fe));
            PPO_TRY(launch_absmax_f32(ctx, buf->feat + off * fe, n * fe, buf->d_feat_absmax));
        } else {
            PPO_TRY(launch_narrow_to_f32(ctx, d_f, feat_bytes, buf->feat + off * fe, n * fe, buf->d_feat_absmax));
        }
    } else {
        PPO_TRY(h2d(ctx, buf->feat + off * fe, feat, (size_t)n * fe * 4));
        PPO_TRY(launch_absmax_f32(ctx, buf->feat + off * fe, n * fe, buf->d_feat_absmax));
    }
    if (mask_bits != nullptr) {
        uint64_t* d_bits = (uint64_t*)((char*)ctx->d_scratch + 64 + act_bytes + feat_scratch);
        PPO_TRY(h2d(ctx, d_bits, mask_bits, bit_words * 8));
        PPO_TRY(launch_mask_from_bits(ctx, d_bits, buf->mask + off * buf->A, n * buf->A));
    } else {
        PPO_TRY(h2d(ctx, buf->mask + off * buf->A, mask, (size_t)n * buf->A * 4));
    }
    PPO_TRY(h2d(ctx, buf->old_prob + off, old_prob, (size_t)n * 4));
    PPO_TRY(h2d(ctx, buf->reward + off, reward, (size_t)n * 4));
    PPO_TRY(h2d(ctx, buf->terminal + off, terminal, (size_t)n));
    PPO_TRY(launch_normalize_bool(ctx, buf->terminal + off, n));   // any non-zero byte is `true`; the scan assumes 0/1
    PPO_TRY(h2d(ctx, d_act, action, (size_t)n * 8));
    PPO_TRY(launch_convert_actions_in(ctx, d_act, buf->action + off, n, buf->A, d_bad));
    PPO_TRY(check_bad_flag(ctx, d_bad, "append: selected action outside 1..A"));
    buf->n += n;
    buf->stats_valid = false;
    buf->returns_valid = false;
    return PPO_OK;
}

}  // namespace

namespace {
// leakyrelu' gate of an fp32 activation (FFMA and tf32 engines gate on `act > 0`)
__global__ void __launch_bounds__(256)
gates_from_f32_kernel(const float* __restrict__ act, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = act[i] > 0.0f ? 1 : 0;
}

__global__ void __launch_bounds__(256)
l2_read_sweep_kernel(const uint4* __restrict__ p, size_t n16, unsigned* sink) {
    unsigned acc = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9e3779b9u) *sink = acc;      // never true for the zero-filled buffer; keeps the loads alive
}
}  // namespace

// L2 flush between timed launches: WRITE 256 MB (> the 126 MB L2), then READ a second 256 MB region.  After the write
// alone the L2 is full of dirty lines whose write-back (up to 126 MB of extra DRAM traffic, ~19 us) would land inside
// the next timed kernel; the read sweep evicts them before the timer starts and leaves only clean lines behind.
int flush_l2(ppo_ctx* ctx) {
    if (!ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;
        PPO_CUDA(cudaMalloc(&ctx->d_flush, 2 * ctx->flush_bytes + 16));
        PPO_CUDA(cudaMemsetAsync(ctx->d_flush, 0, 2 * ctx->flush_bytes + 16, ctx->stream));
    }
    PPO_CUDA(cudaMemsetAsync(ctx->d_flush, 0, ctx->flush_bytes, ctx->stream));
    const uint4* second = reinterpret_cast<const uint4*>((const char*)ctx->d_flush + ctx->flush_bytes);
    l2_read_sweep_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(second, ctx->flush_bytes / 16,
                                                                   reinterpret_cast<unsigned*>((char*)ctx->d_flush + 2 * ctx->flush_bytes));
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo

using namespace ppo;

// =============================================================================================
extern "C" {

const char* ppo_last_error(void) { return ppo::g_err; }
const char* ppo_version(void) { return "ppo_b200 0.1 (sm_100a)"; }

int ppo_ctx_create(int device, ppo_ctx** out) {
    PPO_REQUIRE(out != nullptr, "ctx_create: null out");
    *out = nullptr;
    int count = 0;
    PPO_CUDA(cudaGetDeviceCount(&count));
    PPO_REQUIRE(device >= 0 && device < count, "ctx_create: device %d of %d", device, count);
    PPO_CUDA(cudaSetDevice(device));
    ppo_ctx* c = new (std::nothrow) ppo_ctx();
    if (!c) { set_error("out of host memory"); return PPO_ERR_NOMEM; }
    c->device = device;
    cudaDeviceProp prop;
    PPO_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    PPO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->pinned_doubles = 1 << 12;
    PPO_CUDA(cudaMallocHost((void**)&c->h_pinned, sizeof(double) * (size_t)c->pinned_doubles));
    PPO_CUDA(cudaMalloc((void**)&c->d_step, sizeof(int)));
    PPO_CUDA(cudaMemset(c->d_step, 0, sizeof(int)));
    *out = c;
    return PPO_OK;
}

// children call this from their destroy functions (after releasing their own device memory)
static void ctx_release_child(ppo_ctx* ctx) {
    if (ctx == nullptr) return;
    if (--ctx->children <= 0 && ctx->dead) { ctx->dead = false; ppo_ctx_destroy(ctx); }
}

int ppo_ctx_destroy(ppo_ctx* ctx) {
    if (!ctx) return PPO_OK;
    if (ctx->children > 0) { ctx->dead = true; return PPO_OK; }      // released by the last child (see ppo_ctx::children)
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_destroy(ctx);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    if (ctx->d_step) cudaFree(ctx->d_step);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PPO_OK;
}

int ppo_sync(ppo_ctx* ctx) {
    PPO_TRY(use(ctx));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

int64_t ppo_ctx_launch_count(ppo_ctx* ctx) { return ctx ? ctx->launches : -1; }
void* ppo_ctx_stream(ppo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- communicator -----------------------------------------------------------------------------
int ppo_comm_unique_id(void* id128) {
    PPO_REQUIRE(id128 != nullptr, "null id");
    return nccl_unique_id(id128);
}
int ppo_comm_init(ppo_ctx* ctx, int nranks, int rank, const void* id128) {
    PPO_TRY(use(ctx));
    PPO_REQUIRE(id128 != nullptr, "null id");
    return nccl_init(ctx, nranks, rank, id128);
}
int ppo_comm_destroy(ppo_ctx* ctx) {
    PPO_TRY(use(ctx));
    return nccl_destroy(ctx);
}
int ppo_comm_allreduce_f64(ppo_ctx* ctx, double* host_inout, int n) {
    PPO_TRY(use(ctx));
    PPO_REQUIRE(n >= 0 && host_inout != nullptr, "allreduce: bad args");
    if (ctx->nccl_comm == nullptr || ctx->nranks == 1 || n == 0) return PPO_OK;
    PPO_TRY(ensure_scratch(ctx, (size_t)n * 8));
    PPO_TRY(h2d(ctx, ctx->d_scratch, host_inout, (size_t)n * 8));
    PPO_TRY(nccl_allreduce_f64(ctx, (double*)ctx->d_scratch, n));
    PPO_TRY(d2h(ctx, host_inout, ctx->d_scratch, (size_t)n * 8));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

// ---- rollout buffer ---------------------------------------------------------------------------
int ppo_buffer_create(ppo_ctx* ctx, int64_t capacity, int nf, int nhe, int apa, ppo_buf** out) {
    PPO_TRY(use(ctx));
    PPO_REQUIRE(out != nullptr, "buffer_create: null out");
    *out = nullptr;
    PPO_REQUIRE(capacity >= 1 && capacity < ((int64_t)1 << 31), "buffer_create: capacity %lld", (long long)capacity);
    PPO_REQUIRE(nf >= 1 && nhe >= 1 && apa >= 1, "buffer_create: nf=%d nhe=%d apa=%d", nf, nhe, apa);
    ppo_buf* b = new (std::nothrow) ppo_buf();
    if (!b) { set_error("out of host memory"); return PPO_ERR_NOMEM; }
    b->ctx = ctx; b->cap = capacity; b->nf = nf; b->nhe = nhe; b->apa = apa; b->A = nhe * apa;
    ctx->children += 1;
    const int64_t cap16 = round_up(capacity, SCAN_TILE);   // scan reads whole 16-item chunks only when in range
    int s = PPO_OK;
    if ((s = dev_alloc(&b->feat, (size_t)capacity * nf * nhe)) != PPO_OK ||
        (s = dev_alloc(&b->mask, (size_t)capacity * b->A)) != PPO_OK ||
        (s = dev_alloc(&b->action, (size_t)capacity)) != PPO_OK ||
        (s = dev_alloc(&b->old_prob, (size_t)capacity)) != PPO_OK ||
        (s = dev_alloc(&b->reward, (size_t)cap16)) != PPO_OK ||
        (s = dev_alloc(&b->reward_alt, (size_t)cap16)) != PPO_OK ||
        (s = dev_alloc(&b->terminal, (size_t)cap16)) != PPO_OK ||
        (s = dev_alloc(&b->perm, (size_t)capacity)) != PPO_OK ||
        (s = dev_alloc(&b->d_norm, 2)) != PPO_OK ||
        (s = dev_alloc(&b->d_feat_absmax, 1)) != PPO_OK ||
        (s = dev_alloc(&b->d_tile_stats, (size_t)2 * SCAN_STATS_PER_TILE * ceil_div(capacity, SCAN_TILE))) != PPO_OK) {
        ppo_buffer_destroy(b);
        return s;
    }
    if (cudaMemsetAsync(b->d_feat_absmax, 0, sizeof(unsigned), ctx->stream) != cudaSuccess) {
        ppo_buffer_destroy(b);
        set_error("buffer_create: memset failed");
        return PPO_ERR_CUDA;
    }
    *out = b;
    return PPO_OK;
}

int ppo_buffer_destroy(ppo_buf* b) {
    if (!b) return PPO_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    dev_free(b->feat); dev_free(b->mask); dev_free(b->action); dev_free(b->old_prob); dev_free(b->reward);
    dev_free(b->terminal); dev_free(b->perm); dev_free(b->reward_saved); dev_free(b->reward_alt); dev_free(b->d_norm); dev_free(b->d_tile_stats); dev_free(b->d_feat_absmax);
    free_batch(b->batch);
    ppo_ctx* ctx = b->ctx;
    delete b;
    ctx_release_child(ctx);
    return PPO_OK;
}

int ppo_buffer_append(ppo_buf* buf, int64_t n, const float* feat, const float* mask, const int64_t* action,
                      const float* old_prob, const float* reward, const uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    return append_common(buf, n, feat, 4, mask, action, old_prob, reward, terminal);
}

int ppo_buffer_append_i64(ppo_buf* buf, int64_t n, const int64_t* feat, const float* mask, const int64_t* action,
                          const float* old_prob, const float* reward, const uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    return append_common(buf, n, feat, 8, mask, action, old_prob, reward, terminal);
}

int ppo_buffer_append_i8(ppo_buf* buf, int64_t n, const int8_t* feat, const float* mask, const int64_t* action,
                         const float* old_prob, const float* reward, const uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    return append_common(buf, n, feat, 1, mask, action, old_prob, reward, terminal);
}

int ppo_buffer_append_packed(ppo_buf* buf, int64_t n, const void* feat, int feat_elem_bytes, const uint64_t* mask_bits,
                             const int64_t* action, const float* old_prob, const float* reward, const uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "append: null buffer");
    PPO_REQUIRE(feat_elem_bytes == 1 || feat_elem_bytes == 2 || feat_elem_bytes == 4 || feat_elem_bytes == 8,
                "append_packed: feature element size %d (1 = Int8, 2 = Int16, 4 = Float32, 8 = Int64)", feat_elem_bytes);
    PPO_REQUIRE(mask_bits != nullptr, "append_packed: null mask bits");
    return append_common(buf, n, feat, feat_elem_bytes, nullptr, action, old_prob, reward, terminal, mask_bits);
}

int ppo_buffer_append_i16(ppo_buf* buf, int64_t n, const int16_t* feat, const float* mask, const int64_t* action,
                          const float* old_prob, const float* reward, const uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    return append_common(buf, n, feat, 2, mask, action, old_prob, reward, terminal);
}

int64_t ppo_buffer_length(ppo_buf* buf) { return buf ? buf->n : -1; }

int ppo_buffer_clear(ppo_buf* buf) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    buf->n = 0;
    buf->perm_len = 0;
    buf->stats_valid = false;
    buf->returns_valid = false;
    PPO_TRY(use(buf->ctx));
    PPO_CUDA(cudaMemsetAsync(buf->d_feat_absmax, 0, sizeof(unsigned), buf->ctx->stream));
    return PPO_OK;
}

int ppo_compute_returns(ppo_buf* buf, double discount, int discount_is_f32) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    if (buf->n == 0) return PPO_OK;
    PPO_TRY(ensure_scratch(ctx, scan_scratch_bytes(buf->n)));
    // the K2 statistics ride along only when the normalisation extension is on (the reference has none)
    PPO_TRY(launch_returns_scan(ctx, buf->reward, buf->reward_alt, buf->terminal, buf->n, discount, discount_is_f32,
                                buf->normalize ? buf->d_tile_stats : nullptr, ctx->d_scratch));
    std::swap(buf->reward, buf->reward_alt);      // rollouts.rewards .= returns
    buf->n_tiles_stats = SCAN_STATS_PER_TILE * ceil_div(buf->n, SCAN_TILE);
    buf->stats_valid = buf->normalize != 0;
    buf->returns_valid = true;
    if (buf->normalize)
        PPO_TRY(launch_norm_finalize(ctx, buf->d_tile_stats, buf->n_tiles_stats, buf->n, buf->norm_eps, buf->d_norm));
    return PPO_OK;
}

int ppo_normalize_advantage(ppo_buf* buf, int enable, double eps) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    buf->normalize = enable ? 1 : 0;
    buf->norm_eps = eps;
    if (enable) {
        PPO_REQUIRE(buf->stats_valid || buf->returns_valid, "normalize_advantage: call ppo_compute_returns first");
        if (!buf->stats_valid) {
            buf->n_tiles_stats = SCAN_STATS_PER_TILE * ceil_div(buf->n, SCAN_TILE);
            PPO_TRY(launch_returns_stats(ctx, buf->reward, buf->n, buf->d_tile_stats));
            buf->stats_valid = true;
        }
        PPO_TRY(launch_norm_finalize(ctx, buf->d_tile_stats, buf->n_tiles_stats, buf->n, eps, buf->d_norm));
    }
    return PPO_OK;
}

int ppo_buffer_save_rewards(ppo_buf* buf) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    if (!buf->reward_saved) PPO_TRY(dev_alloc(&buf->reward_saved, (size_t)buf->cap));
    PPO_CUDA(cudaMemcpyAsync(buf->reward_saved, buf->reward, (size_t)buf->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    buf->saved_n = buf->n;
    return PPO_OK;
}

int ppo_buffer_restore_rewards(ppo_buf* buf) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    if (!buf->reward_saved || buf->saved_n != buf->n) {
        set_error("restore_rewards: no snapshot of the current %lld transitions", (long long)buf->n);
        return PPO_ERR_STATE;
    }
    PPO_CUDA(cudaMemcpyAsync(buf->reward, buf->reward_saved, (size_t)buf->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    buf->stats_valid = false;
    buf->returns_valid = false;
    return PPO_OK;
}

int ppo_buffer_read(ppo_buf* buf, int64_t start, int64_t count, float* feat, float* mask, int64_t* action,
                    float* old_prob, float* rewards_or_returns, uint8_t* terminal) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(start >= 0 && count >= 0 && start + count <= buf->n, "buffer_read: range [%lld, %lld) of %lld",
                (long long)start, (long long)(start + count), (long long)buf->n);
    const int64_t fe = (int64_t)buf->nf * buf->nhe;
    if (feat) PPO_TRY(d2h(ctx, feat, buf->feat + start * fe, (size_t)count * fe * 4));
    if (mask) PPO_TRY(d2h(ctx, mask, buf->mask + start * buf->A, (size_t)count * buf->A * 4));
    if (old_prob) PPO_TRY(d2h(ctx, old_prob, buf->old_prob + start, (size_t)count * 4));
    if (rewards_or_returns) PPO_TRY(d2h(ctx, rewards_or_returns, buf->reward + start, (size_t)count * 4));
    if (terminal) PPO_TRY(d2h(ctx, terminal, buf->terminal + start, (size_t)count));
    if (action) {
        PPO_TRY(ensure_scratch(ctx, (size_t)count * 8));
        PPO_TRY(launch_convert_actions_out(ctx, buf->action + start, (int64_t*)ctx->d_scratch, count));
        PPO_TRY(d2h(ctx, action, ctx->d_scratch, (size_t)count * 8));
    }
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

static int permute_with_device_index(ppo_buf* buf, const int* d_idx) {
    // gather every array into fresh storage, then swap pointers (no copy back)
    ppo_ctx* ctx = buf->ctx;
    const int64_t n = buf->n;
    const int64_t fe = (int64_t)buf->nf * buf->nhe;
    float *nfeat = nullptr, *nmask = nullptr, *nprob = nullptr, *nrew = nullptr;
    int* nact = nullptr;
    uint8_t* nterm = nullptr;
    PPO_TRY(dev_alloc(&nfeat, (size_t)buf->cap * fe));
    PPO_TRY(dev_alloc(&nmask, (size_t)buf->cap * buf->A));
    PPO_TRY(dev_alloc(&nact, (size_t)buf->cap));
    PPO_TRY(dev_alloc(&nprob, (size_t)buf->cap));
    PPO_TRY(dev_alloc(&nrew, (size_t)round_up(buf->cap, SCAN_TILE)));
    PPO_TRY(dev_alloc(&nterm, (size_t)round_up(buf->cap, SCAN_TILE)));
    GatherArgs a{};
    a.feat = buf->feat; a.mask = buf->mask; a.action = buf->action; a.old_prob = buf->old_prob; a.ret = buf->reward;
    a.index = d_idx; a.count = n; a.feat_elems = (int)fe; a.mask_elems = buf->A;
    a.feat_out = nfeat; a.mask_out = nmask; a.action_out = nact; a.prob_out = nprob; a.adv_out = nrew;
    a.norm = nullptr;
    PPO_TRY(launch_gather(ctx, a, 0));
    PPO_TRY(launch_permute_inplace_u8(ctx, buf->terminal, nterm, d_idx, n));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    std::swap(buf->feat, nfeat); std::swap(buf->mask, nmask); std::swap(buf->action, nact);
    std::swap(buf->old_prob, nprob); std::swap(buf->reward, nrew); std::swap(buf->terminal, nterm);
    dev_free(nfeat); dev_free(nmask); dev_free(nact); dev_free(nprob); dev_free(nrew); dev_free(nterm);
    buf->stats_valid = false;
    buf->returns_valid = false;
    buf->saved_n = 0;
    return PPO_OK;
}

int ppo_buffer_permute(ppo_buf* buf, const int64_t* idx1, int64_t n) {
    PPO_REQUIRE(buf != nullptr && idx1 != nullptr, "permute: null argument");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(n == buf->n, "permute: length(idx) = %lld != length(rollouts) = %lld", (long long)n,
                (long long)buf->n);   // @assert length(idx) == length(rollouts), rollout_buffer.jl:82
    if (n == 0) return PPO_OK;
    PPO_TRY(ensure_scratch(ctx, 64 + (size_t)n * 12));
    int* d_bad = (int*)ctx->d_scratch;
    int64_t* d_i64 = (int64_t*)((char*)ctx->d_scratch + 64);
    int* d_idx = (int*)(d_i64 + n);
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PPO_TRY(h2d(ctx, d_i64, idx1, (size_t)n * 8));
    PPO_TRY(launch_perm_from_host(ctx, d_i64, d_idx, n, buf->n, d_bad));
    PPO_TRY(check_bad_flag(ctx, d_bad, "permute: index outside 1..length(rollouts)"));
    return permute_with_device_index(buf, d_idx);
}

int ppo_buffer_shuffle(ppo_buf* buf, uint64_t seed) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    if (buf->n == 0) return PPO_OK;
    PPO_TRY(ensure_scratch(ctx, (size_t)buf->n * 4));
    int* d_idx = (int*)ctx->d_scratch;
    PPO_TRY(launch_feistel_permutation(ctx, d_idx, buf->n, seed));
    return permute_with_device_index(buf, d_idx);
}

// ---- dataset ----------------------------------------------------------------------------------
int ppo_permutation_set(ppo_buf* buf, const int64_t* perm1, int64_t n) {
    PPO_REQUIRE(buf != nullptr && perm1 != nullptr, "permutation_set: null argument");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(n >= 1 && n <= buf->cap, "permutation_set: n = %lld (capacity %lld)", (long long)n,
                (long long)buf->cap);
    PPO_TRY(ensure_scratch(ctx, 64 + (size_t)n * 8));
    int* d_bad = (int*)ctx->d_scratch;
    int64_t* d_i64 = (int64_t*)((char*)ctx->d_scratch + 64);
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PPO_TRY(h2d(ctx, d_i64, perm1, (size_t)n * 8));
    PPO_TRY(launch_perm_from_host(ctx, d_i64, buf->perm, n, buf->n, d_bad));
    buf->perm_len = 0;
    PPO_TRY(check_bad_flag(ctx, d_bad, "permutation_set: index outside 1..length(dataset)"));
    buf->perm_len = n;
    return PPO_OK;
}

int ppo_permutation_generate(ppo_buf* buf, uint64_t seed, int64_t* perm1_out) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(buf->n >= 1, "permutation_generate: empty buffer");
    PPO_TRY(launch_feistel_permutation(ctx, buf->perm, buf->n, seed));
    buf->perm_len = buf->n;
    if (perm1_out) {
        PPO_TRY(ensure_scratch(ctx, (size_t)buf->n * 8));
        PPO_TRY(launch_perm_to_i64(ctx, buf->perm, (int64_t*)ctx->d_scratch, buf->n));
        PPO_TRY(d2h(ctx, perm1_out, ctx->d_scratch, (size_t)buf->n * 8));
        PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PPO_OK;
}

int ppo_gather_device(ppo_buf* buf, int64_t start, int64_t count, int variant) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    if (buf->perm_len == 0) { set_error("gather: no permutation set"); return PPO_ERR_STATE; }
    PPO_REQUIRE(start >= 0 && count >= 1 && start + count <= buf->perm_len, "gather: range [%lld, %lld) of %lld",
                (long long)start, (long long)(start + count), (long long)buf->perm_len);
    PPO_TRY(ensure_batch(ctx, buf->batch, count, buf->nf * buf->nhe, buf->A));
    return gather_into(buf, buf->batch, buf->perm + start, count, variant);
}

int ppo_batch_read(ppo_buf* buf, int64_t count, float* feat_out, float* mask_out, int64_t* action_out,
                   float* prob_out, float* returns_out) {
    PPO_REQUIRE(buf != nullptr, "null buffer");
    PPO_TRY(use(buf->ctx));
    PPO_REQUIRE(count >= 0 && count <= buf->batch.cap, "batch_read: count %lld > staged rows %lld", (long long)count,
                (long long)buf->batch.cap);
    return read_batch_to_host(buf->ctx, buf->batch, count, buf->nf * buf->nhe, buf->A, feat_out, mask_out,
                              action_out, prob_out, returns_out);
}

int ppo_gather(ppo_buf* buf, int64_t start, int64_t count, float* feat_out, float* mask_out, int64_t* action_out,
               float* prob_out, float* returns_out) {
    PPO_TRY(ppo_gather_device(buf, start, count, 0));
    return read_batch_to_host(buf->ctx, buf->batch, count, buf->nf * buf->nhe, buf->A, feat_out, mask_out,
                              action_out, prob_out, returns_out);
}

int ppo_gather_indices(ppo_buf* buf, const int64_t* idx1, int64_t count, float* feat_out, float* mask_out,
                       int64_t* action_out, float* prob_out, float* returns_out) {
    PPO_REQUIRE(buf != nullptr && idx1 != nullptr, "gather_indices: null argument");
    ppo_ctx* ctx = buf->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(count >= 1, "gather_indices: empty index vector");
    PPO_TRY(ensure_batch(ctx, buf->batch, count, buf->nf * buf->nhe, buf->A));
    PPO_TRY(ensure_scratch(ctx, 64 + (size_t)count * 12));
    int* d_bad = (int*)ctx->d_scratch;
    int64_t* d_i64 = (int64_t*)((char*)ctx->d_scratch + 64);
    int* d_idx = (int*)(d_i64 + count);
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PPO_TRY(h2d(ctx, d_i64, idx1, (size_t)count * 8));
    PPO_TRY(launch_perm_from_host(ctx, d_i64, d_idx, count, buf->n, d_bad));
    PPO_TRY(check_bad_flag(ctx, d_bad, "dataset index outside 1..length(dataset)"));
    PPO_TRY(gather_into(buf, buf->batch, d_idx, count, 0));
    // the scratch holding d_idx is reused by read_batch_to_host only after the gather is enqueued
    // on the same stream, so ordering is preserved.
    return read_batch_to_host(ctx, buf->batch, count, buf->nf * buf->nhe, buf->A, feat_out, mask_out, action_out,
                              prob_out, returns_out);
}

// ---- policy -------------------------------------------------------------------------------------
int ppo_policy_create(ppo_ctx* ctx, int n_layers, const int* dims, const float* const* W, const float* const* b,
                      float leaky_slope, ppo_policy** out) {
    PPO_TRY(use(ctx));
    PPO_REQUIRE(out != nullptr, "policy_create: null out");
    *out = nullptr;
    PPO_REQUIRE(n_layers >= 1 && n_layers <= 64 && dims != nullptr, "policy_create: n_layers = %d", n_layers);
    for (int l = 0; l <= n_layers; ++l) PPO_REQUIRE(dims[l] >= 1, "policy_create: dims[%d] = %d", l, dims[l]);
    ppo_policy* p = new (std::nothrow) ppo_policy();
    if (!p) { set_error("out of host memory"); return PPO_ERR_NOMEM; }
    p->ctx = ctx; p->L = n_layers; p->slope = leaky_slope;
    ctx->children += 1;
    p->dims.assign(dims, dims + n_layers + 1);
    int64_t off = 0;
    for (int l = 0; l < n_layers; ++l) {
        // Flux.params order, densely packed (no padding): kernels that use vector loads check the alignment of a
        // segment at run time and fall back to scalar accesses
        p->w_off.push_back(off); off += (int64_t)dims[l] * dims[l + 1];
        p->b_off.push_back(off); off += dims[l + 1];
    }
    p->P = off;
    int s;
    if ((s = dev_alloc(&p->params, (size_t)p->P)) != PPO_OK || (s = dev_alloc(&p->grads, (size_t)p->P)) != PPO_OK) {
        ppo_policy_destroy(p);
        return s;
    }
    PPO_CUDA(cudaMemsetAsync(p->grads, 0, (size_t)p->P * 4, ctx->stream));
    if (W != nullptr && b != nullptr) {
        s = ppo_policy_write(p, W, b);
        if (s != PPO_OK) { ppo_policy_destroy(p); return s; }
    } else {
        PPO_CUDA(cudaMemsetAsync(p->params, 0, (size_t)p->P * 4, ctx->stream));
    }
    // default engine: the fastest one whose shape contract the policy meets (every binding sees the same default)
    s = ppo_policy_set_gemm_mode(p, PPO_GEMM_AUTO);
    if (s != PPO_OK) { ppo_policy_destroy(p); return s; }
    *out = p;
    return PPO_OK;
}

int ppo_policy_destroy(ppo_policy* p) {
    if (!p) return PPO_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    free_workspace(p);
    tc_destroy(p);
    f16_destroy(p);
    p2p_destroy(p);
    dev_free(p->params); dev_free(p->grads); dev_free(p->d_loss_partials); dev_free(p->d_loss_hist);
    free_batch(p->hbatch);
    ppo_ctx* ctx = p->ctx;
    delete p;
    ctx_release_child(ctx);
    return PPO_OK;
}

int ppo_policy_read(ppo_policy* p, float* const* W, float* const* b) {
    PPO_REQUIRE(p != nullptr && W != nullptr && b != nullptr, "policy_read: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    for (int l = 0; l < p->L; ++l) {
        PPO_REQUIRE(W[l] != nullptr && b[l] != nullptr, "policy_read: null layer %d", l);
        PPO_TRY(d2h(ctx, W[l], p->params + p->w_off[l], (size_t)p->dims[l] * p->dims[l + 1] * 4));
        PPO_TRY(d2h(ctx, b[l], p->params + p->b_off[l], (size_t)p->dims[l + 1] * 4));
    }
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

int ppo_policy_write(ppo_policy* p, const float* const* W, const float* const* b) {
    PPO_REQUIRE(p != nullptr && W != nullptr && b != nullptr, "policy_write: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    for (int l = 0; l < p->L; ++l) {
        PPO_REQUIRE(W[l] != nullptr && b[l] != nullptr, "policy_write: null layer %d", l);
        PPO_TRY(h2d(ctx, p->params + p->w_off[l], W[l], (size_t)p->dims[l] * p->dims[l + 1] * 4));
        PPO_TRY(h2d(ctx, p->params + p->b_off[l], b[l], (size_t)p->dims[l + 1] * 4));
    }
    PPO_TRY(refresh_engine_weights(p));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

int ppo_policy_set_gemm_mode(ppo_policy* p, int mode) {
    PPO_REQUIRE(p != nullptr, "null policy");
    PPO_TRY(use(p->ctx));
    if (mode == PPO_GEMM_AUTO) {
        // the engines refuse shapes outside their contract with PPO_ERR_INVALID and no side effect: walk down the list
        if (f16_prepare(p) == PPO_OK) mode = PPO_GEMM_F16X3_TC;
        else if (tc_prepare(p, PPO_GEMM_TF32X3_TC) == PPO_OK) mode = PPO_GEMM_TF32X3_TC;
        else mode = PPO_GEMM_FP32_SIMT;
        p->gemm_mode = mode;
        return refresh_engine_weights(p);
    }
    PPO_REQUIRE(mode == PPO_GEMM_FP32_SIMT || mode == PPO_GEMM_TF32X3_TC || mode == PPO_GEMM_BF16_TC ||
                mode == PPO_GEMM_F16X3_TC, "set_gemm_mode: unknown mode %d", mode);
    if (mode == PPO_GEMM_F16X3_TC) PPO_TRY(f16_prepare(p));
    else if (mode != PPO_GEMM_FP32_SIMT) PPO_TRY(tc_prepare(p, mode));
    p->gemm_mode = mode;
    return refresh_engine_weights(p);
}

int ppo_policy_p2p_export(ppo_policy* p, void* handle64) {
    PPO_REQUIRE(p != nullptr && handle64 != nullptr, "p2p_export: null argument");
    PPO_TRY(use(p->ctx));
    return p2p_export(p, handle64);
}
int ppo_policy_p2p_connect(ppo_policy* p, int nranks, int rank, const void* handles) {
    PPO_REQUIRE(p != nullptr && handles != nullptr, "p2p_connect: null argument");
    PPO_TRY(use(p->ctx));
    PPO_REQUIRE(p->ctx->nranks == nranks && p->ctx->rank == rank, "p2p_connect: communicator is rank %d of %d", p->ctx->rank,
                p->ctx->nranks);
    return p2p_connect(p, nranks, rank, handles);
}

int ppo_policy_get_gemm_mode(ppo_policy* p) { return p ? p->gemm_mode : PPO_ERR_INVALID; }

int ppo_policy_p2p_wait(ppo_policy* p, int64_t* total_ns, int64_t* waits, int reset) {
    PPO_REQUIRE(p != nullptr && total_ns != nullptr && waits != nullptr, "p2p_wait: null argument");
    PPO_TRY(use(p->ctx));
    return p2p_wait_stats(p, total_ns, waits, reset);
}

int ppo_policy_set_token_compaction(ppo_policy* p, int enable) {
    PPO_REQUIRE(p != nullptr, "set_token_compaction: null policy");
    p->compact_tokens = enable ? 1 : 0;
    return PPO_OK;
}

int ppo_policy_active_tokens(ppo_policy* p, int64_t* active_out) {
    PPO_REQUIRE(p != nullptr && active_out != nullptr, "active_tokens: null argument");
    PPO_TRY(use(p->ctx));
    *active_out = -1;
    if (p->gemm_mode == PPO_GEMM_F16X3_TC && p->f16 != nullptr) return f16_active_tokens(p, active_out);
    return PPO_OK;
}

int ppo_policy_read_gates(ppo_policy* p, int layer, int64_t rows, uint8_t* gates_out) {
    PPO_REQUIRE(p != nullptr && gates_out != nullptr, "read_gates: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(layer >= 1 && layer <= p->L - 1, "read_gates: hidden layer %d of %d", layer, p->L - 1);
    PPO_REQUIRE(rows >= 1 && rows <= p->ws_tokens, "read_gates: %lld rows, the last pass held %lld", (long long)rows,
                (long long)p->ws_tokens);
    const int64_t n = rows * p->dims[layer];
    PPO_TRY(ensure_scratch(ctx, (size_t)n));
    uint8_t* d_out = (uint8_t*)ctx->d_scratch;
    if (p->gemm_mode == PPO_GEMM_F16X3_TC) {
        PPO_TRY(f16_read_gates(p, layer, rows, d_out));
    } else {
        const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
        gates_from_f32_kernel<<<blocks, 256, 0, ctx->stream>>>(p->act[layer], n, d_out);
        ctx->launches += 1;
        PPO_CUDA(cudaGetLastError());
    }
    PPO_TRY(d2h(ctx, gates_out, d_out, (size_t)n));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

int64_t ppo_policy_num_params(ppo_policy* p) { return p ? p->P : -1; }

static int upload_host_batch(ppo_policy* p, int64_t nb, int nhe, const float* feat, const float* mask) {
    ppo_ctx* ctx = p->ctx;
    const int nf = p->dims[0];
    const int A = nhe * p->dims[p->L];
    PPO_TRY(ensure_batch(ctx, p->hbatch, nb, nf * nhe, A));
    PPO_TRY(h2d(ctx, p->hbatch.feat, feat, (size_t)nb * nhe * nf * 4));
    PPO_TRY(h2d(ctx, p->hbatch.mask, mask, (size_t)nb * A * 4));
    return PPO_OK;
}

// forward + masked softmax of nb host states; the probabilities [nb][A] are left in ctx->d_scratch
static int probabilities_to_scratch(ppo_policy* p, int64_t nb, int nhe, const float* feat, const float* mask, size_t extra_scratch) {
    ppo_ctx* ctx = p->ctx;
    const int A = nhe * p->dims[p->L];
    const int64_t M = nb * nhe;
    PPO_TRY(upload_host_batch(p, nb, nhe, feat, mask));
    PPO_CUDA(cudaMemsetAsync(p->hbatch.action, 0, (size_t)nb * 4, ctx->stream));
    PPO_CUDA(cudaMemsetAsync(p->hbatch.adv, 0, (size_t)nb * 4, ctx->stream));
    // old_prob = 1.0f bit pattern is not memset-able; adv = 0 makes the ratio irrelevant, but keep it finite
    PPO_CUDA(cudaMemsetAsync(p->hbatch.old_prob, 0x3f, (size_t)nb * 4, ctx->stream));
    PPO_TRY(ensure_workspace(p, M));
    PPO_TRY(ensure_loss_buffers(p, nb, A, 1));
    PPO_TRY(ensure_scratch(ctx, (size_t)nb * A * 4 + extra_scratch));
    PPO_TRY(policy_forward(p, p->hbatch.feat, M, p->hbatch.mask));
    return launch_loss(ctx, p->act[p->L], p->hbatch.mask, p->hbatch.action, p->hbatch.old_prob, p->hbatch.adv, nb, A,
                       0.0, 0.0, 1.0 / (double)nb, nullptr, p->d_loss_partials, p->d_loss_hist, (float*)ctx->d_scratch);
}

int ppo_batch_action_probabilities(ppo_policy* p, int64_t nb, int nhe, const float* feat, const float* mask,
                                   float* probs_out) {
    PPO_REQUIRE(p != nullptr && feat && mask && probs_out, "batch_action_probabilities: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(nb >= 1 && nhe >= 1, "batch_action_probabilities: nb=%lld nhe=%d", (long long)nb, nhe);
    const int A = nhe * p->dims[p->L];
    PPO_TRY(probabilities_to_scratch(p, nb, nhe, feat, mask, 0));
    PPO_TRY(d2h(ctx, probs_out, ctx->d_scratch, (size_t)nb * A * 4));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

int ppo_sample_actions(ppo_policy* p, int64_t nb, int nhe, const float* feat, const float* mask, uint64_t seed,
                       int64_t* action1_out, float* prob_out, float* probs_out) {
    PPO_REQUIRE(p != nullptr && feat && mask && action1_out && prob_out, "sample_actions: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(nb >= 1 && nhe >= 1, "sample_actions: nb=%lld nhe=%d", (long long)nb, nhe);
    const int A = nhe * p->dims[p->L];
    const size_t pbytes = round_up((int64_t)nb * A * 4, 16);
    PPO_TRY(probabilities_to_scratch(p, nb, nhe, feat, mask, (size_t)nb * 16 + 32));
    int64_t* d_act = (int64_t*)((char*)ctx->d_scratch + pbytes);
    float* d_prob = (float*)((char*)ctx->d_scratch + pbytes + (size_t)nb * 8);
    PPO_TRY(launch_sample_actions(ctx, (const float*)ctx->d_scratch, nb, A, seed, d_act, d_prob));
    PPO_TRY(d2h(ctx, action1_out, d_act, (size_t)nb * 8));
    PPO_TRY(d2h(ctx, prob_out, d_prob, (size_t)nb * 4));
    if (probs_out) PPO_TRY(d2h(ctx, probs_out, ctx->d_scratch, (size_t)nb * A * 4));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

// ---- optimiser ------------------------------------------------------------------------------------
int ppo_adam_create(ppo_policy* p, double eta, double beta1, double beta2, double eps, ppo_opt** out) {
    PPO_REQUIRE(p != nullptr && out != nullptr, "adam_create: null argument");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    *out = nullptr;
    ppo_opt* o = new (std::nothrow) ppo_opt();
    if (!o) { set_error("out of host memory"); return PPO_ERR_NOMEM; }
    o->ctx = ctx; o->policy = p; o->eta = eta; o->beta1 = beta1; o->beta2 = beta2; o->eps = eps;
    ctx->children += 1;
    int s;
    if ((s = dev_alloc(&o->m, (size_t)p->P)) != PPO_OK || (s = dev_alloc(&o->v, (size_t)p->P)) != PPO_OK ||
        (s = dev_alloc(&o->d_bp, 2)) != PPO_OK) {
        ppo_adam_destroy(o);
        return s;
    }
    PPO_CUDA(cudaMemsetAsync(o->m, 0, (size_t)p->P * 4, ctx->stream));
    PPO_CUDA(cudaMemsetAsync(o->v, 0, (size_t)p->P * 4, ctx->stream));
    double bp[2] = {beta1, beta2};   // Flux state starts at (beta1, beta2), i.e. t = 1
    PPO_CUDA(cudaMemcpyAsync(o->d_bp, bp, sizeof(bp), cudaMemcpyHostToDevice, ctx->stream));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = o;
    return PPO_OK;
}

int ppo_adam_destroy(ppo_opt* o) {
    if (!o) return PPO_OK;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    dev_free(o->m); dev_free(o->v); dev_free(o->d_bp);
    ppo_ctx* ctx = o->ctx;
    delete o;
    ctx_release_child(ctx);
    return PPO_OK;
}

int ppo_adam_set_eta(ppo_opt* o, double eta) {
    PPO_REQUIRE(o != nullptr, "null optimiser");
    o->eta = eta;
    return PPO_OK;
}
double ppo_adam_get_eta(ppo_opt* o) { return o ? o->eta : 0.0; }

int ppo_adam_update(ppo_opt* o, const float* grad_flat) {
    PPO_REQUIRE(o != nullptr && grad_flat != nullptr, "adam_update: null argument");
    ppo_policy* p = o->policy;
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    PPO_TRY(h2d(ctx, p->grads, grad_flat, (size_t)p->P * 4));
    PPO_TRY(launch_adam(ctx, p->params, o->m, o->v, p->grads, p->P, o->eta, o->beta1, o->beta2, o->eps, o->d_bp, 1.0f));
    PPO_TRY(refresh_engine_weights(p));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPO_OK;
}

// ---- loss / update ----------------------------------------------------------------------------------
int ppo_loss_from_logits(ppo_ctx* ctx, int64_t nb, int A, const float* logits, const float* mask,
                         const int64_t* action1, const float* old_prob, const float* advantage, double epsilon,
                         double entropy_weight, double* ppoloss, double* entropyloss, float* dlogits_out) {
    PPO_TRY(use(ctx));
    PPO_REQUIRE(logits && mask && action1 && old_prob && advantage, "loss_from_logits: null input");
    PPO_REQUIRE(nb >= 1 && A >= 1, "loss_from_logits: nb=%lld A=%d", (long long)nb, A);
    const size_t mat = (size_t)round_up(nb * A * 4, 256), vec = (size_t)round_up(nb * 4, 256);
    const int64_t blocks = loss_num_blocks(nb, A);
    const size_t total = 256 + 3 * mat + 3 * vec + (size_t)round_up(nb * 8, 256) + (size_t)blocks * 16 + 64;
    PPO_TRY(ensure_scratch(ctx, total));
    char* s = (char*)ctx->d_scratch;
    int* d_bad = (int*)s; s += 256;
    float* d_logits = (float*)s; s += mat;
    float* d_mask = (float*)s; s += mat;
    float* d_dl = (float*)s; s += mat;
    int* d_act = (int*)s; s += vec;
    float* d_old = (float*)s; s += vec;
    float* d_adv = (float*)s; s += vec;
    int64_t* d_a64 = (int64_t*)s; s += round_up(nb * 8, 256);
    double* d_part = (double*)s; s += (size_t)blocks * 16;
    double* d_out = (double*)s;
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PPO_TRY(h2d(ctx, d_logits, logits, (size_t)nb * A * 4));
    PPO_TRY(h2d(ctx, d_mask, mask, (size_t)nb * A * 4));
    PPO_TRY(h2d(ctx, d_old, old_prob, (size_t)nb * 4));
    PPO_TRY(h2d(ctx, d_adv, advantage, (size_t)nb * 4));
    PPO_TRY(h2d(ctx, d_a64, action1, (size_t)nb * 8));
    PPO_TRY(launch_convert_actions_in(ctx, d_a64, d_act, nb, A, d_bad));
    PPO_TRY(launch_loss(ctx, d_logits, d_mask, d_act, d_old, d_adv, nb, A, epsilon, entropy_weight, 1.0 / (double)nb,
                        dlogits_out ? d_dl : nullptr, d_part, d_out, nullptr));
    double out[2];
    PPO_TRY(d2h(ctx, out, d_out, sizeof(out)));
    if (dlogits_out) PPO_TRY(d2h(ctx, dlogits_out, d_dl, (size_t)nb * A * 4));
    PPO_TRY(check_bad_flag(ctx, d_bad, "loss_from_logits: action outside 1..A"));
    if (ppoloss) *ppoloss = out[0];
    if (entropyloss) *entropyloss = out[1];
    return PPO_OK;
}

static int finish_step(ppo_policy* p, int64_t slot, double entropy_weight, double* ppoloss, double* entw,
                       float* grads_out) {
    ppo_ctx* ctx = p->ctx;
    double out[2];
    PPO_TRY(d2h(ctx, out, p->d_loss_hist + 2 * slot, sizeof(out)));
    if (grads_out) PPO_TRY(d2h(ctx, grads_out, p->grads, (size_t)p->P * 4));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->nccl_comm != nullptr && ctx->nranks > 1) PPO_TRY(ppo_comm_allreduce_f64(ctx, out, 2));
    if (ppoloss) *ppoloss = out[0];
    if (entw) *entw = out[1] * entropy_weight;
    return PPO_OK;
}

int ppo_step_batch_host(ppo_policy* p, ppo_opt* opt, int64_t nb, int nhe, const float* feat, const float* mask,
                        const int64_t* linear_action_index, const float* old_prob, const float* advantage,
                        double epsilon, double entropy_weight, double* ppoloss, double* entropyloss_weighted,
                        float* grads_out) {
    PPO_REQUIRE(p != nullptr && feat && mask && linear_action_index && old_prob && advantage,
                "step_batch: null argument");
    PPO_REQUIRE(opt == nullptr || opt->policy == p, "step_batch: optimiser belongs to another policy");
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    PPO_REQUIRE(nb >= 1 && nhe >= 1, "step_batch: nb=%lld nhe=%d", (long long)nb, nhe);
    const int A = nhe * p->dims[p->L];
    PPO_TRY(upload_host_batch(p, nb, nhe, feat, mask));
    PPO_TRY(ensure_scratch(ctx, 64 + (size_t)nb * 8));
    int* d_bad = (int*)ctx->d_scratch;
    int64_t* d_lin = (int64_t*)((char*)ctx->d_scratch + 64);
    PPO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PPO_TRY(h2d(ctx, d_lin, linear_action_index, (size_t)nb * 8));
    PPO_TRY(launch_linear_index_in(ctx, d_lin, p->hbatch.action, nb, A, d_bad));
    PPO_TRY(h2d(ctx, p->hbatch.old_prob, old_prob, (size_t)nb * 4));
    PPO_TRY(h2d(ctx, p->hbatch.adv, advantage, (size_t)nb * 4));
    PPO_TRY(check_bad_flag(ctx, d_bad, "step_batch: linear action index outside its column"));
    double cnt = (double)nb;
    if (ctx->nccl_comm != nullptr && ctx->nranks > 1) PPO_TRY(ppo_comm_allreduce_f64(ctx, &cnt, 1));
    PPO_TRY(step_core(p, opt, p->hbatch, nb, nhe, epsilon, entropy_weight, 1.0 / cnt, 0));
    return finish_step(p, 0, entropy_weight, ppoloss, entropyloss_weighted, grads_out);
}

int ppo_step_batch(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, int64_t start, int64_t count, double epsilon,
                   double entropy_weight, double* ppoloss, double* entropyloss_weighted, float* grads_out) {
    PPO_REQUIRE(p != nullptr && buf != nullptr, "step_batch: null argument");
    PPO_REQUIRE(opt == nullptr || opt->policy == p, "step_batch: optimiser belongs to another policy");
    PPO_REQUIRE(p->ctx == buf->ctx, "step_batch: policy and buffer live on different contexts");
    PPO_REQUIRE(p->dims[0] == buf->nf && p->dims[p->L] == buf->apa, "step_batch: policy %d->%d vs buffer nf=%d apa=%d",
                p->dims[0], p->dims[p->L], buf->nf, buf->apa);
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(ppo_gather_device(buf, start, count, 0));
    double cnt = (double)count;
    if (ctx->nccl_comm != nullptr && ctx->nranks > 1) PPO_TRY(ppo_comm_allreduce_f64(ctx, &cnt, 1));
    PPO_TRY(step_core(p, opt, buf->batch, count, buf->nhe, epsilon, entropy_weight, 1.0 / cnt, 0));
    return finish_step(p, 0, entropy_weight, ppoloss, entropyloss_weighted, grads_out);
}

int ppo_step_epoch(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, double epsilon, int64_t batch_size,
                   double entropy_weight, double* ppoloss_mean, double* entropyloss_mean) {
    PPO_REQUIRE(p != nullptr && opt != nullptr && buf != nullptr, "step_epoch: null argument");
    PPO_REQUIRE(opt->policy == p, "step_epoch: optimiser belongs to another policy");
    PPO_REQUIRE(p->ctx == buf->ctx, "step_epoch: policy and buffer live on different contexts");
    PPO_REQUIRE(p->dims[0] == buf->nf && p->dims[p->L] == buf->apa, "step_epoch: policy %d->%d vs buffer nf=%d apa=%d",
                p->dims[0], p->dims[p->L], buf->nf, buf->apa);
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(use(ctx));
    const int64_t num_data = buf->n;
    // @assert 1 <= batch_size <= num_data, src/train.jl:88
    PPO_REQUIRE(batch_size >= 1 && batch_size <= num_data, "step_epoch: need 1 <= batch_size (%lld) <= num_data (%lld)",
                (long long)batch_size, (long long)num_data);
    if (buf->perm_len != num_data) {
        set_error("step_epoch: permutation of length %lld set, dataset has %lld", (long long)buf->perm_len,
                  (long long)num_data);
        return PPO_ERR_STATE;
    }
    const int64_t nbatches = ceil_div(num_data, batch_size);
    if (2 * nbatches > ctx->pinned_doubles) {        // the loss history is read back through pinned staging: grow it on demand
        PPO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_pinned) PPO_CUDA(cudaFreeHost(ctx->h_pinned));
        ctx->h_pinned = nullptr;
        ctx->pinned_doubles = 0;
        const int64_t want = round_up(2 * nbatches, 4096);
        PPO_CUDA(cudaMallocHost((void**)&ctx->h_pinned, sizeof(double) * (size_t)want));
        ctx->pinned_doubles = want;
    }
    const bool dp = ctx->nccl_comm != nullptr && ctx->nranks > 1;
    // global row count of every minibatch (local count when single-GPU)
    std::vector<double> counts((size_t)nbatches + 1);
    for (int64_t k = 0; k < nbatches; ++k)
        counts[(size_t)k] = (double)std::min(batch_size, num_data - k * batch_size);
    counts[(size_t)nbatches] = (double)nbatches;
    if (dp) {
        PPO_TRY(ppo_comm_allreduce_f64(ctx, counts.data(), (int)nbatches + 1));
        PPO_REQUIRE(counts[(size_t)nbatches] == (double)nbatches * ctx->nranks,
                    "step_epoch: ranks disagree on the number of minibatches");
    }
    PPO_TRY(ensure_batch(ctx, buf->batch, batch_size, buf->nf * buf->nhe, buf->A));
    PPO_TRY(ensure_loss_buffers(p, batch_size, buf->A, nbatches));
    auto run_batch = [&](int64_t k) -> int {
        const int64_t start = k * batch_size;
        const int64_t count = std::min(batch_size, num_data - start);
        PPO_TRY(gather_into(buf, buf->batch, buf->perm + start, count, 0));
        return step_core(p, opt, buf->batch, count, buf->nhe, epsilon, entropy_weight, 1.0 / counts[(size_t)k], k);
    };
    // All full minibatches of an epoch launch the same ~40 kernels with the same arguments except for the minibatch
    // number, so the loop body is captured ONCE into a CUDA graph that reads the minibatch number from a device
    // counter, and replayed: one graph launch per minibatch instead of ~40 kernel launches + ~20 host-side tensor
    // map encodings.  This is what bounds the small-minibatch regime (the reference's batch_size = 32).  The first
    // minibatch runs un-captured (it sizes every workspace); the ragged last minibatch too.
    const int64_t n_full = num_data / batch_size;
    bool same_global = true;
    for (int64_t k = 1; k < n_full; ++k) same_global = same_global && counts[(size_t)k] == counts[0];
    const char* no_graph = getenv("PPO_B200_NO_GRAPH");
    const bool use_graph = n_full >= 4 && same_global && !(no_graph && no_graph[0] == '1');
    PPO_TRY(run_batch(0));
    int64_t k = 1;
    if (use_graph) {
        const int one = 1;
        PPO_CUDA(cudaMemcpyAsync(ctx->d_step, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int64_t launches_before = ctx->launches;
        PPO_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int st = gather_into(buf, buf->batch, buf->perm, batch_size, 0, ctx->d_step, batch_size);
        if (st == PPO_OK)
            st = step_core(p, opt, buf->batch, batch_size, buf->nhe, epsilon, entropy_weight, 1.0 / counts[0], 0, ctx->d_step,
                           ctx->d_step);
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (st != PPO_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        PPO_CUDA(ce);
        PPO_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        for (; k < n_full; ++k) {
            cudaError_t le = cudaGraphLaunch(exec, ctx->stream);
            if (le != cudaSuccess) {
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                PPO_CUDA(le);
            }
        }
        // kernels executed = captured kernel nodes x replays (the capture itself executed nothing)
        const int64_t nodes = ctx->launches - launches_before;
        ctx->launches = launches_before + nodes * (n_full - 1);
        PPO_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
    }
    for (; k < nbatches; ++k) PPO_TRY(run_batch(k));
    if (dp) PPO_TRY(nccl_allreduce_f64(ctx, p->d_loss_hist, 2 * nbatches));
    if (dp) PPO_TRY(p2p_check(p));
    PPO_TRY(d2h(ctx, ctx->h_pinned, p->d_loss_hist, (size_t)nbatches * 16));
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    // Flux.mean(ppo_loss_history), Flux.mean(entropy_loss_history): unweighted by batch length, :127
    double sp = 0.0, se = 0.0;
    for (int64_t k = 0; k < nbatches; ++k) {
        sp += ctx->h_pinned[2 * k];
        se += ctx->h_pinned[2 * k + 1] * entropy_weight;
    }
    if (ppoloss_mean) *ppoloss_mean = sp / (double)nbatches;
    if (entropyloss_mean) *entropyloss_mean = se / (double)nbatches;
    return PPO_OK;
}

int ppo_train(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, double epsilon, int64_t batch_size, int num_epochs,
              double entropy_weight, uint64_t seed, double* ppo_hist, double* entropy_hist, double* lr_hist) {
    PPO_REQUIRE(p != nullptr && opt != nullptr && buf != nullptr, "train: null argument");
    PPO_REQUIRE(num_epochs >= 0, "train: num_epochs = %d", num_epochs);
    for (int e = 0; e < num_epochs; ++e) {
        PPO_TRY(ppo_permutation_generate(buf, seed + (uint64_t)e + 1, nullptr));
        double a = 0.0, b = 0.0;
        PPO_TRY(ppo_step_epoch(p, opt, buf, epsilon, batch_size, entropy_weight, &a, &b));
        if (ppo_hist) ppo_hist[e] = a;
        if (entropy_hist) entropy_hist[e] = b;
        if (lr_hist) lr_hist[e] = opt->eta;
    }
    return PPO_OK;
}

}  // extern "C"
