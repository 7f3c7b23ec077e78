// bench_hooks.cu — ppo_bench_kernel: time ONE kernel of the hot path in isolation on a synthetic,
// device-resident problem (CUDA events on the ctx stream, optional L2 flush between launches).
// Used by bench.py for the per-kernel roofline numbers; not part of the reference-facing API.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "gemm_f16.cuh"

namespace ppo {
namespace {

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

// kind 0: uniform [-1,1) floats; 1: small integer rewards [-4,4]; 2: mask (0 / -inf, p=.25 per group of 16,
// group 0 of each row of `row` elements unmasked); 3: probabilities in (0.05, 1]
__global__ void fill_f32_kernel(float* p, int64_t n, int kind, int row, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = hash32(i * 0x9E3779B97F4A7C15ull + seed);
        float u = (float)(h >> 8) * (1.0f / 16777216.0f);
        float v;
        if (kind == 0) v = 2.0f * u - 1.0f;
        else if (kind == 1) v = (float)((int)(h % 9u) - 4);
        else if (kind == 2) {
            int64_t col = i % row, grp = col / 16;
            uint32_t hg = hash32((uint64_t)((i / row) * 4096 + grp) * 0x9E3779B97F4A7C15ull + seed);
            v = (grp != 0 && (hg & 3u) == 0u) ? -INFINITY : 0.0f;
        } else v = 0.05f + 0.95f * u;
        p[i] = v;
    }
}
__global__ void fill_terminal_kernel(uint8_t* t, int64_t n, int mean_len, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        t[i] = (hash32(i * 0x9E3779B97F4A7C15ull + seed) % (uint32_t)mean_len) == 0u || i == n - 1;
}
__global__ void fill_zero_i32_kernel(int* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = 0;
}

__global__ void axpy_kernel(float* y, const float* x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i];
}
int launch_axpy(ppo_ctx* ctx, float* y, const float* x, int64_t n) {
    axpy_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 16), 256, 0, ctx->stream>>>(y, x, n);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

struct Scope {  // frees everything on exit
    std::vector<void*> ptrs;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scope() {
        for (void* p : ptrs) cudaFree(p);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    template <typename T>
    int alloc(T** p, size_t count) {
        *p = nullptr;
        PPO_CUDA(cudaMalloc((void**)p, (count ? count : 1) * sizeof(T)));
        ptrs.push_back(*p);
        return PPO_OK;
    }
};

int fill(ppo_ctx* ctx, float* p, int64_t n, int kind, int row, uint64_t seed) {
    fill_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 16), 256, 0, ctx->stream>>>(p, n, kind, row, seed);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

template <typename F>
int time_loop(ppo_ctx* ctx, Scope& sc, int iters, int flush, F&& launch, double* ms_out) {
    PPO_CUDA(cudaEventCreate(&sc.e0));
    PPO_CUDA(cudaEventCreate(&sc.e1));
    const char* wenv = getenv("PPO_BENCH_WARMUP");          // profiling runs set 0 so that ncu sees one launch per kernel
    const int warm = wenv ? atoi(wenv) : 3;
    for (int w = 0; w < warm; ++w) PPO_TRY(launch());
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (flush) PPO_TRY(flush_l2(ctx));
        PPO_CUDA(cudaEventRecord(sc.e0, ctx->stream));
        PPO_TRY(launch());
        PPO_CUDA(cudaEventRecord(sc.e1, ctx->stream));
        PPO_CUDA(cudaEventSynchronize(sc.e1));
        float ms = 0.0f;
        PPO_CUDA(cudaEventElapsedTime(&ms, sc.e0, sc.e1));
        total += ms;
    }
    *ms_out = total / iters;
    return PPO_OK;
}

}  // namespace
}  // namespace ppo

using namespace ppo;

extern "C" int ppo_bench_kernel(ppo_ctx* ctx, const char* which, int64_t n, int a, int b, int c, int iters,
                                int flush_l2_flag, double* ms_out, double* work_out) {
    PPO_REQUIRE(ctx && which && ms_out && work_out, "bench_kernel: null argument");
    PPO_CUDA(cudaSetDevice(ctx->device));
    PPO_REQUIRE(n >= 1 && iters >= 1, "bench_kernel: n=%lld iters=%d", (long long)n, iters);
    Scope sc;
    const std::string w(which);
    if (w == "scan" || w == "scan_norm") {      // scan_norm: with the K2 statistics of the normalisation extension fused in
        float *r, *o; uint8_t* t; double* stats;
        const int64_t n16 = round_up(n, SCAN_TILE);
        PPO_TRY(sc.alloc(&r, (size_t)n16)); PPO_TRY(sc.alloc(&o, (size_t)n16)); PPO_TRY(sc.alloc(&t, (size_t)n16));
        PPO_TRY(sc.alloc(&stats, (size_t)2 * SCAN_STATS_PER_TILE * ceil_div(n, SCAN_TILE)));
        void* scratch; PPO_TRY(sc.alloc((char**)&scratch, scan_scratch_bytes(n)));
        PPO_TRY(fill(ctx, r, n, 1, 1, 11));
        fill_terminal_kernel<<<148 * 8, 256, 0, ctx->stream>>>(t, n, a > 0 ? a : 15, 12);
        const double disc = (b == 0) ? 1.0 : 0.99;
        PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                          [&]() { return launch_returns_scan(ctx, r, o, t, n, disc, 0, w == "scan_norm" ? stats : nullptr, scratch); }, ms_out));
        *work_out = 9.0 * (double)n;
        return PPO_OK;
    }
    if (w == "shuffle") {
        int* perm; PPO_TRY(sc.alloc(&perm, (size_t)n));
        PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, [&]() { return launch_feistel_permutation(ctx, perm, n, 77); }, ms_out));
        *work_out = 4.0 * (double)n;
        return PPO_OK;
    }
    if (w == "gather0" || w == "gather1") {
        // n buffer rows, a = feature floats per row, b = mask floats per row, c = rows gathered per launch
        const int64_t cnt = c;
        PPO_REQUIRE(a >= 1 && b >= 1 && cnt >= 1 && cnt <= n, "bench gather: bad shape");
        float *feat, *mask, *prob, *ret, *fo, *mo, *po, *ao; int *act, *perm, *acto;
        PPO_TRY(sc.alloc(&feat, (size_t)n * a)); PPO_TRY(sc.alloc(&mask, (size_t)n * b));
        PPO_TRY(sc.alloc(&prob, (size_t)n)); PPO_TRY(sc.alloc(&ret, (size_t)n)); PPO_TRY(sc.alloc(&act, (size_t)n));
        PPO_TRY(sc.alloc(&perm, (size_t)n));
        PPO_TRY(sc.alloc(&fo, (size_t)cnt * a)); PPO_TRY(sc.alloc(&mo, (size_t)cnt * b));
        PPO_TRY(sc.alloc(&po, (size_t)cnt)); PPO_TRY(sc.alloc(&ao, (size_t)cnt)); PPO_TRY(sc.alloc(&acto, (size_t)cnt));
        PPO_TRY(fill(ctx, feat, n * a, 0, 1, 1)); PPO_TRY(fill(ctx, mask, n * b, 2, b, 2));
        PPO_TRY(fill(ctx, prob, n, 3, 1, 3)); PPO_TRY(fill(ctx, ret, n, 1, 1, 4));
        fill_zero_i32_kernel<<<148 * 8, 256, 0, ctx->stream>>>(act, n);
        PPO_TRY(launch_feistel_permutation(ctx, perm, n, 5));
        GatherArgs g{};
        g.feat = feat; g.mask = mask; g.action = act; g.old_prob = prob; g.ret = ret; g.count = cnt;
        g.feat_elems = a; g.mask_elems = b; g.feat_out = fo; g.mask_out = mo; g.action_out = acto; g.prob_out = po;
        g.adv_out = ao; g.norm = nullptr;
        int64_t pos = 0;
        const int variant = (w == "gather1") ? 1 : 0;
        auto launch = [&]() -> int {
            g.index = perm + pos;
            pos = (pos + cnt + cnt <= n) ? pos + cnt : 0;   // walk through the epoch's minibatches
            return launch_gather(ctx, g, variant);
        };
        PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, launch, ms_out));
        *work_out = (double)cnt * (4.0 + 2.0 * (4.0 * a + 4.0 * b + 12.0));
        return PPO_OK;
    }
    if (w == "loss") {
        const int A = a; const int64_t nb = n;
        float *lg, *mk, *old, *adv, *dl; int* act; double *part, *out;
        PPO_TRY(sc.alloc(&lg, (size_t)nb * A)); PPO_TRY(sc.alloc(&mk, (size_t)nb * A)); PPO_TRY(sc.alloc(&dl, (size_t)nb * A));
        PPO_TRY(sc.alloc(&old, (size_t)nb)); PPO_TRY(sc.alloc(&adv, (size_t)nb)); PPO_TRY(sc.alloc(&act, (size_t)nb));
        PPO_TRY(sc.alloc(&part, (size_t)2 * loss_num_blocks(nb, A))); PPO_TRY(sc.alloc(&out, 2));
        PPO_TRY(fill(ctx, lg, nb * A, 0, 1, 1)); PPO_TRY(fill(ctx, mk, nb * A, 2, A, 2));
        PPO_TRY(fill(ctx, old, nb, 3, 1, 3)); PPO_TRY(fill(ctx, adv, nb, 1, 1, 4));
        fill_zero_i32_kernel<<<148 * 8, 256, 0, ctx->stream>>>(act, nb);
        auto launch = [&]() -> int {
            return launch_loss(ctx, lg, mk, act, old, adv, nb, A, 0.05, 0.01, 1.0 / (double)nb, dl, part, out, nullptr);
        };
        PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, launch, ms_out));
        *work_out = (double)nb * (12.0 * A + 12.0);
        return PPO_OK;
    }
    if (w == "adam") {
        float *x, *m, *v, *g; double* bp;
        PPO_TRY(sc.alloc(&x, (size_t)n)); PPO_TRY(sc.alloc(&m, (size_t)n)); PPO_TRY(sc.alloc(&v, (size_t)n));
        PPO_TRY(sc.alloc(&g, (size_t)n)); PPO_TRY(sc.alloc(&bp, 2));
        PPO_TRY(fill(ctx, x, n, 0, 1, 1)); PPO_TRY(fill(ctx, g, n, 0, 1, 2));
        PPO_CUDA(cudaMemsetAsync(m, 0, (size_t)n * 4, ctx->stream)); PPO_CUDA(cudaMemsetAsync(v, 0, (size_t)n * 4, ctx->stream));
        double h[2] = {0.9, 0.999};
        PPO_CUDA(cudaMemcpyAsync(bp, h, 16, cudaMemcpyHostToDevice, ctx->stream));
        PPO_CUDA(cudaStreamSynchronize(ctx->stream));
        auto launch = [&]() -> int { return launch_adam(ctx, x, m, v, g, n, 1e-4, 0.9, 0.999, 1e-8, bp, 1.0f); };
        PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, launch, ms_out));
        *work_out = 28.0 * (double)n;
        return PPO_OK;
    }
    if (w == "head_fwd" || w == "head_bwd") {
        const int64_t M = n; const int K = a, N = b;
        float *H, *W, *bias, *lg, *dH, *dW, *db, *part;
        PPO_TRY(sc.alloc(&H, (size_t)M * K)); PPO_TRY(sc.alloc(&W, (size_t)K * N)); PPO_TRY(sc.alloc(&bias, (size_t)N));
        PPO_TRY(sc.alloc(&lg, (size_t)M * N)); PPO_TRY(sc.alloc(&dH, (size_t)M * K)); PPO_TRY(sc.alloc(&dW, (size_t)K * N));
        PPO_TRY(sc.alloc(&db, (size_t)N));
        const size_t pb = wgrad_partial_bytes(M, K, N);
        PPO_TRY(sc.alloc((char**)&part, pb));
        PPO_TRY(fill(ctx, H, M * K, 0, 1, 1)); PPO_TRY(fill(ctx, W, (int64_t)K * N, 0, 1, 2));
        PPO_TRY(fill(ctx, bias, N, 0, 1, 3)); PPO_TRY(fill(ctx, lg, M * N, 0, 1, 4));
        if (w == "head_fwd") {
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, [&]() { return launch_head_fwd(ctx, H, nullptr, W, bias, lg, M, K, N); }, ms_out));
            *work_out = (double)M * (4.0 * K + 4.0 * N);
        } else {
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return launch_head_bwd(ctx, H, nullptr, lg, W, dH, nullptr, dW, db, nullptr, M, K, N, 0.01f, part, pb, true); }, ms_out));
            *work_out = (double)M * (8.0 * K + 4.0 * N);
        }
        return PPO_OK;
    }
    if (w == "gemm_fwd" || w == "gemm_dgrad" || w == "gemm_wgrad") {
        // fp32 SIMT engine; the tensor-core engines are timed through "tc_*" below
        const int64_t M = n; const int K = a, N = b;
        float *X, *W, *bias, *Y, *dX, *dW, *db, *part;
        PPO_TRY(sc.alloc(&X, (size_t)M * K)); PPO_TRY(sc.alloc(&W, (size_t)K * N)); PPO_TRY(sc.alloc(&bias, (size_t)N));
        PPO_TRY(sc.alloc(&Y, (size_t)M * N)); PPO_TRY(sc.alloc(&dX, (size_t)M * K)); PPO_TRY(sc.alloc(&dW, (size_t)K * N));
        PPO_TRY(sc.alloc(&db, (size_t)N));
        const size_t pb = wgrad_partial_bytes(M, K, N);
        PPO_TRY(sc.alloc((char**)&part, pb));
        PPO_TRY(fill(ctx, X, M * K, 0, 1, 1)); PPO_TRY(fill(ctx, W, (int64_t)K * N, 0, 1, 2));
        PPO_TRY(fill(ctx, bias, N, 0, 1, 3)); PPO_TRY(fill(ctx, Y, M * N, 0, 1, 4));
        if (w == "gemm_fwd")
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, [&]() { return launch_linear_fwd_simt(ctx, X, W, bias, Y, M, K, N, true, 0.01f); }, ms_out));
        else if (w == "gemm_dgrad")
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, [&]() { return launch_linear_dgrad_simt(ctx, Y, W, X, dX, M, K, N, 0.01f); }, ms_out));
        else
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag, [&]() { return launch_linear_wgrad_simt(ctx, X, Y, dW, db, M, K, N, part, pb); }, ms_out));
        *work_out = 2.0 * (double)M * K * N;
        return PPO_OK;
    }
    if (w == "tc1_fwd" || w == "tc1_dgrad" || w == "tc1_wgrad") {
        const int64_t M = n; const int K = a, N = b;
        float *X, *Xl, *W, *Wh, *Wl, *WTh, *WTl, *bias, *Y, *Yl, *dX, *dXl, *dW, *db, *part;
        PPO_TRY(sc.alloc(&X, (size_t)M * K + 64)); PPO_TRY(sc.alloc(&Xl, (size_t)M * K + 64));
        PPO_TRY(sc.alloc(&W, (size_t)K * N)); PPO_TRY(sc.alloc(&Wh, (size_t)K * N)); PPO_TRY(sc.alloc(&Wl, (size_t)K * N));
        PPO_TRY(sc.alloc(&WTh, (size_t)K * N)); PPO_TRY(sc.alloc(&WTl, (size_t)K * N)); PPO_TRY(sc.alloc(&bias, (size_t)N));
        PPO_TRY(sc.alloc(&Y, (size_t)M * N)); PPO_TRY(sc.alloc(&Yl, (size_t)M * N));
        PPO_TRY(sc.alloc(&dX, (size_t)M * K)); PPO_TRY(sc.alloc(&dXl, (size_t)M * K)); PPO_TRY(sc.alloc(&dW, (size_t)K * N));
        PPO_TRY(sc.alloc(&db, (size_t)N));
        const size_t pb = tc_test_partial_bytes(ctx, M, K, N);
        PPO_TRY(sc.alloc((char**)&part, pb));
        PPO_TRY(fill(ctx, X, M * K, 0, 1, 1)); PPO_TRY(fill(ctx, W, (int64_t)K * N, 0, 1, 2));
        PPO_TRY(fill(ctx, bias, N, 0, 1, 3)); PPO_TRY(fill(ctx, Y, M * N, 0, 1, 4));
        PPO_TRY(tc_test_split(ctx, X, X, Xl, M * K)); PPO_TRY(tc_test_split(ctx, Y, Y, Yl, M * N));
        PPO_TRY(tc_test_weight_prep(ctx, W, Wh, Wl, WTh, WTl, K, N));
        if (w == "tc1_fwd")
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return tc_test_fwd(ctx, X, Xl, WTh, WTl, bias, Y, Yl, M, K, N, 1, 0.01f); }, ms_out));
        else if (w == "tc1_dgrad")
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return tc_test_dgrad(ctx, Y, Yl, Wh, Wl, X, dX, dXl, M, K, N, 0.01f, part, db); }, ms_out));
        else
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return tc_test_wgrad(ctx, X, Xl, Y, Yl, dW, db, part, pb, M, K, N); }, ms_out));
        *work_out = 2.0 * (double)M * K * N;
        return PPO_OK;
    }
    if (w == "tc3_fwd" || w == "tc3_dgrad" || w == "tc3_wgrad" || w == "head16_fwd" || w == "head16_bwd") {
        // fp16-split engine: operands as scaled fp16 hi/lo pairs (uniform [-1, 1) data; output bound K + 1)
        const int64_t M = n; const int K = a, N = b;
        float *X, *W, *bias, *Y, *dW, *db, *part, *sc_, *lg;
        __half *Xh, *Xl, *Wh, *Wl, *WTh, *WTl, *Yh, *Yl, *dXh, *dXl;
        unsigned *st, *sgn;
        PPO_TRY(sc.alloc(&sgn, std::max(f16_test_sign_words(M, (K + 31) / 32 * 32), f16_test_sign_words(M, (N + 31) / 32 * 32)) + 64));
        PPO_TRY(sc.alloc(&X, (size_t)std::max<int64_t>(M * K, M * N))); PPO_TRY(sc.alloc(&W, (size_t)K * N));
        PPO_TRY(sc.alloc(&bias, (size_t)std::max(K, N))); PPO_TRY(sc.alloc(&Y, (size_t)M * N));
        PPO_TRY(sc.alloc(&Xh, (size_t)M * K + 128)); PPO_TRY(sc.alloc(&Xl, (size_t)M * K + 128));
        PPO_TRY(sc.alloc(&Wh, (size_t)K * N)); PPO_TRY(sc.alloc(&Wl, (size_t)K * N));
        PPO_TRY(sc.alloc(&WTh, (size_t)K * N)); PPO_TRY(sc.alloc(&WTl, (size_t)K * N));
        // output pairs as the engine allocates them: hi and lo planes of one allocation, the lo plane a whole number of
        // rows behind (one 3-D TMA store box covers both)
        {
            const size_t py = (size_t)round_up((int64_t)((size_t)M * N + 128) * 2, (int64_t)N * 2 * 128) / 2;
            const size_t px = (size_t)round_up((int64_t)((size_t)M * K + 128) * 2, (int64_t)K * 2 * 128) / 2;
            PPO_TRY(sc.alloc(&Yh, 2 * py)); Yl = Yh + py;
            PPO_TRY(sc.alloc(&dXh, 2 * px)); dXl = dXh + px;
        }
        PPO_TRY(sc.alloc(&dW, (size_t)K * N)); PPO_TRY(sc.alloc(&db, (size_t)std::max(K, N)));
        PPO_TRY(sc.alloc(&sc_, 16)); PPO_TRY(sc.alloc(&st, 4)); PPO_TRY(sc.alloc(&lg, (size_t)M * 4));
        const size_t pb = std::max(f16_test_partial_bytes(ctx, M, K, N), f16_test_head_partial_bytes(M, K, N <= 4 ? N : 4));
        PPO_TRY(sc.alloc((char**)&part, pb));
        PPO_CUDA(cudaMemsetAsync(st, 0, 16, ctx->stream));
        PPO_TRY(fill(ctx, X, M * K, 0, 1, 1)); PPO_TRY(fill(ctx, W, (int64_t)K * N, 0, 1, 2));
        PPO_TRY(fill(ctx, bias, std::max(K, N), 0, 1, 3)); PPO_TRY(fill(ctx, Y, M * N, 0, 1, 4));
        float *sX = sc_, *sW = sc_ + 2, *sY = sc_ + 4, *sO = sc_ + 6;
        PPO_TRY(f16_test_operand(ctx, X, Xh, Xl, M * K, sX, st));
        PPO_TRY(f16_test_operand(ctx, Y, Yh, Yl, M * N, sY, st));
        PPO_REQUIRE(K % 32 == 0 || w != "tc3_dgrad", "bench f16: K %% 32 for dgrad");
        if (K % 32 == 0) PPO_TRY(f16_test_signbits(ctx, X, M, K, sgn));
        if (w == "head16_fwd" || w == "head16_bwd") {
            PPO_REQUIRE(N <= 4, "bench head16: N <= 4");
            PPO_TRY(fill(ctx, lg, M * N, 0, 1, 5));
            PPO_TRY(f16_test_set_scale(ctx, sO, (float)N + 1.0f));
            if (w == "head16_fwd") {
                PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                                  [&]() { return f16_test_head_fwd(ctx, Xh, Xl, W, bias, lg, M, K, N, sX); }, ms_out));
                *work_out = (double)M * (4.0 * K + 4.0 * N);
            } else {
                PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                                  [&]() { return f16_test_head_bwd(ctx, Xh, Xl, lg, W, dXh, dXl, dW, db, bias, M, K, N, 0.01f, part, pb, sX, sO); }, ms_out));
                *work_out = (double)M * (8.0 * K + 4.0 * N);
            }
            return PPO_OK;
        }
        PPO_TRY(f16_test_weight(ctx, W, Wh, Wl, WTh, WTl, K, N, sW, st));
        if (w == "tc3_fwd") {
            PPO_TRY(f16_test_set_scale(ctx, sO, (float)K + 1.0f));
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return f16_test_fwd(ctx, Xh, Xl, WTh, WTl, bias, Yh, Yl, sgn, M, K, N, 1, 0.01f, sX, sW, sO); }, ms_out));
        } else if (w == "tc3_dgrad") {
            PPO_TRY(f16_test_set_scale(ctx, sO, (float)N));
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return f16_test_dgrad(ctx, Yh, Yl, Wh, Wl, sgn, dXh, dXl, M, K, N, 0.01f, part, db, sY, sW, sO); }, ms_out));
        } else {
            PPO_TRY(time_loop(ctx, sc, iters, flush_l2_flag,
                              [&]() { return f16_test_wgrad(ctx, Xh, Xl, Yh, Yl, dW, part, pb, M, K, N, sX, sY); }, ms_out));
        }
        *work_out = 2.0 * (double)M * K * N;
        return PPO_OK;
    }
    set_error("bench_kernel: unknown kernel '%s'", which);
    return PPO_ERR_INVALID;
}

extern "C" int ppo_dense_op(ppo_ctx* ctx, int mode, int op, int64_t M, int K, int N, const float* X, const float* W,
                            const float* bias, const float* dY, float slope, float* out, float* out2) {
    PPO_REQUIRE(ctx && X && W && out, "dense_op: null argument");
    PPO_REQUIRE(op >= 0 && op <= 2 && M >= 1 && K >= 1 && N >= 1, "dense_op: bad op/shape");
    PPO_REQUIRE(op == 0 || dY != nullptr, "dense_op: dY required");
    PPO_REQUIRE(mode == PPO_GEMM_FP32_SIMT || mode == PPO_GEMM_TF32X3_TC || mode == PPO_GEMM_F16X3_TC,
                "dense_op: unsupported engine %d", mode);
    PPO_CUDA(cudaSetDevice(ctx->device));
    Scope sc;
    float *dXp, *dWp, *dB, *dDY = nullptr, *dOut, *dOut2;
    PPO_TRY(sc.alloc(&dXp, (size_t)M * K)); PPO_TRY(sc.alloc(&dWp, (size_t)K * N)); PPO_TRY(sc.alloc(&dB, (size_t)N));
    PPO_TRY(sc.alloc(&dOut2, (size_t)N));
    const size_t out_elems = op == 0 ? (size_t)M * N : (op == 1 ? (size_t)M * K : (size_t)K * N);
    PPO_TRY(sc.alloc(&dOut, out_elems));
    cudaStream_t s = ctx->stream;
    PPO_CUDA(cudaMemcpyAsync(dXp, X, (size_t)M * K * 4, cudaMemcpyHostToDevice, s));
    PPO_CUDA(cudaMemcpyAsync(dWp, W, (size_t)K * N * 4, cudaMemcpyHostToDevice, s));
    if (bias) PPO_CUDA(cudaMemcpyAsync(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice, s));
    else PPO_CUDA(cudaMemsetAsync(dB, 0, (size_t)N * 4, s));
    if (dY) {
        PPO_TRY(sc.alloc(&dDY, (size_t)M * N));
        PPO_CUDA(cudaMemcpyAsync(dDY, dY, (size_t)M * N * 4, cudaMemcpyHostToDevice, s));
    }
    const bool act = slope >= 0.0f;
    if (mode == PPO_GEMM_FP32_SIMT) {
        float* part; const size_t pb = wgrad_partial_bytes(M, K, N);
        PPO_TRY(sc.alloc((char**)&part, pb));
        if (op == 0) PPO_TRY(launch_linear_fwd_simt(ctx, dXp, dWp, dB, dOut, M, K, N, act, slope));
        else if (op == 1) PPO_TRY(launch_linear_dgrad_simt(ctx, dDY, dWp, dXp, dOut, M, K, N, slope));
        else PPO_TRY(launch_linear_wgrad_simt(ctx, dXp, dDY, dOut, dOut2, M, K, N, part, pb));
    } else if (mode == PPO_GEMM_F16X3_TC) {
        // fp16-split engine: operands as scaled fp16 hi/lo pairs; output scale from the same guaranteed bound the
        // policy path uses (max|X| * max column abs-sum of W + max|b|, resp. max|dY| * max row abs-sum), host-evaluated
        __half *Xh, *Xl, *Wh, *Wl, *WTh, *WTl, *DYh = nullptr, *DYl = nullptr, *Oh, *Ol;
        float *scs, *part; unsigned* st;
        PPO_TRY(sc.alloc(&Xh, (size_t)M * K + 128)); PPO_TRY(sc.alloc(&Xl, (size_t)M * K + 128));
        PPO_TRY(sc.alloc(&Wh, (size_t)K * N)); PPO_TRY(sc.alloc(&Wl, (size_t)K * N));
        PPO_TRY(sc.alloc(&WTh, (size_t)K * N)); PPO_TRY(sc.alloc(&WTl, (size_t)K * N));
        PPO_TRY(sc.alloc(&Oh, out_elems + 128)); PPO_TRY(sc.alloc(&Ol, out_elems + 128));
        PPO_TRY(sc.alloc(&scs, 16)); PPO_TRY(sc.alloc(&st, 4));
        PPO_CUDA(cudaMemsetAsync(st, 0, 16, s));
        float *sX = scs, *sW = scs + 2, *sDY = scs + 4, *sO = scs + 6;
        PPO_TRY(f16_test_operand(ctx, dXp, Xh, Xl, M * K, sX, st));
        PPO_TRY(f16_test_weight(ctx, dWp, Wh, Wl, WTh, WTl, K, N, sW, st));
        if (dDY) {
            PPO_TRY(sc.alloc(&DYh, (size_t)M * N + 128)); PPO_TRY(sc.alloc(&DYl, (size_t)M * N + 128));
            PPO_TRY(f16_test_operand(ctx, dDY, DYh, DYl, M * N, sDY, st));
        }
        double xmax = 0.0, dymax = 0.0, bmax = 0.0, colmax = 0.0, rowmax = 0.0;
        for (int64_t i = 0; i < M * K; ++i) xmax = std::max(xmax, (double)fabsf(X[i]));
        if (dY) for (int64_t i = 0; i < M * N; ++i) dymax = std::max(dymax, (double)fabsf(dY[i]));
        if (bias) for (int n = 0; n < N; ++n) bmax = std::max(bmax, (double)fabsf(bias[n]));
        {
            std::vector<double> col(N, 0.0);
            for (int k = 0; k < K; ++k) {
                double r = 0.0;
                for (int n = 0; n < N; ++n) { const double a = fabsf(W[(size_t)k * N + n]); r += a; col[n] += a; }
                rowmax = std::max(rowmax, r);
            }
            for (int n = 0; n < N; ++n) colmax = std::max(colmax, col[n]);
        }
        if (op == 0) {
            PPO_TRY(f16_test_set_scale(ctx, sO, (float)((xmax * colmax + bmax) * 1.001)));
            PPO_TRY(f16_test_fwd(ctx, Xh, Xl, WTh, WTl, dB, Oh, Ol, nullptr, M, K, N, act ? 1 : 0, slope, sX, sW, sO));
            PPO_TRY(f16_test_join(ctx, Oh, Ol, (int64_t)out_elems, sO, dOut));
        } else if (op == 1) {
            float *cs_scratch, *cs_out;
            PPO_TRY(sc.alloc((char**)&cs_scratch, f16_test_partial_bytes(ctx, M, K, N)));
            PPO_TRY(sc.alloc(&cs_out, (size_t)K));
            PPO_TRY(f16_test_set_scale(ctx, sO, (float)(dymax * rowmax * 1.001)));
            unsigned* gate;
            PPO_TRY(sc.alloc(&gate, f16_test_sign_words(M, K) + 64));
            PPO_TRY(f16_test_signbits(ctx, dXp, M, K, gate));
            PPO_TRY(f16_test_dgrad(ctx, DYh, DYl, Wh, Wl, gate, Oh, Ol, M, K, N, slope, cs_scratch, cs_out, sDY, sW, sO));
            PPO_TRY(f16_test_join(ctx, Oh, Ol, (int64_t)out_elems, sO, dOut));
            if (out2) PPO_CUDA(cudaMemcpyAsync(out2, cs_out, (size_t)K * 4, cudaMemcpyDeviceToHost, s));
        } else {
            const size_t pb = f16_test_partial_bytes(ctx, M, K, N);
            PPO_TRY(sc.alloc((char**)&part, pb));
            PPO_TRY(f16_test_wgrad(ctx, Xh, Xl, DYh, DYl, dOut, part, pb, M, K, N, sX, sDY));
            PPO_CUDA(cudaMemsetAsync(dOut2, 0, (size_t)N * 4, s));     // the bias gradient comes from the kernel that produced dY
        }
    } else {
        // tensor-core engine: operands as tf32 hi/lo pairs; the result is hi + lo of the output pair
        float *Xh, *Xl, *Wh, *Wl, *WTh, *WTl, *DYh = nullptr, *DYl = nullptr, *Ol, *part;
        PPO_TRY(sc.alloc(&Xh, (size_t)M * K + 64)); PPO_TRY(sc.alloc(&Xl, (size_t)M * K + 64));
        PPO_TRY(sc.alloc(&Wh, (size_t)K * N)); PPO_TRY(sc.alloc(&Wl, (size_t)K * N));
        PPO_TRY(sc.alloc(&WTh, (size_t)K * N)); PPO_TRY(sc.alloc(&WTl, (size_t)K * N)); PPO_TRY(sc.alloc(&Ol, out_elems));
        PPO_TRY(tc_test_split(ctx, dXp, Xh, Xl, M * K));
        PPO_TRY(tc_test_weight_prep(ctx, dWp, Wh, Wl, WTh, WTl, K, N));
        if (dDY) {
            PPO_TRY(sc.alloc(&DYh, (size_t)M * N)); PPO_TRY(sc.alloc(&DYl, (size_t)M * N));
            PPO_TRY(tc_test_split(ctx, dDY, DYh, DYl, M * N));
        }
        if (op == 0) {
            PPO_TRY(tc_test_fwd(ctx, Xh, Xl, WTh, WTl, dB, dOut, Ol, M, K, N, act ? 1 : 0, slope));
            PPO_TRY(launch_axpy(ctx, dOut, Ol, (int64_t)out_elems));
        } else if (op == 1) {
            float *cs_scratch, *cs_out;
            PPO_TRY(sc.alloc((char**)&cs_scratch, tc_test_partial_bytes(ctx, M, K, N)));
            PPO_TRY(sc.alloc(&cs_out, (size_t)K));
            PPO_TRY(tc_test_dgrad(ctx, DYh, DYl, Wh, Wl, dXp, dOut, Ol, M, K, N, slope, cs_scratch, cs_out));
            PPO_TRY(launch_axpy(ctx, dOut, Ol, (int64_t)out_elems));
            if (out2) PPO_CUDA(cudaMemcpyAsync(out2, cs_out, (size_t)K * 4, cudaMemcpyDeviceToHost, s));
        } else {
            const size_t pb = tc_test_partial_bytes(ctx, M, K, N);
            PPO_TRY(sc.alloc((char**)&part, pb));
            PPO_TRY(tc_test_wgrad(ctx, Xh, Xl, DYh, DYl, dOut, dOut2, part, pb, M, K, N));
        }
    }
    PPO_CUDA(cudaMemcpyAsync(out, dOut, out_elems * 4, cudaMemcpyDeviceToHost, s));
    if (op == 2 && out2) PPO_CUDA(cudaMemcpyAsync(out2, dOut2, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    PPO_CUDA(cudaStreamSynchronize(s));
    return PPO_OK;
}
