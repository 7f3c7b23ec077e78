// gemm_simt.cu — K5/K7 on fp32 FFMA pipes (PPO_GEMM_FP32_SIMT) + the skinny policy head.
//
// The policy MLP (reference test/policy.jl:9-31: Chain(Dense(nf,H,leakyrelu), ..., Dense(H,apa)))
// is applied per token, so one minibatch is a GEMM with M = nb*nhe rows:
//     fwd    Y[M,N]  = act(X[M,K] W[K,N] + b)            Dense forward (W = bytes of Julia's [out,in])
//     dgrad  dX[M,K] = (dY[M,N] W^T) .* leakyrelu'(X)    Zygote pullback of Dense input + activation
//     wgrad  dW[K,N] = X^T dY,  db[N] = colsum(dY)        Zygote pullback of Dense weight/bias
// This file is the true-fp32 engine: classic 128x128x8 register-blocked tiles, deterministic
// (fixed summation order, split-K partials folded in a fixed order, no float atomics).  It is
// the numerical anchor the tcgen05 engines (gemm_tc.cu) are tested against, and the path used
// for shapes the tensor-core kernels do not cover.  The head (H -> apa, N <= 8) is HBM-bound and
// has its own streaming kernels: head_fwd (one warp per token) and head_bwd, which produces dH,
// dW and db in a single pass over the activations.
#include "common.cuh"

namespace ppo {

namespace {

constexpr int BM = 128, BN = 128, BK = 8, GT = 256;

enum { EPI_BIAS_ACT = 0, EPI_DGRAD = 1, EPI_PARTIAL = 2 };

struct GemmArgs {
    const float* A; const float* B; float* C;
    int64_t M, N, K;          // C is M x N, contraction length K
    int64_t lda, ldb, ldc;
    const float* bias;        // EPI_BIAS_ACT
    const float* Hprev;       // EPI_DGRAD: activation whose derivative gates dX (same shape as C)
    float slope; int act;
    int64_t k_per_split;      // EPI_PARTIAL: contraction range per blockIdx.z
    float* colsum;            // EPI_PARTIAL: per-split column sums of B (bias gradient), or nullptr
};

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(GT)
sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    int64_t kbeg = 0, kend = g.K;
    if (EPI == EPI_PARTIAL) {
        kbeg = (int64_t)blockIdx.z * g.k_per_split;
        kend = kbeg + g.k_per_split < g.K ? kbeg + g.k_per_split : g.K;
    }
    const int ty = tid / 16, tx = tid % 16;

    float ra[4], rb[4];
    auto load_a = [&](int64_t k0) {
        if (A_KC) {
            const int row = tid >> 1, kq = (tid & 1) * 4;
            const int64_t m = m0 + row;
            const float* p = g.A + m * g.lda + k0 + kq;
            if (m < g.M && k0 + kq + 3 < kend && ((g.lda & 3) == 0) && (((uintptr_t)p & 15) == 0)) {
                float4 v = *reinterpret_cast<const float4*>(p);
                ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) ra[i] = (m < g.M && k0 + kq + i < kend) ? p[i] : 0.0f;
            }
        } else {
            const int kk = tid >> 5, mq = (tid & 31) * 4;
            const int64_t k = k0 + kk, m = m0 + mq;
            const float* p = g.A + k * g.lda + m;
            if (k < kend && m + 3 < g.M && ((g.lda & 3) == 0) && (((uintptr_t)p & 15) == 0)) {
                float4 v = *reinterpret_cast<const float4*>(p);
                ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) ra[i] = (k < kend && m + i < g.M) ? p[i] : 0.0f;
            }
        }
    };
    auto load_b = [&](int64_t k0) {
        if (B_KC) {
            const int row = tid >> 1, kq = (tid & 1) * 4;
            const int64_t n = n0 + row;
            const float* p = g.B + n * g.ldb + k0 + kq;
            if (n < g.N && k0 + kq + 3 < kend && ((g.ldb & 3) == 0) && (((uintptr_t)p & 15) == 0)) {
                float4 v = *reinterpret_cast<const float4*>(p);
                rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) rb[i] = (n < g.N && k0 + kq + i < kend) ? p[i] : 0.0f;
            }
        } else {
            const int kk = tid >> 5, nq = (tid & 31) * 4;
            const int64_t k = k0 + kk, n = n0 + nq;
            const float* p = g.B + k * g.ldb + n;
            if (k < kend && n + 3 < g.N && ((g.ldb & 3) == 0) && (((uintptr_t)p & 15) == 0)) {
                float4 v = *reinterpret_cast<const float4*>(p);
                rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) rb[i] = (k < kend && n + i < g.N) ? p[i] : 0.0f;
            }
        }
    };
    auto store_a = [&](int buf) {
        if (A_KC) {
            const int row = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) As[buf][kq + i][row] = ra[i];
        } else {
            const int kk = tid >> 5, mq = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&As[buf][kk][mq]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        }
    };
    auto store_b = [&](int buf) {
        if (B_KC) {
            const int row = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[buf][kq + i][row] = rb[i];
        } else {
            const int kk = tid >> 5, nq = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    float bsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = 0.0f;
    const bool do_colsum = (EPI == EPI_PARTIAL) && g.colsum != nullptr && blockIdx.x == 0;

    int buf = 0;
    if (kbeg < kend) {
        load_a(kbeg); load_b(kbeg);
        store_a(0); store_b(0);
    }
    __syncthreads();
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (do_colsum) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[j] += bv[j];
            }
        }
        if (more) { store_a(buf ^ 1); store_b(buf ^ 1); }
        __syncthreads();
        buf ^= 1;
    }

    float* C = g.C;
    if (EPI == EPI_PARTIAL) C += (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int64_t n = n0 + tx * 4 + jh * 64;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][jh * 4 + j];
            if (EPI == EPI_BIAS_ACT) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < g.N) {
                        float x = v[j] + (g.bias ? g.bias[n + j] : 0.0f);
                        v[j] = (g.act && !(x > 0.0f)) ? g.slope * x : x;
                    }
                }
            } else if (EPI == EPI_DGRAD) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < g.N) {
                        const float h = g.Hprev[m * g.ldc + n + j];
                        v[j] = (h > 0.0f) ? v[j] : g.slope * v[j];
                    }
                }
            }
            float* cp = C + m * g.ldc + n;
            if (n + 3 < g.N && ((g.ldc & 3) == 0) && (((uintptr_t)cp & 15) == 0)) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.N) cp[j] = v[j];
            }
        }
    }
    if (do_colsum && ty == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t n = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
            if (n < g.N) g.colsum[(int64_t)blockIdx.z * g.N + n] = bsum[j];
        }
    }
}

// out[i] = sum_z partial[z*stride + i], fixed order
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int64_t splits, int64_t stride, int64_t count,
                       float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int64_t z = 0; z < splits; ++z) s += partial[z * stride + i];
        out[i] = s;
    }
}

// ---- policy head: N = apa <= 8 -----------------------------------------------------------
constexpr int HEAD_NMAX = 8;

// logits[m][n] = sum_k (H[m][k] (+ Hlo[m][k])) W[k][n] + b[n]: one warp per token row, two rows in flight.
// W is staged in shared memory as [c][n][k/4] so that the 32 lanes (consecutive k/4) read consecutive words.
template <int N, bool HAS_LO>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ H, const float* __restrict__ Hlo, const float* __restrict__ W,
                const float* __restrict__ bias, float* __restrict__ logits, int64_t M, int K) {
    extern __shared__ __align__(16) float sW[];   // [4][N][K/4]
    const int KV = K >> 2;
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        sW[((k & 3) * N + n) * KV + (k >> 2)] = W[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t m = w0; m < M; m += 2 * warps) {
        const int64_t m2 = m + warps;
        const bool two = m2 < M;
        const float4* hp0 = reinterpret_cast<const float4*>(H + m * K);
        const float4* hp1 = reinterpret_cast<const float4*>(H + (two ? m2 : m) * K);
        const float4* lp0 = HAS_LO ? reinterpret_cast<const float4*>(Hlo + m * K) : nullptr;
        const float4* lp1 = HAS_LO ? reinterpret_cast<const float4*>(Hlo + (two ? m2 : m) * K) : nullptr;
        float a0[N], a1[N];
#pragma unroll
        for (int n = 0; n < N; ++n) { a0[n] = 0.0f; a1[n] = 0.0f; }
        for (int kv = lane; kv < KV; kv += 32) {
            float4 h0 = __ldcs(hp0 + kv), h1 = __ldcs(hp1 + kv);
            if (HAS_LO) {   // exact activation = hi + lo
                const float4 l0 = __ldcs(lp0 + kv), l1 = __ldcs(lp1 + kv);
                h0.x += l0.x; h0.y += l0.y; h0.z += l0.z; h0.w += l0.w;
                h1.x += l1.x; h1.y += l1.y; h1.z += l1.z; h1.w += l1.w;
            }
            const float v0[4] = {h0.x, h0.y, h0.z, h0.w};
            const float v1[4] = {h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    const float w = sW[(c * N + n) * KV + kv];
                    a0[n] = fmaf(v0[c], w, a0[n]);
                    a1[n] = fmaf(v1[c], w, a1[n]);
                }
        }
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                a0[n] += __shfl_xor_sync(0xffffffffu, a0[n], d);
                a1[n] += __shfl_xor_sync(0xffffffffu, a1[n], d);
            }
        if (lane == 0) {
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float bn = bias ? bias[n] : 0.0f;
                logits[m * N + n] = a0[n] + bn;
                if (two) logits[m2 * N + n] = a1[n] + bn;
            }
        }
    }
}

__device__ __forceinline__ void tf32_split_dev(float a, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - hi));
    lo = __uint_as_float(l);
}

// One pass over H (= hi (+ lo)): dH = (dlogits W^T) .* leakyrelu'(H), written either exact or as a tf32 hi/lo
// pair (tensor-core mode); per-CTA partials of dW[k][n], db[n] and of the column sums of dH (= the bias
// gradient of the layer below).  Partial layout per CTA: [K*N dW][N db][K colsum(dH)].
template <int N, int KPT, bool HAS_LO>   // KPT = ceil(K / 256) columns per thread
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ H, const float* __restrict__ Hlo, const float* __restrict__ dlogits,
                const float* __restrict__ W, float* __restrict__ dH, float* __restrict__ dHlo,
                float* __restrict__ partial, int64_t M, int K, float slope, int64_t rows_per_cta, int need_dH) {
    const int tid = threadIdx.x;
    float w[KPT][N], aw[KPT][N], ab[N], cs[KPT];
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
        cs[q] = 0.0f;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const int k = tid + q * 256;
            w[q][n] = (k < K) ? W[k * N + n] : 0.0f;
            aw[q][n] = 0.0f;
        }
    }
#pragma unroll
    for (int n = 0; n < N; ++n) ab[n] = 0.0f;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    constexpr int RU = 4;
    for (int64_t m = r0; m < r1; m += RU) {
        float h[RU][KPT], d[RU][N];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const bool rv = m + u < r1;
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const int k = tid + q * 256;
                float x = 0.0f;
                if (rv && k < K) {
                    x = __ldcs(H + (m + u) * K + k);
                    if (HAS_LO) x += __ldcs(Hlo + (m + u) * K + k);
                }
                h[u][q] = x;
            }
#pragma unroll
            for (int n = 0; n < N; ++n) d[u][n] = rv ? __ldg(dlogits + (m + u) * N + n) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const bool rv = m + u < r1;
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const int k = tid + q * 256;
                float dx = 0.0f;
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    dx = fmaf(d[u][n], w[q][n], dx);
                    aw[q][n] = fmaf(h[u][q], d[u][n], aw[q][n]);
                }
                if (need_dH && rv && k < K) {
                    const float g = (h[u][q] > 0.0f) ? dx : slope * dx;
                    cs[q] += g;
                    if (dHlo != nullptr) {
                        float hi, lo;
                        tf32_split_dev(g, hi, lo);
                        dH[(m + u) * K + k] = hi;
                        dHlo[(m + u) * K + k] = lo;
                    } else {
                        dH[(m + u) * K + k] = g;
                    }
                }
            }
            if (tid == 0) {
#pragma unroll
                for (int n = 0; n < N; ++n) ab[n] += d[u][n];
            }
        }
    }
    float* pw = partial + (int64_t)blockIdx.x * ((int64_t)K * N + N + K);
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
        const int k = tid + q * 256;
        if (k < K) {
#pragma unroll
            for (int n = 0; n < N; ++n) pw[k * N + n] = aw[q][n];
            pw[(int64_t)K * N + N + k] = cs[q];
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) pw[(int64_t)K * N + n] = ab[n];
    }
}

inline int64_t wgrad_splits(int64_t M, int K, int N) {
    int64_t tiles = ceil_div(K, BM) * ceil_div(N, BN);
    int64_t s = ceil_div(148 * 4, tiles);
    int64_t maxs = ceil_div(M, 512);
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    return s;
}

inline int64_t head_ctas(int64_t M) {
    int64_t c = 148 * 4;
    int64_t maxc = ceil_div(M, 64);
    return c < maxc ? c : (maxc < 1 ? 1 : maxc);
}

}  // namespace

size_t wgrad_partial_bytes(int64_t M, int K, int N) {
    size_t a = (size_t)wgrad_splits(M, K, N) * ((size_t)K * N + N) * sizeof(float);
    size_t b = (size_t)head_ctas(M) * ((size_t)K * N + N + K) * sizeof(float);
    return a > b ? a : b;
}

int launch_linear_fwd_simt(ppo_ctx* ctx, const float* X, const float* W, const float* bias, float* Y, int64_t M,
                           int K, int N, bool act, float slope) {
    GemmArgs g{};
    g.A = X; g.B = W; g.C = Y; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = N; g.ldc = N;
    g.bias = bias; g.slope = slope; g.act = act ? 1 : 0;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), 1);
    sgemm_kernel<true, false, EPI_BIAS_ACT><<<grid, GT, 0, ctx->stream>>>(g);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_linear_dgrad_simt(ppo_ctx* ctx, const float* dY, const float* W, const float* Hprev, float* dX,
                             int64_t M, int K, int N, float slope) {
    // dX[M,K] = dY[M,N] * W[K,N]^T, gated by leakyrelu'(Hprev[M,K])
    GemmArgs g{};
    g.A = dY; g.B = W; g.C = dX; g.M = M; g.N = K; g.K = N; g.lda = N; g.ldb = N; g.ldc = K;
    g.Hprev = Hprev; g.slope = slope;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(K, BN), 1);
    sgemm_kernel<true, true, EPI_DGRAD><<<grid, GT, 0, ctx->stream>>>(g);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_linear_wgrad_simt(ppo_ctx* ctx, const float* X, const float* dY, float* dW, float* db, int64_t M, int K,
                             int N, float* partial, size_t partial_bytes) {
    // dW[K,N] = X[M,K]^T dY[M,N]; db[N] = colsum(dY)
    const int64_t splits = wgrad_splits(M, K, N);
    const size_t need = (size_t)splits * ((size_t)K * N + N) * sizeof(float);
    PPO_REQUIRE(need <= partial_bytes, "wgrad: partial buffer too small (%zu > %zu)", need, partial_bytes);
    GemmArgs g{};
    g.A = X; g.B = dY; g.C = partial; g.M = K; g.N = N; g.K = M; g.lda = K; g.ldb = N; g.ldc = N;
    g.k_per_split = round_up(ceil_div(M, splits), BK);
    g.colsum = partial + (size_t)splits * K * N;
    dim3 grid((unsigned)ceil_div(K, BM), (unsigned)ceil_div(N, BN), (unsigned)splits);
    sgemm_kernel<false, false, EPI_PARTIAL><<<grid, GT, 0, ctx->stream>>>(g);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    const int64_t cnt = (int64_t)K * N;
    reduce_partials_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, ctx->stream>>>(partial, splits, cnt, cnt, dW);
    reduce_partials_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, ctx->stream>>>(g.colsum, splits, N, N, db);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_head_fwd(ppo_ctx* ctx, const float* H, const float* Hlo, const float* W, const float* bias, float* logits,
                    int64_t M, int K, int N) {
    if (N > HEAD_NMAX || (K & 3) != 0 || (size_t)K * N * 4 > 48 * 1024 || ((uintptr_t)H & 15) != 0) {
        PPO_REQUIRE(Hlo == nullptr, "head_fwd: hi/lo activations need N <= %d and K %% 4 == 0", HEAD_NMAX);
        return launch_linear_fwd_simt(ctx, H, W, bias, logits, M, K, N, false, 0.0f);
    }
    int64_t blocks = ceil_div(M, 8);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)K * N * sizeof(float);
#define PPO_HEAD_FWD(N_)                                                                                          \
    case N_:                                                                                                      \
        if (Hlo) head_fwd_kernel<N_, true><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H, Hlo, W, bias, logits, M, K); \
        else head_fwd_kernel<N_, false><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H, Hlo, W, bias, logits, M, K);    \
        break
    switch (N) {
        PPO_HEAD_FWD(1); PPO_HEAD_FWD(2); PPO_HEAD_FWD(3); PPO_HEAD_FWD(4);
        PPO_HEAD_FWD(5); PPO_HEAD_FWD(6); PPO_HEAD_FWD(7); PPO_HEAD_FWD(8);
    }
#undef PPO_HEAD_FWD
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_head_bwd(ppo_ctx* ctx, const float* H, const float* Hlo, const float* dlogits, const float* W, float* dH,
                    float* dHlo, float* dW, float* db, float* db_below, int64_t M, int K, int N, float slope,
                    float* partial, size_t partial_bytes, bool need_dH) {
    if (N > 4 || K > 1024) {
        // generic path through the tile GEMMs (fp32 engine only)
        PPO_REQUIRE(Hlo == nullptr && dHlo == nullptr && db_below == nullptr,
                    "head_bwd: hi/lo activations need N <= 4 and K <= 1024");
        if (need_dH) PPO_TRY(launch_linear_dgrad_simt(ctx, dlogits, W, H, dH, M, K, N, slope));
        return launch_linear_wgrad_simt(ctx, H, dlogits, dW, db, M, K, N, partial, partial_bytes);
    }
    const int64_t ctas = head_ctas(M);
    const int64_t rows = ceil_div(M, ctas);
    const int64_t stride = (int64_t)K * N + N + K;
    const size_t need = (size_t)ctas * stride * sizeof(float);
    PPO_REQUIRE(need <= partial_bytes, "head_bwd: partial buffer too small (%zu > %zu)", need, partial_bytes);
    const int kpt = (int)ceil_div(K, 256);
#define PPO_HEAD_BWD(N_, Q_)                                                                                   \
    if (N == N_ && kpt == Q_) {                                                                                \
        if (Hlo)                                                                                               \
            head_bwd_kernel<N_, Q_, true><<<(unsigned)ctas, 256, 0, ctx->stream>>>(H, Hlo, dlogits, W, dH, dHlo, partial, \
                                                                                   M, K, slope, rows, need_dH ? 1 : 0); \
        else                                                                                                   \
            head_bwd_kernel<N_, Q_, false><<<(unsigned)ctas, 256, 0, ctx->stream>>>(H, Hlo, dlogits, W, dH, dHlo, partial, \
                                                                                    M, K, slope, rows, need_dH ? 1 : 0); \
    }
    PPO_HEAD_BWD(1, 1) PPO_HEAD_BWD(1, 2) PPO_HEAD_BWD(1, 3) PPO_HEAD_BWD(1, 4)
    PPO_HEAD_BWD(2, 1) PPO_HEAD_BWD(2, 2) PPO_HEAD_BWD(2, 3) PPO_HEAD_BWD(2, 4)
    PPO_HEAD_BWD(3, 1) PPO_HEAD_BWD(3, 2) PPO_HEAD_BWD(3, 3) PPO_HEAD_BWD(3, 4)
    PPO_HEAD_BWD(4, 1) PPO_HEAD_BWD(4, 2) PPO_HEAD_BWD(4, 3) PPO_HEAD_BWD(4, 4)
#undef PPO_HEAD_BWD
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    const int64_t cnt = (int64_t)K * N;
    reduce_partials_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, ctx->stream>>>(partial, ctas, stride, cnt, dW);
    reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(partial + cnt, ctas, stride, N, db);
    ctx->launches += 2;
    if (db_below != nullptr && need_dH) {
        reduce_partials_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, ctx->stream>>>(partial + cnt + N, ctas, stride, K, db_below);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
