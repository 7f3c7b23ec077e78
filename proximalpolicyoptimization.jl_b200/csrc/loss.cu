// loss.cu — K6: fused masked-softmax + probability ratio + clipped surrogate + smoothed entropy,
// forward AND the gradient w.r.t. the logits, in one pass over HBM.
//
// Replaces, per minibatch (reference file:line):
//   batch_action_probabilities  test/quad_game_utilities.jl:73-79  probs = softmax(logits + mask, dims=1)
//   get_linear_action_index     src/train.jl:48-52                 sel_b = probs[a_b, b]
//   simplified_ppo_clip         src/train.jl:1-7                   clip_b = (1 +- eps) adv_b   (Float64)
//   smoothed_entropy            src/train.jl:21-26                 p~ = p + fl32(1f-8/A); H_b = -sum p~ log p~
//   ppo_loss_with_entropy       src/train.jl:35-46                 ppoloss = -mean min(sel/old*adv, clip)
//   Zygote pullback of the above at src/train.jl:67-79:
//       g[a]  = w_ent/nb (log p~[a] + 1) - [a == a_b][gain_b <= clip_b] adv_b/(nb old_b)
//       dz[a] = p[a] (g[a] - sum_a' p[a'] g[a'])                     (masked => p = 0 => dz = 0 exactly)
// Layout: logits/mask/dlogits are [nb][A] row-major (= Julia [A, nb]); a group of G lanes owns one
// sample and each lane holds V float4 of it in registers, so every byte is read once and written
// once (12*A + 12 B/sample); reductions are warp shuffles; per-block partial sums are folded in
// a fixed order by a finalize kernel (deterministic, no float atomics).
#include "common.cuh"

namespace ppo {

namespace {

constexpr int LOSS_THREADS = 256;

template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// exp(x) for x <= 0 with ~1 ulp error in 6 instructions: 2^(x log2 e) with the product carried in two floats,
// so the argument reduction error of the bare ex2.approx path (|x| * 2^-24) is removed.
__device__ __forceinline__ float fast_exp_neg(float x) {
    const float L2E_HI = 1.4426950216293335f, L2E_LO = 1.9259629911e-8f;
    const float t = x * L2E_HI;
    const float r = fmaf(x, L2E_LO, fmaf(x, L2E_HI, -t));        // low part of x * log2(e)
    const float e = exp2f(t);                                       // ex2.approx (2 ulp)
    return (e == 0.0f) ? 0.0f : fmaf(e * 0.6931471805599453f, r, e);   // e 2^r ~= e (1 + r ln 2); exp(-inf) = 0 exactly
}
__device__ __forceinline__ float fast_log(float x) { return __logf(x); }   // lg2.approx * ln 2: abs err ~1e-7 on [-24, 0]

struct LossArgs {
    const float* logits; const float* mask; const int* action; const float* old_prob; const float* adv;
    int64_t nb; int A;
    double eps_clip; float c_ent;      // w_ent / nb_global
    float inv_nb;                      // 1 / nb_global
    float smooth_over_A;               // fl32(1f-8 / A)
    float* dlogits; float* probs_out; double* partials;
    unsigned* dl_absmax;               // optional: max |dlogits| as the bit pattern of a non-negative float (atomicMax)
};

// max |dlogits| of the block -> atomicMax (order-independent, deterministic): the fp16-split engine derives the scale of
// the gradient operands from it, so the separate abs-max pass over dlogits is not needed
__device__ __forceinline__ void block_absmax_store(float m, unsigned* out) {
    if (out == nullptr) return;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

__device__ __forceinline__ void block_reduce_store(double a, double b, double* partials) {
    __shared__ double sa[LOSS_THREADS / 32], sb[LOSS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, d);
        b += __shfl_down_sync(0xffffffffu, b, d);
    }
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int k = 0; k < LOSS_THREADS / 32; ++k) { x += sa[k]; y += sb[k]; }
        partials[2 * (int64_t)blockIdx.x] = x;
        partials[2 * (int64_t)blockIdx.x + 1] = y;
    }
}

// per-sample scalar tail shared by both kernels
__device__ __forceinline__ void ppo_terms(float sel, float old, float adv, double eps_clip, float inv_nb,
                                          double& mn, float& coef) {
    const float gain = __fmul_rn(__fdiv_rn(sel, old), adv);                 // (sel / old) * adv, fp32
    const double clip = (adv >= 0.0f ? (1.0 + eps_clip) : (1.0 - eps_clip)) * (double)adv;
    const double gd = (double)gain;
    const bool active = !(clip < gd);   // min(gain, clip) keeps gain on ties (Base.min)
    mn = active ? gd : clip;
    coef = active ? __fmul_rn(__fdiv_rn(adv, old), inv_nb) : 0.0f;
}

template <int G, int V>
__global__ void __launch_bounds__(LOSS_THREADS)
loss_vec_kernel(LossArgs a) {
    constexpr int SPB = LOSS_THREADS / G;   // samples per block iteration
    const int gl = threadIdx.x % G;         // lane within the group
    const int gi = threadIdx.x / G;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned group_base = lane - (lane % G);
    double acc_ppo = 0.0, acc_ent = 0.0;
    float amax = 0.0f;

    for (int64_t s0 = (int64_t)blockIdx.x * SPB; s0 < a.nb; s0 += (int64_t)gridDim.x * SPB) {
        const int64_t b = s0 + gi;
        const bool valid = b < a.nb;
        const int64_t bb = valid ? b : (a.nb - 1);   // keep the whole warp converged for shuffles
        const float4* zp = reinterpret_cast<const float4*>(a.logits + bb * a.A);
        const float4* mp = reinterpret_cast<const float4*>(a.mask + bb * a.A);
        float z[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 x = __ldcs(zp + gl + G * v);
            const float4 m = __ldcs(mp + gl + G * v);
            z[v][0] = x.x + m.x; z[v][1] = x.y + m.y; z[v][2] = x.z + m.z; z[v][3] = x.w + m.w;
        }
        float mx = -INFINITY;
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) mx = fmaxf(mx, z[v][c]);
        mx = group_max<G>(mx);
        float sum = 0.0f;
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) { z[v][c] = fast_exp_neg(z[v][c] - mx); sum += z[v][c]; }
        sum = group_sum<G>(sum);
        const float inv_sum = __frcp_rn(sum);

        const int act = a.action[bb];
        const int act_vec = act >> 2, act_c = act & 3;
        const int owner = act_vec % G, owner_v = act_vec / G;
        float g[V][4];
        float ent = 0.0f, sel_local = 0.0f;
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float p = z[v][c] * inv_sum;
                z[v][c] = p;
                const float ps = p + a.smooth_over_A;
                const float lg = fast_log(ps);
                ent = fmaf(ps, lg, ent);
                g[v][c] = a.c_ent * (lg + 1.0f);
                if (v == owner_v && c == act_c) sel_local = p;
            }
        ent = group_sum<G>(ent);
        const float sel = __shfl_sync(0xffffffffu, sel_local, group_base + owner);

        double mn; float coef;
        ppo_terms(sel, a.old_prob[bb], a.adv[bb], a.eps_clip, a.inv_nb, mn, coef);
        if (gl == owner) {
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (v == owner_v && c == act_c) g[v][c] -= coef;
        }
        float dot = 0.0f;
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) dot = fmaf(z[v][c], g[v][c], dot);
        dot = group_sum<G>(dot);

        if (valid) {
            if (a.dlogits != nullptr) {
                float4* dp = reinterpret_cast<float4*>(a.dlogits + b * a.A);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4 d4 = make_float4(z[v][0] * (g[v][0] - dot), z[v][1] * (g[v][1] - dot),
                                                  z[v][2] * (g[v][2] - dot), z[v][3] * (g[v][3] - dot));
                    dp[gl + G * v] = d4;
                    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(d4.x), fabsf(d4.y))), fmaxf(fabsf(d4.z), fabsf(d4.w)));
                }
            }
            if (a.probs_out != nullptr) {
                float4* pp = reinterpret_cast<float4*>(a.probs_out + b * a.A);
#pragma unroll
                for (int v = 0; v < V; ++v) pp[gl + G * v] = make_float4(z[v][0], z[v][1], z[v][2], z[v][3]);
            }
            if (gl == 0) { acc_ppo += mn; acc_ent += (double)(-ent); }
        }
    }
    block_absmax_store(amax, a.dl_absmax);
    block_reduce_store(acc_ppo, acc_ent, a.partials);
}

// any A: one warp per sample, three passes over the (cache-resident) column
__global__ void __launch_bounds__(LOSS_THREADS)
loss_generic_kernel(LossArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPB = LOSS_THREADS / 32;
    double acc_ppo = 0.0, acc_ent = 0.0;
    float amax = 0.0f;
    for (int64_t b = (int64_t)blockIdx.x * WPB + warp; b < a.nb; b += (int64_t)gridDim.x * WPB) {
        const float* zp = a.logits + b * a.A;
        const float* mp = a.mask + b * a.A;
        float mx = -INFINITY;
        for (int i = lane; i < a.A; i += 32) mx = fmaxf(mx, zp[i] + mp[i]);
        mx = group_max<32>(mx);
        float sum = 0.0f;
        for (int i = lane; i < a.A; i += 32) sum += expf(zp[i] + mp[i] - mx);
        sum = group_sum<32>(sum);
        const int act = a.action[b];
        float ent = 0.0f, dot = 0.0f;
        double mn; float coef;
        {
            const float sel = __fdiv_rn(expf(zp[act] + mp[act] - mx), sum);
            ppo_terms(sel, a.old_prob[b], a.adv[b], a.eps_clip, a.inv_nb, mn, coef);
        }
        for (int i = lane; i < a.A; i += 32) {
            const float p = __fdiv_rn(expf(zp[i] + mp[i] - mx), sum);
            const float ps = p + a.smooth_over_A;
            const float lg = logf(ps);
            ent = fmaf(ps, lg, ent);
            float g = a.c_ent * (lg + 1.0f);
            if (i == act) g -= coef;
            dot = fmaf(p, g, dot);
        }
        ent = group_sum<32>(ent);
        dot = group_sum<32>(dot);
        for (int i = lane; i < a.A; i += 32) {
            const float p = __fdiv_rn(expf(zp[i] + mp[i] - mx), sum);
            const float ps = p + a.smooth_over_A;
            float g = a.c_ent * (logf(ps) + 1.0f);
            if (i == act) g -= coef;
            if (a.dlogits != nullptr) { const float dz = p * (g - dot); a.dlogits[b * a.A + i] = dz; amax = fmaxf(amax, fabsf(dz)); }
            if (a.probs_out != nullptr) a.probs_out[b * a.A + i] = p;
        }
        if (lane == 0) { acc_ppo += mn; acc_ent += (double)(-ent); }
    }
    block_absmax_store(amax, a.dl_absmax);
    block_reduce_store(acc_ppo, acc_ent, a.partials);
}

// out[0] = ppoloss = -(1/nb) sum min(gain, clip);  out[1] = entropyloss = -(1/nb) sum H_b
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const double* __restrict__ partials, int64_t blocks, double inv_nb, double* __restrict__ out2,
                     const int* __restrict__ step) {
    if (step != nullptr) out2 += 2 * (int64_t)(*step);
    __shared__ double s1[256], s2[256];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < blocks; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { s1[threadIdx.x] += s1[threadIdx.x + d]; s2[threadIdx.x] += s2[threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = -s1[0] * inv_nb; out2[1] = -s2[0] * inv_nb; }
}

struct Cfg { int G, V; };
// A = 4 G V floats per sample: G lanes (a power of two <= 32) x V float4 per lane.  Prefer few lanes with several
// float4 each: the per-sample work (three shuffle reductions, the clip/ratio tail in Float64) is then shared by
// fewer lanes, which is what bounds the kernel at small A.
inline bool pick_cfg(int A, Cfg& c) {
    if (A % 4 != 0) return false;
    const int q = A / 4;
    for (int V = 4; V >= 1; --V) {
        if (q % V) continue;
        const int G = q / V;
        if (G <= 32 && (G & (G - 1)) == 0) { c.G = G; c.V = V; return true; }
    }
    return false;
}

}  // namespace

int64_t loss_num_blocks(int64_t nb, int A) {
    Cfg c;
    int64_t spb = pick_cfg(A, c) ? LOSS_THREADS / c.G : LOSS_THREADS / 32;
    int64_t blocks = ceil_div(nb > 0 ? nb : 1, spb);
    const int64_t cap = 148 * 8 * 2;
    return blocks > cap ? cap : blocks;
}

int launch_loss(ppo_ctx* ctx, const float* logits, const float* mask, const int* action, const float* old_prob,
                const float* adv, int64_t nb, int A, double epsilon, double entropy_weight, double inv_nb_global,
                float* dlogits, double* partials, double* loss_out2, float* probs_out, const int* step,
                unsigned* dl_absmax) {
    PPO_REQUIRE(nb >= 1 && A >= 1, "loss: nb=%lld A=%d", (long long)nb, A);
    LossArgs a;
    a.logits = logits; a.mask = mask; a.action = action; a.old_prob = old_prob; a.adv = adv;
    a.nb = nb; a.A = A; a.eps_clip = epsilon;
    a.inv_nb = (float)inv_nb_global;
    a.c_ent = (float)entropy_weight * (float)inv_nb_global;
    a.smooth_over_A = 1e-8f / (float)A;
    a.dlogits = dlogits; a.probs_out = probs_out; a.partials = partials;
    a.dl_absmax = dlogits != nullptr ? dl_absmax : nullptr;
    const int64_t blocks = loss_num_blocks(nb, A);
    Cfg c;
    const bool aligned = ((uintptr_t)logits % 16 == 0) && ((uintptr_t)mask % 16 == 0) &&
                         (dlogits == nullptr || (uintptr_t)dlogits % 16 == 0) &&
                         (probs_out == nullptr || (uintptr_t)probs_out % 16 == 0);
    if (pick_cfg(A, c) && aligned) {
#define PPO_LOSS_CASE(G_, V_) \
    if (c.G == G_ && c.V == V_) loss_vec_kernel<G_, V_><<<(unsigned)blocks, LOSS_THREADS, 0, ctx->stream>>>(a)
        PPO_LOSS_CASE(1, 1); PPO_LOSS_CASE(1, 2); PPO_LOSS_CASE(1, 3); PPO_LOSS_CASE(1, 4);
        PPO_LOSS_CASE(2, 3); PPO_LOSS_CASE(2, 4); PPO_LOSS_CASE(4, 3); PPO_LOSS_CASE(4, 4);
        PPO_LOSS_CASE(8, 3); PPO_LOSS_CASE(8, 4); PPO_LOSS_CASE(16, 3); PPO_LOSS_CASE(16, 4);
        PPO_LOSS_CASE(32, 3); PPO_LOSS_CASE(32, 4);
#undef PPO_LOSS_CASE
    } else {
        loss_generic_kernel<<<(unsigned)blocks, LOSS_THREADS, 0, ctx->stream>>>(a);
    }
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    loss_finalize_kernel<<<1, 256, 0, ctx->stream>>>(partials, blocks, inv_nb_global, loss_out2, step);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}


// ---- batched rollout inference -------------------------------------------------------
// One categorical draw per state by inverse CDF over the probabilities the softmax kernel produced, as
// `rand(Categorical(ap))` does for one state on the host (reference src/collect_rollouts.jl:5-7): sequential Float32
// cumulative sum, first index whose cumulative probability exceeds the Float32 uniform draw.  The draw of state i is
// output i of the splitmix64 stream of `seed` (counter-based: s_i = seed + (i + 1) * golden), top 24 bits.
// Masked actions have probability exactly 0 and can never be returned: if rounding leaves the total below the draw,
// the last action with non-zero probability is returned.
__global__ void __launch_bounds__(128)
sample_actions_kernel(const float* __restrict__ probs, int64_t nb, int A, uint64_t seed, int64_t* __restrict__ action1,
                      float* __restrict__ prob_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    const float* p = probs + i * A;
    float c = p[0];
    int a = 0, last = p[0] > 0.0f ? 0 : -1;
    while (c <= u && a < A - 1) {
        ++a;
        const float pa = p[a];
        c = __fadd_rn(c, pa);
        if (pa > 0.0f) last = a;
    }
    if (!(p[a] > 0.0f)) a = last < 0 ? a : last;
    action1[i] = (int64_t)a + 1;
    prob_out[i] = p[a];
}

int launch_sample_actions(ppo_ctx* ctx, const float* probs, int64_t nb, int A, uint64_t seed, int64_t* action1, float* prob_out) {
    sample_actions_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, ctx->stream>>>(probs, nb, A, seed, action1, prob_out);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
