// shuffle.cu — K3: device-side minibatch permutation.
//
// Replaces `file_indices = randperm(num_data)` (reference src/train.jl:93) and `shuffle(1:N)`
// (src/rollout_buffer.jl:90-93).  Julia's randperm is stdlib code outside the reference tree,
// sequential and version dependent, so the contract is: (a) any host-supplied permutation is
// honoured bit-exactly (launch_perm_from_host), and (b) the device draws perm[i] = walk(i) with
// a counter-based cycle-walking generalised Feistel bijection whose CPU restatement
// (oracle/ppo_oracle.py:feistel_permutation, oracle/ppo_oracle_c.c) is bit-identical.
// Embarrassingly parallel, no key sort: 4 B written per index (int32, 0-based on the device).
#include "common.cuh"

namespace ppo {

namespace {

constexpr int FEISTEL_ROUNDS = 10;

struct FeistelKeys {
    uint32_t k[FEISTEL_ROUNDS];
};

__host__ __device__ inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

inline uint64_t splitmix64_next(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
feistel_kernel(int* __restrict__ perm0, int64_t n, FeistelKeys keys, int abits, int bbits) {
    const uint32_t amask = (1u << abits) - 1u, bmask = (1u << bbits) - 1u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)i;
        do {
            uint32_t a = (uint32_t)(x >> bbits), b = (uint32_t)x & bmask;
#pragma unroll
            for (int r = 0; r < FEISTEL_ROUNDS; ++r) {
                if ((r & 1) == 0) a ^= fmix32(b ^ keys.k[r]) & amask;
                else              b ^= fmix32(a ^ keys.k[r]) & bmask;
            }
            x = ((uint64_t)a << bbits) | b;
        } while (x >= (uint64_t)n);
        perm0[i] = (int)x;
    }
}

// Int64 1-based (Julia) -> int32 0-based, range-checked against [1, limit]
__global__ void __launch_bounds__(256)
perm_in_kernel(const int64_t* __restrict__ p1, int* __restrict__ p0, int64_t n, int64_t limit, int* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = p1[i];
        if (v < 1 || v > limit) { atomicExch(bad, 1); v = 1; }
        p0[i] = (int)(v - 1);
    }
}

__global__ void __launch_bounds__(256)
perm_out_kernel(const int* __restrict__ p0, int64_t* __restrict__ p1, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p1[i] = (int64_t)p0[i] + 1;
}

inline unsigned grid_for(ppo_ctx* ctx, int64_t n, int per_block) {
    int64_t want = ceil_div(n, per_block);
    int64_t cap = (int64_t)ctx->num_sms * 8;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

int launch_feistel_permutation(ppo_ctx* ctx, int* perm0, int64_t n, uint64_t seed) {
    if (n <= 0) return PPO_OK;
    PPO_REQUIRE(n < ((int64_t)1 << 31), "permutation: n must be < 2^31");
    FeistelKeys keys;
    uint64_t s = seed;
    for (int r = 0; r < FEISTEL_ROUNDS; ++r) keys.k[r] = (uint32_t)splitmix64_next(s);
    int bits = 2;
    while (((int64_t)1 << bits) < n) ++bits;
    int abits = bits / 2, bbits = bits - abits;
    feistel_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(perm0, n, keys, abits, bbits);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_perm_from_host(ppo_ctx* ctx, const int64_t* d_perm1, int* perm0, int64_t n, int64_t limit,
                          int* d_bad) {
    if (n <= 0) return PPO_OK;
    perm_in_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(d_perm1, perm0, n, limit, d_bad);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_perm_to_i64(ppo_ctx* ctx, const int* perm0, int64_t* d_perm1, int64_t n) {
    if (n <= 0) return PPO_OK;
    perm_out_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(perm0, d_perm1, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
