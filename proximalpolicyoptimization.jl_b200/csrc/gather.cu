// gather.cu — K4: minibatch gather (and the small index/dtype conversion kernels).
//
// Replaces get_batch (reference src/rollout_buffer.jl:117-133) + batch_state
// (test/quad_game_utilities.jl:26-33): rows perm[start .. start+count) of the SoA buffer are
// copied into a contiguous minibatch
//     feat_out[b] = feat[idx[b]]   (nhe*nf floats, one contiguous record per transition)
//     mask_out[b] = mask[idx[b]]   (A floats)
//     action/old_prob/returns scalars (returns -> advantage; identity or the K2 normalisation)
// Pure byte movement: 4 + 2*R bytes per sample, R = 4*nf*nhe + 4*A + 12, HBM-bound, bit-exact.
//
// Two implementations, selectable for measurement:
//   variant 0: one warp per record, 128-bit LDG/STG, 8 independent loads in flight per lane.
//   variant 1: TMA bulk copies (cp.async.bulk global->shared->global, SASS UBLKCP) driven by one
//              elected lane per warp through an mbarrier ring: no register staging at all.
#include "common.cuh"

namespace ppo {

namespace {

// ------------------------------------------------------------------------------------------
// variant 0: vectorised LDG/STG
// ------------------------------------------------------------------------------------------
template <int UNROLL>
__device__ __forceinline__ void warp_copy_vec4(const float4* __restrict__ s, float4* __restrict__ d,
                                               int nvec, int lane) {
    int i = lane;
    for (; i + (UNROLL - 1) * 32 < nvec; i += UNROLL * 32) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(s + i + u * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) d[i + u * 32] = v[u];
    }
    for (; i < nvec; i += 32) d[i] = __ldcs(s + i);
}

// the per-sample scalars of a record, by three lanes of the warp that copies it (no second launch)
__device__ __forceinline__ void gather_record_scalars(const GatherArgs& a, int64_t rec, int64_t src, int lane) {
    if (lane == 0 && a.action_out) a.action_out[rec] = a.action[src];
    if (lane == 1 && a.prob_out) a.prob_out[rec] = a.old_prob[src];
    if (lane == 2 && a.adv_out) {
        const float r = a.ret[src];
        a.adv_out[rec] = (a.norm != nullptr) ? (r - a.norm[0]) * a.norm[1] : r;
    }
}

__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(GatherArgs a) {
    if (a.step != nullptr) a.index += (int64_t)(*a.step) * a.step_stride;   // minibatch number lives on the device (CUDA graph replay)
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int fvec = a.feat_elems >> 2, mvec = a.mask_elems >> 2;
    for (int64_t rec = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; rec < a.count; rec += warps) {
        const int64_t src = a.index[rec];
        warp_copy_vec4<8>(reinterpret_cast<const float4*>(a.feat + src * a.feat_elems),
                          reinterpret_cast<float4*>(a.feat_out + rec * a.feat_elems), fvec, lane);
        warp_copy_vec4<2>(reinterpret_cast<const float4*>(a.mask + src * a.mask_elems),
                          reinterpret_cast<float4*>(a.mask_out + rec * a.mask_elems), mvec, lane);
        gather_record_scalars(a, rec, src, lane);
    }
}

__global__ void __launch_bounds__(256)
gather_rows_scalar_kernel(GatherArgs a) {
    if (a.step != nullptr) a.index += (int64_t)(*a.step) * a.step_stride;
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t rec = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; rec < a.count; rec += warps) {
        const int64_t src = a.index[rec];
        const float* fs = a.feat + src * a.feat_elems;
        float* fd = a.feat_out + rec * a.feat_elems;
        for (int i = lane; i < a.feat_elems; i += 32) fd[i] = fs[i];
        const float* ms = a.mask + src * a.mask_elems;
        float* md = a.mask_out + rec * a.mask_elems;
        for (int i = lane; i < a.mask_elems; i += 32) md[i] = ms[i];
        gather_record_scalars(a, rec, src, lane);
    }
}

// per-sample scalars: one thread per sample, coalesced stores
__global__ void __launch_bounds__(256)
gather_scalars_kernel(GatherArgs a) {
    if (a.step != nullptr) a.index += (int64_t)(*a.step) * a.step_stride;
    float mu = 0.0f, inv = 1.0f;
    if (a.norm != nullptr) { mu = a.norm[0]; inv = a.norm[1]; }
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.count;
         b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t src = a.index[b];
        if (a.action_out) a.action_out[b] = a.action[src];
        if (a.prob_out) a.prob_out[b] = a.old_prob[src];
        if (a.adv_out) {
            float r = a.ret[src];
            a.adv_out[b] = (a.norm != nullptr) ? (r - mu) * inv : r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// variant 1: TMA bulk copy ring (cp.async.bulk, no tensor map needed for 1-D records)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int BULK_WARPS = 4;
constexpr int BULK_MAX_STAGES = 8;

// dynamic smem: [BULK_WARPS][stages][rec_bytes] then mbarriers
__global__ void __launch_bounds__(BULK_WARPS * 32)
gather_rows_bulk_kernel(GatherArgs a, int stages, int rec_bytes_padded) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (a.step != nullptr) a.index += (int64_t)(*a.step) * a.step_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)BULK_WARPS * stages * rec_bytes_padded);
    unsigned char* my = smem + (size_t)warp * stages * rec_bytes_padded;
    uint64_t* mybar = bars + warp * BULK_MAX_STAGES;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(mybar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane != 0) return;

    const uint32_t fbytes = (uint32_t)a.feat_elems * 4u, mbytes = (uint32_t)a.mask_elems * 4u;
    const int64_t gw = (int64_t)blockIdx.x * BULK_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * BULK_WARPS;
    // records of this warp: gw, gw+nw, ...
    const int64_t mine = (a.count > gw) ? (a.count - gw + nw - 1) / nw : 0;

    auto issue_load = [&](int64_t k) {
        const int s = (int)(k % stages);
        const int64_t rec = gw + k * nw;
        const int64_t src = a.index[rec];
        unsigned char* dst = my + (size_t)s * rec_bytes_padded;
        mbar_expect_tx(mybar + s, fbytes + mbytes);
        bulk_g2s(dst, a.feat + src * a.feat_elems, fbytes, mybar + s);
        bulk_g2s(dst + fbytes, a.mask + src * a.mask_elems, mbytes, mybar + s);
    };

    const int64_t pro = mine < stages ? mine : stages;
    for (int64_t k = 0; k < pro; ++k) issue_load(k);
    for (int64_t k = 0; k < mine; ++k) {
        const int s = (int)(k % stages);
        const uint32_t parity = (uint32_t)((k / stages) & 1);
        mbar_wait(mybar + s, parity);
        const int64_t rec = gw + k * nw;
        unsigned char* src = my + (size_t)s * rec_bytes_padded;
        bulk_s2g(a.feat_out + rec * a.feat_elems, src, fbytes);
        bulk_s2g(a.mask_out + rec * a.mask_elems, src + fbytes, mbytes);
        bulk_commit();
        // refill the stage used one iteration ago: its store group is the second newest
        if (k >= 1 && (k - 1) + stages < mine) {
            bulk_wait_read<1>();
            issue_load(k - 1 + stages);
        }
    }
    bulk_wait_all<0>();
}

// ------------------------------------------------------------------------------------------
// small conversion kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
permute_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ idx,
                  int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256)
actions_in_kernel(const int64_t* __restrict__ a1, int* __restrict__ a0, int64_t n, int A, int* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = a1[i];
        if (v < 1 || v > A) { atomicExch(bad, 1); v = 1; }
        a0[i] = (int)(v - 1);
    }
}

__global__ void __launch_bounds__(256)
actions_out_kernel(const int* __restrict__ a0, int64_t* __restrict__ a1, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        a1[i] = (int64_t)a0[i] + 1;
}

// linear (1-based, column-major into probs[A, nb]) -> 0-based action within the column;
// get_linear_action_index, reference src/train.jl:48-52, inverted and range-checked.
__global__ void __launch_bounds__(256)
linear_in_kernel(const int64_t* __restrict__ lin1, int* __restrict__ a0, int64_t n, int A, int* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = lin1[i] - 1 - i * (int64_t)A;
        if (v < 0 || v >= A) { atomicExch(bad, 1); v = 0; }
        a0[i] = (int)v;
    }
}

__global__ void step_advance_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256)
normalize_bool_kernel(uint8_t* __restrict__ t, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        t[i] = t[i] != 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256)
i64_to_f32_kernel(const int64_t* __restrict__ s, float* __restrict__ d, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        d[i] = (float)s[i];
}

// narrow integer features -> Float32 (exact), 16 elements per thread: 16 / 32 bytes in, 64 bytes out
// block-wide max of a non-negative float -> atomicMax on its bit pattern (order-independent, deterministic)
__device__ __forceinline__ void block_absmax_to(float m, unsigned* out) {
    if (out == nullptr) return;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

// max |x| of a float array (the rollout buffer keeps the abs-max of its features up to date at every append, so that the
// fp16-split engine has a bound for a minibatch's features without a pass over them)
__global__ void __launch_bounds__(256)
absmax_f32_kernel(const float* __restrict__ x, int64_t n, unsigned* out) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
    block_absmax_to(m, out);
}

template <typename T>
__global__ void __launch_bounds__(256)
narrow_to_f32_kernel(const T* __restrict__ s, float* __restrict__ d, int64_t n, unsigned* absmax_out) {
    float amax = 0.0f;
    const int64_t n16 = n >> 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        T v[16];
        if (sizeof(T) == 1) {
            *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(s) + i);
        } else {
            reinterpret_cast<uint4*>(v)[0] = __ldcs(reinterpret_cast<const uint4*>(s) + 2 * i);
            reinterpret_cast<uint4*>(v)[1] = __ldcs(reinterpret_cast<const uint4*>(s) + 2 * i + 1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(d)[4 * i + q] = make_float4((float)v[4 * q], (float)v[4 * q + 1], (float)v[4 * q + 2], (float)v[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < 16; ++q) amax = fmaxf(amax, fabsf((float)v[q]));
    }
    const int64_t tail = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && tail < n) { d[tail] = (float)s[tail]; amax = fmaxf(amax, fabsf((float)s[tail])); }      // < 16 leftover elements
    block_absmax_to(amax, absmax_out);
}

// action mask from one bit per action (1 = allowed -> 0.0f, 0 = masked -> -Inf32): bit i of the stream is bit (i & 63)
// of word i >> 6 (the layout of a Julia BitMatrix's chunks)
__global__ void __launch_bounds__(256)
mask_from_bits_kernel(const uint64_t* __restrict__ bits, float* __restrict__ mask, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mask[i] = ((__ldg(bits + (i >> 6)) >> (i & 63)) & 1ull) ? 0.0f : -INFINITY;
}

// element-wise variant for a destination that is not 16-byte aligned (an append at an odd element offset)
template <typename T>
__global__ void __launch_bounds__(256)
narrow_to_f32_scalar_kernel(const T* __restrict__ s, float* __restrict__ d, int64_t n, unsigned* absmax_out) {
    float amax = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        d[i] = (float)s[i];
        amax = fmaxf(amax, fabsf((float)s[i]));
    }
    block_absmax_to(amax, absmax_out);
}

inline unsigned grid_for(ppo_ctx* ctx, int64_t n, int per_block, int waves = 8) {
    int64_t want = ceil_div(n, per_block);
    int64_t cap = (int64_t)ctx->num_sms * waves;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

int launch_gather(ppo_ctx* ctx, const GatherArgs& a, int variant) {
    if (a.count <= 0) return PPO_OK;
    const bool vec_ok = (a.feat_elems % 4 == 0) && (a.mask_elems % 4 == 0) &&
                        ((uintptr_t)a.feat % 16 == 0) && ((uintptr_t)a.mask % 16 == 0) &&
                        ((uintptr_t)a.feat_out % 16 == 0) && ((uintptr_t)a.mask_out % 16 == 0);
    bool scalars_done = false;
    if (a.feat_out != nullptr) {
        if (variant == 1 && vec_ok) {
            const int rec = (a.feat_elems + a.mask_elems) * 4;
            const int rec_pad = (int)round_up(rec, 128);
            int stages = (int)(200 * 1024 / ((int64_t)BULK_WARPS * rec_pad));
            if (stages > BULK_MAX_STAGES) stages = BULK_MAX_STAGES;
            PPO_REQUIRE(stages >= 2, "gather(bulk): record of %d bytes too large for the shared-memory ring", rec);
            size_t smem = (size_t)BULK_WARPS * stages * rec_pad + BULK_WARPS * BULK_MAX_STAGES * sizeof(uint64_t);
            PPO_CUDA(cudaFuncSetAttribute(gather_rows_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            int per_sm = (int)((220 * 1024) / smem);
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 4) per_sm = 4;
            int64_t blocks = (int64_t)ctx->num_sms * per_sm;
            int64_t need = ceil_div(a.count, BULK_WARPS);
            if (blocks > need) blocks = need;
            gather_rows_bulk_kernel<<<(unsigned)blocks, BULK_WARPS * 32, smem, ctx->stream>>>(a, stages, rec_pad);
        } else if (vec_ok) {
            // 8 warps per block; enough blocks for ~2 waves at 8 blocks/SM
            gather_rows_vec_kernel<<<grid_for(ctx, a.count, 8, 16), 256, 0, ctx->stream>>>(a);
            scalars_done = true;
        } else {
            gather_rows_scalar_kernel<<<grid_for(ctx, a.count, 8, 16), 256, 0, ctx->stream>>>(a);
            scalars_done = true;
        }
        ctx->launches += 1;
        PPO_CUDA(cudaGetLastError());
    }
    if (!scalars_done && (a.action_out || a.prob_out || a.adv_out)) {
        gather_scalars_kernel<<<grid_for(ctx, a.count, 256), 256, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        PPO_CUDA(cudaGetLastError());
    }
    return PPO_OK;
}

int launch_permute_inplace_u8(ppo_ctx* ctx, const uint8_t* src, uint8_t* dst, const int* idx, int64_t n) {
    if (n <= 0) return PPO_OK;
    permute_u8_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(src, dst, idx, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_convert_actions_in(ppo_ctx* ctx, const int64_t* a1, int* a0, int64_t n, int A, int* d_bad) {
    if (n <= 0) return PPO_OK;
    actions_in_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(a1, a0, n, A, d_bad);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_convert_actions_out(ppo_ctx* ctx, const int* a0, int64_t* a1, int64_t n) {
    if (n <= 0) return PPO_OK;
    actions_out_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(a0, a1, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_linear_index_in(ppo_ctx* ctx, const int64_t* lin1, int* a0, int64_t n, int A, int* d_bad) {
    if (n <= 0) return PPO_OK;
    linear_in_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(lin1, a0, n, A, d_bad);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_step_advance(ppo_ctx* ctx, int* d_step) {
    step_advance_kernel<<<1, 1, 0, ctx->stream>>>(d_step);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_normalize_bool(ppo_ctx* ctx, uint8_t* t, int64_t n) {
    if (n <= 0) return PPO_OK;
    normalize_bool_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_i64_to_f32(ppo_ctx* ctx, const int64_t* src, float* dst, int64_t n) {
    if (n <= 0) return PPO_OK;
    i64_to_f32_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(src, dst, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_mask_from_bits(ppo_ctx* ctx, const uint64_t* bits, float* mask, int64_t n) {
    if (n <= 0) return PPO_OK;
    mask_from_bits_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(bits, mask, n);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_absmax_f32(ppo_ctx* ctx, const float* x, int64_t n, unsigned* out) {
    if (n <= 0) return PPO_OK;
    absmax_f32_kernel<<<grid_for(ctx, n, 1024, 8), 256, 0, ctx->stream>>>(x, n, out);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_narrow_to_f32(ppo_ctx* ctx, const void* src, int elem_bytes, float* dst, int64_t n, unsigned* absmax_out) {
    if (n <= 0) return PPO_OK;
    if (((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
        const unsigned grid = grid_for(ctx, n / 16 + 1, 256, 16);
        if (elem_bytes == 1) narrow_to_f32_kernel<int8_t><<<grid, 256, 0, ctx->stream>>>((const int8_t*)src, dst, n, absmax_out);
        else narrow_to_f32_kernel<int16_t><<<grid, 256, 0, ctx->stream>>>((const int16_t*)src, dst, n, absmax_out);
    } else {
        const unsigned grid = grid_for(ctx, n, 256, 16);
        if (elem_bytes == 1) narrow_to_f32_scalar_kernel<int8_t><<<grid, 256, 0, ctx->stream>>>((const int8_t*)src, dst, n, absmax_out);
        else narrow_to_f32_scalar_kernel<int16_t><<<grid, 256, 0, ctx->stream>>>((const int16_t*)src, dst, n, absmax_out);
    }
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
