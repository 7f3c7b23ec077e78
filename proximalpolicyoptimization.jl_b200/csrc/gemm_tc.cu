// gemm_tc.cu — K5/K7 on the 5th-generation tensor cores: hand-written tcgen05 + TMEM + TMA kernels.
//
// The policy MLP (reference test/policy.jl:9-31) is the only dense contraction on the hot path.
// The reference is Float32 end to end, and parity is stated at 1e-5 relative, which a single
// TF32/BF16 pass (~1e-3) cannot meet.  PPO_GEMM_TF32X3_TC therefore runs the error-compensated
// 3-pass split on tcgen05.mma.kind::tf32 with fp32 accumulation in TMEM:
//        a = a_hi + a_lo,   a_hi = the top 19 bits of a (what the tensor core reads from an fp32 word),
//                           a_lo = rna_tf32(a - a_hi)                     (exactly representable)
//        A B ~= A_hi B_hi + A_hi B_lo + A_lo B_hi                        (error ~2^-21 per product)
// "hi" operands are simply the fp32 tensors themselves; every producer (GEMM epilogue, split
// kernel) writes the companion "lo" tensor next to its output.
//
//   kk kernel   (forward, dgrad): D[M,N] = sum_k A[M,k] B[N,k], both operands K-major.
//       persistent CTAs, 6 warps: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle),
//       warp 1 = single-thread tcgen05.mma issuer (UMMA 128 x BN x 8, 3 passes per k-step) and
//       TMEM owner (2 x BN fp32 accumulator columns, double buffered), warps 2-5 = epilogue
//       (tcgen05.ld -> bias/leakyrelu or leakyrelu' gate -> hi/lo split -> swizzled smem -> TMA store).
//   mn kernel   (wgrad): D[Kin,Nout] = sum_m X[m,Kin] dY[m,Nout]; both operands MN-major straight
//       from the row-major activations (3-D tensor maps lay the 32-column blocks out as the
//       canonical MN-major SWIZZLE_128B atoms), split over m across CTAs, partials reduced in a
//       fixed order (deterministic).
//
// Layer shapes the tensor-core kernels do not cover fall back to the fp32 FFMA CUDA kernels of
// gemm_simt.cu layer by layer (still on the GPU; never a CPU path).
#include <cuda.h>

#include <algorithm>

#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace ppo {

namespace {

// error-compensated split: hi = rna_tf32(a) (low 13 mantissa bits zero, so the tensor core reads it
// exactly), lo = rna_tf32(a - hi) (|lo| <= 2^-11 |a|, signed and unbiased)
__device__ __forceinline__ void tf32_split(float a, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - hi));
    lo = __uint_as_float(l);
}

// ---------------------------------------------------------------------------------------------
// Tile configuration shared by both kernels.
//
// The tensor core adds into its fp32 accumulator with round-toward-zero, so the error of an
// accumulation chain grows linearly with its length (measured: ~3e-8 of the accumulator per MMA).
// Both kernels therefore cut the contraction into chunks of TC_CHUNK_KB k-blocks (64 elements =
// 24 MMAs), let the tensor core accumulate ONE chunk in a TMEM accumulator stage, and have the
// epilogue warps fold the finished chunk into fp32 registers with round-to-nearest adds while the
// MMAs of the next chunk fill the other stage.  Together with the rna hi/lo split this brings the
// GEMMs to fp32-grade error (~5e-7) at tensor-core speed.
// ---------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                   // 32 tf32 = one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_CHUNK_KB = 2;              // k-blocks per TMEM accumulation chain
constexpr int TC_EPI_THREADS = 256;         // 8 epilogue warps: 4 lane quarters x 2 column halves
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 4;   // 16 KB

enum { TC_EPI_FWD = 0, TC_EPI_DGRAD = 1 };

struct KKParams {
    int M, N, K;
    int tiles_m, tiles_n, k_blocks;
    int epi;
    int act;                 // fwd: apply leakyrelu
    float slope;
    const float* bias;       // fwd
    const float* gate;       // dgrad: activation whose sign gates the gradient, [M][N]
    float* colsum_partial;   // dgrad: [tiles_m][4][N] column sums of the output per 32-row quarter (or nullptr)
};

template <int BN>
struct KKSmem {
    static constexpr int B_TILE_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    // 4 independent store groups (row half x column half, 2 warps each); each stages 64 rows x 16 columns
    // of hi and of lo (8 KB) for its own TMA stores
    static constexpr int STAGING_BYTES = 4 * 2 * 64 * 16 * 4;
    static constexpr int TOTAL = TC_STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// fold one finished TMEM accumulation chain into the thread's fp32 running sums
template <int CPT>
__device__ __forceinline__ void drain_chunk(uint32_t taddr, float* s) {
#pragma unroll
    for (int c4 = 0; c4 < CPT / 32; ++c4) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)(c4 * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[c4 * 32 + j] += v[j];
    }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kk_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const __grid_constant__ CUtensorMap tmC_hi, const __grid_constant__ CUtensorMap tmC_lo,
                  const KKParams p) {
    using S = KKSmem<BN>;
    constexpr int CPT = BN / 2;   // columns per epilogue thread
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* staging = smem + TC_STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES);
    uint64_t* full = bars;                    // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;   // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int chunks_per_tile = (p.k_blocks + TC_CHUNK_KB - 1) / TC_CHUNK_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, TC_EPI_THREADS); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        tma_prefetch_desc(&tmC_hi); tma_prefetch_desc(&tmC_lo);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.tiles_n) * TC_BM, n0 = (tile % p.tiles_n) * BN;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    unsigned char* st = smem + stage * S::STAGE_BYTES;
                    mbar_expect_tx(full + stage, (uint32_t)S::STAGE_BYTES);
                    const int k0 = kb * TC_BK;
                    tma_load_2d(st, &tmA_hi, k0, m0, full + stage);
                    tma_load_2d(st + A_TILE_BYTES, &tmA_lo, k0, m0, full + stage);
                    tma_load_2d(st + 2 * A_TILE_BYTES, &tmB_hi, k0, n0, full + stage);
                    tma_load_2d(st + 2 * A_TILE_BYTES + S::B_TILE_BYTES, &tmB_lo, k0, n0, full + stage);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            uint32_t cc = 0;   // accumulation chains issued so far
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.k_blocks; ++cc) {
                    const int acc = (int)(cc & 1u);
                    mbar_wait(tempty + acc, ((cc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    const int kb_end = kb + TC_CHUNK_KB < p.k_blocks ? kb + TC_CHUNK_KB : p.k_blocks;
                    for (int kc = 0; kb < kb_end; ++kb, ++kc) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                        const uint64_t a_hi = make_desc(sa, 16, 1024);
                        const uint64_t a_lo = make_desc(sa + A_TILE_BYTES, 16, 1024);
                        const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES, 16, 1024);
                        const uint64_t b_lo = make_desc(sa + 2 * A_TILE_BYTES + S::B_TILE_BYTES, 16, 1024);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            const uint64_t koff = (uint64_t)(k * 2);   // 8 tf32 = 32 bytes = 2 x 16 B
                            umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc | k) != 0 ? 1u : 0u);
                            umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                            umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                        }
                        tc_commit(empty + stage);            // frees the smem slot when these MMAs retire
                        if (kb == kb_end - 1) tc_commit(tfull + acc);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps = 4 lane quarters x 2 column halves =====================
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which BN/2 columns
        const int row_in_tile = q * 32 + lane;
        // store group = (row half, column half): two warps, 64 rows x CPT columns, own staging + own TMA stores
        const int rh = q >> 1;
        const int grp = rh * 2 + half;
        const int rg = (q & 1) * 32 + lane;            // row within the group's 64 rows
        float* st_hi = reinterpret_cast<float*>(staging + grp * 8192);
        float* st_lo = st_hi + 64 * 16;
        const bool storer = ((q & 1) == 0) && lane == 0;
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.tiles_n) * TC_BM, n0 = (tile % p.tiles_n) * BN;
            const int row = m0 + row_in_tile;
            float s[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) s[j] = 0.0f;
            if (p.epi == TC_EPI_DGRAD && row < p.M) {
                // the gate values (this thread's CPT columns of the activation) are needed only after the whole
                // contraction: pull their lines towards the SM now so that the epilogue does not stall the MMA issuer
                const char* gp = reinterpret_cast<const char*>(p.gate + (size_t)row * p.N + n0 + half * CPT);
#pragma unroll
                for (int l = 0; l < CPT * 4 / 128; ++l)
                    if (n0 + half * CPT + l * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + l * 128));
            }
            for (int ch = 0; ch < chunks_per_tile; ++ch, ++cc) {
                const int acc = (int)(cc & 1u);
                mbar_wait(tfull + acc, (cc >> 1) & 1u);
                tc_fence_after();
                drain_chunk<CPT>(tmem_base + (uint32_t)(acc * BN + half * CPT) + ((uint32_t)(q * 32) << 16), s);
                tc_fence_before();
                mbar_arrive(tempty + acc);
            }
            // ---- bias / activation (or gradient gate), hi/lo split, staged TMA store (16 columns at a time) ----
#pragma unroll
            for (int g = 0; g < CPT / 16; ++g) {
                const int col0 = n0 + half * CPT + g * 16;
                float* v = s + g * 16;
                if (p.epi == TC_EPI_FWD) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = col0 + j;
                        const float x = v[j] + ((p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.0f);
                        v[j] = (p.act && !(x > 0.0f)) ? p.slope * x : x;
                    }
                } else if (row < p.M) {
                    const float* gp = p.gate + (size_t)row * p.N + col0;
                    if (col0 + 16 <= p.N) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 h = __ldg(reinterpret_cast<const float4*>(gp) + j4);
                            v[4 * j4 + 0] = (h.x > 0.0f) ? v[4 * j4 + 0] : p.slope * v[4 * j4 + 0];
                            v[4 * j4 + 1] = (h.y > 0.0f) ? v[4 * j4 + 1] : p.slope * v[4 * j4 + 1];
                            v[4 * j4 + 2] = (h.z > 0.0f) ? v[4 * j4 + 2] : p.slope * v[4 * j4 + 2];
                            v[4 * j4 + 3] = (h.w > 0.0f) ? v[4 * j4 + 3] : p.slope * v[4 * j4 + 3];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (col0 + j < p.N) v[j] = (__ldg(gp + j) > 0.0f) ? v[j] : p.slope * v[j];
                    }
                }
                if (p.colsum_partial != nullptr) {
                    // column sums of this warp's 32 rows x 16 columns with a halving butterfly (16 shuffles):
                    // = the bias gradient of the layer below, fused here so that dX is never re-read for it
                    float k8[8], k4[4], k2[2], k1;
                    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float send = b4 ? v[i] : v[i + 8];
                        k8[i] = (b4 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float send = b3 ? k8[i] : k8[i + 4];
                        k4[i] = (b3 ? k8[i + 4] : k8[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = b2 ? k4[i] : k4[i + 2];
                        k2[i] = (b2 ? k4[i + 2] : k4[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    {
                        const float send = b1 ? k2[0] : k2[1];
                        k1 = (b1 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
                    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
                    const int cidx = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
                    if ((lane & 1) == 0 && col0 + cidx < p.N)
                        p.colsum_partial[((size_t)(tile / p.tiles_n) * 4 + q) * p.N + col0 + cidx] = k1;
                }
                if (storer) bulk_wait_read0();        // the group's previous TMA stores have read its staging tile
                named_bar_sync(1 + grp, 64);
                // 64-byte rows, SWIZZLE_64B: 16-byte chunk j4 of row r goes to chunk j4 ^ ((r >> 1) & 3)
                float4* rhp = reinterpret_cast<float4*>(st_hi + rg * 16);
                float4* rlp = reinterpret_cast<float4*>(st_lo + rg * 16);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float4 h, l;
                    tf32_split(v[4 * j4 + 0], h.x, l.x); tf32_split(v[4 * j4 + 1], h.y, l.y);
                    tf32_split(v[4 * j4 + 2], h.z, l.z); tf32_split(v[4 * j4 + 3], h.w, l.w);
                    const int sw = j4 ^ ((rg >> 1) & 3);
                    rhp[sw] = h;
                    rlp[sw] = l;
                }
                fence_proxy_async_smem();
                named_bar_sync(1 + grp, 64);
                if (storer) {
                    tma_store_2d(&tmC_hi, st_hi, col0, m0 + rh * 64);
                    tma_store_2d(&tmC_lo, st_lo, col0, m0 + rh * 64);
                    bulk_commit();
                }
            }
        }
        if (storer) bulk_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------
// mn kernel: wgrad.  D[Kin(128-tile), Nout(BN-tile)] = sum over the CTA's m-range of X[m,:]^T dY[m,:]
// ---------------------------------------------------------------------------------------------
struct MNParams {
    int Kin, Nout;
    int64_t M;
    int tiles_k, tiles_n, splits;
    int64_t rows_per_split;   // multiple of TC_CHUNK_KB * TC_BK
    float* partial;           // [splits][Kin][Nout]
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_mn_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const MNParams p) {
    constexpr int B_TILE_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    constexpr int CPT = BN / 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_STAGES;
    uint64_t* tfull = bars + 2 * TC_STAGES;   // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x % (p.tiles_k * p.tiles_n);
    const int split = blockIdx.x / (p.tiles_k * p.tiles_n);
    const int kin0 = (tile / p.tiles_n) * TC_BM, n0 = (tile % p.tiles_n) * BN;
    const int64_t r_begin = (int64_t)split * p.rows_per_split;
    int64_t r_end = r_begin + p.rows_per_split;
    if (r_end > p.M) r_end = p.M;
    const int k_blocks = r_end > r_begin ? (int)((r_end - r_begin + TC_BK - 1) / TC_BK) : 0;
    const int chunks = (k_blocks + TC_CHUNK_KB - 1) / TC_CHUNK_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, TC_EPI_THREADS); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(empty + stage, phase ^ 1);
                unsigned char* st = smem + stage * STAGE_BYTES;
                mbar_expect_tx(full + stage, (uint32_t)STAGE_BYTES);
                const int r0 = (int)(r_begin + (int64_t)kb * TC_BK);
                // box {32 cols, 32 rows, 4 (or BN/32) column blocks}: rows beyond M are zero-filled
                tma_load_3d(st, &tmA_hi, 0, r0, kin0 / 32, full + stage);
                tma_load_3d(st + A_TILE_BYTES, &tmA_lo, 0, r0, kin0 / 32, full + stage);
                tma_load_3d(st + 2 * A_TILE_BYTES, &tmB_hi, 0, r0, n0 / 32, full + stage);
                tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tmB_lo, 0, r0, n0 / 32, full + stage);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TC_BM, BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            int kb = 0;
            for (uint32_t cc = 0; kb < k_blocks; ++cc) {
                const int acc = (int)(cc & 1u);
                mbar_wait(tempty + acc, ((cc >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                const int kb_end = (kb + TC_CHUNK_KB < k_blocks) ? kb + TC_CHUNK_KB : k_blocks;
                for (int kc = 0; kb < kb_end; ++kb, ++kc) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    // MN-major tf32 = SWIZZLE_128B_BASE32B: an atom is 4 k-rows x 128 B (32 columns).  The
                    // 32-column blocks are TC_BK rows * 128 B = 4096 B apart (LBO); consecutive 4-row atoms along
                    // k are 512 B apart (SBO); one MMA (K = 8) consumes two of them = 1024 B per k-step.
                    const uint64_t a_hi = make_desc(sa, 4096, 512, 1);
                    const uint64_t a_lo = make_desc(sa + A_TILE_BYTES, 4096, 512, 1);
                    const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES, 4096, 512, 1);
                    const uint64_t b_lo = make_desc(sa + 2 * A_TILE_BYTES + B_TILE_BYTES, 4096, 512, 1);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)(k * 64);   // 1024 B per k-step
                        umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc | k) != 0 ? 1u : 0u);
                        umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                        umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                    }
                    tc_commit(empty + stage);
                    if (kb == kb_end - 1) tc_commit(tfull + acc);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = kin0 + q * 32 + lane;
        float s[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) s[j] = 0.0f;
        for (uint32_t cc = 0; cc < (uint32_t)chunks; ++cc) {
            const int acc = (int)(cc & 1u);
            mbar_wait(tfull + acc, (cc >> 1) & 1u);
            tc_fence_after();
            drain_chunk<CPT>(tmem_base + (uint32_t)(acc * BN + half * CPT) + ((uint32_t)(q * 32) << 16), s);
            tc_fence_before();
            mbar_arrive(tempty + acc);
        }
        if (row < p.Kin) {
            float* out = p.partial + ((size_t)split * p.Kin + (size_t)row) * p.Nout;
            const int col0 = n0 + half * CPT;
            if (col0 + CPT <= p.Nout && (p.Nout & 3) == 0) {
#pragma unroll
                for (int j4 = 0; j4 < CPT / 4; ++j4)
                    reinterpret_cast<float4*>(out + col0)[j4] = make_float4(s[4 * j4], s[4 * j4 + 1], s[4 * j4 + 2], s[4 * j4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    if (col0 + j < p.Nout) out[col0 + j] = s[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
// x -> (hi, lo); hi may alias x (in-place split of a tensor produced by an fp32 FFMA kernel)
__global__ void __launch_bounds__(256)
split_kernel(const float* x, float* hi, float* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 h, l;
        tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}
__global__ void __launch_bounds__(256)
split_tail_kernel(const float* x, float* hi, float* __restrict__ lo, int64_t begin, int64_t n) {
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float h, l; tf32_split(x[i], h, l); hi[i] = h; lo[i] = l; }
}

// W[K][N] -> hi/lo of W and of W^T[N][K]
__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ W, float* __restrict__ W_hi, float* __restrict__ W_lo,
                   float* __restrict__ WT_hi, float* __restrict__ WT_lo, int K, int N) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        const float v = (k < K && n < N) ? W[(size_t)k * N + n] : 0.0f;
        tile[r][tx] = v;
        if (k < K && n < N) { float h, l; tf32_split(v, h, l); W_hi[(size_t)k * N + n] = h; W_lo[(size_t)k * N + n] = l; }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        if (n < N && k < K) {
            float h, l; tf32_split(tile[tx][r], h, l);
            WT_hi[(size_t)n * K + k] = h;
            WT_lo[(size_t)n * K + k] = l;
        }
    }
}

__global__ void __launch_bounds__(256)
tc_reduce_partials_kernel(const float* __restrict__ partial, int splits, int64_t stride, int64_t count, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < splits; ++z) s += partial[(size_t)z * stride + i];
        out[i] = s;
    }
}

// stage 1 of folding the dgrad epilogue's per-quarter column sums: block (x = column block, y = row chunk)
__global__ void __launch_bounds__(256)
colsum_fold_kernel(const float* __restrict__ part, int64_t rows, int N, int64_t rows_per_chunk, float* __restrict__ out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += part[r * N + n]; s1 += part[(r + 1) * N + n]; s2 += part[(r + 2) * N + n]; s3 += part[(r + 3) * N + n];
    }
    for (; r < r1; ++r) s0 += part[r * N + n];
    out[(size_t)blockIdx.y * N + n] = (s0 + s1) + (s2 + s3);
}

// column sums of dY_hi + dY_lo [M][N] (bias gradient): per-CTA partials over a row range
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dY, const float* __restrict__ dYlo, int64_t M, int N, int64_t rows_per_cta,
              float* __restrict__ partial) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    for (int n = threadIdx.x; n < N; n += 256) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int64_t r = r0;
        if (dYlo != nullptr) {
            for (; r + 1 < r1; r += 2) {
                s0 += dY[r * N + n]; s1 += dYlo[r * N + n]; s2 += dY[(r + 1) * N + n]; s3 += dYlo[(r + 1) * N + n];
            }
            for (; r < r1; ++r) { s0 += dY[r * N + n]; s1 += dYlo[r * N + n]; }
        } else {
            for (; r + 3 < r1; r += 4) {
                s0 += dY[r * N + n]; s1 += dY[(r + 1) * N + n]; s2 += dY[(r + 2) * N + n]; s3 += dY[(r + 3) * N + n];
            }
            for (; r < r1; ++r) s0 += dY[r * N + n];
        }
        partial[(size_t)blockIdx.x * N + n] = (s0 + s2) + (s1 + s3);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_encode() {
    if (g_encode) return PPO_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PPO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return PPO_ERR_CUDA;
    }
    g_encode = (EncodeTiledFn)fn;
    return PPO_OK;
}

// 2-D map over a row-major [rows][cols] fp32 matrix: box {32 cols, box_rows} with 128B swizzle (operand loads) or
// {16 cols, box_rows} with 64B swizzle (epilogue stores)
int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows, int box_cols = 32) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d %lld x %lld) failed: %d", (long long)rows, (long long)cols, (int)r); return PPO_ERR_CUDA; }
    return PPO_OK;
}
// 3-D view of a row-major [rows][cols] matrix as {32, rows, cols/32}: one box = `blocks` column blocks of
// TC_BK rows each, i.e. the canonical MN-major tf32 operand layout (128B swizzle with 32-byte atoms)
int make_map_mn(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int blocks) {
    cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)((cols + 31) / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)TC_BK, (cuuint32_t)blocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(mn %lld x %lld) failed: %d", (long long)rows, (long long)cols, (int)r); return PPO_ERR_CUDA; }
    return PPO_OK;
}

struct TcLayer {
    float* W_hi = nullptr;   // [K][N]  (dgrad B operand, K-major in N)
    float* W_lo = nullptr;
    float* WT_hi = nullptr;  // [N][K]  (forward B operand)
    float* WT_lo = nullptr;
};

struct TcState {
    int mode = 0;
    std::vector<TcLayer> layers;
    // hi/lo companions of the workspace tensors (allocated for `tokens`).  p->act[l] and p->dact[i]
    // hold the hi parts in this mode.
    int64_t tokens = 0;
    float* x_hi = nullptr;                 // input features of the current minibatch [M][dims[0]] (+ slack)
    float* x_lo = nullptr;
    std::vector<float*> act_lo;            // act_lo[l] for l = 1..L-1
    float* dact_lo[2] = {nullptr, nullptr};
    float* partial = nullptr;
    size_t partial_bytes = 0;
};

TcState* state(ppo_policy* p) { return reinterpret_cast<TcState*>(p->tc); }

int split(ppo_ctx* ctx, const float* x, float* hi, float* lo, int64_t n) {
    const int64_t n4 = n / 4;
    if (n4 > 0) {
        int64_t blocks = std::min<int64_t>(ceil_div(n4, 256), (int64_t)ctx->num_sms * 16);
        split_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, hi, lo, n4);
        ctx->launches += 1;
    }
    if (n4 * 4 < n) {
        split_tail_kernel<<<1, 256, 0, ctx->stream>>>(x, hi, lo, n4 * 4, n);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int wgrad_splits_tc(int64_t M, int tiles, int num_sms) {
    int s = std::max(1, num_sms / tiles);
    const int64_t max_s = std::max<int64_t>(1, M / (TC_BK * 8));
    if (s > max_s) s = (int)max_s;
    return s;
}

size_t tc_partial_bytes(int64_t tokens, int K, int N, int num_sms) {
    const int BN = N > 128 ? 256 : 128;
    const int tiles = (int)(ceil_div(K, TC_BM) * ceil_div(N, BN));
    const size_t a = (size_t)wgrad_splits_tc(tokens, tiles, num_sms) * K * N * 4;
    const size_t b = (size_t)num_sms * 4 * N * 4;   // stand-alone colsum partials
    const size_t c = ((size_t)4 * ceil_div(tokens, TC_BM) + 64) * (size_t)std::max(K, N) * 4;   // dgrad-epilogue column sums
    return std::max(a, std::max(b, c));
}

int ensure_tc_workspace(ppo_policy* p, int64_t tokens) {
    TcState* st = state(p);
    if (tokens <= st->tokens) return PPO_OK;
    ppo_ctx* ctx = p->ctx;
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    auto fr = [](float*& q) { if (q) cudaFree(q); q = nullptr; };
    fr(st->x_hi); fr(st->x_lo); fr(st->dact_lo[0]); fr(st->dact_lo[1]); fr(st->partial);
    for (auto& a : st->act_lo) fr(a);
    const int L = p->L;
    int hmax = 1;
    for (int l = 1; l < L; ++l) hmax = std::max(hmax, p->dims[l]);
    st->act_lo.assign(L + 1, nullptr);
    // +256 B of slack: the MN-major 3-D view reads whole 32-column blocks of the last row
    PPO_CUDA(cudaMalloc((void**)&st->x_hi, (size_t)tokens * p->dims[0] * 4 + 256));
    PPO_CUDA(cudaMalloc((void**)&st->x_lo, (size_t)tokens * p->dims[0] * 4 + 256));
    for (int l = 1; l < L; ++l) PPO_CUDA(cudaMalloc((void**)&st->act_lo[l], (size_t)tokens * p->dims[l] * 4));
    PPO_CUDA(cudaMalloc((void**)&st->dact_lo[0], (size_t)tokens * hmax * 4));
    PPO_CUDA(cudaMalloc((void**)&st->dact_lo[1], (size_t)tokens * hmax * 4));
    size_t pb = 16;
    for (int l = 0; l + 1 < L; ++l) pb = std::max(pb, tc_partial_bytes(tokens, p->dims[l], p->dims[l + 1], ctx->num_sms));
    PPO_CUDA(cudaMalloc((void**)&st->partial, pb));
    st->partial_bytes = pb;
    st->tokens = tokens;
    return PPO_OK;
}

template <int BN>
int launch_kk(ppo_ctx* ctx, const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, float* C_lo,
              int64_t M, int N, int K, const KKParams& base) {
    CUtensorMap mA, mAl, mB, mBl, mC, mCl;
    PPO_TRY(make_map_2d(&mA, A, M, K, TC_BM));
    PPO_TRY(make_map_2d(&mAl, A_lo, M, K, TC_BM));
    PPO_TRY(make_map_2d(&mB, B, N, K, BN));
    PPO_TRY(make_map_2d(&mBl, B_lo, N, K, BN));
    PPO_TRY(make_map_2d(&mC, C, M, N, 64, 16));      // store boxes: 64 rows x 16 columns, SWIZZLE_64B
    PPO_TRY(make_map_2d(&mCl, C_lo, M, N, 64, 16));
    KKParams p = base;
    p.M = (int)M; p.N = N; p.K = K;
    p.tiles_m = (int)ceil_div(M, TC_BM); p.tiles_n = (int)ceil_div(N, BN); p.k_blocks = (int)ceil_div(K, TC_BK);
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = std::min(tiles, ctx->num_sms);
    const size_t smem = KKSmem<BN>::TOTAL;
    PPO_CUDA(cudaFuncSetAttribute(tc_gemm_kk_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kk_kernel<BN><<<grid, TC_THREADS, smem, ctx->stream>>>(mA, mAl, mB, mBl, mC, mCl, p);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int kk_dispatch(ppo_ctx* ctx, const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, float* C_lo,
                int64_t M, int N, int K, const KKParams& base) {
    PPO_REQUIRE(M < ((int64_t)1 << 31), "tc gemm: M too large");
    PPO_REQUIRE(K % 4 == 0 && N % 16 == 0, "tc gemm: K %% 4 and N %% 16 required (K=%d N=%d)", K, N);
    if (N > 128) return launch_kk<256>(ctx, A, A_lo, B, B_lo, C, C_lo, M, N, K, base);
    return launch_kk<128>(ctx, A, A_lo, B, B_lo, C, C_lo, M, N, K, base);
}

template <int BN>
int launch_mn(ppo_ctx* ctx, const float* X, const float* X_lo, const float* dY, const float* dY_lo, int64_t M, int Kin,
              int Nout, float* partial, int splits) {
    CUtensorMap mA, mAl, mB, mBl;
    PPO_TRY(make_map_mn(&mA, X, M, Kin, TC_BM / 32));
    PPO_TRY(make_map_mn(&mAl, X_lo, M, Kin, TC_BM / 32));
    PPO_TRY(make_map_mn(&mB, dY, M, Nout, BN / 32));
    PPO_TRY(make_map_mn(&mBl, dY_lo, M, Nout, BN / 32));
    MNParams p;
    p.Kin = Kin; p.Nout = Nout; p.M = M;
    p.tiles_k = (int)ceil_div(Kin, TC_BM); p.tiles_n = (int)ceil_div(Nout, BN); p.splits = splits;
    p.rows_per_split = round_up(ceil_div(M, splits), TC_CHUNK_KB * TC_BK);
    p.partial = partial;
    const size_t smem = (size_t)TC_STAGES * (2 * A_TILE_BYTES + 2 * BN * TC_BK * 4) + 1024 + 256;
    PPO_CUDA(cudaFuncSetAttribute(tc_gemm_mn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.tiles_k * p.tiles_n * splits;
    tc_gemm_mn_kernel<BN><<<grid, TC_THREADS, smem, ctx->stream>>>(mA, mAl, mB, mBl, p);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// fold [rows][N] per-quarter column sums (rows = 4 * tiles_m) into out[N]; scratch holds 64 * N floats
int fold_colsum(ppo_ctx* ctx, const float* part, int64_t rows, int N, float* scratch, float* out) {
    const int chunks = (int)std::min<int64_t>(64, rows);
    const int64_t rpc = ceil_div(rows, chunks);
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)chunks);
    colsum_fold_kernel<<<grid, 256, 0, ctx->stream>>>(part, rows, N, rpc, scratch);
    tc_reduce_partials_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, ctx->stream>>>(scratch, chunks, N, N, out);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int wgrad_tc(ppo_ctx* ctx, const float* X, const float* X_lo, const float* dY, const float* dY_lo, float* dW, float* db,
             float* partial, size_t partial_bytes, int64_t M, int K, int N) {
    PPO_REQUIRE(K % 4 == 0 && N % 32 == 0, "tc wgrad: K %% 4 and N %% 32 required (K=%d N=%d)", K, N);
    const int BN = N > 128 ? 256 : 128;
    const int tiles = (int)(ceil_div(K, TC_BM) * ceil_div(N, BN));
    const int splits = wgrad_splits_tc(M, tiles, ctx->num_sms);
    PPO_REQUIRE(tc_partial_bytes(M, K, N, ctx->num_sms) <= partial_bytes, "tc wgrad: partial buffer too small");
    if (BN == 256) PPO_TRY(launch_mn<256>(ctx, X, X_lo, dY, dY_lo, M, K, N, partial, splits));
    else PPO_TRY(launch_mn<128>(ctx, X, X_lo, dY, dY_lo, M, K, N, partial, splits));
    const int64_t cnt = (int64_t)K * N;
    tc_reduce_partials_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, ctx->stream>>>(partial, splits, cnt, cnt, dW);
    ctx->launches += 1;
    if (db != nullptr) {
        const int ctas = (int)std::min<int64_t>((int64_t)ctx->num_sms * 4, ceil_div(M, 64));
        const int64_t rows = ceil_div(M, ctas);
        colsum_kernel<<<ctas, 256, 0, ctx->stream>>>(dY, dY_lo, M, N, rows, partial);
        tc_reduce_partials_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, ctx->stream>>>(partial, ctas, N, N, db);
        ctx->launches += 2;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// interface used by abi.cu
// ---------------------------------------------------------------------------------------------
int tc_prepare(ppo_policy* p, int mode) {
    if (mode == PPO_GEMM_BF16_TC) {
        set_error("PPO_GEMM_BF16_TC is not built in this round (a single bf16 pass cannot meet the 1e-5 parity bound); "
                  "use PPO_GEMM_TF32X3_TC");
        return PPO_ERR_STATE;
    }
    PPO_TRY(load_encode());
    // every hidden layer must be expressible on the tensor-core kernels; otherwise refuse loudly
    for (int l = 0; l + 1 < p->L; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        PPO_REQUIRE(K % 4 == 0 && N % 32 == 0 && (l == 0 || K % 32 == 0),
                    "tensor-core engine: layer %d (%d -> %d) needs in %% 4 == 0 and hidden widths %% 32 == 0; "
                    "use PPO_GEMM_FP32_SIMT for this policy", l, K, N);
    }
    if (p->tc == nullptr) p->tc = new TcState();
    TcState* st = state(p);
    st->mode = mode;
    if (st->layers.empty()) {
        st->layers.resize(p->L);
        for (int l = 0; l + 1 < p->L; ++l) {      // hidden layers only; the head has its own streaming kernels
            const size_t n = (size_t)p->dims[l] * p->dims[l + 1] * 4;
            TcLayer& ly = st->layers[l];
            PPO_CUDA(cudaMalloc((void**)&ly.W_hi, n));
            PPO_CUDA(cudaMalloc((void**)&ly.W_lo, n));
            PPO_CUDA(cudaMalloc((void**)&ly.WT_hi, n));
            PPO_CUDA(cudaMalloc((void**)&ly.WT_lo, n));
        }
    }
    return PPO_OK;
}

int tc_refresh_weights(ppo_policy* p) {
    TcState* st = state(p);
    PPO_REQUIRE(st != nullptr, "tensor-core engine not prepared");
    ppo_ctx* ctx = p->ctx;
    for (int l = 0; l + 1 < p->L; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        TcLayer& ly = st->layers[l];
        dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(K, 32));
        weight_prep_kernel<<<grid, 256, 0, ctx->stream>>>(p->params + p->w_off[l], ly.W_hi, ly.W_lo, ly.WT_hi, ly.WT_lo, K, N);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

const float* tc_act_lo(ppo_policy* p, int l) {
    TcState* st = state(p);
    if (st == nullptr || l < 1 || l >= (int)st->act_lo.size()) return nullptr;
    return st->act_lo[l];
}

int tc_linear_fwd(ppo_policy* p, int l, const float* X, float* Y, int64_t M) {
    TcState* st = state(p);
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(ensure_tc_workspace(p, p->ws_tokens > M ? p->ws_tokens : M));
    const int K = p->dims[l], N = p->dims[l + 1];
    TcLayer& ly = st->layers[l];
    const float *A_hi, *A_lo;
    if (l == 0) {
        PPO_TRY(split(ctx, X, st->x_hi, st->x_lo, M * K));
        A_hi = st->x_hi; A_lo = st->x_lo;
    } else {
        A_hi = X; A_lo = st->act_lo[l];          // X == p->act[l] holds the hi part
    }
    KKParams kp{};
    kp.epi = TC_EPI_FWD; kp.act = 1; kp.slope = p->slope; kp.bias = p->params + p->b_off[l]; kp.gate = nullptr;
    return kk_dispatch(ctx, A_hi, A_lo, ly.WT_hi, ly.WT_lo, Y, st->act_lo[l + 1], M, N, K, kp);
}

int tc_linear_bwd(ppo_policy* p, int l, const float* X, const float* dY, float* dX, float* dW, float* db_below, int64_t M) {
    // dY arrives as a tf32 hi/lo pair (written by head_bwd or by the dgrad epilogue of the layer above), and its
    // column sums (this layer's bias gradient) were already produced by that same kernel.
    TcState* st = state(p);
    ppo_ctx* ctx = p->ctx;
    PPO_TRY(ensure_tc_workspace(p, p->ws_tokens > M ? p->ws_tokens : M));
    const int K = p->dims[l], N = p->dims[l + 1];
    TcLayer& ly = st->layers[l];
    PPO_REQUIRE(dY == p->dact[0] || dY == p->dact[1], "tc_linear_bwd: unexpected gradient buffer");
    const float* dY_lo = (dY == p->dact[0]) ? st->dact_lo[0] : st->dact_lo[1];
    const float* X_hi = (l == 0) ? st->x_hi : X;
    const float* X_lo = (l == 0) ? st->x_lo : st->act_lo[l];
    PPO_TRY(wgrad_tc(ctx, X_hi, X_lo, dY, dY_lo, dW, nullptr, st->partial, st->partial_bytes, M, K, N));
    if (dX != nullptr) {
        float* dX_lo = (dX == p->dact[0]) ? st->dact_lo[0] : st->dact_lo[1];
        KKParams kp{};
        kp.epi = TC_EPI_DGRAD; kp.act = 0; kp.slope = p->slope; kp.bias = nullptr; kp.gate = X_hi;
        kp.colsum_partial = (db_below != nullptr) ? st->partial : nullptr;
        // dX[M, K] = dY[M, N] * W[K, N]^T : A = dY (K-major in N), B = W rows (K-major in N)
        PPO_TRY(kk_dispatch(ctx, dY, dY_lo, ly.W_hi, ly.W_lo, dX, dX_lo, M, K, N, kp));
        if (db_below != nullptr) {
            const int64_t rows = 4 * ceil_div(M, TC_BM);
            PPO_TRY(fold_colsum(ctx, st->partial, rows, K, st->partial + (size_t)rows * K, db_below));
        }
    }
    return PPO_OK;
}

float* tc_dact_lo(ppo_policy* p, const float* dact) {
    TcState* st = state(p);
    if (st == nullptr) return nullptr;
    return (dact == p->dact[0]) ? st->dact_lo[0] : st->dact_lo[1];
}

void tc_destroy(ppo_policy* p) {
    TcState* st = state(p);
    if (!st) return;
    auto fr = [](float*& q) { if (q) cudaFree(q); q = nullptr; };
    for (auto& ly : st->layers) { fr(ly.W_hi); fr(ly.W_lo); fr(ly.WT_hi); fr(ly.WT_lo); }
    fr(st->x_hi); fr(st->x_lo); fr(st->dact_lo[0]); fr(st->dact_lo[1]); fr(st->partial);
    for (auto& a : st->act_lo) fr(a);
    delete st;
    p->tc = nullptr;
}

// ---------------------------------------------------------------------------------------------
// stand-alone GEMM entry points for tests and per-kernel benches (device pointers, hi/lo pairs)
// ---------------------------------------------------------------------------------------------
int tc_test_fwd(ppo_ctx* ctx, const float* X_hi, const float* X_lo, const float* WT_hi, const float* WT_lo, const float* bias,
                float* Y_hi, float* Y_lo, int64_t M, int K, int N, int act, float slope) {
    PPO_TRY(load_encode());
    KKParams kp{};
    kp.epi = TC_EPI_FWD; kp.act = act; kp.slope = slope; kp.bias = bias;
    return kk_dispatch(ctx, X_hi, X_lo, WT_hi, WT_lo, Y_hi, Y_lo, M, N, K, kp);
}
int tc_test_dgrad(ppo_ctx* ctx, const float* dY_hi, const float* dY_lo, const float* W_hi, const float* W_lo, const float* gate,
                  float* dX_hi, float* dX_lo, int64_t M, int K, int N, float slope, float* colsum_scratch, float* colsum_out) {
    PPO_TRY(load_encode());
    KKParams kp{};
    kp.epi = TC_EPI_DGRAD; kp.slope = slope; kp.gate = gate;
    kp.colsum_partial = colsum_out ? colsum_scratch : nullptr;
    PPO_TRY(kk_dispatch(ctx, dY_hi, dY_lo, W_hi, W_lo, dX_hi, dX_lo, M, K, N, kp));
    if (colsum_out) {
        const int64_t rows = 4 * ceil_div(M, TC_BM);
        PPO_TRY(fold_colsum(ctx, colsum_scratch, rows, K, colsum_scratch + (size_t)rows * K, colsum_out));
    }
    return PPO_OK;
}
int tc_test_wgrad(ppo_ctx* ctx, const float* X_hi, const float* X_lo, const float* dY_hi, const float* dY_lo, float* dW,
                  float* db, float* partial, size_t partial_bytes, int64_t M, int K, int N) {
    PPO_TRY(load_encode());
    return wgrad_tc(ctx, X_hi, X_lo, dY_hi, dY_lo, dW, db, partial, partial_bytes, M, K, N);
}
int tc_test_split(ppo_ctx* ctx, const float* x, float* hi, float* lo, int64_t n) { return split(ctx, x, hi, lo, n); }
int tc_test_weight_prep(ppo_ctx* ctx, const float* W, float* W_hi, float* W_lo, float* WT_hi, float* WT_lo, int K, int N) {
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(K, 32));
    weight_prep_kernel<<<grid, 256, 0, ctx->stream>>>(W, W_hi, W_lo, WT_hi, WT_lo, K, N);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
size_t tc_test_partial_bytes(ppo_ctx* ctx, int64_t M, int K, int N) { return tc_partial_bytes(M, K, N, ctx->num_sms); }

}  // namespace ppo
