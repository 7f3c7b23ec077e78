// gemm_f16.cu — K5/K7 engine PPO_GEMM_F16X3_TC: fp32-grade GEMMs on tcgen05.mma.kind::f16.
//
// The reference's policy MLP (test/policy.jl:9-31) is Float32 end to end and parity is stated at 1e-5.
// kind::f16 runs at twice the kind::tf32 rate and its operands take half the bytes, so this engine keeps
// every GEMM operand as an fp16 PAIR of a power-of-two-scaled tensor:
//        T~ = T * 2^e,     hi = fp16_rn(T~),     lo = fp16_rn(T~ - hi)           (22 significant bits)
//        A B ~= (A_hi B_hi + A_hi B_lo + A_lo B_hi) * 2^-(eA + eB)              (error ~2^-22 per product)
// i.e. the same error-compensated 3-pass scheme as the tf32 engine (gemm_tc.cu) at twice the tensor rate
// and HALF the HBM/L2 bytes (a pair costs 4 B/element, like the plain fp32 tensor).
//
// fp16 has a 5-bit exponent, so each tensor carries its own exponent e, chosen ON THE DEVICE from a
// guaranteed upper bound B >= max|T| such that B * 2^e <= 2^14 (no overflow, ever):
//   * inputs / weights / dlogits: exact abs-max (absmax_kernel, wstats_kernel);
//   * forward activations:  B(act[l+1]) = B(act[l]) * max_n sum_k |W[k][n]| + max|b|     (plan_fwd_kernel)
//   * backward activations: B(dZ[l-1])  = B(dZ[l])  * max_k sum_n |W[k][n]|              (plan_bwd_kernel)
// The bounds are loose by ~2^4 per layer, which costs nothing: fp16 subnormals keep the absolute error of a
// scaled element <= 2^-25, i.e. <= 2^-25 / 2^14-looseness of the tensor's max (tests hold 1e-5 = 2^-16.6).
// Scales are exact powers of two, so scaling/unscaling introduces no rounding.
//
// Kernels (all sm_100a, hand-written):
//   f16_gemm_kk_kernel  forward / dgrad: persistent, warp-specialised (TMA warp, one MMA thread, 16 epilogue
//                       warps), 64-element k-blocks (128-byte swizzle rows), contraction cut into 256-element
//                       chains that the epilogue warps fold into fp32 registers (the tensor core accumulates
//                       with round-toward-zero; the fold compensates the resulting deficit), epilogue = unscale
//                       + bias + leakyrelu + sign bits (or sign-bit gate + bias-gradient column sums), rescale,
//                       packed fp16 split, per-warp swizzled staging, per-warp TMA stores.  K > 128: CTA PAIRS
//                       (2-CTA clusters, tcgen05.mma.cta_group::2, UMMA 256 x 256 x 16, each CTA stages half of
//                       the weight tile); K <= 128: 128 x 128 tiles with 4 TMEM stages (epilogue-bound shapes).
//   f16_gemm_mn_kernel  wgrad: both operands MN-major straight from the row-major activations (3-D tensor
//                       maps produce the canonical MN-major SWIZZLE_128B atoms), split over rows, fixed-order
//                       reduction (deterministic).
//   head_fwd16 / head_bwd16  the Dense(H, apa) head (N = apa <= 4 is too narrow for a UMMA tile): streaming
//                       kernels over the fp16 pairs.
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>

#include "gemm_f16.cuh"
#include "tc_ptx.cuh"

namespace ppo {

namespace {

constexpr int F_BM = 128;
constexpr int F_BK = 64;                    // 64 fp16 = one 128-byte swizzle row
constexpr int F_STAGES = 2;
// mn kernel: k-blocks per TMEM accumulation chain (128 elements = 24 MMAs).  256-element chains were measured at 1 % faster
// (1437 vs 1449 us sustained at M = 2^20, K = N = 512: the kernel is not drain-bound) and are not worth the accuracy margin.
constexpr int F_CHUNK_KB = 2;
// kk kernel: 256 elements = 48 MMAs per chain.  The two TMEM stages then hold a whole K = 512 tile of look-ahead, so
// the MMA issuer keeps running while the epilogue warps finish the previous tile (with 128-element chains the tensor
// pipe idled ~45 % of the time waiting for them: profiles/r01_f16_ncu.md).  The longer chain's round-toward-zero
// deficit is removed by the compensated fold (KK16Params::rz_comp): measured gradient error vs fp64 4-5e-6 of each
// tensor's max-abs (fp32 FFMA engine: 2-4e-6; 128-element chains without compensation: 5-6e-6).
constexpr int F_KK_CHUNK_KB = 4;
// calibrated on the C3-width policy gradient (scripts/f16_matrix.sh): the scale bias of the gradient against the fp64
// oracle crosses zero at 1.3e-8 .. 1.7e-8 per MMA for 128- and 256-element chains alike (theory for an fp32 adder that
// truncates every add: 0.72 * 2^-24 / 2 = 2.1e-8)
constexpr float F_RZ_COMP = 1.5e-8f;
constexpr int F_EPI_THREADS = 256;          // mn kernel: 8 epilogue warps = 4 lane quarters x 2 column halves
constexpr int F_THREADS = 64 + F_EPI_THREADS;
constexpr int F_KK_EPI_THREADS = 512;       // kk kernel: 16 epilogue warps = 4 lane quarters x 4 column quarters (the
constexpr int F_KK_THREADS = 64 + F_KK_EPI_THREADS;   // per-element epilogue is issue/latency bound: more warps, fewer columns each)
constexpr int FA_TILE_BYTES = F_BM * F_BK * 2;   // 16 KB
constexpr int F_MAX_LAYERS = 8;              // table size
// Depth contract.  A tensor's exponent comes from an a-priori bound (plan_fwd / plan_bwd) that runs ahead of the real
// abs-max by ~0.87 sqrt(K) per layer (2^4.3 at K = 512, 2^4.8 at K = 1024: max column L1 norm of a Glorot matrix against
// the ~unit gain of the layer).  A scaled element keeps an absolute error <= 2^-25 even when the looseness pushes it into
// fp16's subnormal range, i.e. <= looseness * 2^-39 of the tensor's max; 1e-5 = 2^-16.6 therefore allows a total
// looseness of 2^22: four hidden layers of width <= 1024 (4 x 4.8 = 19.2 bits) are inside, deeper policies are refused
// (PPO_GEMM_AUTO then picks the tf32 engine, whose fp32-range operands need no scaling).
constexpr int F_MAX_HIDDEN = 4;
constexpr int F_MAX_WIDTH = 1024;
constexpr int F_EXP_TARGET = 14;            // bound * 2^e <= 2^14 (fp16 max is ~2^16)
constexpr int F_EXP_CLAMP = 60;

enum { F_EPI_FWD = 0, F_EPI_DGRAD = 1 };

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// hi = fp16_rn(a), lo = fp16_rn(a - hi); a is already scaled into fp16 range
__device__ __forceinline__ void f16_split(float a, __half& hi, __half& lo) {
    hi = __float2half_rn(a);
    lo = __float2half_rn(a - __half2float(hi));
}
// two elements at once (cvt.rn.f16x2.f32): returns the packed hi pair and lo pair
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void unpack8(const uint4& q, float* x) {
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        x[2 * i] = f.x; x[2 * i + 1] = f.y;
    }
}

// sign-bit words of an activation [M][N]: tiled so that the 32 rows a warp owns are contiguous (coalesced 128-byte
// stores / loads from the row-per-thread epilogues): word (row, w) lives at ((row / 128) * (N / 32) + w) * 128 + row % 128
__host__ __device__ inline size_t sign_index(int64_t row, int w, int nwords) {
    return ((size_t)(row >> 7) * nwords + w) * 128 + (size_t)(row & 127);
}
inline size_t sign_words(int64_t rows, int N) { return (size_t)ceil_div(rows, 128) * 128 * (N / 32); }

struct KK16Params {
    int M, N, K;             // M: rows the host sized grids / tensor maps for (>= the rows actually in use)
    // rows actually in use, read on the device (token compaction: the number of active tokens of a minibatch is only known
    // there); nullptr = M.  Tiles past that count are not computed; the rows between it and the end of its tile are
    // (dgrad writes zeros there and keeps them out of the column sums).
    const int* M_dev;
    int tiles_m, tiles_n, k_blocks;
    int epi;
    int act;                 // fwd: apply leakyrelu
    float slope;
    const float* bias;       // fwd
    const uint32_t* gate;    // dgrad: sign bits of the activation that gates the gradient (bit = act > 0), sign_index layout
    uint32_t* signs_out;     // fwd: sign bits of the output activation (rows padded to a multiple of 128)
    float* colsum_partial;   // dgrad: [tiles_m][4][N] column sums of the (unscaled) output per 32-row quarter
    const float* sc_a;       // {scale, 1/scale} of the A operand
    const float* sc_b;
    const float* sc_c;       // of the output
    int store3d;             // the output's hi and lo planes are stored by ONE 3-D TMA box per 16-column group (tmC_hi is the
                             // 3-D map {16 cols, 32 rows, 2 planes}); 0: two 2-D stores
    int chunk_kb;            // k-blocks per TMEM accumulation chain
    // The tensor core adds each MMA result into its fp32 accumulator with round-toward-zero: every add loses on average
    // ~0.72 * 2^-24 of the running sum, always towards zero, so a chain of n MMAs comes out short by ~ n/2 of that (the
    // running sum grows along the chain).  The fold multiplies each finished chain by 1 + rz_comp * n to remove this
    // systematic part (the random part stays); 0 disables it.
    float rz_comp;
};

// TWO: a CTA pair (one cluster) shares one UMMA of M = 256 (cta_group::2): each CTA stages its own 128 rows of A and
// HALF of the B tile, so the operand bytes per SM and MMA cycle drop from 64 to 43 B/clk (the L2 -> SM feed is what
// bounds the single-CTA kernel at K = 512)
template <int BN, bool TWO = false>
struct KK16Smem {
    static constexpr int B_ROWS = TWO ? BN / 2 : BN;    // rows of the B tile this CTA stages
    static constexpr int B_TILE_BYTES = B_ROWS * F_BK * 2;
    static constexpr int STAGE_BYTES = 2 * FA_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int NST = (BN > 128 && !TWO) ? 2 : 3;   // operand pipeline stages (96 KB / 64 KB each)
    static constexpr int NACC = BN > 128 ? 2 : 4;            // TMEM accumulator stages (all 512 columns in use)
    // every epilogue warp stages its own 32 rows x 16 columns of hi and of lo (2 x 1 KB) and issues its own TMA
    // stores: no cross-warp barrier in the epilogue, and the hi / lo buffers alternate so that a store's smem read
    // overlaps the conversion of the other half
    static constexpr int STAGING_BYTES = 16 * 2 * 32 * 16 * 2;
    static constexpr int TOTAL = NST * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// 16 columns per TMEM load: the 16-warp kk epilogue runs at 96 registers per thread
// `comp` compensates the tensor core's round-toward-zero accumulation (see KK16Params::rz_comp)
template <int CPT>
__device__ __forceinline__ void drain_chunk16_narrow(uint32_t taddr, float* s, float comp) {
#pragma unroll
    for (int c = 0; c < CPT / 16; ++c) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)(c * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) s[c * 16 + j] = fmaf(v[j], comp, s[c * 16 + j]);
    }
}

template <int CPT>
__device__ __forceinline__ void drain_chunk16(uint32_t taddr, float* s, float comp) {
#pragma unroll
    for (int c4 = 0; c4 < CPT / 32; ++c4) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)(c4 * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[c4 * 32 + j] = fmaf(v[j], comp, s[c4 * 32 + j]);
    }
}

// column sums of a warp's 32 rows x 16 columns (one column set per thread register) with a halving
// butterfly (16 shuffles); lanes with (lane & 1) == 0 end up holding the sum of column cidx
__device__ __forceinline__ void colsum16(const float* v, int lane, float& out, int& cidx) {
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
    out = k1;
    cidx = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
}

template <int BN, bool TWO>
__global__ void __launch_bounds__(F_KK_THREADS, 1)
f16_gemm_kk_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const __grid_constant__ CUtensorMap tmC_hi, const __grid_constant__ CUtensorMap tmC_lo,
                   const KK16Params p) {
    using S = KK16Smem<BN, TWO>;
    constexpr int CPT = BN / 4;   // columns per epilogue thread
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NST = S::NST, NACC = S::NACC;
    unsigned char* staging = smem + NST * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES);
    uint64_t* full = bars;                    // [NST]
    uint64_t* empty = bars + NST;             // [NST]
    uint64_t* tfull = bars + 2 * NST;         // [NACC]
    uint64_t* tempty = tfull + NACC;          // [NACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work units: single-CTA mode = 128 x BN tiles over the grid; pair mode = 256 x BN tiles over the clusters, CTA rank r
    // of the pair owning rows [128 r, 128 r + 128) of the unit (an odd last m-tile leaves rank 1 a fully out-of-range
    // tile: TMA zero-fills its loads and clips its stores)
    const int rank = TWO ? (int)cluster_ctarank() : 0;
    const int unit0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // rows in use (token compaction: known only on the device).  Rows between that count and the end of the last tile are
    // computed like any other: the producers keep them finite (zeros / bounded stale values, see split16_rows_kernel), and
    // every consumer either stops at the row count or multiplies them by exact zeros.
    const int M_rows = p.M_dev != nullptr ? min(__ldg(p.M_dev), p.M) : p.M;
    const int tiles_m = (M_rows + F_BM - 1) / F_BM;
    const int num_tiles = (TWO ? (tiles_m + 1) / 2 : tiles_m) * p.tiles_n;
    const int chunks_per_tile = (p.k_blocks + p.chunk_kb - 1) / p.chunk_kb;
    auto m_tile_of = [&](int unit) -> int { return TWO ? 2 * (unit / p.tiles_n) + rank : unit / p.tiles_n; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        // pair mode: one arrival per epilogue WARP of both CTAs, all on the leader's barrier
        for (int a = 0; a < NACC; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, TWO ? 2 * (F_KK_EPI_THREADS / 32) : F_KK_EPI_THREADS); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        tma_prefetch_desc(&tmC_hi); tma_prefetch_desc(&tmC_lo);
    }
    if (warp == 1) { if (TWO) tmem_alloc_2sm(tmem_slot, NACC * BN); else tmem_alloc(tmem_slot, NACC * BN); }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_step) {
                const int m0 = m_tile_of(tile) * F_BM, n0 = (tile % p.tiles_n) * BN;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    unsigned char* st = smem + stage * S::STAGE_BYTES;
                    const int k0 = kb * F_BK;
                    if (TWO) {
                        // the leader's barrier counts the bytes of both CTAs (its own expect_tx may come after the
                        // peer's first complete_tx: the phase cannot end before the leader's arrival)
                        if (rank == 0) mbar_expect_tx(full + stage, 2u * (uint32_t)S::STAGE_BYTES);
                        const int nb0 = n0 + rank * (BN / 2);
                        tma_load_2d_2sm(st, &tmA_hi, k0, m0, full + stage);
                        tma_load_2d_2sm(st + FA_TILE_BYTES, &tmA_lo, k0, m0, full + stage);
                        tma_load_2d_2sm(st + 2 * FA_TILE_BYTES, &tmB_hi, k0, nb0, full + stage);
                        tma_load_2d_2sm(st + 2 * FA_TILE_BYTES + S::B_TILE_BYTES, &tmB_lo, k0, nb0, full + stage);
                    } else {
                        mbar_expect_tx(full + stage, (uint32_t)S::STAGE_BYTES);
                        tma_load_2d(st, &tmA_hi, k0, m0, full + stage);
                        tma_load_2d(st + FA_TILE_BYTES, &tmA_lo, k0, m0, full + stage);
                        tma_load_2d(st + 2 * FA_TILE_BYTES, &tmB_hi, k0, n0, full + stage);
                        tma_load_2d(st + 2 * FA_TILE_BYTES + S::B_TILE_BYTES, &tmB_lo, k0, n0, full + stage);
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread; in pair mode the leader CTA's) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(TWO ? 2 * F_BM : F_BM, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            uint32_t cc = 0;   // accumulation chains issued so far
            for (int tile = unit0; tile < num_tiles; tile += unit_step) {
                for (int kb = 0; kb < p.k_blocks; ++cc) {
                    const int acc = (int)(cc % (uint32_t)NACC);
                    mbar_wait(tempty + acc, ((cc / (uint32_t)NACC) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    const int kb_end = kb + p.chunk_kb < p.k_blocks ? kb + p.chunk_kb : p.k_blocks;
                    for (int kc = 0; kb < kb_end; ++kb, ++kc) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                        const uint64_t a_hi = make_desc(sa, 16, 1024);
                        const uint64_t a_lo = make_desc(sa + FA_TILE_BYTES, 16, 1024);
                        const uint64_t b_hi = make_desc(sa + 2 * FA_TILE_BYTES, 16, 1024);
                        const uint64_t b_lo = make_desc(sa + 2 * FA_TILE_BYTES + S::B_TILE_BYTES, 16, 1024);
#pragma unroll
                        for (int k = 0; k < F_BK / 16; ++k) {
                            const uint64_t koff = (uint64_t)(k * 2);   // 16 fp16 = 32 bytes = 2 x 16 B
                            if (TWO) {
                                umma_f16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc | k) != 0 ? 1u : 0u);
                                umma_f16_2sm(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                                umma_f16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                            } else {
                                umma_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc | k) != 0 ? 1u : 0u);
                                umma_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                                umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                            }
                        }
                        if (TWO) {       // both CTAs' producers / epilogues are released by the same commit
                            tc_commit_2sm(empty + stage, (uint16_t)3);
                            if (kb == kb_end - 1) tc_commit_2sm(tfull + acc, (uint16_t)3);
                        } else {
                            tc_commit(empty + stage);            // frees the smem slot when these MMAs retire
                            if (kb == kb_end - 1) tc_commit(tfull + acc);
                        }
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: 16 warps = 4 lane quarters x 4 column quarters =====================
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int cq = (warp - 2) >> 2;                // which BN/4 columns
        const int row_in_tile = q * 32 + lane;
        uint4* st_hi = reinterpret_cast<uint4*>(staging + (warp - 2) * 2048);
        uint4* st_lo = st_hi + 32 * 2;                 // + 1024 B
        const int swz = (lane >> 2) & 1;               // SWIZZLE_32B: 16-byte chunk c of row r goes to chunk c ^ ((r >> 2) & 1)
        // every scale is a power of two, so folding them into one another changes no rounding: the accumulator is folded
        // straight into the OUTPUT's scaled domain (x 2^-(eA + eB) x 2^eC), the bias joins it pre-scaled
        const float cscale = __ldg(p.sc_c);
        const float uc = __ldg(p.sc_a + 1) * __ldg(p.sc_b + 1) * cscale;
        const float inv_cscale = __ldg(p.sc_c + 1);
        // leakyrelu(x) = max(x, slope x) for 0 <= slope <= 1 (one FMNMX instead of a compare-select); act = 0: identity
        const float slope_eff = p.act ? p.slope : 1.0f;
        const bool slope_unit = slope_eff >= 0.0f && slope_eff <= 1.0f;
        uint32_t cc = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_step) {
            const int mt = m_tile_of(tile);
            const int m0 = mt * F_BM, n0 = (tile % p.tiles_n) * BN;
            const int row = m0 + row_in_tile;
            const bool mt_valid = mt < tiles_m;
            float s[CPT];
            if (p.epi == F_EPI_FWD && p.bias != nullptr) {
                // the accumulation starts from the (pre-scaled) bias, fetched before the contraction finishes
#pragma unroll
                for (int g = 0; g < CPT / 32; ++g) {
                    const int col0 = n0 + cq * CPT + g * 32;
                    if (col0 + 32 <= p.N) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j4);
                            s[g * 32 + 4 * j4 + 0] = b4.x * cscale; s[g * 32 + 4 * j4 + 1] = b4.y * cscale;
                            s[g * 32 + 4 * j4 + 2] = b4.z * cscale; s[g * 32 + 4 * j4 + 3] = b4.w * cscale;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) s[g * 32 + j] = (col0 + j < p.N) ? __ldg(p.bias + col0 + j) * cscale : 0.0f;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j) s[j] = 0.0f;
            }
            // dgrad: rows beyond the data contribute zeros (also to the column sums)
            const float row_uc = (p.epi == F_EPI_DGRAD && row >= M_rows) ? 0.0f : uc;
            // dgrad: this thread's gate bits (one word per 32 columns), fetched before the contraction finishes
            uint32_t gbits[CPT / 32];
            if (p.epi == F_EPI_DGRAD) {
#pragma unroll
                for (int g = 0; g < CPT / 32; ++g) {
                    const int col0 = n0 + cq * CPT + g * 32;
                    gbits[g] = (row < M_rows && col0 < p.N) ? __ldg(p.gate + sign_index(row, col0 >> 5, p.N >> 5)) : 0u;
                }
            }
            for (int ch = 0; ch < chunks_per_tile; ++ch, ++cc) {
                const int acc = (int)(cc % (uint32_t)NACC);
                mbar_wait(tfull + acc, (cc / (uint32_t)NACC) & 1u);
                tc_fence_after();
                const int kb_in_chunk = (ch + 1) * p.chunk_kb <= p.k_blocks ? p.chunk_kb : p.k_blocks - ch * p.chunk_kb;
                drain_chunk16_narrow<CPT>(tmem_base + (uint32_t)(acc * BN + cq * CPT) + ((uint32_t)(q * 32) << 16), s,
                                          (1.0f + p.rz_comp * (float)(3 * (F_BK / 16) * kb_in_chunk)) * row_uc);
                tc_fence_before();
                if (TWO) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty + acc);
                } else {
                    mbar_arrive(tempty + acc);
                }
            }
            // ---- unscale, bias / activation (or gradient gate), rescale, fp16 split, staged TMA store (16 columns at a time)
#pragma unroll
            for (int g = 0; g < CPT / 32; ++g) {
                const int col0 = n0 + cq * CPT + g * 32;
                float* v = s + g * 32;
                if (p.epi == F_EPI_FWD) {
                    // v = scaled pre-activation (bias included): sign bit, leakyrelu
                    uint32_t w = 0u;
                    if (slope_unit && p.signs_out == nullptr) {      // (the last hidden layer: the head gates on sign(hi))
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], slope_eff * v[j]);
                    } else if (slope_unit) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = v[j];
                            asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, 0f00000000;\n@p or.b32 %0, %0, %2;\n}" : "+r"(w) : "f"(x), "r"(1u << j));
                            v[j] = fmaxf(x, slope_eff * x);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = v[j];
                            const bool pos = x > 0.0f;
                            w |= (pos ? 1u : 0u) << j;
                            v[j] = pos ? x : slope_eff * x;
                        }
                    }
                    if (p.signs_out != nullptr && col0 < p.N && mt_valid) p.signs_out[sign_index(row, col0 >> 5, p.N >> 5)] = w;
                } else {
                    // v = scaled dY W^T: the gate picks 1 or the slope
                    const uint32_t w = gbits[g];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : v[j] * p.slope;
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float* vv = v + 16 * hh;
                    if (p.colsum_partial != nullptr) {
                        // = the bias gradient of the layer below, fused here so that dX is never re-read for it
                        float cs; int cidx;
                        colsum16(vv, lane, cs, cidx);
                        if ((lane & 1) == 0 && col0 + 16 * hh + cidx < p.N && mt_valid)
                            p.colsum_partial[((size_t)mt * 4 + q) * p.N + col0 + 16 * hh + cidx] = cs * inv_cscale;
                    }
                    uint4 h[2], l[2];
#pragma unroll
                    for (int j8 = 0; j8 < 2; ++j8) {
                        f16_split2(vv[8 * j8 + 0], vv[8 * j8 + 1], h[j8].x, l[j8].x);
                        f16_split2(vv[8 * j8 + 2], vv[8 * j8 + 3], h[j8].y, l[j8].y);
                        f16_split2(vv[8 * j8 + 4], vv[8 * j8 + 5], h[j8].z, l[j8].z);
                        f16_split2(vv[8 * j8 + 6], vv[8 * j8 + 7], h[j8].w, l[j8].w);
                    }
                    if (p.store3d) {
                        // one 3-D box {16 columns, 32 rows, hi | lo plane} per group: half as many TMA operations; the
                        // single staging pair is free again long before the next group's maths is done
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                        st_hi[lane * 2 + (0 ^ swz)] = h[0];
                        st_hi[lane * 2 + (1 ^ swz)] = h[1];
                        st_lo[lane * 2 + (0 ^ swz)] = l[0];
                        st_lo[lane * 2 + (1 ^ swz)] = l[1];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&tmC_hi, st_hi, col0 + 16 * hh, m0 + q * 32, 0); bulk_commit(); }
                    } else {
                    // hi: wait until the previous hi store has read its buffer (the lo store issued after it may still run)
                    if (lane == 0) bulk_wait_read1();
                    __syncwarp();
                    st_hi[lane * 2 + (0 ^ swz)] = h[0];
                    st_hi[lane * 2 + (1 ^ swz)] = h[1];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmC_hi, st_hi, col0 + 16 * hh, m0 + q * 32); bulk_commit(); }
                    if (lane == 0) bulk_wait_read1();
                    __syncwarp();
                    st_lo[lane * 2 + (0 ^ swz)] = l[0];
                    st_lo[lane * 2 + (1 ^ swz)] = l[1];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmC_lo, st_lo, col0 + 16 * hh, m0 + q * 32); bulk_commit(); }
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();      // neither CTA leaves while the pair's MMAs / commits may still touch the other
    if (warp == 1) { if (TWO) tmem_dealloc_2sm(tmem_base, NACC * BN); else tmem_dealloc(tmem_base, NACC * BN); }
}

// ---------------------------------------------------------------------------------------------
// mn kernel: wgrad.  D[Kin(128-tile), Nout(BN-tile)] = sum over the CTA's row range of X[m,:]^T dY[m,:]
// (both scaled; the reduction kernel multiplies by 2^-(eX + edY))
// ---------------------------------------------------------------------------------------------
struct MN16Params {
    int Kin, Nout;
    int64_t M;
    const int* M_dev;         // rows actually in use (device-side, see KK16Params::M_dev); nullptr = M
    int tiles_k, tiles_n, splits;
    int64_t rows_per_split;   // multiple of F_CHUNK_KB * F_BK (recomputed on the device when M_dev is set)
    float* partial;           // [splits][Kin][Nout]
};

template <int BN>
__global__ void __launch_bounds__(F_THREADS, 1)
f16_gemm_mn_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const MN16Params p) {
    constexpr int B_TILE_BYTES = BN * F_BK * 2;
    constexpr int STAGE_BYTES = 2 * FA_TILE_BYTES + 2 * B_TILE_BYTES;
    constexpr int CPT = BN / 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + F_STAGES;
    uint64_t* tfull = bars + 2 * F_STAGES;   // [2]
    uint64_t* tempty = tfull + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x % (p.tiles_k * p.tiles_n);
    const int split = blockIdx.x / (p.tiles_k * p.tiles_n);
    const int kin0 = (tile / p.tiles_n) * F_BM, n0 = (tile % p.tiles_n) * BN;
    int64_t M_rows = p.M, rows_per_split = p.rows_per_split;
    if (p.M_dev != nullptr) {
        M_rows = min((int64_t)__ldg(p.M_dev), p.M);
        constexpr int64_t G = F_CHUNK_KB * F_BK;
        rows_per_split = ((M_rows + p.splits - 1) / p.splits + G - 1) / G * G;
    }
    const int64_t r_begin = (int64_t)split * rows_per_split;
    int64_t r_end = r_begin + rows_per_split;
    if (r_end > M_rows) r_end = M_rows;
    const int k_blocks = r_end > r_begin ? (int)((r_end - r_begin + F_BK - 1) / F_BK) : 0;
    const int chunks = (k_blocks + F_CHUNK_KB - 1) / F_CHUNK_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < F_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, F_EPI_THREADS); }
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
                const int r0 = (int)(r_begin + (int64_t)kb * F_BK);
                // box {64 cols, 64 rows, 2 (or BN/64) column blocks}: rows beyond M are zero-filled
                tma_load_3d(st, &tmA_hi, 0, r0, kin0 / 64, full + stage);
                tma_load_3d(st + FA_TILE_BYTES, &tmA_lo, 0, r0, kin0 / 64, full + stage);
                tma_load_3d(st + 2 * FA_TILE_BYTES, &tmB_hi, 0, r0, n0 / 64, full + stage);
                tma_load_3d(st + 2 * FA_TILE_BYTES + B_TILE_BYTES, &tmB_lo, 0, r0, n0 / 64, full + stage);
                if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(F_BM, BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            int kb = 0;
            for (uint32_t cc = 0; kb < k_blocks; ++cc) {
                const int acc = (int)(cc & 1u);
                mbar_wait(tempty + acc, ((cc >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                const int kb_end = (kb + F_CHUNK_KB < k_blocks) ? kb + F_CHUNK_KB : k_blocks;
                for (int kc = 0; kb < kb_end; ++kb, ++kc) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    // MN-major fp16, SWIZZLE_128B: an atom is 8 k-rows x 128 B (64 columns).  The 64-column blocks
                    // are F_BK rows * 128 B = 8192 B apart (LBO); consecutive 8-row atoms along k are 1024 B apart
                    // (SBO); one MMA (K = 16) consumes two of them = 2048 B per k-step.
                    const uint64_t a_hi = make_desc(sa, 8192, 1024);
                    const uint64_t a_lo = make_desc(sa + FA_TILE_BYTES, 8192, 1024);
                    const uint64_t b_hi = make_desc(sa + 2 * FA_TILE_BYTES, 8192, 1024);
                    const uint64_t b_lo = make_desc(sa + 2 * FA_TILE_BYTES + B_TILE_BYTES, 8192, 1024);
#pragma unroll
                    for (int k = 0; k < F_BK / 16; ++k) {
                        const uint64_t koff = (uint64_t)(k * 128);   // 2048 B per k-step
                        umma_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc | k) != 0 ? 1u : 0u);
                        umma_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                        umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                    }
                    tc_commit(empty + stage);
                    if (kb == kb_end - 1) tc_commit(tfull + acc);
                    if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
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
            const int kb_in_chunk = ((int)cc + 1) * F_CHUNK_KB <= k_blocks ? F_CHUNK_KB : k_blocks - (int)cc * F_CHUNK_KB;
            drain_chunk16<CPT>(tmem_base + (uint32_t)(acc * BN + half * CPT) + ((uint32_t)(q * 32) << 16), s,
                               1.0f + F_RZ_COMP * (float)(3 * (F_BK / 16) * kb_in_chunk));
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
// scales: statistics, planning, splitting
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_max_256(float m) {
    __shared__ float sm[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = sm[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = fmaxf(r, sm[i]);
    __syncthreads();
    return r;
}

// out = max(out, max |x|) as the bit pattern of a non-negative float (order-preserving, deterministic)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, unsigned* out) {
    float m = 0.0f;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(x[(n4 << 2) + threadIdx.x]));
    m = block_max_256(m);
    if (threadIdx.x == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

__device__ __forceinline__ void write_scale(float* sc, float bound) {
    int e = 0;
    if (bound > 0.0f && bound < INFINITY) {
        int ex;
        frexpf(bound, &ex);                       // bound = m * 2^ex, m in [0.5, 1)  =>  bound <= 2^ex
        e = F_EXP_TARGET - ex;
        e = e < -F_EXP_CLAMP ? -F_EXP_CLAMP : (e > F_EXP_CLAMP ? F_EXP_CLAMP : e);
    }
    sc[0] = ldexpf(1.0f, e);
    sc[1] = ldexpf(1.0f, -e);
}

struct F16LayerTable {
    int L;                          // Dense layers (hidden + head)
    int K[F_MAX_LAYERS], N[F_MAX_LAYERS];
    long long w_off[F_MAX_LAYERS], b_off[F_MAX_LAYERS];
};

// stats[l][0..3] = {max |W|, max_n sum_k |W[k][n]|, max_k sum_n |W[k][n]|, max |b|}; grid (blocks, L).
// Column blocks: 256 threads = 32 columns x 8 row groups (coalesced 128-byte rows, 8 partial sums per column folded in
// a fixed order); row blocks: one warp per row, lanes stride over the columns, fixed-order shuffle fold.  All maxima go
// through atomicMax on the bit pattern (order-independent), so the statistics are deterministic.
__global__ void __launch_bounds__(256)
wstats_kernel(const float* __restrict__ params, F16LayerTable t, unsigned* stats) {
    __shared__ float part[8][33];
    const int l = blockIdx.y;
    const int K = t.K[l], N = t.N[l];
    const int nbc = (N + 31) / 32, nbr = (K + 7) / 8;
    if ((int)blockIdx.x >= nbc + nbr) return;
    const float* W = params + t.w_off[l];
    const float* b = params + t.b_off[l];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    float amax = 0.0f, colmax = 0.0f, rowmax = 0.0f, bmax = 0.0f;
    if ((int)blockIdx.x < nbc) {
        const int n = blockIdx.x * 32 + lane;
        float s = 0.0f;
        if (n < N)
            for (int k = grp; k < K; k += 8) { const float a = fabsf(W[(size_t)k * N + n]); s += a; amax = fmaxf(amax, a); }
        part[grp][lane] = s;
        __syncthreads();
        if (grp == 0 && n < N) {
            float c = 0.0f;
#pragma unroll
            for (int g = 0; g < 8; ++g) c += part[g][lane];
            colmax = c;
            bmax = fabsf(b[n]);
        }
    } else {
        const int k = (blockIdx.x - nbc) * 8 + grp;
        float s = 0.0f;
        if (k < K)
            for (int n = lane; n < N; n += 32) s += fabsf(W[(size_t)k * N + n]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        rowmax = s;
    }
    amax = block_max_256(amax); colmax = block_max_256(colmax); rowmax = block_max_256(rowmax); bmax = block_max_256(bmax);
    if (threadIdx.x == 0) {
        unsigned* st = stats + 4 * l;
        if (amax > 0.0f) atomicMax(st + 0, __float_as_uint(amax));
        if (colmax > 0.0f) atomicMax(st + 1, __float_as_uint(colmax));
        if (rowmax > 0.0f) atomicMax(st + 2, __float_as_uint(rowmax));
        if (bmax > 0.0f) atomicMax(st + 3, __float_as_uint(bmax));
    }
}

// scale slots (pairs {scale, 1/scale}): 0 = X0, l = act[l] (l = 1..L-1), L + l = W_l (l = 0..L-2),
// 2L - 1 + l = dZ_l, the gradient at the output of hidden layer l (l = 0..L-2)
__host__ __device__ inline int sc_act(int l) { return l; }
__host__ __device__ inline int sc_w(int L, int l) { return L + l; }
__host__ __device__ inline int sc_dz(int L, int l) { return 2 * L - 1 + l; }

__global__ void plan_w_kernel(int L, const unsigned* wstats, float* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int l = 0; l + 1 < L; ++l) write_scale(sc + 2 * sc_w(L, l), __uint_as_float(wstats[4 * l]));
}
// Forward plan: scales of X0 and of every hidden activation from a bound of the input features and the weight statistics.
// The bound is the minibatch's own abs-max (st_x0: consumed and cleared) or, when the rollout buffer supplies one, the
// abs-max of everything the buffer holds (feat_bound: read only).  Also clears the abs-max word of dlogits, which the
// loss kernel of this minibatch accumulates into (plan_bwd reads it, nobody clears it after).
struct PlanFwd {
    int L;                       // 0: nothing to do
    float slope;
    unsigned* st_x0;
    const unsigned* feat_bound;
    const unsigned* wstats;
    float* sc;
    unsigned* st_dl;
};
__device__ __forceinline__ void plan_fwd_body(const PlanFwd& pf) {
    const int L = pf.L;
    float B;
    if (pf.feat_bound != nullptr) B = __uint_as_float(__ldcg(pf.feat_bound));
    else { B = __uint_as_float(*pf.st_x0); *pf.st_x0 = 0u; }
    if (pf.st_dl != nullptr) *pf.st_dl = 0u;
    write_scale(pf.sc + 2 * sc_act(0), B);
    const float amp = fmaxf(1.0f, fabsf(pf.slope));
    for (int l = 0; l + 1 < L; ++l) {
        B = (B * __uint_as_float(pf.wstats[4 * l + 1]) + __uint_as_float(pf.wstats[4 * l + 3])) * amp;
        B *= 1.0009765625f;      // the bound itself is evaluated in fp32
        write_scale(pf.sc + 2 * sc_act(l + 1), B);
    }
}
__global__ void plan_fwd_kernel(PlanFwd pf) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    plan_fwd_body(pf);
}
// Backward plan: G(dZ_{l-1}) = G(dZ_l) * max_k sum_n |W_l[k][n]| from the abs-max of dlogits (dZ_{L-1} = dlogits); `write`:
// store the scales (one block does); returns the scale pair of dZ_{L-2}, the head's output
__device__ __forceinline__ float2 plan_bwd_body(int L, float slope, const unsigned* st_dl, const unsigned* wstats, float* sc, bool write) {
    float G = __uint_as_float(__ldcg(st_dl));
    const float amp = fmaxf(1.0f, fabsf(slope));
    float2 head = make_float2(1.0f, 1.0f);
    for (int l = L - 1; l >= 1; --l) {      // dZ_{l-1} = (dZ_l W_l^T) .* act'
        G = G * __uint_as_float(__ldcg(wstats + 4 * l + 2)) * amp * 1.0009765625f;
        float tmp[2];
        write_scale(tmp, G);
        if (write) { sc[2 * sc_dz(L, l - 1)] = tmp[0]; sc[2 * sc_dz(L, l - 1) + 1] = tmp[1]; }
        if (l == L - 1) head = make_float2(tmp[0], tmp[1]);
    }
    return head;
}
// stand-alone (the abs-max word came from absmax_kernel; consumed and cleared here)
__global__ void plan_bwd_kernel(int L, float slope, unsigned* st_dl, const unsigned* wstats, float* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    plan_bwd_body(L, slope, st_dl, wstats, sc, true);
    *st_dl = 0u;
}
__global__ void scale_from_stat_kernel(unsigned* st, float* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    write_scale(sc, __uint_as_float(*st));
    *st = 0u;
}
__global__ void scale_from_bound_kernel(float bound, float* sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) write_scale(sc, bound);
}

// x * scale -> (hi, lo), 8 elements per thread
__global__ void __launch_bounds__(256)
split16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n8, const float* sc) {
    const float s = __ldg(sc);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f16_split(v[j] * s, h[j], l[j]);
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
    }
}
__global__ void __launch_bounds__(256)
split16_tail_kernel(const float* x, __half* hi, __half* lo, int64_t begin, int64_t n, const float* sc) {
    const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { __half h, l; f16_split(x[i] * __ldg(sc), h, l); hi[i] = h; lo[i] = l; }
}
// (hi + lo) / scale -> fp32
__global__ void __launch_bounds__(256)
join16_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int64_t n, const float* sc, float* __restrict__ out) {
    const float inv = __ldg(sc + 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (__half2float(hi[i]) + __half2float(lo[i])) * inv;
}

// sign bits of an fp32 matrix [M][N] (N % 32 == 0) in the sign_index layout (tests / benches build gates with it)
__global__ void __launch_bounds__(256)
signbits_kernel(const float* __restrict__ x, int64_t M, int N, uint32_t* __restrict__ out) {
    const int nw = N >> 5;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * nw; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / nw;
        const int w = (int)(i % nw);
        uint32_t b = 0u;
        for (int j = 0; j < 32; ++j) b |= (x[row * N + w * 32 + j] > 0.0f ? 1u : 0u) << j;
        out[sign_index(row, w, nw)] = b;
    }
}

// the leakyrelu' gates the backward pass uses, unpacked to one byte per activation (parity tests compare them with the
// oracle's own pre-activation signs): from the sign-bit words of a hidden activation ...
__global__ void __launch_bounds__(256)
gates_from_signbits_kernel(const uint32_t* __restrict__ signs, int64_t M, int N, uint8_t* __restrict__ out) {
    const int nw = N >> 5;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / N;
        const int col = (int)(i % N);
        out[i] = (uint8_t)((signs[sign_index(row, col >> 5, nw)] >> (col & 31)) & 1u);
    }
}
// ... or, for the last hidden activation, from sign(hi) exactly as head_bwd16_kernel gates
__global__ void __launch_bounds__(256)
gates_from_hi_kernel(const __half* __restrict__ hi, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __half2float(hi[i]) > 0.0f ? 1 : 0;
}

// W[K][N] * scale -> hi/lo of W and of W^T[N][K]
__global__ void __launch_bounds__(256)
weight_prep16_kernel(const float* __restrict__ W, __half* __restrict__ W_hi, __half* __restrict__ W_lo,
                     __half* __restrict__ WT_hi, __half* __restrict__ WT_lo, int K, int N, const float* sc) {
    __shared__ float tile[32][33];
    const float s = __ldg(sc);
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        const float v = (k < K && n < N) ? W[(size_t)k * N + n] * s : 0.0f;
        tile[r][tx] = v;
        if (k < K && n < N) { __half h, l; f16_split(v, h, l); W_hi[(size_t)k * N + n] = h; W_lo[(size_t)k * N + n] = l; }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        if (n < N && k < K) {
            __half h, l; f16_split(tile[tx][r], h, l);
            WT_hi[(size_t)n * K + k] = h;
            WT_lo[(size_t)n * K + k] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused optimiser step + operand refresh: ONE launch does what adam_kernel, adam_tick_kernel, wstats_kernel, plan_w_kernel,
// L-1 weight_prep16_kernel launches and step_advance_kernel did one after the other (8 launches of 2-11 us each: at the
// reference's small minibatches, and at the 8 192-row minibatches of the 8-GPU runs, such fixed costs are what is left).
// All CTAs are co-resident (grid <= number of SMs) and meet at two grid-wide barriers:
//   phase A  Adam (Flux.update!, src/train.jl:81; the maths of adam.cu, element for element) -- or, with data parallelism
//            over peer memory, the rank-ordered sum of the published gradients first (dp_p2p.cu);
//   phase B  weight statistics of the UPDATED parameters (exactly wstats_kernel's sums, same order of additions);
//   phase C  scales (plan_w), fp16 hi/lo copies of W and W^T (weight_prep16), beta powers / exchange epoch / minibatch counter.
// ---------------------------------------------------------------------------------------------
struct F16RefreshArgs {
    F16LayerTable t;
    __half* W_hi[F_MAX_LAYERS]; __half* W_lo[F_MAX_LAYERS]; __half* WT_hi[F_MAX_LAYERS]; __half* WT_lo[F_MAX_LAYERS];
    float* params;
    unsigned* wstats;        // [4 L]
    float* sc;
    unsigned* bar;           // [2] grid-barrier counters (zero between launches)
    // Adam (m == nullptr: refresh only)
    float* m; float* v; const float* grads; long long P;
    double eta, b1, b2, eps; double* bp;
    // peer-memory gradient exchange (peer_xchg == nullptr: local gradient)
    const float* const* peer_xchg; long long Ppad; const unsigned* flags; unsigned* p2p_state; int nranks; float* grads_out;
    int* d_step;             // minibatch counter to advance (or nullptr)
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the (co-resident) grid arrive; the counter is reset by block 0 once the NEXT barrier has been passed
__device__ __forceinline__ void grid_barrier(unsigned* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_gpu_u32(ctr) < gridDim.x) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
f16_adam_refresh_kernel(const F16RefreshArgs a) {
    __shared__ float part[8][33];
    __shared__ float tile[32][33];
    __shared__ int s_abort;
    const int L = a.t.L;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.bar[1] = 0u;      // (the previous launch is over; nobody is at barrier 2 yet)
    // ---------------- phase A: Adam ----------------
    if (a.m != nullptr) {
        if (threadIdx.x == 0) s_abort = 0;
        unsigned epoch = 0u;
        if (a.peer_xchg != nullptr) {
            epoch = a.p2p_state[0];
            // a peer that never publishes must not hang the GPU: see p2p_adam_kernel (dp_p2p.cu)
            if (threadIdx.x == 0) {
                int abort_ = __ldcg(a.p2p_state + 2) != 0u;
                const long long t0 = clock64();
                unsigned long long g0 = 0ull;
                if (blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
                for (int q = 0; q < a.nranks && !abort_; ++q)
                    while (ld_acquire_sys_u32(a.flags + q) < epoch + 1u) {
                        if (clock64() - t0 > 8000000000ll) { a.p2p_state[2] = 1u; abort_ = 1; break; }
                        __nanosleep(64);
                    }
                if (blockIdx.x == 0) {      // diagnostic: time this rank spent waiting for its peers' gradients (ppo_policy_p2p_wait)
                    unsigned long long g1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
                    unsigned long long* acc = reinterpret_cast<unsigned long long*>(a.p2p_state + 4);
                    acc[0] += g1 - g0;
                    acc[1] += 1ull;
                }
                s_abort = abort_;
            }
        }
        __syncthreads();
        if (!s_abort) {
            const size_t off = (size_t)(epoch & 1u) * (size_t)a.Ppad;
            const double b1 = a.b1, b2 = a.b2;
            const double b1p = a.bp[0], b2p = a.bp[1];
            const double om1 = 1.0 - b1, om2 = 1.0 - b2;
            const double c1 = 1.0 - b1p, c2 = 1.0 - b2p;
            auto adam1 = [&](long long i, float g) {
                const double gi = (double)g;
                const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)a.m[i]), __dmul_rn(om1, gi));
                const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)a.v[i]), __dmul_rn(__dmul_rn(om2, gi), gi));
                a.m[i] = mt;
                a.v[i] = vt;
                const double den = __dadd_rn(sqrt((double)vt / c2), a.eps);
                const float d = (float)__dmul_rn(((double)mt / c1) / den, a.eta);
                a.params[i] = __fsub_rn(a.params[i], d);
            };
            if (a.peer_xchg != nullptr) {
                // four parameters per thread and round: all 16-byte loads of the G peers' copies are in flight together
                // (NVLink latency is paid once per round, not once per element); the sum runs in rank order on every
                // rank, so the weights stay bit-identical across ranks
                const long long n4 = a.P >> 2;
                for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
                    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    float4 pv[8];
                    for (int q0 = 0; q0 < a.nranks; q0 += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q0 + u < a.nranks) pv[u] = __ldcv(reinterpret_cast<const float4*>(a.peer_xchg[q0 + u] + off) + i4);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q0 + u < a.nranks) { acc.x += pv[u].x; acc.y += pv[u].y; acc.z += pv[u].z; acc.w += pv[u].w; }
                    }
                    if (a.grads_out != nullptr) reinterpret_cast<float4*>(a.grads_out)[i4] = acc;
                    adam1(4 * i4, acc.x); adam1(4 * i4 + 1, acc.y); adam1(4 * i4 + 2, acc.z); adam1(4 * i4 + 3, acc.w);
                }
                if (blockIdx.x == 0 && (long long)threadIdx.x < (a.P & 3)) {
                    const long long i = (n4 << 2) + threadIdx.x;
                    float g = 0.0f;
                    for (int q = 0; q < a.nranks; ++q) g += __ldcv(a.peer_xchg[q] + off + i);
                    if (a.grads_out != nullptr) a.grads_out[i] = g;
                    adam1(i, g);
                }
            } else {
                for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.P; i += (long long)gridDim.x * blockDim.x)
                    adam1(i, a.grads[i]);
            }
        }
    }
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) a.wstats[i] = 0u;
    grid_barrier(a.bar + 0);
    // ---------------- phase B: weight statistics (wstats_kernel's work items, spread over the grid) ----------------
    {
        const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
        int base = 0;
        for (int l = 0; l < L; ++l) {
            const int K = a.t.K[l], N = a.t.N[l];
            const int nbc = (N + 31) / 32, nbr = (K + 7) / 8;
            const float* W = a.params + a.t.w_off[l];
            const float* b = a.params + a.t.b_off[l];
            // items [base, base + nbc + nbr) belong to layer l; item j is handled by block (j % gridDim.x)
            int first = (int)blockIdx.x - base % (int)gridDim.x;
            if (first < 0) first += (int)gridDim.x;
            for (int it = first; it < nbc + nbr; it += (int)gridDim.x) {
                float amax = 0.0f, colmax = 0.0f, rowmax = 0.0f, bmax = 0.0f;
                if (it < nbc) {
                    const int n = it * 32 + lane;
                    float s = 0.0f;
                    if (n < N)
                        for (int k = grp; k < K; k += 8) { const float x = fabsf(W[(size_t)k * N + n]); s += x; amax = fmaxf(amax, x); }
                    part[grp][lane] = s;
                    __syncthreads();
                    if (grp == 0 && n < N) {
                        float c = 0.0f;
#pragma unroll
                        for (int g = 0; g < 8; ++g) c += part[g][lane];
                        colmax = c;
                        bmax = fabsf(b[n]);
                    }
                } else {
                    const int k = (it - nbc) * 8 + grp;
                    float s = 0.0f;
                    if (k < K)
                        for (int n = lane; n < N; n += 32) s += fabsf(W[(size_t)k * N + n]);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                    rowmax = s;
                }
                amax = block_max_256(amax); colmax = block_max_256(colmax); rowmax = block_max_256(rowmax); bmax = block_max_256(bmax);
                if (threadIdx.x == 0) {
                    unsigned* st = a.wstats + 4 * l;
                    if (amax > 0.0f) atomicMax(st + 0, __float_as_uint(amax));
                    if (colmax > 0.0f) atomicMax(st + 1, __float_as_uint(colmax));
                    if (rowmax > 0.0f) atomicMax(st + 2, __float_as_uint(rowmax));
                    if (bmax > 0.0f) atomicMax(st + 3, __float_as_uint(bmax));
                }
            }
            base += nbc + nbr;
        }
    }
    grid_barrier(a.bar + 1);
    // ---------------- phase C: scales, operand copies, counters ----------------
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.bar[0] = 0u;                       // every block has left barrier 1
        for (int l = 0; l + 1 < L; ++l) write_scale(a.sc + 2 * sc_w(L, l), __uint_as_float(__ldcg(a.wstats + 4 * l)));
        if (a.m != nullptr && !s_abort) { a.bp[0] *= a.b1; a.bp[1] *= a.b2; }
        if (a.peer_xchg != nullptr) a.p2p_state[0] += 1u;
        if (a.d_step != nullptr) *a.d_step += 1;
    }
    {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        int base = 0;
        for (int l = 0; l + 1 < L; ++l) {
            const int K = a.t.K[l], N = a.t.N[l];
            const int tn = (N + 31) / 32, tk = (K + 31) / 32;
            const float* W = a.params + a.t.w_off[l];
            float sc2[2];
            write_scale(sc2, __uint_as_float(__ldcg(a.wstats + 4 * l)));      // (every block derives the same scale)
            const float s = sc2[0];
            int first = (int)blockIdx.x - base % (int)gridDim.x;
            if (first < 0) first += (int)gridDim.x;
            for (int it = first; it < tn * tk; it += (int)gridDim.x) {
                const int k0 = (it / tn) * 32, n0 = (it % tn) * 32;
                __syncthreads();
                for (int r = ty; r < 32; r += 8) {
                    const int k = k0 + r, n = n0 + tx;
                    const float v = (k < K && n < N) ? W[(size_t)k * N + n] * s : 0.0f;
                    tile[r][tx] = v;
                    if (k < K && n < N) { __half h, lo; f16_split(v, h, lo); a.W_hi[l][(size_t)k * N + n] = h; a.W_lo[l][(size_t)k * N + n] = lo; }
                }
                __syncthreads();
                for (int r = ty; r < 32; r += 8) {
                    const int n = n0 + r, k = k0 + tx;
                    if (n < N && k < K) {
                        __half h, lo; f16_split(tile[tx][r], h, lo);
                        a.WT_hi[l][(size_t)n * K + k] = h;
                        a.WT_lo[l][(size_t)n * K + k] = lo;
                    }
                }
            }
            base += tn * tk;
        }
    }
}

// out[i] = (sum_z partial[z*stride + i]) * inv(sa) * inv(sb); fixed order: four interleaved chains over z (independent
// loads in flight), folded as (s0 + s1) + (s2 + s3)
__global__ void __launch_bounds__(256)
f16_reduce_kernel(const float* __restrict__ partial, int splits, int64_t stride, int64_t count, float* __restrict__ out,
                  const float* sa, const float* sb) {
    const float u = (sa ? __ldg(sa + 1) : 1.0f) * (sb ? __ldg(sb + 1) : 1.0f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int z = 0;
        for (; z + 3 < splits; z += 4) {
            s0 += partial[(size_t)z * stride + i];
            s1 += partial[(size_t)(z + 1) * stride + i];
            s2 += partial[(size_t)(z + 2) * stride + i];
            s3 += partial[(size_t)(z + 3) * stride + i];
        }
        for (; z < splits; ++z) s0 += partial[(size_t)z * stride + i];
        out[i] = ((s0 + s1) + (s2 + s3)) * u;
    }
}

// last stage of the head's partial fold: sum the `chunks` rows of scratch [chunks][K*N + N + K] and scatter the three
// slices to dW [K*N], db [N] and (optionally) the bias gradient of the layer below [K]
__global__ void __launch_bounds__(256)
head_finish_kernel(const float* __restrict__ scratch, int chunks, int K, int N, float* __restrict__ dW, float* __restrict__ db,
                   float* __restrict__ db_below) {
    const int stride = K * N + N + K;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= stride) return;
    float s0 = 0.0f, s1 = 0.0f;
    int z = 0;
    for (; z + 1 < chunks; z += 2) { s0 += scratch[(size_t)z * stride + i]; s1 += scratch[(size_t)(z + 1) * stride + i]; }
    if (z < chunks) s0 += scratch[(size_t)z * stride + i];
    const float s = s0 + s1;
    if (i < K * N) dW[i] = s;
    else if (i < K * N + N) db[i - K * N] = s;
    else if (db_below != nullptr) db_below[i - K * N - N] = s;
}

// stage 1 of folding the dgrad epilogue's per-quarter column sums: block (x = column block, y = row chunk)
__global__ void __launch_bounds__(256)
f16_colsum_fold_kernel(const float* __restrict__ part, int64_t rows, int N, int64_t rows_per_chunk, float* __restrict__ out,
                       const int* __restrict__ M_dev = nullptr) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    if (M_dev != nullptr) {      // per-quarter column sums of a dgrad over *M_dev rows: 4 partial rows per 128-row tile
        const int64_t r = 4 * (((int64_t)__ldg(M_dev) + F_BM - 1) / F_BM);
        rows = r < rows ? r : rows;
        rows_per_chunk = (rows + gridDim.y - 1) / gridDim.y;
    }
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

// ---------------------------------------------------------------------------------------------
// Deferred reductions of the backward pass.  Every wgrad leaves split-K partials, every dgrad per-quarter column sums
// (the bias gradient of the layer below), the head per-CTA partials; folding each right after its producer cost 9 small
// launches per minibatch (f16_reduce x3, colsum_fold + reduce x2, colsum_fold + head_finish).  With separate partial
// regions all of them are folded by TWO launches at the end of the backward pass, in exactly the same order of additions:
//   stage 1: every job's rows are cut into `chunks` row ranges, each folded with four interleaved chains; a job with one
//            chunk (split-K partials) is finished here (x the operands' inverse scales);
//   stage 2: the chunk sums are folded (four chains; the head: two chains + scatter to dW / db / db_below).
// ---------------------------------------------------------------------------------------------
enum { FOLD_DIRECT = 0, FOLD_COLSUM = 1, FOLD_HEAD = 2 };
struct FoldJob {
    const float* src;        // [rows][N]
    float* scratch;          // [chunks][N] (kinds 1, 2)
    float* dst;              // kind 0: [N] (final); kind 1: [N]; kind 2: dW [K * Nh]
    float* db; float* db_below;   // kind 2
    const float* sa; const float* sb;   // kind 0: scale pairs of the two operands
    int N, rows, chunks, kind;
    int rows_from_tokens;    // kind 1: rows = 4 * ceil(active tokens / 128) when the pass was compacted
    int K, Nh;               // kind 2
    int begin1, begin2;      // first block of the job in stage 1 / stage 2
};
constexpr int FOLD_MAX_JOBS = 12;
struct FoldJobs { int n; int blocks1, blocks2; const int* M_dev; FoldJob j[FOLD_MAX_JOBS]; };

// columns per stage-1 block: column sums have many rows and few columns (4 * tiles_m x K): 64-column blocks whose four
// 64-thread groups take every fourth row keep hundreds of blocks busy with 64 chunks; the others use one column per thread
__host__ __device__ inline int fold_block_cols(int kind) { return kind == FOLD_COLSUM ? 64 : 256; }

__global__ void __launch_bounds__(256)
fold_stage1_kernel(const FoldJobs js) {
    __shared__ float sub[4][64];
    int ji = 0;
    while (ji + 1 < js.n && (int)blockIdx.x >= js.j[ji + 1].begin1) ++ji;
    const FoldJob& jb = js.j[ji];
    const int local = (int)blockIdx.x - jb.begin1;
    const int chunk = local % jb.chunks, cb = local / jb.chunks;
    int64_t rows = jb.rows;
    if (jb.rows_from_tokens && js.M_dev != nullptr) {
        const int64_t r = 4 * (((int64_t)__ldg(js.M_dev) + F_BM - 1) / F_BM);
        rows = r < rows ? r : rows;
    }
    const int64_t rpc = (rows + jb.chunks - 1) / jb.chunks;
    const int64_t r0 = (int64_t)chunk * rpc;
    const int64_t r1 = r0 + rpc < rows ? r0 + rpc : rows;
    const float* part = jb.src;
    const int64_t N = jb.N;
    if (jb.kind == FOLD_COLSUM) {
        const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
        const int n = cb * 64 + c;
        float s0 = 0.0f, s1 = 0.0f;
        if (n < jb.N) {
            int64_t r = r0 + g;
            for (; r + 4 < r1; r += 8) { s0 += part[r * N + n]; s1 += part[(r + 4) * N + n]; }
            if (r < r1) s0 += part[r * N + n];
        }
        sub[g][c] = s0 + s1;
        __syncthreads();
        if (g == 0 && n < jb.N) jb.scratch[(size_t)chunk * N + n] = (sub[0][c] + sub[1][c]) + (sub[2][c] + sub[3][c]);
        return;
    }
    const int n = cb * 256 + threadIdx.x;
    if (n >= jb.N) return;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += part[r * N + n]; s1 += part[(r + 1) * N + n]; s2 += part[(r + 2) * N + n]; s3 += part[(r + 3) * N + n];
    }
    for (; r < r1; ++r) s0 += part[r * N + n];
    const float s = (s0 + s1) + (s2 + s3);
    if (jb.kind == FOLD_DIRECT) jb.dst[n] = s * (__ldg(jb.sa + 1) * __ldg(jb.sb + 1));
    else jb.scratch[(size_t)chunk * N + n] = s;
}

__global__ void __launch_bounds__(256)
fold_stage2_kernel(const FoldJobs js) {
    int ji = -1;
    for (int k = 0; k < js.n; ++k) {
        const FoldJob& c = js.j[k];
        if (c.kind != FOLD_DIRECT && (int)blockIdx.x >= c.begin2 && (int)blockIdx.x < c.begin2 + (c.N + 255) / 256) ji = k;
    }
    if (ji < 0) return;
    const FoldJob& jb = js.j[ji];
    const int i = ((int)blockIdx.x - jb.begin2) * 256 + threadIdx.x;
    if (i >= jb.N) return;
    const float* sc = jb.scratch;
    const int64_t N = jb.N;
    const int chunks = jb.chunks;
    if (jb.kind == FOLD_COLSUM) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int z = 0;
        for (; z + 3 < chunks; z += 4) {
            s0 += sc[(size_t)z * N + i]; s1 += sc[(size_t)(z + 1) * N + i]; s2 += sc[(size_t)(z + 2) * N + i]; s3 += sc[(size_t)(z + 3) * N + i];
        }
        for (; z < chunks; ++z) s0 += sc[(size_t)z * N + i];
        jb.dst[i] = (s0 + s1) + (s2 + s3);
    } else {
        float s0 = 0.0f, s1 = 0.0f;
        int z = 0;
        for (; z + 1 < chunks; z += 2) { s0 += sc[(size_t)z * N + i]; s1 += sc[(size_t)(z + 1) * N + i]; }
        if (z < chunks) s0 += sc[(size_t)z * N + i];
        const float s = s0 + s1;
        const int KN = jb.K * jb.Nh;
        if (i < KN) jb.dst[i] = s;
        else if (i < KN + jb.Nh) jb.db[i - KN] = s;
        else if (jb.db_below != nullptr) jb.db_below[i - KN - jb.Nh] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// token compaction.  A token (half-edge) all of whose apa actions carry a -Inf mask has probability exactly 0 for each of
// them (softmax(logits + mask), test/quad_game_utilities.jl:73-79): its logits never reach the loss and its dlogits are
// exactly 0, so its rows contribute exact zeros to every weight / bias gradient.  The reference still pushes those rows
// through the MLP (inactive quads of a padded mesh: test/quad_game_utilities.jl:39-44; padded states:
// examples/triangle/distance_weighted/triangle_utilities.jl:31-55).  Here the MLP runs on the ACTIVE tokens only: a
// deterministic two-pass stream compaction builds row -> token (ascending), the input split gathers those rows, every
// GEMM reads the row count from device memory, and the head scatters the logits back to their dense positions.
// ---------------------------------------------------------------------------------------------
constexpr int TOK_THREADS = 1024;
constexpr int TOK_PER_THREAD = 4;
constexpr int TOK_PER_BLOCK = TOK_THREADS * TOK_PER_THREAD;

__device__ __forceinline__ bool token_active(const float* __restrict__ mask, int64_t t, int apa) {
    bool a = false;
    if (apa == 4) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + t);
        a = (m.x != -INFINITY) | (m.y != -INFINITY) | (m.z != -INFINITY) | (m.w != -INFINITY);
    } else {
        for (int j = 0; j < apa; ++j) a |= (__ldg(mask + t * apa + j) != -INFINITY);
    }
    return a;
}
// block-wide sum of one int per thread (1024 threads); every thread gets the total
__device__ __forceinline__ int block_sum_1024(int v, int* sm /*[32]*/) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = sm[threadIdx.x & 31];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
    return r;
}
__global__ void __launch_bounds__(TOK_THREADS)
token_count_kernel(const float* __restrict__ mask, int64_t M, int apa, int* __restrict__ blk_counts) {
    __shared__ int sm[32];
    const int64_t t0 = (int64_t)blockIdx.x * TOK_PER_BLOCK + (int64_t)threadIdx.x * TOK_PER_THREAD;
    int c = 0;
#pragma unroll
    for (int j = 0; j < TOK_PER_THREAD; ++j)
        if (t0 + j < M) c += token_active(mask, t0 + j, apa) ? 1 : 0;
    c = block_sum_1024(c, sm);
    if (threadIdx.x == 0) blk_counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(TOK_THREADS)
token_compact_kernel(const float* __restrict__ mask, int64_t M, int apa, const int* __restrict__ blk_counts,
                     int* __restrict__ tok_of_row, int* __restrict__ rows_out, const PlanFwd pf) {
    __shared__ int sm[32];
    __shared__ int wsum[32];
    if (pf.L > 0 && blockIdx.x == 0 && threadIdx.x == 0) plan_fwd_body(pf);      // (one launch less than a plan kernel of its own)
    int before = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += TOK_THREADS) before += blk_counts[i];
    before = block_sum_1024(before, sm);
    const int64_t t0 = (int64_t)blockIdx.x * TOK_PER_BLOCK + (int64_t)threadIdx.x * TOK_PER_THREAD;
    bool a[TOK_PER_THREAD];
    int c = 0;
#pragma unroll
    for (int j = 0; j < TOK_PER_THREAD; ++j) {
        a[j] = (t0 + j < M) && token_active(mask, t0 + j, apa);
        c += a[j] ? 1 : 0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int wprefix = 0, total = 0;
    {
        const int wv = wsum[lane];
        int wincl = wv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, wincl, d);
            if (lane >= d) wincl += o;
        }
        wprefix = __shfl_sync(0xffffffffu, wincl - wv, warp);
        total = __shfl_sync(0xffffffffu, wincl, 31);
    }
    int r = before + wprefix + incl - c;
#pragma unroll
    for (int j = 0; j < TOK_PER_THREAD; ++j)
        if (a[j]) tok_of_row[r++] = (int)(t0 + j);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *rows_out = before + total;
}
// gathering split: row r of (hi, lo) = features of token tok_of_row[r] * scale; 8 elements per thread
// Rows between the data and the end of the last 256-row unit (what a CTA pair's tile can touch) are zeroed: the forward
// GEMMs compute them like any other row (their activations are leakyrelu(bias): finite and inside the planned bounds).
__global__ void __launch_bounds__(256)
split16_rows_kernel(const float* __restrict__ x, const int* __restrict__ tok_of_row, const int* __restrict__ rows_dev, int K8,
                    __half* __restrict__ hi, __half* __restrict__ lo, const float* sc, int64_t rows_alloc) {
    const float s = __ldg(sc);
    const int64_t rows = __ldg(rows_dev);
    const int64_t n8 = rows * K8;
    {
        const int64_t end = min((rows + 255) / 256 * 256, rows_alloc);
        for (int64_t i = n8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end * K8; i += (int64_t)gridDim.x * blockDim.x) {
            reinterpret_cast<uint4*>(hi)[i] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(lo)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K8;
        const int c = (int)(i - r * K8);
        const float4* src = reinterpret_cast<const float4*>(x + ((int64_t)__ldg(tok_of_row + r) * K8 + c) * 8);
        const float4 a = __ldcs(src), b = __ldcs(src + 1);
        uint4 h, l;
        f16_split2(a.x * s, a.y * s, h.x, l.x);
        f16_split2(a.z * s, a.w * s, h.y, l.y);
        f16_split2(b.x * s, b.y * s, h.z, l.z);
        f16_split2(b.z * s, b.w * s, h.w, l.w);
        reinterpret_cast<uint4*>(hi)[i] = h;
        reinterpret_cast<uint4*>(lo)[i] = l;
    }
}
// gates of compacted rows -> dense token order; tokens the MLP skipped get PPO_GATE_SKIPPED (their gate multiplies an
// exact zero)
__global__ void __launch_bounds__(256)
gates_scatter_kernel(const uint8_t* __restrict__ rows_gates, const int* __restrict__ tok_of_row, const int* __restrict__ rows_dev,
                     int N, uint8_t* __restrict__ out) {
    const int64_t n = (int64_t)__ldg(rows_dev) * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / N;
        out[(int64_t)__ldg(tok_of_row + r) * N + (i - r * N)] = rows_gates[i];
    }
}

// ---------------------------------------------------------------------------------------------
// the policy head Dense(K, N <= 4) on fp16 pairs
// ---------------------------------------------------------------------------------------------
// logits[m][n] = (sum_k (H_hi + H_lo)[m][k] W[k][n]) / scale + b[n]: one warp per token row, two rows in flight.
// W is staged in shared memory as [c][n][k/8] so that the 32 lanes (consecutive k/8) read consecutive words.
template <int N>
__global__ void __launch_bounds__(256)
head_fwd16_kernel(const __half* __restrict__ H_hi, const __half* __restrict__ H_lo, const float* __restrict__ W,
                  const float* __restrict__ bias, float* __restrict__ logits, int64_t M, int K, const float* sc_h,
                  const int* __restrict__ M_dev, const int* __restrict__ tok_of_row) {
    // token compaction: H holds only the active tokens (*M_dev rows); row r is token tok_of_row[r] of the dense logits
    if (M_dev != nullptr) M = min((int64_t)__ldg(M_dev), M);
    extern __shared__ __align__(16) float sW[];   // [8][N][K/8]
    const int KV = K >> 3;
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        sW[((k & 7) * N + n) * KV + (k >> 3)] = W[i];
    }
    __syncthreads();
    const float inv = __ldg(sc_h + 1);
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t m = w0; m < M; m += 2 * warps) {
        const int64_t m2 = m + warps;
        const bool two = m2 < M;
        const uint4* hp0 = reinterpret_cast<const uint4*>(H_hi + m * K);
        const uint4* lp0 = reinterpret_cast<const uint4*>(H_lo + m * K);
        const uint4* hp1 = reinterpret_cast<const uint4*>(H_hi + (two ? m2 : m) * K);
        const uint4* lp1 = reinterpret_cast<const uint4*>(H_lo + (two ? m2 : m) * K);
        float a0[N], a1[N];
#pragma unroll
        for (int n = 0; n < N; ++n) { a0[n] = 0.0f; a1[n] = 0.0f; }
        for (int kv = lane; kv < KV; kv += 32) {
            const uint4 qh0 = __ldcs(hp0 + kv), ql0 = __ldcs(lp0 + kv), qh1 = __ldcs(hp1 + kv), ql1 = __ldcs(lp1 + kv);
            float h0[8], l0[8], h1[8], l1[8];
            unpack8(qh0, h0); unpack8(ql0, l0); unpack8(qh1, h1); unpack8(ql1, l1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float v0 = h0[c] + l0[c], v1 = h1[c] + l1[c];
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    const float w = sW[(c * N + n) * KV + kv];
                    a0[n] = fmaf(v0, w, a0[n]);
                    a1[n] = fmaf(v1, w, a1[n]);
                }
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
            const int64_t o0 = tok_of_row ? (int64_t)__ldg(tok_of_row + m) : m;
            const int64_t o1 = (two && tok_of_row) ? (int64_t)__ldg(tok_of_row + m2) : m2;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float bn = bias ? bias[n] : 0.0f;
                logits[o0 * N + n] = fmaf(a0[n], inv, bn);
                if (two) logits[o1 * N + n] = fmaf(a1[n], inv, bn);
            }
        }
    }
}

// One pass over H (fp16 pair): dH = (dlogits W^T) .* leakyrelu'(H) written as a scaled fp16 pair; per-CTA partials
// of dW[k][n], db[n] and of the column sums of dH (= the bias gradient of the layer below).
// Partial layout per CTA: [K*N dW][N db][K colsum(dH)].  A thread owns 4 consecutive columns (8-byte loads and
// stores); K/4 threads cover a row, so a CTA walks 256 / (K/4) rows at once, RU of them in flight per thread.
template <int N>
__global__ void __launch_bounds__(256, 2)
head_bwd16_kernel(const __half* __restrict__ H_hi, const __half* __restrict__ H_lo,
                  const float* __restrict__ dlogits, const float* __restrict__ W, __half* __restrict__ dH_hi, __half* __restrict__ dH_lo,
                  float* __restrict__ partial, int64_t M, int K, float slope, int64_t rows_per_cta, int need_dH,
                  const float* sc_h, const float* sc_dh, const int* __restrict__ M_dev, const int* __restrict__ tok_of_row,
                  int plan_L, const unsigned* __restrict__ plan_st_dl, const unsigned* __restrict__ plan_wstats, float* plan_sc) {
    extern __shared__ float red[];                     // [rpp][K*N + N + K]
    // plan_L > 0: the backward plan (scales of every activation gradient from max |dlogits|, which the loss kernel left in
    // plan_st_dl) is evaluated here by every CTA for itself -- a handful of multiplies -- and stored by CTA 0 for the
    // kernels that follow, instead of an abs-max pass over dlogits and a one-thread plan kernel in front of this one
    __shared__ float s_plan_scale;
    if (plan_L > 0) {
        if (threadIdx.x == 0) s_plan_scale = plan_bwd_body(plan_L, slope, plan_st_dl, plan_wstats, plan_sc, blockIdx.x == 0).x;
        __syncthreads();
    }
    const int64_t M_alloc = M;
    if (M_dev != nullptr) {      // token compaction: rows = active tokens, row r reads the dlogits of token tok_of_row[r]
        M = min((int64_t)__ldg(M_dev), M);
        rows_per_cta = (M + gridDim.x - 1) / gridDim.x;
    }
    const int tid = threadIdx.x;
    const int TPR = K >> 2;
    const int rpp = 256 / TPR;
    const int rg = tid / TPR, ct = tid - rg * TPR;
    const bool active = rg < rpp;
    const int k = 4 * ct;
    const float inv_h = __ldg(sc_h + 1);
    const float s_dh = need_dH ? (plan_L > 0 ? s_plan_scale : __ldg(sc_dh)) : 1.0f;
    float w[4][N], aw[4][N], ab[N], cs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        cs[c] = 0.0f;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            w[c][n] = active ? W[(k + c) * N + n] : 0.0f;
            aw[c][n] = 0.0f;
        }
    }
#pragma unroll
    for (int n = 0; n < N; ++n) ab[n] = 0.0f;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    constexpr int RU = 6;
    for (int64_t m = r0 + rg; m < r1 && active; m += (int64_t)RU * rpp) {
        uint2 hh[RU], hl[RU];
        float d[RU][N];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int64_t row = m + (int64_t)u * rpp;
            const bool rv = row < r1;
            hh[u] = make_uint2(0u, 0u); hl[u] = make_uint2(0u, 0u);
            if (rv) {
                hh[u] = __ldcs(reinterpret_cast<const uint2*>(H_hi + row * K + k));
                hl[u] = __ldcs(reinterpret_cast<const uint2*>(H_lo + row * K + k));
            }
            const int64_t tok = (rv && tok_of_row != nullptr) ? (int64_t)__ldg(tok_of_row + row) : row;
#pragma unroll
            for (int n = 0; n < N; ++n) d[u][n] = rv ? __ldg(dlogits + tok * N + n) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int64_t row = m + (int64_t)u * rpp;
            const bool rv = row < r1;
            const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hh[u].x));
            const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hh[u].y));
            const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&hl[u].x));
            const float2 l23 = __half22float2(*reinterpret_cast<const __half2*>(&hl[u].y));
            // gate on sign(hi): a positive activation below 2^-25 of its tensor's scale rounds to hi = 0 and takes the
            // other leakyrelu' branch -- the same class of event as a rounding-level flip of a pre-activation at 0
            const float gate[4] = {h01.x, h01.y, h23.x, h23.y};
            const float hv[4] = {h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y};   // scaled activation
            float gv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float dx = 0.0f;
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    dx = fmaf(d[u][n], w[c][n], dx);
                    aw[c][n] = fmaf(hv[c], d[u][n], aw[c][n]);
                }
                const float g = (gate[c] > 0.0f) ? dx : slope * dx;
                cs[c] += g;                                       // d = 0 for rows out of range
                gv[c] = g * s_dh;
            }
            if (need_dH && rv) {
                uint2 oh, ol;
                f16_split2(gv[0], gv[1], oh.x, ol.x);
                f16_split2(gv[2], gv[3], oh.y, ol.y);
                *reinterpret_cast<uint2*>(dH_hi + row * K + k) = oh;
                *reinterpret_cast<uint2*>(dH_lo + row * K + k) = ol;
            }
            if (ct == 0) {
#pragma unroll
                for (int n = 0; n < N; ++n) ab[n] += d[u][n];
            }
        }
    }
    // fold the row groups (fixed order) and write this CTA's partials
    const int stride = K * N + N + K;
    if (active) {
        float* mine = red + (size_t)rg * stride;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int n = 0; n < N; ++n) mine[(k + c) * N + n] = aw[c][n] * inv_h;
            mine[K * N + N + k + c] = cs[c];
        }
        if (ct == 0) {
#pragma unroll
            for (int n = 0; n < N; ++n) mine[K * N + n] = ab[n];
        }
    }
    __syncthreads();
    float* pw = partial + (int64_t)blockIdx.x * stride;
    for (int i = tid; i < stride; i += 256) {
        float s = 0.0f;
        for (int g = 0; g < rpp; ++g) s += red[(size_t)g * stride + i];
        pw[i] = s;
    }
    // the wgrad contraction of the layer below reads dH in whole 128-row groups: zero the rows between the data and the
    // next multiple of 128 (compaction leaves stale rows of an earlier, larger minibatch there)
    if (need_dH && M_dev != nullptr && blockIdx.x == gridDim.x - 1) {
        const int64_t end = min((M + 127) / 128 * 128, M_alloc);
        const int64_t n4 = (end - M) * (int64_t)(K >> 2);
        for (int64_t i = tid; i < n4; i += 256) {
            *reinterpret_cast<uint2*>(dH_hi + M * K + 4 * i) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(dH_lo + M * K + 4 * i) = make_uint2(0u, 0u);
        }
    }
}

inline int64_t head16_ctas(int64_t M, int num_sms) {
    const int64_t c = (int64_t)num_sms * 4;
    const int64_t maxc = ceil_div(M, 64);
    return c < maxc ? c : (maxc < 1 ? 1 : maxc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode16 = nullptr;

int load_encode16() {
    if (g_encode16) return PPO_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PPO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return PPO_ERR_CUDA;
    }
    g_encode16 = (EncodeTiledFn)fn;
    return PPO_OK;
}

// 2-D map over a row-major [rows][cols] fp16 matrix: box {64 cols, box_rows} with 128B swizzle (operand loads) or
// {16 cols, box_rows} with 32B swizzle (epilogue stores)
int make_map16_2d(CUtensorMap* m, const __half* base, int64_t rows, int64_t cols, int box_rows, int box_cols = 64) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode16(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f16 2d %lld x %lld) failed: %d", (long long)rows, (long long)cols, (int)r); return PPO_ERR_CUDA; }
    return PPO_OK;
}
// store map over BOTH planes of an output pair: {cols, rows, 2} with the lo plane `plane_bytes` after the hi plane; box
// {16 columns, 32 rows, 2 planes} = the 2 x 1 KB a warp stages (SWIZZLE_32B)
int make_map16_store3d(CUtensorMap* m, const __half* hi, int64_t rows, int64_t cols, int64_t plane_bytes) {
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_bytes};
    cuuint32_t box[3] = {16, 32, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode16(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)hi, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PPO_OK : PPO_ERR_CUDA;      // (the caller falls back to two 2-D maps)
}
// 3-D view of a row-major [rows][cols] fp16 matrix as {64, rows, cols/64}: one box = `blocks` column blocks of
// F_BK rows each, i.e. the canonical MN-major SWIZZLE_128B operand layout
int make_map16_mn(CUtensorMap* m, const __half* base, int64_t rows, int64_t cols, int blocks) {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)((cols + 63) / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)F_BK, (cuuint32_t)blocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode16(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f16 mn %lld x %lld) failed: %d", (long long)rows, (long long)cols, (int)r); return PPO_ERR_CUDA; }
    return PPO_OK;
}

struct F16Layer {
    __half* W_hi = nullptr;   // [K][N]  (dgrad B operand)
    __half* W_lo = nullptr;
    __half* WT_hi = nullptr;  // [N][K]  (forward B operand)
    __half* WT_lo = nullptr;
};

struct F16State {
    std::vector<F16Layer> layers;
    F16LayerTable table{};
    int64_t tokens = 0;
    __half* x_hi = nullptr;                  // scaled minibatch features [M][dims[0]] (+ slack)
    __half* x_lo = nullptr;
    std::vector<__half*> act_hi, act_lo;     // act[l], l = 1..L-1
    std::vector<uint32_t*> act_sign;         // sign bits of act[l] ([M][dims[l]/32]): the leakyrelu' gate of the backward pass
    __half* dz_hi[2] = {nullptr, nullptr};   // ping-pong activation gradients [M][Hmax]
    __half* dz_lo[2] = {nullptr, nullptr};
    float* partial = nullptr;
    size_t partial_bytes = 0;
    float* sc = nullptr;                     // scale pairs
    unsigned* st = nullptr;                  // [0] abs-max X0, [1] abs-max dlogits, [2 + 4 l + j] weight statistics
    // token compaction (see token_compact_kernel)
    int* tok_of_row = nullptr;               // [tokens] row -> token of the dense minibatch, ascending
    int* blk_counts = nullptr;               // [ceil(tokens / TOK_PER_BLOCK)] active tokens per block
    int* d_rows = nullptr;                   // number of active tokens of the current minibatch
    unsigned* bar = nullptr;                 // grid-barrier counters of f16_adam_refresh_kernel
    bool compact = false;                    // the last forward pass ran compacted (the backward pass follows it)
};

F16State* state(ppo_policy* p) { return reinterpret_cast<F16State*>(p->f16); }

template <typename T>
void fr(T*& q) { if (q) cudaFree(q); q = nullptr; }

int wgrad_splits16(int64_t M, int tiles, int num_sms) {
    int s = std::max(1, num_sms / tiles);
    const int64_t max_s = std::max<int64_t>(1, M / (F_BK * 8));
    if (s > max_s) s = (int)max_s;
    return s;
}

size_t head16_partial_bytes(int64_t M, int K, int N, int num_sms) {
    return (size_t)(head16_ctas(M, num_sms) + 64) * ((size_t)K * N + N + K) * sizeof(float);   // partials + fold scratch
}

size_t f16_wgrad_partial_bytes(int64_t tokens, int K, int N, int num_sms) {
    const int BN = N > 128 ? 256 : 128;
    const int tiles = (int)(ceil_div(K, F_BM) * ceil_div(N, BN));
    return (size_t)wgrad_splits16(tokens, tiles, num_sms) * K * N * 4;
}
// row chunks of a two-stage fold: both stages walk their rows serially (a dependent chain of L2 round trips), so the
// chunk count balances them: ~sqrt(rows), at most 64
inline int colsum_chunks(int64_t rows) {
    int c = 1;
    while ((int64_t)c * c < rows && c < 64) ++c;
    return c;
}
// dgrad-epilogue column sums [4 * tiles_m][K] + the fold's scratch [64][K]
size_t f16_colsum_partial_bytes(int64_t tokens, int K) { return ((size_t)4 * ceil_div(tokens, F_BM) + 64) * (size_t)K * 4; }
size_t f16_partial_bytes(int64_t tokens, int K, int N, int num_sms) {
    return std::max(f16_wgrad_partial_bytes(tokens, K, N, num_sms), f16_colsum_partial_bytes(tokens, std::max(K, N)));
}

int launch_absmax(ppo_ctx* ctx, const float* x, int64_t n, unsigned* out) {
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(n / 4 + 1, 256), (int64_t)ctx->num_sms * 8));
    absmax_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, n, out);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_split16(ppo_ctx* ctx, const float* x, __half* hi, __half* lo, int64_t n, const float* sc) {
    const int64_t n8 = n / 8;
    if (n8 > 0) {
        const int64_t blocks = std::min<int64_t>(ceil_div(n8, 256), (int64_t)ctx->num_sms * 16);
        split16_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, hi, lo, n8, sc);
        ctx->launches += 1;
    }
    if (n8 * 8 < n) {
        split16_tail_kernel<<<1, 256, 0, ctx->stream>>>(x, hi, lo, n8 * 8, n, sc);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

template <int BN, bool TWO>
int launch_kk16(ppo_ctx* ctx, const __half* A, const __half* A_lo, const __half* B, const __half* B_lo, __half* C, __half* C_lo,
                int64_t M, int N, int K, const KK16Params& base) {
    CUtensorMap mA, mAl, mB, mBl, mC, mCl;
    PPO_TRY(make_map16_2d(&mA, A, M, K, F_BM));
    PPO_TRY(make_map16_2d(&mAl, A_lo, M, K, F_BM));
    PPO_TRY(make_map16_2d(&mB, B, N, K, KK16Smem<BN, TWO>::B_ROWS));
    PPO_TRY(make_map16_2d(&mBl, B_lo, N, K, KK16Smem<BN, TWO>::B_ROWS));
    // store boxes: 32 rows x 16 columns (one warp's), SWIZZLE_32B; hi and lo planes in one 3-D box when the lo plane lies
    // a whole number of rows behind the hi plane (the engine's own buffers do), else one 2-D map each
    const int64_t plane = (int64_t)((const char*)C_lo - (const char*)C);
    static const int no3d = getenv("PPO_F16_STORE2D") ? atoi(getenv("PPO_F16_STORE2D")) : 0;      // tuning experiments
    bool store3d = !no3d && plane > 0 && plane % ((int64_t)N * 2) == 0 && plane % 16 == 0 && plane < ((int64_t)1 << 40) &&
                   make_map16_store3d(&mC, C, M, N, plane) == PPO_OK;
    if (!store3d) PPO_TRY(make_map16_2d(&mC, C, M, N, 32, 16));
    PPO_TRY(make_map16_2d(&mCl, C_lo, M, N, 32, 16));
    KK16Params p = base;
    p.store3d = store3d ? 1 : 0;
    static const int env_chunk = getenv("PPO_F16_KK_CHUNK") ? atoi(getenv("PPO_F16_KK_CHUNK")) : 0;      // tuning experiments
    static const float env_comp = getenv("PPO_F16_RZ_COMP") ? (float)atof(getenv("PPO_F16_RZ_COMP")) : -1.0f;
    p.chunk_kb = env_chunk > 0 ? env_chunk : F_KK_CHUNK_KB;
    p.rz_comp = env_comp >= 0.0f ? env_comp : F_RZ_COMP;
    p.M = (int)M; p.N = N; p.K = K;
    p.tiles_m = (int)ceil_div(M, F_BM); p.tiles_n = (int)ceil_div(N, BN); p.k_blocks = (int)ceil_div(K, F_BK);
    const size_t smem = KK16Smem<BN, TWO>::TOTAL;
    PPO_CUDA(cudaFuncSetAttribute(f16_gemm_kk_kernel<BN, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (TWO) {
        const int units = ((p.tiles_m + 1) / 2) * p.tiles_n;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(F_KK_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // persistent pairs: as many clusters as can be co-resident (an SM without a free partner in its TPC stays idle)
        static int max_clusters = 0;
        if (max_clusters == 0) {
            cfg.gridDim = dim3((unsigned)(2 * (ctx->num_sms / 2)));
            int nc = 0;
            PPO_CUDA(cudaOccupancyMaxActiveClusters(&nc, f16_gemm_kk_kernel<BN, TWO>, &cfg));
            max_clusters = std::max(1, std::min(nc, ctx->num_sms / 2));
        }
        cfg.gridDim = dim3((unsigned)(2 * std::min(units, max_clusters)));
        PPO_CUDA(cudaLaunchKernelEx(&cfg, f16_gemm_kk_kernel<BN, TWO>, mA, mAl, mB, mBl, mC, mCl, p));
    } else {
        const int tiles = p.tiles_m * p.tiles_n;
        const int grid = std::min(tiles, ctx->num_sms);
        f16_gemm_kk_kernel<BN, TWO><<<grid, F_KK_THREADS, smem, ctx->stream>>>(mA, mAl, mB, mBl, mC, mCl, p);
    }
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int kk16_dispatch(ppo_ctx* ctx, const __half* A, const __half* A_lo, const __half* B, const __half* B_lo, __half* C, __half* C_lo,
                  int64_t M, int N, int K, const KK16Params& base) {
    PPO_REQUIRE(M < ((int64_t)1 << 31), "f16 gemm: M too large");
    PPO_REQUIRE(K % 8 == 0 && N % 32 == 0, "f16 gemm: K %% 8 and N %% 32 required (K=%d N=%d)", K, N);
    // tile width: 128 x 256 tiles need the fewest operand bytes per MMA (the kernel is L2-feed bound at K = 512); for a
    // short contraction (K <= 128: the first layer) the epilogue dominates and 128 x 128 tiles with 4 TMEM stages keep
    // it off the critical path (measured 0.56 vs 0.66 ms at M = 2^20, K = 64, N = 512)
    static const int force128 = getenv("PPO_F16_BN128") ? atoi(getenv("PPO_F16_BN128")) : -1;   // tuning experiments
    const bool wide = force128 >= 0 ? !force128 : (K > 128);
    // CTA pairs (cta_group::2) halve the weight-tile bytes per SM: 1.35 vs 1.43 ms at M = 2^20, K = N = 512
    static const int pair_env = getenv("PPO_F16_2SM") ? atoi(getenv("PPO_F16_2SM")) : -1;         // tuning experiments
    const bool pair = pair_env >= 0 ? pair_env != 0 : (M > F_BM);
    if (N > 128 && wide && pair) return launch_kk16<256, true>(ctx, A, A_lo, B, B_lo, C, C_lo, M, N, K, base);
    if (N > 128 && wide) return launch_kk16<256, false>(ctx, A, A_lo, B, B_lo, C, C_lo, M, N, K, base);
    return launch_kk16<128, false>(ctx, A, A_lo, B, B_lo, C, C_lo, M, N, K, base);
}

template <int BN>
int launch_mn16(ppo_ctx* ctx, const __half* X, const __half* X_lo, const __half* dY, const __half* dY_lo, int64_t M, int Kin,
                int Nout, float* partial, int splits, const int* M_dev) {
    CUtensorMap mA, mAl, mB, mBl;
    PPO_TRY(make_map16_mn(&mA, X, M, Kin, F_BM / 64));
    PPO_TRY(make_map16_mn(&mAl, X_lo, M, Kin, F_BM / 64));
    PPO_TRY(make_map16_mn(&mB, dY, M, Nout, BN / 64));
    PPO_TRY(make_map16_mn(&mBl, dY_lo, M, Nout, BN / 64));
    MN16Params p;
    p.Kin = Kin; p.Nout = Nout; p.M = M; p.M_dev = M_dev;
    p.tiles_k = (int)ceil_div(Kin, F_BM); p.tiles_n = (int)ceil_div(Nout, BN); p.splits = splits;
    p.rows_per_split = round_up(ceil_div(M, splits), F_CHUNK_KB * F_BK);
    p.partial = partial;
    const size_t smem = (size_t)F_STAGES * (2 * FA_TILE_BYTES + 2 * BN * F_BK * 2) + 1024 + 256;
    PPO_CUDA(cudaFuncSetAttribute(f16_gemm_mn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.tiles_k * p.tiles_n * splits;
    f16_gemm_mn_kernel<BN><<<grid, F_THREADS, smem, ctx->stream>>>(mA, mAl, mB, mBl, p);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// the deferred reductions of one backward pass: producers append jobs, fold_flush launches the two stages
struct FoldList {
    FoldJobs js{};
    int add(const FoldJob& j) {
        PPO_REQUIRE(js.n < FOLD_MAX_JOBS, "fold list full");
        js.j[js.n++] = j;
        return PPO_OK;
    }
};
int fold_flush(ppo_ctx* ctx, FoldList& fl, const int* M_dev) {
    FoldJobs& js = fl.js;
    if (js.n == 0) return PPO_OK;
    js.M_dev = M_dev;
    int b1 = 0, b2 = 0;
    for (int i = 0; i < js.n; ++i) {
        FoldJob& j = js.j[i];
        j.begin1 = b1; j.begin2 = b2;
        b1 += (int)ceil_div(j.N, fold_block_cols(j.kind)) * j.chunks;
        if (j.kind != FOLD_DIRECT) b2 += (int)ceil_div(j.N, 256);
    }
    js.blocks1 = b1; js.blocks2 = b2;
    fold_stage1_kernel<<<(unsigned)b1, 256, 0, ctx->stream>>>(js);
    ctx->launches += 1;
    if (b2 > 0) {
        fold_stage2_kernel<<<(unsigned)b2, 256, 0, ctx->stream>>>(js);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// fold [rows][N] per-quarter column sums (rows = 4 * tiles_m) into out[N]; scratch holds 64 * N floats
int fold_colsum16(ppo_ctx* ctx, const float* part, int64_t rows, int N, float* scratch, float* out, const int* M_dev = nullptr,
                  FoldList* defer = nullptr) {
    const int chunks = colsum_chunks(rows);
    if (defer != nullptr) {
        FoldJob j{};
        j.src = part; j.scratch = scratch; j.dst = out; j.N = N; j.rows = (int)rows; j.chunks = chunks; j.kind = FOLD_COLSUM;
        j.rows_from_tokens = 1;
        return defer->add(j);
    }
    const int64_t rpc = ceil_div(rows, chunks);
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)chunks);
    f16_colsum_fold_kernel<<<grid, 256, 0, ctx->stream>>>(part, rows, N, rpc, scratch, M_dev);
    f16_reduce_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, ctx->stream>>>(scratch, chunks, N, N, out, nullptr, nullptr);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int wgrad16(ppo_ctx* ctx, const __half* X, const __half* X_lo, const __half* dY, const __half* dY_lo, float* dW,
            float* partial, size_t partial_bytes, int64_t M, int K, int N, const float* sc_x, const float* sc_dy,
            const int* M_dev = nullptr, FoldList* defer = nullptr) {
    PPO_REQUIRE(K % 8 == 0 && N % 32 == 0, "f16 wgrad: K %% 8 and N %% 32 required (K=%d N=%d)", K, N);
    const int BN = N > 128 ? 256 : 128;
    const int tiles = (int)(ceil_div(K, F_BM) * ceil_div(N, BN));
    const int splits = wgrad_splits16(M, tiles, ctx->num_sms);
    PPO_REQUIRE((size_t)splits * K * N * 4 <= partial_bytes, "f16 wgrad: partial buffer too small");
    if (BN == 256) PPO_TRY(launch_mn16<256>(ctx, X, X_lo, dY, dY_lo, M, K, N, partial, splits, M_dev));
    else PPO_TRY(launch_mn16<128>(ctx, X, X_lo, dY, dY_lo, M, K, N, partial, splits, M_dev));
    const int64_t cnt = (int64_t)K * N;
    if (defer != nullptr) {
        FoldJob j{};
        j.src = partial; j.dst = dW; j.sa = sc_x; j.sb = sc_dy; j.N = (int)cnt; j.rows = splits; j.chunks = 1; j.kind = FOLD_DIRECT;
        return defer->add(j);
    }
    f16_reduce_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, ctx->stream>>>(partial, splits, cnt, cnt, dW, sc_x, sc_dy);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int head_fwd16(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* W, const float* bias, float* logits,
               int64_t M, int K, int N, const float* sc_h, const int* M_dev = nullptr, const int* tok_of_row = nullptr) {
    PPO_REQUIRE(N >= 1 && N <= 4 && K % 8 == 0 && (size_t)K * N * 4 <= 48 * 1024, "f16 head_fwd: needs N <= 4, K %% 8 == 0 (K=%d N=%d)", K, N);
    int64_t blocks = ceil_div(M, 8);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)K * N * sizeof(float);
    switch (N) {
        case 1: head_fwd16_kernel<1><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H_hi, H_lo, W, bias, logits, M, K, sc_h, M_dev, tok_of_row); break;
        case 2: head_fwd16_kernel<2><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H_hi, H_lo, W, bias, logits, M, K, sc_h, M_dev, tok_of_row); break;
        case 3: head_fwd16_kernel<3><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H_hi, H_lo, W, bias, logits, M, K, sc_h, M_dev, tok_of_row); break;
        default: head_fwd16_kernel<4><<<(unsigned)blocks, 256, smem, ctx->stream>>>(H_hi, H_lo, W, bias, logits, M, K, sc_h, M_dev, tok_of_row); break;
    }
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int head_bwd16(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* dlogits, const float* W, __half* dH_hi,
               __half* dH_lo, float* dW, float* db, float* db_below, int64_t M, int K, int N, float slope, float* partial,
               size_t partial_bytes, const float* sc_h, const float* sc_dh, const int* M_dev = nullptr,
               const int* tok_of_row = nullptr, FoldList* defer = nullptr, int plan_L = 0, const unsigned* plan_st_dl = nullptr,
               const unsigned* plan_wstats = nullptr, float* plan_sc = nullptr) {
    PPO_REQUIRE(N >= 1 && N <= 4 && K % 4 == 0 && K >= 4 && K <= 1024, "f16 head_bwd: needs N <= 4, K %% 4 == 0, K <= 1024 (K=%d N=%d)", K, N);
    const int64_t ctas = head16_ctas(M, ctx->num_sms);
    const int64_t rows = ceil_div(M, ctas);
    const int64_t stride = (int64_t)K * N + N + K;
    PPO_REQUIRE((size_t)(ctas + 64) * stride * sizeof(float) <= partial_bytes, "f16 head_bwd: partial buffer too small");
    const int need_dH = dH_hi != nullptr ? 1 : 0;
    const int rpp = 256 / (K / 4);
    const size_t smem = (size_t)rpp * stride * sizeof(float);
    switch (N) {
        case 1: head_bwd16_kernel<1><<<(unsigned)ctas, 256, smem, ctx->stream>>>(H_hi, H_lo, dlogits, W, dH_hi, dH_lo, partial, M, K, slope, rows, need_dH, sc_h, sc_dh, M_dev, tok_of_row, plan_L, plan_st_dl, plan_wstats, plan_sc); break;
        case 2: head_bwd16_kernel<2><<<(unsigned)ctas, 256, smem, ctx->stream>>>(H_hi, H_lo, dlogits, W, dH_hi, dH_lo, partial, M, K, slope, rows, need_dH, sc_h, sc_dh, M_dev, tok_of_row, plan_L, plan_st_dl, plan_wstats, plan_sc); break;
        case 3: head_bwd16_kernel<3><<<(unsigned)ctas, 256, smem, ctx->stream>>>(H_hi, H_lo, dlogits, W, dH_hi, dH_lo, partial, M, K, slope, rows, need_dH, sc_h, sc_dh, M_dev, tok_of_row, plan_L, plan_st_dl, plan_wstats, plan_sc); break;
        default: head_bwd16_kernel<4><<<(unsigned)ctas, 256, smem, ctx->stream>>>(H_hi, H_lo, dlogits, W, dH_hi, dH_lo, partial, M, K, slope, rows, need_dH, sc_h, sc_dh, M_dev, tok_of_row, plan_L, plan_st_dl, plan_wstats, plan_sc); break;
    }
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    // fold the per-CTA partials [ctas][stride] in two parallel stages (a single pass over 592 partials per output is a
    // 60 us serial chain: it dominated the small-minibatch regime of the multi-GPU runs)
    const int chunks = colsum_chunks(ctas);
    const int64_t rpc = ceil_div(ctas, chunks);
    float* scratch = partial + (size_t)ctas * stride;
    if (defer != nullptr) {
        FoldJob j{};
        j.src = partial; j.scratch = scratch; j.dst = dW; j.db = db; j.db_below = (db_below != nullptr && need_dH) ? db_below : nullptr;
        j.N = (int)stride; j.rows = (int)ctas; j.chunks = chunks; j.kind = FOLD_HEAD; j.K = K; j.Nh = N;
        return defer->add(j);
    }
    f16_colsum_fold_kernel<<<dim3((unsigned)ceil_div(stride, 256), (unsigned)chunks), 256, 0, ctx->stream>>>(partial, ctas, (int)stride,
                                                                                                        rpc, scratch);
    head_finish_kernel<<<(unsigned)ceil_div(stride, 256), 256, 0, ctx->stream>>>(scratch, chunks, K, N, dW, db,
                                                                              (db_below != nullptr && need_dH) ? db_below : nullptr);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int ensure_f16_workspace(ppo_policy* p, int64_t tokens) {
    F16State* st = state(p);
    if (tokens <= st->tokens) return PPO_OK;
    ppo_ctx* ctx = p->ctx;
    PPO_CUDA(cudaStreamSynchronize(ctx->stream));
    fr(st->x_hi); fr(st->x_lo); fr(st->dz_hi[0]); fr(st->dz_hi[1]); st->dz_lo[0] = st->dz_lo[1] = nullptr; fr(st->partial);      // (lo planes live in the hi allocations)
    fr(st->tok_of_row); fr(st->blk_counts);
    for (auto& a : st->act_hi) fr(a);
    for (auto& a : st->act_lo) a = nullptr;
    for (auto& a : st->act_sign) fr(a);
    const int L = p->L;
    int hmax = 1;
    for (int l = 1; l < L; ++l) hmax = std::max(hmax, p->dims[l]);
    st->act_hi.assign(L + 1, nullptr);
    st->act_lo.assign(L + 1, nullptr);
    st->act_sign.assign(L + 1, nullptr);
    // +256 B of slack: the MN-major 3-D view reads whole 64-column blocks of the last row.  Every operand buffer starts
    // out as zeros: with token compaction the kernels read whole 64 / 128-row blocks past the active rows, and what they
    // find there (zeros, or finite values of an earlier minibatch) is multiplied by exact zeros.
    const size_t slack = 256;
    auto zalloc = [&](__half** q, size_t bytes) -> int {
        PPO_CUDA(cudaMalloc((void**)q, bytes));
        PPO_CUDA(cudaMemsetAsync(*q, 0, bytes, ctx->stream));
        return PPO_OK;
    };
    PPO_TRY(zalloc(&st->x_hi, (size_t)tokens * p->dims[0] * 2 + slack));
    PPO_TRY(zalloc(&st->x_lo, (size_t)tokens * p->dims[0] * 2 + slack));
    // The hi and lo halves of an output pair are two planes of ONE allocation, the lo plane a whole number of rows (of
    // every width the buffer is used with) behind the hi plane: the GEMM epilogues then store both with one 3-D TMA box
    auto lcm = [](int64_t a, int64_t b) { int64_t x = a, y = b; while (y) { const int64_t t = x % y; x = y; y = t; } return a / x * b; };
    auto pair_alloc = [&](__half** hi, __half** lo, size_t bytes, int64_t row_bytes_lcm) -> int {
        const size_t plane = (size_t)round_up((int64_t)bytes, lcm(row_bytes_lcm, 256));
        PPO_TRY(zalloc(hi, 2 * plane));
        *lo = reinterpret_cast<__half*>(reinterpret_cast<char*>(*hi) + plane);
        return PPO_OK;
    };
    int64_t hid_lcm = 2;
    for (int l = 1; l < L; ++l) hid_lcm = lcm(hid_lcm, (int64_t)p->dims[l] * 2);
    for (int l = 1; l < L; ++l) {
        PPO_TRY(pair_alloc(&st->act_hi[l], &st->act_lo[l], (size_t)tokens * p->dims[l] * 2 + slack, (int64_t)p->dims[l] * 2));
        if (l < L - 1) PPO_CUDA(cudaMalloc((void**)&st->act_sign[l], sign_words(tokens, p->dims[l]) * 4));   // (the head gates on sign(hi))
    }
    for (int i = 0; i < 2; ++i)
        PPO_TRY(pair_alloc(&st->dz_hi[i], &st->dz_lo[i], (size_t)tokens * hmax * 2 + slack, hid_lcm));
    PPO_CUDA(cudaMalloc((void**)&st->tok_of_row, (size_t)tokens * sizeof(int)));
    PPO_CUDA(cudaMalloc((void**)&st->blk_counts, (size_t)ceil_div(tokens, TOK_PER_BLOCK) * sizeof(int)));
    if (st->d_rows == nullptr) PPO_CUDA(cudaMalloc((void**)&st->d_rows, sizeof(int)));
    // one region per producer of the backward pass (head, every wgrad, every dgrad's column sums): their reductions are
    // deferred to the end of the pass (fold_flush), so no two of them may share memory
    size_t pb = 256;
    for (int l = 0; l + 1 < L; ++l) pb += round_up(f16_wgrad_partial_bytes(tokens, p->dims[l], p->dims[l + 1], ctx->num_sms), 256);
    for (int l = 1; l + 1 < L; ++l) pb += round_up(f16_colsum_partial_bytes(tokens, p->dims[l]), 256);
    pb += round_up(head16_partial_bytes(tokens, p->dims[L - 1], p->dims[L], ctx->num_sms), 256);
    PPO_CUDA(cudaMalloc((void**)&st->partial, pb));
    st->partial_bytes = pb;
    st->tokens = tokens;
    return PPO_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// interface used by abi.cu
// ---------------------------------------------------------------------------------------------
int f16_prepare(ppo_policy* p) {
    PPO_TRY(load_encode16());
    const int L = p->L;
    static_assert(F_MAX_HIDDEN + 1 <= F_MAX_LAYERS, "layer table too small");
    PPO_REQUIRE(L >= 2 && L - 1 <= F_MAX_HIDDEN, "fp16-split engine: needs 1..%d hidden layers (have %d): the a-priori scale "
                "bounds lose 1e-5 parity beyond that depth; use PPO_GEMM_TF32X3_TC or PPO_GEMM_FP32_SIMT", F_MAX_HIDDEN, L - 1);
    for (int l = 0; l + 1 < L; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        PPO_REQUIRE(N <= F_MAX_WIDTH, "fp16-split engine: hidden width %d > %d", N, F_MAX_WIDTH);
        PPO_REQUIRE(K % 8 == 0 && N % 32 == 0 && (l == 0 || K % 32 == 0),
                    "fp16-split engine: layer %d (%d -> %d) needs in %% 8 == 0 and hidden widths %% 32 == 0; "
                    "use PPO_GEMM_FP32_SIMT for this policy", l, K, N);
    }
    PPO_REQUIRE(p->dims[L] <= 4 && p->dims[L - 1] <= 1024, "fp16-split engine: head %d -> %d needs out <= 4 and in <= 1024",
                p->dims[L - 1], p->dims[L]);
    if (p->f16 == nullptr) p->f16 = new F16State();
    F16State* st = state(p);
    if (st->layers.empty()) {
        st->layers.resize(L);
        for (int l = 0; l + 1 < L; ++l) {      // hidden layers only; the head has its own streaming kernels
            const size_t n = (size_t)p->dims[l] * p->dims[l + 1] * 2;
            F16Layer& ly = st->layers[l];
            PPO_CUDA(cudaMalloc((void**)&ly.W_hi, n));
            PPO_CUDA(cudaMalloc((void**)&ly.W_lo, n));
            PPO_CUDA(cudaMalloc((void**)&ly.WT_hi, n));
            PPO_CUDA(cudaMalloc((void**)&ly.WT_lo, n));
        }
        st->table.L = L;
        for (int l = 0; l < L; ++l) {
            st->table.K[l] = p->dims[l]; st->table.N[l] = p->dims[l + 1];
            st->table.w_off[l] = p->w_off[l]; st->table.b_off[l] = p->b_off[l];
        }
        const size_t nsc = (size_t)2 * (3 * L), nst = (size_t)2 + 4 * L;
        PPO_CUDA(cudaMalloc((void**)&st->sc, nsc * 4));
        PPO_CUDA(cudaMalloc((void**)&st->st, nst * 4));
        PPO_CUDA(cudaMemsetAsync(st->sc, 0, nsc * 4, p->ctx->stream));
        PPO_CUDA(cudaMemsetAsync(st->st, 0, nst * 4, p->ctx->stream));
        PPO_CUDA(cudaMalloc((void**)&st->bar, 64));
        PPO_CUDA(cudaMemsetAsync(st->bar, 0, 64, p->ctx->stream));
    }
    return PPO_OK;
}

// optimiser step (opt != nullptr) + weight statistics + scales + fp16 operand copies in ONE launch (see
// f16_adam_refresh_kernel).  xv != nullptr: the gradient is the rank-ordered sum of the peers' published copies.
// d_step: minibatch counter to advance at the end (CUDA-graph replay of the epoch loop), or nullptr.
int f16_adam_refresh(ppo_policy* p, ppo_opt* opt, const P2PView* xv, int* d_step) {
    F16State* st = state(p);
    PPO_REQUIRE(st != nullptr, "fp16-split engine not prepared");
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    F16RefreshArgs a{};
    a.t = st->table;
    for (int l = 0; l + 1 < L; ++l) {
        a.W_hi[l] = st->layers[l].W_hi; a.W_lo[l] = st->layers[l].W_lo;
        a.WT_hi[l] = st->layers[l].WT_hi; a.WT_lo[l] = st->layers[l].WT_lo;
    }
    a.params = p->params; a.wstats = st->st + 2; a.sc = st->sc; a.bar = st->bar;
    if (opt != nullptr) {
        a.m = opt->m; a.v = opt->v; a.grads = p->grads; a.P = p->P;
        a.eta = opt->eta; a.b1 = opt->beta1; a.b2 = opt->beta2; a.eps = opt->eps; a.bp = opt->d_bp;
    }
    if (xv != nullptr) {
        PPO_REQUIRE(opt != nullptr, "fused refresh: the peer-memory exchange needs the optimiser");
        a.peer_xchg = xv->peer_xchg; a.Ppad = xv->Ppad; a.flags = xv->flags; a.p2p_state = xv->state; a.nranks = xv->nranks;
        a.grads_out = p->grads;
    }
    a.d_step = d_step;
    // every CTA must be resident at once (grid-wide barriers): a few per SM (the Adam phase is latency-bound Float64
    // maths, more resident warps hide it), never more than the occupancy calculator grants
    static int per_sm = 0;
    if (per_sm == 0) {
        int occ = 0;
        PPO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f16_adam_refresh_kernel, 256, 0));
        per_sm = std::max(1, std::min(4, occ));
    }
    // (small policies: fewer CTAs make the two grid barriers cheaper than the work they separate)
    const int64_t ctas = std::min<int64_t>((int64_t)ctx->num_sms * per_sm, std::max<int64_t>(32, ceil_div(p->P, 1024)));
    f16_adam_refresh_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int f16_refresh_weights(ppo_policy* p) { return f16_adam_refresh(p, nullptr, nullptr, nullptr); }

int f16_forward(ppo_policy* p, const float* X, int64_t M, const float* mask, const unsigned* feat_bound) {
    F16State* st = state(p);
    PPO_REQUIRE(st != nullptr, "fp16-split engine not prepared");
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    PPO_TRY(ensure_f16_workspace(p, p->ws_tokens > M ? p->ws_tokens : M));
    // a bound of |X| is all the plan needs: the rollout buffer's running abs-max when it supplies one, else a pass over X
    if (feat_bound == nullptr) PPO_TRY(launch_absmax(ctx, X, M * p->dims[0], st->st + 0));
    PlanFwd pf{L, p->slope, st->st + 0, feat_bound, st->st + 2, st->sc, st->st + 1};
    // token compaction (mask: the minibatch's [rows][nhe * apa] action mask, or nullptr = run every token)
    st->compact = mask != nullptr && p->compact_tokens != 0 && M >= 1;
    if (!st->compact) {       // (compacted: token_compact_kernel evaluates the plan)
        plan_fwd_kernel<<<1, 32, 0, ctx->stream>>>(pf);
        ctx->launches += 1;
    }
    const int* M_dev = st->compact ? st->d_rows : nullptr;
    const int* tok = st->compact ? st->tok_of_row : nullptr;
    if (st->compact) {
        const int apa = p->dims[L];
        const unsigned nblk = (unsigned)ceil_div(M, TOK_PER_BLOCK);
        token_count_kernel<<<nblk, TOK_THREADS, 0, ctx->stream>>>(mask, M, apa, st->blk_counts);
        token_compact_kernel<<<nblk, TOK_THREADS, 0, ctx->stream>>>(mask, M, apa, st->blk_counts, st->tok_of_row, st->d_rows, pf);
        const int K8 = p->dims[0] / 8;
        const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(M * K8, 256), (int64_t)ctx->num_sms * 16));
        split16_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(X, tok, M_dev, K8, st->x_hi, st->x_lo, st->sc + 2 * sc_act(0), M);
        ctx->launches += 3;
        PPO_CUDA(cudaGetLastError());
    } else {
        PPO_TRY(launch_split16(ctx, X, st->x_hi, st->x_lo, M * p->dims[0], st->sc + 2 * sc_act(0)));
    }
    for (int l = 0; l + 1 < L; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        F16Layer& ly = st->layers[l];
        KK16Params kp{};
        kp.epi = F_EPI_FWD; kp.act = 1; kp.slope = p->slope; kp.bias = p->params + p->b_off[l];
        kp.M_dev = M_dev;
        kp.sc_a = st->sc + 2 * sc_act(l); kp.sc_b = st->sc + 2 * sc_w(L, l); kp.sc_c = st->sc + 2 * sc_act(l + 1);
        const __half* A_hi = (l == 0) ? st->x_hi : st->act_hi[l];
        const __half* A_lo = (l == 0) ? st->x_lo : st->act_lo[l];
        kp.signs_out = st->act_sign[l + 1];      // nullptr for the last hidden layer
        PPO_TRY(kk16_dispatch(ctx, A_hi, A_lo, ly.WT_hi, ly.WT_lo, st->act_hi[l + 1], st->act_lo[l + 1], M, N, K, kp));
    }
    // (compacted: the logits of skipped tokens keep whatever finite value they held; the mask turns them into -Inf)
    return head_fwd16(ctx, st->act_hi[L - 1], st->act_lo[L - 1], p->params + p->w_off[L - 1], p->params + p->b_off[L - 1],
                      p->act[L], M, p->dims[L - 1], p->dims[L], st->sc + 2 * sc_act(L - 1), M_dev, tok);
}

int f16_backward(ppo_policy* p, int64_t M, bool dl_stat_ready) {
    F16State* st = state(p);
    PPO_REQUIRE(st != nullptr, "fp16-split engine not prepared");
    ppo_ctx* ctx = p->ctx;
    const int L = p->L;
    PPO_REQUIRE(M <= st->tokens, "fp16-split engine: backward without a forward of the same minibatch");
    const int* M_dev = st->compact ? st->d_rows : nullptr;
    const int* tok = st->compact ? st->tok_of_row : nullptr;
    // scales of the activation gradients: from max |dlogits|.  dl_stat_ready: the loss kernel already left it in st[1]
    // and the head kernel evaluates the plan itself; otherwise an abs-max pass and the plan kernel run first
    if (!dl_stat_ready) {
        PPO_TRY(launch_absmax(ctx, p->dlogits, M * p->dims[L], st->st + 1));
        plan_bwd_kernel<<<1, 32, 0, ctx->stream>>>(L, p->slope, st->st + 1, st->st + 2, st->sc);
        ctx->launches += 1;
    }
    int pp = 0;
    // The reductions of the partials (split-K, column sums, head) are deferred to two launches at the end of the pass --
    // except under the per-layer overlapped NCCL all-reduce, which wants every layer's gradient as early as possible.
    const bool per_layer = dp_overlap(ctx);
    FoldList fl;
    FoldList* defer = per_layer ? nullptr : &fl;
    char* arena = reinterpret_cast<char*>(st->partial);
    size_t arena_off = 0;
    auto region = [&](size_t bytes) -> float* {
        float* r = reinterpret_cast<float*>(arena + arena_off);
        if (defer != nullptr) arena_off += round_up(bytes, 256);      // (eager reductions reuse the same memory)
        return r;
    };
    // head: dZ_{L-2}, dW_head, db_head and db_{L-2} = colsum(dZ_{L-2})
    {
        const size_t hb = head16_partial_bytes(M, p->dims[L - 1], p->dims[L], ctx->num_sms);
        float* part = region(hb);
        PPO_TRY(head_bwd16(ctx, st->act_hi[L - 1], st->act_lo[L - 1], p->dlogits, p->params + p->w_off[L - 1], st->dz_hi[pp],
                           st->dz_lo[pp], p->grads + p->w_off[L - 1], p->grads + p->b_off[L - 1], p->grads + p->b_off[L - 2], M,
                           p->dims[L - 1], p->dims[L], p->slope, part, st->partial_bytes - (size_t)((char*)part - arena),
                           st->sc + 2 * sc_act(L - 1), st->sc + 2 * sc_dz(L, L - 2), M_dev, tok, defer,
                           dl_stat_ready ? L : 0, st->st + 1, st->st + 2, st->sc));
    }
    // data parallelism: a layer's slice of the flat gradient vector (dW_l, db_l: contiguous in Flux.params order) is
    // all-reduced on the communication stream as soon as it is complete, while the layers below still compute
    if (per_layer) PPO_TRY(grads_ready(ctx, p->grads + p->w_off[L - 1], p->P - p->w_off[L - 1]));
    for (int l = L - 2; l >= 0; --l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        F16Layer& ly = st->layers[l];
        const __half* X_hi = (l == 0) ? st->x_hi : st->act_hi[l];
        const __half* X_lo = (l == 0) ? st->x_lo : st->act_lo[l];
        const float* sc_dy = st->sc + 2 * sc_dz(L, l);
        {
            float* part = region(f16_wgrad_partial_bytes(M, K, N, ctx->num_sms));
            PPO_TRY(wgrad16(ctx, X_hi, X_lo, st->dz_hi[pp], st->dz_lo[pp], p->grads + p->w_off[l], part,
                            st->partial_bytes - (size_t)((char*)part - arena), M, K, N, st->sc + 2 * sc_act(l), sc_dy, M_dev, defer));
        }
        if (per_layer) PPO_TRY(grads_ready(ctx, p->grads + p->w_off[l], p->w_off[l + 1] - p->w_off[l]));      // dW_l and db_l (db_l came from above)
        if (l > 0) {
            float* part = region(f16_colsum_partial_bytes(M, K));
            KK16Params kp{};
            kp.epi = F_EPI_DGRAD; kp.act = 0; kp.slope = p->slope; kp.gate = st->act_sign[l];
            kp.colsum_partial = part;
            kp.M_dev = M_dev;
            kp.sc_a = sc_dy; kp.sc_b = st->sc + 2 * sc_w(L, l); kp.sc_c = st->sc + 2 * sc_dz(L, l - 1);
            // dX[M, K] = dY[M, N] * W[K, N]^T : A = dY (K-major in N), B = W rows (K-major in N)
            PPO_TRY(kk16_dispatch(ctx, st->dz_hi[pp], st->dz_lo[pp], ly.W_hi, ly.W_lo, st->dz_hi[pp ^ 1], st->dz_lo[pp ^ 1], M, K, N, kp));
            const int64_t rows = 4 * ceil_div(M, F_BM);
            PPO_TRY(fold_colsum16(ctx, part, rows, K, part + (size_t)rows * K, p->grads + p->b_off[l - 1], M_dev, defer));
            pp ^= 1;
        }
    }
    PPO_REQUIRE(arena_off <= st->partial_bytes, "fp16-split engine: partial arena too small");
    if (defer != nullptr) PPO_TRY(fold_flush(ctx, fl, M_dev));
    return PPO_OK;
}

// gates of hidden activation l (1..L-1) of the last forward pass -> d_out[M][dims[l]] (1 = positive branch)
int f16_read_gates(ppo_policy* p, int l, int64_t M, uint8_t* d_out) {
    F16State* st = state(p);
    PPO_REQUIRE(st != nullptr && M <= st->tokens, "fp16-split engine: no forward pass of %lld tokens to read gates from", (long long)M);
    ppo_ctx* ctx = p->ctx;
    const int N = p->dims[l];
    const int64_t n = M * N;
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
    uint8_t* rows_out = d_out;
    if (st->compact) {      // gates of the compacted rows go through the (idle) gradient ping-pong buffer, then to token order
        rows_out = reinterpret_cast<uint8_t*>(st->dz_hi[0]);
        PPO_CUDA(cudaMemsetAsync(d_out, PPO_GATE_SKIPPED, (size_t)n, ctx->stream));
    }
    if (l < p->L - 1) gates_from_signbits_kernel<<<blocks, 256, 0, ctx->stream>>>(st->act_sign[l], M, N, rows_out);
    else gates_from_hi_kernel<<<blocks, 256, 0, ctx->stream>>>(st->act_hi[l], n, rows_out);
    ctx->launches += 1;
    if (st->compact) {
        gates_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(rows_out, st->tok_of_row, st->d_rows, N, d_out);
        ctx->launches += 1;
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// the device word the loss kernel accumulates max |dlogits| into (bit pattern of a non-negative float)
unsigned* f16_dlogits_stat(ppo_policy* p) {
    F16State* st = state(p);
    return st != nullptr ? st->st + 1 : nullptr;
}

// active tokens of the last forward pass (-1: it ran every token)
int f16_active_tokens(ppo_policy* p, int64_t* out) {
    F16State* st = state(p);
    *out = -1;
    if (st == nullptr || !st->compact) return PPO_OK;
    int v = 0;
    PPO_CUDA(cudaMemcpyAsync(&v, st->d_rows, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
    PPO_CUDA(cudaStreamSynchronize(p->ctx->stream));
    *out = v;
    return PPO_OK;
}

void f16_destroy(ppo_policy* p) {
    F16State* st = state(p);
    if (!st) return;
    for (auto& ly : st->layers) { fr(ly.W_hi); fr(ly.W_lo); fr(ly.WT_hi); fr(ly.WT_lo); }
    fr(st->x_hi); fr(st->x_lo); fr(st->dz_hi[0]); fr(st->dz_hi[1]); st->dz_lo[0] = st->dz_lo[1] = nullptr; fr(st->partial);      // (lo planes live in the hi allocations)
    for (auto& a : st->act_hi) fr(a);
    for (auto& a : st->act_lo) a = nullptr;
    for (auto& a : st->act_sign) fr(a);
    fr(st->sc); fr(st->st); fr(st->tok_of_row); fr(st->blk_counts); fr(st->d_rows); fr(st->bar);
    delete st;
    p->f16 = nullptr;
}

// ---------------------------------------------------------------------------------------------
// stand-alone entry points for tests and per-kernel benches (device pointers)
// ---------------------------------------------------------------------------------------------
int f16_test_operand(ppo_ctx* ctx, const float* x, __half* hi, __half* lo, int64_t n, float* sc, unsigned* st) {
    PPO_TRY(launch_absmax(ctx, x, n, st));
    scale_from_stat_kernel<<<1, 32, 0, ctx->stream>>>(st, sc);
    ctx->launches += 1;
    return launch_split16(ctx, x, hi, lo, n, sc);
}
int f16_test_weight(ppo_ctx* ctx, const float* W, __half* W_hi, __half* W_lo, __half* WT_hi, __half* WT_lo, int K, int N,
                    float* sc, unsigned* st) {
    PPO_TRY(launch_absmax(ctx, W, (int64_t)K * N, st));
    scale_from_stat_kernel<<<1, 32, 0, ctx->stream>>>(st, sc);
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(K, 32));
    weight_prep16_kernel<<<grid, 256, 0, ctx->stream>>>(W, W_hi, W_lo, WT_hi, WT_lo, K, N, sc);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
int f16_test_set_scale(ppo_ctx* ctx, float* sc, float bound) {
    scale_from_bound_kernel<<<1, 32, 0, ctx->stream>>>(bound, sc);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
int f16_test_fwd(ppo_ctx* ctx, const __half* X_hi, const __half* X_lo, const __half* WT_hi, const __half* WT_lo,
                 const float* bias, __half* Y_hi, __half* Y_lo, uint32_t* Y_sign, int64_t M, int K, int N, int act, float slope,
                 const float* sc_x, const float* sc_w, const float* sc_y) {
    PPO_TRY(load_encode16());
    KK16Params kp{};
    kp.epi = F_EPI_FWD; kp.act = act; kp.slope = slope; kp.bias = bias; kp.signs_out = Y_sign;
    kp.sc_a = sc_x; kp.sc_b = sc_w; kp.sc_c = sc_y;
    return kk16_dispatch(ctx, X_hi, X_lo, WT_hi, WT_lo, Y_hi, Y_lo, M, N, K, kp);
}
int f16_test_dgrad(ppo_ctx* ctx, const __half* dY_hi, const __half* dY_lo, const __half* W_hi, const __half* W_lo,
                   const uint32_t* gate, __half* dX_hi, __half* dX_lo, int64_t M, int K, int N, float slope,
                   float* colsum_scratch, float* colsum_out, const float* sc_dy, const float* sc_w, const float* sc_dx) {
    PPO_TRY(load_encode16());
    KK16Params kp{};
    kp.epi = F_EPI_DGRAD; kp.slope = slope; kp.gate = gate;
    kp.colsum_partial = colsum_out ? colsum_scratch : nullptr;
    kp.sc_a = sc_dy; kp.sc_b = sc_w; kp.sc_c = sc_dx;
    PPO_TRY(kk16_dispatch(ctx, dY_hi, dY_lo, W_hi, W_lo, dX_hi, dX_lo, M, K, N, kp));
    if (colsum_out) {
        const int64_t rows = 4 * ceil_div(M, F_BM);
        PPO_TRY(fold_colsum16(ctx, colsum_scratch, rows, K, colsum_scratch + (size_t)rows * K, colsum_out));
    }
    return PPO_OK;
}
int f16_test_wgrad(ppo_ctx* ctx, const __half* X_hi, const __half* X_lo, const __half* dY_hi, const __half* dY_lo, float* dW,
                   float* partial, size_t partial_bytes, int64_t M, int K, int N, const float* sc_x, const float* sc_dy) {
    PPO_TRY(load_encode16());
    return wgrad16(ctx, X_hi, X_lo, dY_hi, dY_lo, dW, partial, partial_bytes, M, K, N, sc_x, sc_dy);
}
int f16_test_join(ppo_ctx* ctx, const __half* hi, const __half* lo, int64_t n, const float* sc, float* out) {
    join16_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16), 256, 0, ctx->stream>>>(hi, lo, n, sc, out);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
int f16_test_signbits(ppo_ctx* ctx, const float* x, int64_t M, int N, uint32_t* out) {
    PPO_REQUIRE(N % 32 == 0, "signbits: N %% 32");
    signbits_kernel<<<(unsigned)std::min<int64_t>(ceil_div(M * (N / 32), 256), (int64_t)ctx->num_sms * 16), 256, 0, ctx->stream>>>(x, M, N, out);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
size_t f16_test_sign_words(int64_t M, int N) { return sign_words(M, N); }
size_t f16_test_partial_bytes(ppo_ctx* ctx, int64_t M, int K, int N) { return f16_partial_bytes(M, K, N, ctx->num_sms); }
int f16_test_head_fwd(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* W, const float* bias, float* logits,
                      int64_t M, int K, int N, const float* sc_h) {
    return head_fwd16(ctx, H_hi, H_lo, W, bias, logits, M, K, N, sc_h);
}
int f16_test_head_bwd(ppo_ctx* ctx, const __half* H_hi, const __half* H_lo, const float* dlogits, const float* W,
                      __half* dH_hi, __half* dH_lo, float* dW, float* db, float* db_below, int64_t M, int K, int N,
                      float slope, float* partial, size_t partial_bytes, const float* sc_h, const float* sc_dh) {
    return head_bwd16(ctx, H_hi, H_lo, dlogits, W, dH_hi, dH_lo, dW, db, db_below, M, K, N, slope, partial, partial_bytes, sc_h, sc_dh);
}
size_t f16_test_head_partial_bytes(int64_t M, int K, int N) { return head16_partial_bytes(M, K, N, 148); }

}  // namespace ppo
