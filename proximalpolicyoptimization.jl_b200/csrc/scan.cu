// scan.cu — K1: segmented reverse scan for discounted returns; K2: advantage statistics.
//
// Replaces compute_returns (reference src/collect_rollouts.jl:26-42), called in place on
// `rollouts.rewards` by compute_state_value! (src/rollout_buffer.jl:55-64):
//     v <- rewards[i] + discount * (terminal[i] ? 0 : v_next)        for i = n .. 1
// Each transition is the affine map c -> b_i + a_i c with (a_i, b_i) = (terminal_i ? 0 : g, r_i);
// the scan composes maps right to left.  One pass over HBM (read 4 B reward + 1 B terminal,
// write 4 B return = 9 B/transition): single-pass chained scan with decoupled look-back,
// tiles handed out right-to-left by an atomic ticket so a tile only ever waits on tiles that
// are already resident.  The carry is Float64 exactly like Julia's promoted `v` (the reference
// passes a Float64 discount everywhere); inside a thread's 16-item chunk the recurrence is the
// reference's serial loop with unfused multiply/add, so any chunk that starts right of an
// episode end is bit-identical to the serial result, and with discount == 1 and integer
// rewards every value is exact.
//
// Memory layout: rewards float[n], terminal uint8[n]; thread t of a tile owns 16 consecutive
// transitions = 4 x LDG.128 + 1 x LDG.128 in flight, 4 x STG.128 out.
#include "common.cuh"

namespace ppo {

namespace {

struct Map {  // c -> B + A*c
    double A, B;
};

__device__ __forceinline__ Map compose(const Map& left, const Map& right) {
    // left(right(c)) = left.B + left.A*(right.B + right.A*c)
    Map m;
    m.A = left.A * right.A;
    m.B = __dadd_rn(left.B, __dmul_rn(left.A, right.B));
    return m;
}

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
    asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct ScanScratch {
    int* counter;
    int* flags;     // 0 = empty, 1 = aggregate published, 2 = inclusive value published
    double* aggA;
    double* aggB;
    double* incl;
};

__host__ __device__ inline ScanScratch carve(void* scratch, int64_t tiles) {
    ScanScratch s;
    char* p = (char*)scratch;
    s.counter = (int*)p;
    s.flags = (int*)(p + 16);
    size_t off = 16 + (size_t)((tiles * 4 + 15) / 16) * 16;
    s.aggA = (double*)(p + off);
    s.aggB = s.aggA + tiles;
    s.incl = s.aggB + tiles;
    return s;
}

template <bool F32CARRY>
__global__ void __launch_bounds__(SCAN_THREADS)
returns_scan_kernel(float* __restrict__ rew, const uint8_t* __restrict__ term, int64_t n, double g,
                    int tiles, ScanScratch sc, double* __restrict__ tile_stats) {
    __shared__ int s_tile;
    __shared__ double sA[SCAN_THREADS / 32], sB[SCAN_THREADS / 32];
    __shared__ double s_carry;
    __shared__ double s_sum[SCAN_THREADS / 32], s_sq[SCAN_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(sc.counter, 1);
    __syncthreads();
    const int t = s_tile;                       // ticket: 0 = right-most tile
    const int tile_idx = tiles - 1 - t;         // position from the left
    const int64_t base = (int64_t)tile_idx * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;

    float r[SCAN_ITEMS];
    uint32_t tm[SCAN_ITEMS / 4];
    if (base + SCAN_ITEMS <= n) {
        const float4* rp = reinterpret_cast<const float4*>(rew + base);
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            float4 v = __ldcs(rp + q);
            r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
        uint4 tv = __ldcs(reinterpret_cast<const uint4*>(term + base));
        tm[0] = tv.x; tm[1] = tv.y; tm[2] = tv.z; tm[3] = tv.w;
    } else {
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) tm[q] = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int64_t idx = base + i;
            bool in = idx < n;
            r[i] = in ? rew[idx] : 0.0f;
            uint32_t tb = in ? (term[idx] != 0) : 1u;   // padding behaves like an episode end
            tm[i >> 2] |= tb << (8 * (i & 3));
        }
    }
    auto is_term = [&](int i) -> bool { return ((tm[i >> 2] >> (8 * (i & 3))) & 0xffu) != 0; };

    // 1. per-thread aggregate map, right to left (B is the serial recurrence with zero carry)
    Map me{1.0, 0.0};
#pragma unroll
    for (int i = SCAN_ITEMS - 1; i >= 0; --i) {
        double a = is_term(i) ? 0.0 : g;
        me.B = __dadd_rn((double)r[i], __dmul_rn(a, me.B));
        me.A = a * me.A;
    }

    // 2. reverse inclusive scan over the lanes of the warp (lane l: lanes l..31)
    Map inc = me;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Map o;
        o.A = __shfl_down_sync(0xffffffffu, inc.A, d);
        o.B = __shfl_down_sync(0xffffffffu, inc.B, d);
        if (lane + d < 32) inc = compose(inc, o);
    }
    Map lane_excl;  // lanes l+1..31
    lane_excl.A = __shfl_down_sync(0xffffffffu, inc.A, 1);
    lane_excl.B = __shfl_down_sync(0xffffffffu, inc.B, 1);
    if (lane == 31) { lane_excl.A = 1.0; lane_excl.B = 0.0; }
    if (lane == 0) { sA[warp] = inc.A; sB[warp] = inc.B; }
    __syncthreads();

    // 3. maps of the warps to the right of this one (warp+1 .. last)
    constexpr int NW = SCAN_THREADS / 32;
    Map warp_excl{1.0, 0.0};
    for (int k = NW - 1; k > warp; --k) warp_excl = compose(Map{sA[k], sB[k]}, warp_excl);

    // 4. tile aggregate, publication and look-back (thread 0)
    if (tid == 0) {
        Map tile = compose(Map{sA[0], sB[0]}, warp_excl);
        double carry = 0.0;
        if (t == 0) {
            sc.incl[0] = tile.B;
            __threadfence();
            st_release_i32(sc.flags + 0, 2);
        } else {
            if (tile.A == 0.0) {   // an episode ends inside the tile: inclusive value known already
                sc.incl[t] = tile.B;
                __threadfence();
                st_release_i32(sc.flags + t, 2);
            } else {
                sc.aggA[t] = tile.A; sc.aggB[t] = tile.B;
                __threadfence();
                st_release_i32(sc.flags + t, 1);
            }
            Map cur{1.0, 0.0};
            int j = t - 1;
            while (true) {
                int f;
                while ((f = ld_acquire_i32(sc.flags + j)) == 0) { __nanosleep(20); }
                if (f == 2) { carry = __dadd_rn(cur.B, __dmul_rn(cur.A, __ldcg(sc.incl + j))); break; }
                cur = compose(cur, Map{__ldcg(sc.aggA + j), __ldcg(sc.aggB + j)});
                if (cur.A == 0.0) { carry = cur.B; break; }
                --j;   // j >= 0 always holds: ticket 0 publishes an inclusive value
            }
            if (tile.A != 0.0) {
                sc.incl[t] = __dadd_rn(tile.B, __dmul_rn(tile.A, carry));
                __threadfence();
                st_release_i32(sc.flags + t, 2);
            }
        }
        s_carry = carry;
    }
    __syncthreads();

    // 5. thread's incoming carry, then the reference's serial recurrence over its 16 items
    double c = s_carry;
    c = __dadd_rn(warp_excl.B, __dmul_rn(warp_excl.A, c));
    c = __dadd_rn(lane_excl.B, __dmul_rn(lane_excl.A, c));
    double lsum = 0.0, lsq = 0.0;
    if (F32CARRY) {
        float gf = (float)g;
        float v = (float)c;
#pragma unroll
        for (int i = SCAN_ITEMS - 1; i >= 0; --i) {
            if (is_term(i)) v = 0.0f;
            v = __fadd_rn(r[i], __fmul_rn(gf, v));
            r[i] = v;
        }
    } else {
        double v = c;
#pragma unroll
        for (int i = SCAN_ITEMS - 1; i >= 0; --i) {
            if (is_term(i)) v = 0.0;
            v = __dadd_rn((double)r[i], __dmul_rn(g, v));
            r[i] = (float)v;
        }
    }
    if (base + SCAN_ITEMS <= n) {
        float4* op = reinterpret_cast<float4*>(rew + base);
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            op[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { double x = r[i]; lsum += x; lsq += x * x; }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            if (base + i < n) { rew[base + i] = r[i]; double x = r[i]; lsum += x; lsq += x * x; }
        }
    }

    // 6. K2 statistics of the returns in this tile (fixed-order reduction => deterministic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lsum += __shfl_down_sync(0xffffffffu, lsum, d);
        lsq += __shfl_down_sync(0xffffffffu, lsq, d);
    }
    if (lane == 0) { s_sum[warp] = lsum; s_sq[warp] = lsq; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < NW; ++k) { a += s_sum[k]; b += s_sq[k]; }
        tile_stats[2 * (int64_t)tile_idx] = a;
        tile_stats[2 * (int64_t)tile_idx + 1] = b;
    }
}

// K2: fold the per-tile {sum, sumsq} into {mean, 1/(std+eps)} (population std, Float64).
__global__ void __launch_bounds__(1024)
norm_finalize_kernel(const double* __restrict__ tile_stats, int64_t tiles, int64_t n, double eps,
                     float* __restrict__ out2) {
    __shared__ double s1[1024], s2[1024];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < tiles; i += 1024) { a += tile_stats[2 * i]; b += tile_stats[2 * i + 1]; }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { s1[threadIdx.x] += s1[threadIdx.x + d]; s2[threadIdx.x] += s2[threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double mu = s1[0] / (double)n;
        double var = s2[0] / (double)n - mu * mu;
        if (var < 0.0) var = 0.0;
        out2[0] = (float)mu;
        out2[1] = (float)(1.0 / (sqrt(var) + eps));
    }
}

}  // namespace

size_t scan_scratch_bytes(int64_t n) {
    int64_t tiles = ceil_div(n > 0 ? n : 1, SCAN_TILE);
    return 16 + (size_t)round_up(tiles * 4, 16) + (size_t)tiles * 3 * sizeof(double);
}

int launch_returns_scan(ppo_ctx* ctx, float* reward_inout, const uint8_t* terminal, int64_t n,
                        double discount, int discount_is_f32, double* tile_stats, void* scratch) {
    if (n <= 0) return PPO_OK;
    int64_t tiles = ceil_div(n, SCAN_TILE);
    PPO_REQUIRE(tiles < (int64_t)1 << 30, "returns scan: too many tiles");
    ScanScratch sc = carve(scratch, tiles);
    PPO_CUDA(cudaMemsetAsync(scratch, 0, 16 + (size_t)round_up(tiles * 4, 16), ctx->stream));
    if (discount_is_f32)
        returns_scan_kernel<true><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(
            reward_inout, terminal, n, (double)(float)discount, (int)tiles, sc, tile_stats);
    else
        returns_scan_kernel<false><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(
            reward_inout, terminal, n, discount, (int)tiles, sc, tile_stats);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_norm_finalize(ppo_ctx* ctx, const double* tile_stats, int64_t n_tiles, int64_t n, double eps,
                         float* d_norm) {
    norm_finalize_kernel<<<1, 1024, 0, ctx->stream>>>(tile_stats, n_tiles, n, eps, d_norm);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
