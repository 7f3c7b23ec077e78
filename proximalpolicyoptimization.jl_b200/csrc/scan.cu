// scan.cu — K1: segmented reverse scan for discounted returns; K2: advantage statistics.
//
// Replaces compute_returns (reference src/collect_rollouts.jl:26-42), called in place on
// `rollouts.rewards` by compute_state_value! (src/rollout_buffer.jl:55-64):
//     v <- rewards[i] + discount * (terminal[i] ? 0 : v_next)        for i = n .. 1
// Each transition is the affine map c -> b_i + a_i c with (a_i, b_i) = (terminal_i ? 0 : g, r_i);
// the scan composes maps right to left.  One pass over HBM (read 4 B reward + 1 B terminal,
// write 4 B return = 9 B/transition): single-pass chained scan with decoupled look-back whose
// tile is ONE WARP's 512 transitions; warps are persistent and never synchronise with each other
// (see returns_scan_kernel).  Tile ids count from the RIGHT end (the scan runs right to left), so a
// tile only waits on lower-numbered tiles, which co-resident warps with lower ids process first or
// at the same time.  The carry is Float64 exactly like Julia's promoted `v` (the reference passes
// a Float64 discount everywhere); inside a thread's 16-item chunk the recurrence is the
// reference's serial loop with unfused multiply/add, so any chunk that starts right of an
// episode end is bit-identical to the serial result, and with discount == 1 and integer
// rewards every value is exact.
//
// The kernel runs out of place (the rollout buffer swaps its reward/return arrays afterwards, which is what
// `rollouts.rewards .= compute_returns(...)` amounts to).
// Memory layout: rewards float[n], terminal uint8[n]; lane l of a warp owns 16 consecutive
// transitions = 4 x cp.async 16 B + 1 x cp.async 16 B in flight one tile ahead, 4 x STG.128 out.
#include <algorithm>

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

// Decoupled look-back of tile t over the tiles to its right (ids t-1, t-2, ...), executed by one whole warp: lane l
// inspects tile base - l, so one round covers 32 predecessors.  The round ends at the nearest tile with a published
// inclusive value (lane k): carry = (agg_0 o ... o agg_{k-1})(incl_k); without one the 32 aggregates are composed and the
// window moves 32 tiles further.  It runs only inside episodes longer than a tile plus the look-ahead window.
__device__ __forceinline__ double warp_lookback(const ScanScratch sc, int t, int lane) {
    Map cur{1.0, 0.0};
    int base = t - 1;
    for (;;) {
        const int j = base - lane;
        int f;
        unsigned term;
        for (;;) {
            f = j >= 0 ? ld_acquire_i32(sc.flags + j) : 2;     // (tile 0 always publishes an inclusive value)
            term = __ballot_sync(0xffffffffu, f == 2);
            const unsigned need = term ? ((1u << (__ffs(term) - 1)) - 1u) : 0xffffffffu;   // lanes nearer than the terminator
            const unsigned pending = __ballot_sync(0xffffffffu, f == 0) & need;
            if (!pending) break;
            __nanosleep(20);
        }
        const int k = term ? (__ffs(term) - 1) : 32;
        Map m{1.0, 0.0};
        if (lane < k) { m.A = __ldcg(sc.aggA + j); m.B = __ldcg(sc.aggB + j); }
        else if (lane == k) { m.A = 0.0; m.B = j >= 0 ? __ldcg(sc.incl + j) : 0.0; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Map o;
            o.A = __shfl_down_sync(0xffffffffu, m.A, d);
            o.B = __shfl_down_sync(0xffffffffu, m.B, d);
            if (lane + d < 32) m = compose(m, o);
        }
        m.A = __shfl_sync(0xffffffffu, m.A, 0);
        m.B = __shfl_sync(0xffffffffu, m.B, 0);
        cur = compose(cur, m);
        if (cur.A == 0.0) return cur.B;      // reached an inclusive value (or an episode end)
        base -= 32;
    }
}

constexpr int SCAN_LOOK = 128;   // transitions right of a warp tile inspected for an episode end (4 per lane)
constexpr int SCAN_WT = 32 * SCAN_ITEMS;       // transitions per warp tile (512)
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

// float offset of 16-byte piece j (0..3) of chunk c (0..31) inside a warp's region (chunks padded to 20 floats)
__device__ __forceinline__ int sx_off(int c, int j) { return c * 20 + j * 4; }

// one pipeline stage of ONE WARP in shared memory: its 512 rewards in the padded blocked arrangement and its terminals
struct __align__(16) WarpStage {
    float x[32 * 20];        // 32 chunks of 16 items padded to 20 floats (conflict-free 128-bit accesses)
    uint8_t t[SCAN_WT];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent WARPS: the scan tile is one warp's 512 transitions, and a warp never synchronises with the other warps of its
// CTA (the earlier CTA-wide tile spent 28 % of its issue cycles at the per-tile __syncthreads: profiles/r02_scan_ncu.md).
// Warp g scans warp tiles g, g + W, g + 2W, ... (ids count from the RIGHT end, the direction of the scan; W = all
// co-resident warps).  While a tile is being scanned the cp.async copies of the warp's next tile are in flight.
//
// Where a tile's incoming carry comes from:
//   * look-ahead: the 128 transitions right of the tile (prefetched into registers one tile ahead), composed into one
//     affine map.  With episodic data an episode end almost always lies inside it, and the carry is known without
//     waiting for anybody;
//   * otherwise the decoupled look-back over the tiles to the right (aggregates / inclusive values in global memory).
// A tile publishes its aggregate / inclusive value ONLY IF its own first 128 transitions hold no episode end: exactly
// then its left neighbour's look-ahead misses and reads them.  (A tile whose look-ahead hit publishes its inclusive value
// directly, which ends any look-back that reaches it, so nobody ever waits on a tile that does not publish.)
//
// G1: discount == 1 (every in-tree use of the reference): g * v == v exactly, so the multiplies are dropped
// (bit-identical results, fewer Float64 instructions).
template <bool F32CARRY, bool G1>
__global__ void __launch_bounds__(SCAN_THREADS, 4)
returns_scan_kernel(const float* __restrict__ rew, float* __restrict__ out, const uint8_t* __restrict__ term, int64_t n,
                    double g, double g8, int tiles, ScanScratch sc, double* __restrict__ tile_stats) {
    extern __shared__ __align__(16) unsigned char scan_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpStage* stages = reinterpret_cast<WarpStage*>(scan_smem) + 2 * warp;      // this warp's two stages

    // look-ahead window of a tile, one register set ahead of its use
    float4 la_r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    uint32_t la_t = 0u;
    // asynchronous copies of warp tile t into a stage (full tiles only; a partial tile is read directly below) and the
    // plain loads of its look-ahead window
    auto issue = [&](int t, WarpStage& st, float4& lr, uint32_t& lt) {
        const int64_t tile_base = (int64_t)(tiles - 1 - t) * SCAN_WT;
        if (tile_base + SCAN_WT <= n) {
            const float* src = rew + tile_base;
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)      // element e = 128 q + 4 lane of the warp's 512
                cp_async16(&st.x[sx_off(q * 8 + (lane >> 2), lane & 3)], src + q * 128 + 4 * lane);
            cp_async16(&st.t[lane * SCAN_ITEMS], term + tile_base + (int64_t)lane * SCAN_ITEMS);
            const int64_t e0 = tile_base + SCAN_WT + 4 * lane;
            if (e0 + 4 <= n) {
                lr = __ldg(reinterpret_cast<const float4*>(rew + e0));
                lt = __ldg(reinterpret_cast<const uint32_t*>(term + e0));
            }
        }
        cp_async_commit();
    };

    const int wstride = (int)gridDim.x * SCAN_WARPS;
    int t = (int)blockIdx.x * SCAN_WARPS + warp;
    float4 nx_r = la_r; uint32_t nx_t = la_t;
    if (t < tiles) issue(t, stages[0], nx_r, nx_t);
    for (int it = 0; t < tiles; t += wstride, ++it) {
        WarpStage& st = stages[it & 1];
        la_r = nx_r; la_t = nx_t;
        const int t_next = t + wstride;
        if (t_next < tiles) { issue(t_next, stages[(it + 1) & 1], nx_r, nx_t); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();

        const int tile_idx = tiles - 1 - t;         // position from the left
        const int64_t tile_base = (int64_t)tile_idx * SCAN_WT;
        const bool full_tile = tile_base + SCAN_WT <= n;
        const int64_t base = tile_base + (int64_t)lane * SCAN_ITEMS;
        float* wx = st.x;

        float r[SCAN_ITEMS];
        uint32_t tm[SCAN_ITEMS / 4];
        if (full_tile) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(wx + sx_off(lane, q));
                r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
            const uint4 tv = *reinterpret_cast<const uint4*>(&st.t[lane * SCAN_ITEMS]);
            tm[0] = tv.x; tm[1] = tv.y; tm[2] = tv.z; tm[3] = tv.w;
        } else {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) tm[q] = 0;
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; ++i) {
                const int64_t idx = base + i;
                const bool in = idx < n;
                r[i] = in ? rew[idx] : 0.0f;
                const uint32_t tb = in ? (term[idx] != 0) : 1u;   // padding behaves like an episode end
                tm[i >> 2] |= tb << (8 * (i & 3));
            }
            // park the rewards in the thread's own stage slot like a full tile's (re-read in step 5)
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)
                *reinterpret_cast<float4*>(wx + sx_off(lane, q)) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        // one bit per item (terminal bytes are 0/1: the append path normalises them)
        const uint32_t tbits = ((tm[0] * 0x01020408u) >> 24 & 0xfu) | (((tm[1] * 0x01020408u) >> 24 & 0xfu) << 4) |
                               (((tm[2] * 0x01020408u) >> 24 & 0xfu) << 8) | (((tm[3] * 0x01020408u) >> 24 & 0xfu) << 12);
        auto is_term = [&](int i) -> bool { return (tbits >> i) & 1u; };

        // 0. look-ahead: the 128 transitions right of the tile, composed into one map (the same value in every lane)
        double lookA, lookB;
        {
            const int64_t e0 = tile_base + SCAN_WT + 4 * lane;
            Map lk{1.0, 0.0};
            float lr[4]; bool lt[4];
            if (full_tile && e0 + 4 <= n) {
                lr[0] = la_r.x; lr[1] = la_r.y; lr[2] = la_r.z; lr[3] = la_r.w;
#pragma unroll
                for (int j = 0; j < 4; ++j) lt[j] = ((la_t >> (8 * j)) & 0xffu) != 0;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool in = e0 + j < n;
                    lr[j] = in ? rew[e0 + j] : 0.0f;
                    lt[j] = in ? (term[e0 + j] != 0) : true;
                }
            }
            if (G1) {
                // discount 1: the carry is the plain sum of the rewards up to and including the first episode end in the
                // window (Float64 sums of <= 128 Float32 rewards are exact unless their exponents span > 29 bits, so the
                // order of the additions does not matter); found with one ballot instead of a scan of affine maps
                const int firstj = lt[0] ? 0 : (lt[1] ? 1 : (lt[2] ? 2 : (lt[3] ? 3 : 4)));
                const unsigned has = __ballot_sync(0xffffffffu, firstj < 4);
                const int fl = has ? (__ffs(has) - 1) : 32;
                const int upto = lane < fl ? 3 : (lane == fl ? firstj : -1);
                double part = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) part += (j <= upto) ? (double)lr[j] : 0.0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                lk.A = has ? 0.0 : 1.0;
                lk.B = part;
            } else {
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    const double a = lt[j] ? 0.0 : g;
                    lk.B = __dadd_rn((double)lr[j], __dmul_rn(a, lk.B));
                    lk.A = a * lk.A;
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    Map o;
                    o.A = __shfl_down_sync(0xffffffffu, lk.A, d);
                    o.B = __shfl_down_sync(0xffffffffu, lk.B, d);
                    if (lane + d < 32) lk = compose(lk, o);
                }
                lk.A = __shfl_sync(0xffffffffu, lk.A, 0);
                lk.B = __shfl_sync(0xffffffffu, lk.B, 0);
            }
            lookA = lk.A; lookB = lk.B;
        }

        // 1. per-thread aggregate map, right to left (B is the serial recurrence with zero carry).  The 16 items
        //    are two independent chains of 8 (instruction-level parallelism for the dependent fp64 mul/add
        //    pairs), combined as left(right(c)).
        constexpr int HALF = SCAN_ITEMS / 2;
        double Bh = 0.0, Bl = 0.0;
#pragma unroll
        for (int i = HALF - 1; i >= 0; --i) {
            const bool t_hi = is_term(i + HALF), t_lo = is_term(i);
            if (G1) {
                Bh = __dadd_rn((double)r[i + HALF], t_hi ? 0.0 : Bh);
                Bl = __dadd_rn((double)r[i], t_lo ? 0.0 : Bl);
            } else {
                Bh = __dadd_rn((double)r[i + HALF], __dmul_rn(t_hi ? 0.0 : g, Bh));
                Bl = __dadd_rn((double)r[i], __dmul_rn(t_lo ? 0.0 : g, Bl));
            }
        }
        const bool th = (tbits >> HALF) != 0u, tl = (tbits & ((1u << HALF) - 1u)) != 0u;
        const double Ah = th ? 0.0 : g8, Al = tl ? 0.0 : g8;
        Map me;
        me.A = Al * Ah;
        me.B = __dadd_rn(Bl, __dmul_rn(Al, Bh));

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
        Map tile;       // the whole warp tile
        tile.A = __shfl_sync(0xffffffffu, inc.A, 0);
        tile.B = __shfl_sync(0xffffffffu, inc.B, 0);

        // 3. incoming carry; publication for the left neighbour (see the kernel comment)
        const bool publish = (__ballot_sync(0xffffffffu, tbits != 0u) & 0xffu) == 0u;     // no episode end in the first 128
        const bool look_hit = (lookA == 0.0);
        if (publish && tile.A == 0.0 && lane == 0) {      // an episode ends inside the tile: its inclusive value needs no carry
            sc.incl[t] = tile.B;
            st_release_i32(sc.flags + t, 2);
        }
        double carry = lookB;       // look-ahead hit: the value of the recurrence at the first transition right of the tile
        if (!look_hit) {
            carry = 0.0;
            if (t > 0) {
                // decoupled look-back over the tiles to the right: lower tile ids, i.e. earlier iterations of co-resident
                // warps (or lower warp ids in the same iteration), which never wait on this one
                if (publish && tile.A != 0.0 && lane == 0) {
                    sc.aggA[t] = tile.A; sc.aggB[t] = tile.B;
                    st_release_i32(sc.flags + t, 1);
                }
                carry = warp_lookback(sc, t, lane);
            }
        }
        if (publish && tile.A != 0.0 && lane == 0) {
            sc.incl[t] = __dadd_rn(tile.B, __dmul_rn(tile.A, carry));
            st_release_i32(sc.flags + t, 2);
        }

        // 4. thread's incoming carry, then the reference's serial recurrence over its 16 items
        double c = __dadd_rn(lane_excl.B, __dmul_rn(lane_excl.A, carry));
        // the left half starts from the value at item 8 = right-half map applied to the carry
        const double c_lo = __dadd_rn(Bh, __dmul_rn(Ah, c));
        // the rewards are re-read from the stage (the thread's own slot) instead of being held in 16 registers across
        // the scan and the look-back: the kernel runs at 64 registers per thread (4 CTAs per SM)
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"((uint32_t)__cvta_generic_to_shared(wx + sx_off(lane, q))));
            r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
        if (F32CARRY) {
            const float gf = (float)g;
            float vh = (float)c, vl = (float)c_lo;
#pragma unroll
            for (int i = HALF - 1; i >= 0; --i) {
                if (is_term(i + HALF)) vh = 0.0f;
                if (is_term(i)) vl = 0.0f;
                vh = __fadd_rn(r[i + HALF], __fmul_rn(gf, vh));
                vl = __fadd_rn(r[i], __fmul_rn(gf, vl));
                r[i + HALF] = vh;
                r[i] = vl;
            }
        } else {
            double vh = c, vl = c_lo;
#pragma unroll
            for (int i = HALF - 1; i >= 0; --i) {
                if (is_term(i + HALF)) vh = 0.0;
                if (is_term(i)) vl = 0.0;
                vh = __dadd_rn((double)r[i + HALF], G1 ? vh : __dmul_rn(g, vh));
                vl = __dadd_rn((double)r[i], G1 ? vl : __dmul_rn(g, vl));
                r[i + HALF] = (float)vh;
                r[i] = (float)vl;
            }
        }
        float lsum = 0.0f, lsq = 0.0f;
        if (full_tile) {
            // back through the (now free) stage for coalesced 128-bit stores
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)
                *reinterpret_cast<float4*>(wx + sx_off(lane, q)) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
            if (tile_stats != nullptr) {
#pragma unroll
                for (int i = 0; i < SCAN_ITEMS; ++i) { lsum += r[i]; lsq = fmaf(r[i], r[i], lsq); }
            }
            __syncwarp();
            float4* op = reinterpret_cast<float4*>(out + tile_base);
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)
                __stcs(op + q * 32 + lane, *reinterpret_cast<const float4*>(wx + sx_off(q * 8 + (lane >> 2), lane & 3)));
            __syncwarp();
        } else {
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; ++i) {
                if (base + i < n) { out[base + i] = r[i]; lsum += r[i]; lsq = fmaf(r[i], r[i], lsq); }
            }
        }

        // 5. K2 statistics of the returns: one {sum, sumsq} pair per warp tile (fp32 partials over its 512 values,
        //    Float64 from there on, folded in a fixed order by norm_finalize_kernel => deterministic)
        if (tile_stats != nullptr) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                lsum += __shfl_down_sync(0xffffffffu, lsum, d);
                lsq += __shfl_down_sync(0xffffffffu, lsq, d);
            }
            if (lane == 0) {
                tile_stats[2 * (int64_t)tile_idx] = (double)lsum;
                tile_stats[2 * (int64_t)tile_idx + 1] = (double)lsq;
            }
        }
    }
}

// K2 statistics as a stand-alone pass over the returns (used when advantage normalisation is switched on after
// compute_returns ran without it): the same partition and the same order of additions as step 6 of the scan kernel.
__global__ void __launch_bounds__(SCAN_THREADS)
returns_stats_kernel(const float* __restrict__ ret, int64_t n, double* __restrict__ tile_stats) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile_idx = blockIdx.x;
    const int64_t base = tile_idx * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
    float lsum = 0.0f, lsq = 0.0f;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) { const float v = ret[base + i]; lsum += v; lsq = fmaf(v, v, lsq); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lsum += __shfl_down_sync(0xffffffffu, lsum, d);
        lsq += __shfl_down_sync(0xffffffffu, lsq, d);
    }
    if (lane == 0) {
        tile_stats[2 * (tile_idx * (SCAN_THREADS / 32) + warp)] = (double)lsum;
        tile_stats[2 * (tile_idx * (SCAN_THREADS / 32) + warp) + 1] = (double)lsq;
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

// warp tiles of n transitions, rounded up to whole 4 096-transition groups (the K2 statistics keep one {sum, sumsq} pair per
// warp tile of every group, so that norm_finalize / returns_stats see the same partition as before)
static inline int64_t scan_warp_tiles(int64_t n) { return ceil_div(n > 0 ? n : 1, SCAN_TILE) * (SCAN_TILE / (32 * SCAN_ITEMS)); }

size_t scan_scratch_bytes(int64_t n) {
    int64_t tiles = scan_warp_tiles(n);
    return 16 + (size_t)round_up(tiles * 4, 16) + (size_t)tiles * 3 * sizeof(double);
}

int launch_returns_scan(ppo_ctx* ctx, const float* reward_in, float* returns_out, const uint8_t* terminal, int64_t n,
                        double discount, int discount_is_f32, double* tile_stats, void* scratch) {
    PPO_REQUIRE(reward_in != returns_out, "returns scan runs out of place (the look-ahead reads its right neighbours' rewards)");
    if (n <= 0) return PPO_OK;
    int64_t tiles = scan_warp_tiles(n);
    PPO_REQUIRE(tiles < (int64_t)1 << 30, "returns scan: too many tiles");
    ScanScratch sc = carve(scratch, tiles);
    PPO_CUDA(cudaMemsetAsync(scratch, 0, 16 + (size_t)round_up(tiles * 4, 16), ctx->stream));
    const double g = discount_is_f32 ? (double)(float)discount : discount;
    double g8 = 1.0;
    for (int i = 0; i < SCAN_ITEMS / 2; ++i) g8 *= g;  // the same left-to-right product a per-item loop would form
    // persistent grid: every CTA must be co-resident (the look-back waits on lower tile ids)
    const size_t smem = 2 * SCAN_WARPS * sizeof(WarpStage);
    int occ = 0;
    const bool g1 = (g == 1.0);
    auto launch = [&](auto kern) -> int {
        PPO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PPO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SCAN_THREADS, smem));
        PPO_REQUIRE(occ >= 1, "returns scan: kernel does not fit on an SM");
        const int64_t grid = std::min<int64_t>(ceil_div(tiles, SCAN_WARPS), (int64_t)ctx->num_sms * occ);
        kern<<<(unsigned)grid, SCAN_THREADS, smem, ctx->stream>>>(reward_in, returns_out, terminal, n, g, g8, (int)tiles, sc,
                                                                  tile_stats);
        return PPO_OK;
    };
    if (discount_is_f32) PPO_TRY(g1 ? launch(returns_scan_kernel<true, true>) : launch(returns_scan_kernel<true, false>));
    else PPO_TRY(g1 ? launch(returns_scan_kernel<false, true>) : launch(returns_scan_kernel<false, false>));
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

int launch_returns_stats(ppo_ctx* ctx, const float* returns, int64_t n, double* tile_stats) {
    if (n <= 0) return PPO_OK;
    returns_stats_kernel<<<(unsigned)ceil_div(n, SCAN_TILE), SCAN_THREADS, 0, ctx->stream>>>(returns, n, tile_stats);
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
