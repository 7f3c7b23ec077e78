/* CPU oracle (plain C) for the serial / byte-moving parts of the PPO-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs (through oracle/c_oracle.py).  Never loaded by the
 * product package.  Parity status: see the header of oracle/ppo_oracle.py.
 *
 * Each function cites the reference file:line it follows (relative to /root/reference).
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libppo_oracle.so oracle/ppo_oracle_c.c
 * (-O2 without -ffast-math so the floating-point order is the written order).
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

/* compute_returns -- src/collect_rollouts.jl:26-42.  The carry `v` is Float64 when the
 * discount is a Float64 (Julia promotes `rewards[idx] + discount * v`), Float32 otherwise;
 * rounded to Float32 on the store into `values`. */
void oracle_compute_returns(const float* rewards, const uint8_t* terminal, int64_t n,
                            double discount, int discount_is_f32, float* values) {
    if (discount_is_f32) {
        float g = (float)discount;
        volatile float v = 0.0f;
        for (int64_t i = n - 1; i >= 0; --i) {
            if (terminal[i]) v = 0.0f;
            float t = g * v;
            v = rewards[i] + t;
            values[i] = v;
        }
    } else {
        double v = 0.0;
        for (int64_t i = n - 1; i >= 0; --i) {
            if (terminal[i]) v = 0.0;
            v = (double)rewards[i] + discount * v;
            values[i] = (float)v;
        }
    }
}

/* get_batch -- src/rollout_buffer.jl:117-133 + batch_state test/quad_game_utilities.jl:26-33:
 * copy the records named by 1-based `indices` into contiguous batch arrays. */
void oracle_get_batch(const float* feat, const float* mask, const int64_t* actions,
                      const float* probs, const float* returns,
                      int64_t feat_elems, int64_t mask_elems,
                      const int64_t* indices1, int64_t nb,
                      float* feat_out, float* mask_out, int64_t* act_out,
                      float* prob_out, float* ret_out) {
    for (int64_t b = 0; b < nb; ++b) {
        int64_t i = indices1[b] - 1;
        memcpy(feat_out + b * feat_elems, feat + i * feat_elems, (size_t)feat_elems * sizeof(float));
        memcpy(mask_out + b * mask_elems, mask + i * mask_elems, (size_t)mask_elems * sizeof(float));
        act_out[b] = actions[i];
        prob_out[b] = probs[i];
        ret_out[b] = returns[i];
    }
}

/* Device-shuffle contract (K3): scalar restatement of oracle/ppo_oracle.py:feistel_permutation.
 * No reference counterpart for the sequence (Julia randperm, src/train.jl:93, is stdlib). */
#define FEISTEL_ROUNDS 10

static uint64_t splitmix64_next(uint64_t* s) {
    *s += 0x9E3779B97F4A7C15ull;
    uint64_t z = *s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

void oracle_feistel_keys(uint64_t seed, uint32_t* keys) {
    uint64_t s = seed;
    for (int r = 0; r < FEISTEL_ROUNDS; ++r) keys[r] = (uint32_t)splitmix64_next(&s);
}

void oracle_feistel_permutation(int64_t n, uint64_t seed, int64_t* perm0) {
    uint32_t keys[FEISTEL_ROUNDS];
    oracle_feistel_keys(seed, keys);
    int bits = 2;
    while (((int64_t)1 << bits) < n) ++bits;
    int abits = bits / 2, bbits = bits - abits;
    uint32_t amask = (1u << abits) - 1u, bmask = (1u << bbits) - 1u;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = (uint64_t)i;
        do {
            uint32_t a = (uint32_t)(x >> bbits), b = (uint32_t)x & bmask;
            for (int r = 0; r < FEISTEL_ROUNDS; ++r) {
                if ((r & 1) == 0) a ^= fmix32(b ^ keys[r]) & amask;
                else              b ^= fmix32(a ^ keys[r]) & bmask;
            }
            x = ((uint64_t)a << bbits) | b;
        } while (x >= (uint64_t)n);
        perm0[i] = (int64_t)x;
    }
}

/* Flux.Optimise.Adam element update (ASSUMED formula, see oracle/ppo_oracle.py:Adam). */
void oracle_adam(float* x, float* mt, float* vt, const float* g, int64_t n,
                 double eta, double b1, double b2, double eps, double b1p, double b2p) {
    for (int64_t i = 0; i < n; ++i) {
        double gi = (double)g[i];
        mt[i] = (float)(b1 * (double)mt[i] + (1.0 - b1) * gi);
        vt[i] = (float)(b2 * (double)vt[i] + (1.0 - b2) * gi * gi);
        float d = (float)((double)mt[i] / (1.0 - b1p) / (__builtin_sqrt((double)vt[i] / (1.0 - b2p)) + eps) * eta);
        x[i] = x[i] - d;
    }
}
