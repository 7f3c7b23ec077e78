// adam.cu — K8: fused Adam step over the flat parameter vector.
//
// Replaces Flux.update!(optimizer, weights, grad) (reference call site src/train.jl:81) for
// Flux.Optimise.Adam (ASSUMED formula — Flux is not vendored; see oracle/ppo_oracle.py:Adam):
//     mt = b1 mt + (1-b1) g ;  vt = b2 vt + (1-b2) g^2
//     x -= mt / (1 - b1^t) / (sqrt(vt / (1 - b2^t)) + eps) * eta
// eta/beta/eps are Float64 scalars broadcast against Float32 arrays in Julia, so every
// element-wise expression is evaluated in Float64 (unfused, like the CPU) and rounded to Float32
// when stored into mt / vt / delta.  28 B/parameter (read w,g,m,v; write w,m,v); P <= ~0.6 M so the
// kernel is latency-bound, which is why beta^t lives in device memory (graph-capturable, no
// host round trip) and is advanced by a trailing one-thread kernel.
#include "common.cuh"

namespace ppo {
namespace {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            int64_t n, double eta, double b1, double b2, double eps, const double* __restrict__ bp,
            float grad_scale) {
    const double b1p = bp[0], b2p = bp[1];
    const double om1 = 1.0 - b1, om2 = 1.0 - b2;
    const double c1 = 1.0 - b1p, c2 = 1.0 - b2p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double gi = (double)(grad_scale == 1.0f ? g[i] : g[i] * grad_scale);
        const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)m[i]), __dmul_rn(om1, gi));
        const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)v[i]), __dmul_rn(__dmul_rn(om2, gi), gi));
        m[i] = mt;
        v[i] = vt;
        const double den = __dadd_rn(sqrt((double)vt / c2), eps);
        const float d = (float)__dmul_rn(((double)mt / c1) / den, eta);
        x[i] = __fsub_rn(x[i], d);
    }
}

__global__ void adam_tick_kernel(double* bp, double b1, double b2) {
    bp[0] *= b1;
    bp[1] *= b2;
}

}  // namespace

int launch_adam(ppo_ctx* ctx, float* x, float* m, float* v, const float* g, int64_t n, double eta, double b1,
                double b2, double eps, double* d_bp, float grad_scale) {
    if (n <= 0) return PPO_OK;
    int64_t blocks = ceil_div(n, 256);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, m, v, g, n, eta, b1, b2, eps, d_bp, grad_scale);
    adam_tick_kernel<<<1, 1, 0, ctx->stream>>>(d_bp, b1, b2);
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo
