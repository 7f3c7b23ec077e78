// dp_p2p.cu — K8 for data parallelism: gradient all-reduce over NVLink peer memory FUSED into the Adam step.
//
// the only exchange of the sharded PPO update is the sum of the minibatch gradient over the ranks
// (P = 560 644 floats = 2.24 MB at C3), once per minibatch, immediately followed by Flux.update! (src/train.jl:81).
// Through NCCL that message costs ~0.4 ms per minibatch on 8 GPUs (scripts/dp8_matrix.sh), 10x its wire time.  Here every
// rank publishes its gradient in an exchange buffer that its peers map with CUDA IPC, and the Adam kernel itself reads
// the G published copies straight over NVLink / NVSwitch and adds them in rank order:
//
//   p2p_publish_kernel   grads -> xchg[epoch & 1] (local copy), then a system-scope release of `epoch + 1` into the
//                        flag word `flags[rank]` of EVERY rank (peer stores);
//   p2p_adam_kernel      waits (system-scope acquire) until all G flags of this rank show `epoch + 1`, then for every
//                        parameter: g = sum_q xchg_q[epoch & 1][i] (q = 0..G-1, the same order on every rank, so the
//                        weights stay bit-identical across ranks), followed by the Adam maths of adam.cu;
//   p2p_tick_kernel      epoch += 1, beta powers advance.
//
// The exchange buffer is double buffered by the parity of the epoch: a rank overwrites xchg[e & 1] at epoch e + 2, i.e.
// after its Adam of epoch e + 1 saw every peer's flag e + 2, and a peer raises that flag only after its own Adam of
// epoch e (the reader of xchg[e & 1]) has finished — no second handshake is needed.  One process per GPU, one node.
// A spin that exceeds ~4 s (a peer died) raises an error flag instead of hanging the GPU, and the update is skipped.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace ppo {

struct DpP2P {
    int nranks = 0, rank = 0;
    int64_t P = 0, Ppad = 0;
    void* block = nullptr;             // one allocation (one IPC handle): [flags: 256 B][xchg: 2 x Ppad floats]
    unsigned* flags = nullptr;
    float* xchg = nullptr;
    unsigned* d_epoch = nullptr;       // [0] epoch, [1] publish block counter, [2] error flag, [4..7] wait ns / waits (u64 x 2)
    float** d_peer_xchg = nullptr;     // [nranks] device pointers (own entry = local)
    unsigned** d_peer_flags = nullptr;
    std::vector<void*> opened;
    bool connected = false;
};

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
p2p_publish_kernel(const float* __restrict__ grads, int64_t P, float* __restrict__ xchg, int64_t Ppad, unsigned* state,
                   unsigned* const* __restrict__ peer_flags, int nranks, int rank) {
    const unsigned epoch = state[0];
    float* dst = xchg + (size_t)(epoch & 1u) * Ppad;
    const int64_t n4 = P >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(grads)[i];
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) dst[(n4 << 2) + threadIdx.x] = grads[(n4 << 2) + threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(state + 1, 1u);
        if (done == gridDim.x - 1) {          // the last block: everything is written and fenced
            state[1] = 0u;
            __threadfence_system();
            for (int q = 0; q < nranks; ++q) st_release_sys(peer_flags[q] + rank, epoch + 1u);
        }
    }
}

// grads_out != nullptr: also (or, with x == nullptr, only) store the reduced gradient
__global__ void __launch_bounds__(256)
p2p_adam_kernel(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, int64_t n, double eta, double b1, double b2,
                double eps, const double* __restrict__ bp, const float* const* __restrict__ peer_xchg, int64_t Ppad,
                const unsigned* flags, unsigned* state, int nranks, float* __restrict__ grads_out) {
    const unsigned epoch = state[0];
    // A peer that never publishes (it died) must not hang the GPU: after ~4 s the block raises the sticky error word and
    // SKIPS its share of the update (no sum of stale exchange buffers ever reaches the weights or the moments); every later
    // launch sees the word and skips at once, and the host turns it into an error before it reads the losses.
    __shared__ int s_abort;
    if (threadIdx.x == 0) {
        int abort_ = __ldcg(state + 2) != 0u;
        const long long t0 = clock64();
        for (int q = 0; q < nranks && !abort_; ++q)
            while (ld_acquire_sys(flags + q) < epoch + 1u) {
                if (clock64() - t0 > 8000000000ll) { state[2] = 1u; abort_ = 1; break; }
                __nanosleep(64);
            }
        s_abort = abort_;
    }
    __syncthreads();
    if (s_abort) return;
    const size_t off = (size_t)(epoch & 1u) * Ppad;
    const double b1p = bp[0], b2p = bp[1];
    const double om1 = 1.0 - b1, om2 = 1.0 - b2;
    const double c1 = 1.0 - b1p, c2 = 1.0 - b2p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.0f;
        for (int q = 0; q < nranks; ++q) g += __ldcv(peer_xchg[q] + off + i);       // fixed rank order on every rank; uncached (peer memory)
        if (grads_out != nullptr) grads_out[i] = g;
        if (x != nullptr) {
            const double gi = (double)g;
            const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)m[i]), __dmul_rn(om1, gi));
            const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)v[i]), __dmul_rn(__dmul_rn(om2, gi), gi));
            m[i] = mt;
            v[i] = vt;
            const double den = __dadd_rn(sqrt((double)vt / c2), eps);
            const float d = (float)__dmul_rn(((double)mt / c1) / den, eta);
            x[i] = __fsub_rn(x[i], d);
        }
    }
}

__global__ void p2p_tick_kernel(unsigned* state, double* bp, double b1, double b2) {
    state[0] += 1u;
    if (bp != nullptr) { bp[0] *= b1; bp[1] *= b2; }
}

}  // namespace

bool p2p_active(const ppo_policy* p) { return p->dp != nullptr && reinterpret_cast<const DpP2P*>(p->dp)->connected; }

int p2p_export(ppo_policy* p, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (p->dp == nullptr) p->dp = new DpP2P();
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    if (d->block == nullptr) {
        d->P = p->P;
        d->Ppad = round_up(p->P, 64);
        const size_t bytes = 256 + (size_t)2 * d->Ppad * 4;
        PPO_CUDA(cudaMalloc(&d->block, bytes));
        PPO_CUDA(cudaMemset(d->block, 0, bytes));
        d->flags = reinterpret_cast<unsigned*>(d->block);
        d->xchg = reinterpret_cast<float*>((char*)d->block + 256);
        PPO_CUDA(cudaMalloc((void**)&d->d_epoch, 64));
        PPO_CUDA(cudaMemset(d->d_epoch, 0, 64));
    }
    cudaIpcMemHandle_t h;
    PPO_CUDA(cudaIpcGetMemHandle(&h, d->block));
    memcpy(handle64, &h, 64);
    return PPO_OK;
}

int p2p_connect(ppo_policy* p, int nranks, int rank, const void* handles) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    PPO_REQUIRE(d != nullptr && d->block != nullptr, "p2p_connect: call ppo_policy_p2p_export first");
    PPO_REQUIRE(nranks >= 2 && nranks <= 64 && rank >= 0 && rank < nranks, "p2p_connect: rank %d of %d", rank, nranks);
    PPO_REQUIRE(!d->connected, "p2p_connect: already connected");
    std::vector<float*> px((size_t)nranks);
    std::vector<unsigned*> pf((size_t)nranks);
    for (int q = 0; q < nranks; ++q) {
        void* base = d->block;
        if (q != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + (size_t)q * 64, 64);
            PPO_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            d->opened.push_back(base);
        }
        pf[(size_t)q] = reinterpret_cast<unsigned*>(base);
        px[(size_t)q] = reinterpret_cast<float*>((char*)base + 256);
    }
    PPO_CUDA(cudaMalloc((void**)&d->d_peer_xchg, (size_t)nranks * sizeof(float*)));
    PPO_CUDA(cudaMalloc((void**)&d->d_peer_flags, (size_t)nranks * sizeof(unsigned*)));
    PPO_CUDA(cudaMemcpy(d->d_peer_xchg, px.data(), (size_t)nranks * sizeof(float*), cudaMemcpyHostToDevice));
    PPO_CUDA(cudaMemcpy(d->d_peer_flags, pf.data(), (size_t)nranks * sizeof(unsigned*), cudaMemcpyHostToDevice));
    d->nranks = nranks; d->rank = rank;
    d->connected = true;
    p->ctx->p2p_grads = true;
    return PPO_OK;
}

void p2p_destroy(ppo_policy* p) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    if (!d) return;
    for (void* q : d->opened) cudaIpcCloseMemHandle(q);
    if (d->d_peer_xchg) cudaFree(d->d_peer_xchg);
    if (d->d_peer_flags) cudaFree(d->d_peer_flags);
    if (d->d_epoch) cudaFree(d->d_epoch);
    if (d->block) cudaFree(d->block);
    delete d;
    p->dp = nullptr;
}

int p2p_publish(ppo_policy* p, P2PView* view) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    PPO_REQUIRE(d != nullptr && d->connected, "p2p gradient exchange is not connected");
    ppo_ctx* ctx = p->ctx;
    const int64_t n = p->P;
    int64_t blocks = std::min<int64_t>(ceil_div(n, 1024), (int64_t)ctx->num_sms);
    if (blocks < 1) blocks = 1;
    p2p_publish_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p->grads, n, d->xchg, d->Ppad, d->d_epoch, d->d_peer_flags,
                                                                 d->nranks, d->rank);
    ctx->launches += 1;
    PPO_CUDA(cudaGetLastError());
    if (view != nullptr) {
        view->peer_xchg = d->d_peer_xchg; view->Ppad = d->Ppad; view->flags = d->flags; view->state = d->d_epoch;
        view->nranks = d->nranks;
    }
    return PPO_OK;
}

// all-reduce p->grads over the ranks and (opt != nullptr) apply Adam, in one pass over peer memory
int p2p_reduce_and_step(ppo_policy* p, ppo_opt* opt) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    PPO_REQUIRE(d != nullptr && d->connected, "p2p gradient exchange is not connected");
    ppo_ctx* ctx = p->ctx;
    const int64_t n = p->P;
    PPO_TRY(p2p_publish(p, nullptr));
    int64_t ablocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 8);
    if (opt != nullptr) {
        p2p_adam_kernel<<<(unsigned)ablocks, 256, 0, ctx->stream>>>(p->params, opt->m, opt->v, n, opt->eta, opt->beta1, opt->beta2,
                                                                   opt->eps, opt->d_bp, d->d_peer_xchg, d->Ppad, d->flags,
                                                                   d->d_epoch, d->nranks, p->grads);
        p2p_tick_kernel<<<1, 1, 0, ctx->stream>>>(d->d_epoch, opt->d_bp, opt->beta1, opt->beta2);
    } else {
        p2p_adam_kernel<<<(unsigned)ablocks, 256, 0, ctx->stream>>>(nullptr, nullptr, nullptr, n, 0.0, 0.0, 0.0, 0.0, nullptr,
                                                                   d->d_peer_xchg, d->Ppad, d->flags, d->d_epoch, d->nranks,
                                                                   p->grads);
        p2p_tick_kernel<<<1, 1, 0, ctx->stream>>>(d->d_epoch, nullptr, 0.0, 0.0);
    }
    ctx->launches += 2;
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// diagnostic: total time (ns) this rank's optimiser kernel waited for its peers' gradients, and the number of waits
int p2p_wait_stats(ppo_policy* p, int64_t* total_ns, int64_t* waits, int reset) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    *total_ns = 0; *waits = 0;
    if (!d || !d->connected) return PPO_OK;
    unsigned long long v[2] = {0, 0};
    PPO_CUDA(cudaMemcpyAsync(v, d->d_epoch + 4, sizeof(v), cudaMemcpyDeviceToHost, p->ctx->stream));
    PPO_CUDA(cudaStreamSynchronize(p->ctx->stream));
    *total_ns = (int64_t)v[0]; *waits = (int64_t)v[1];
    if (reset) PPO_CUDA(cudaMemsetAsync(d->d_epoch + 4, 0, 16, p->ctx->stream));
    return PPO_OK;
}

// non-zero after a spin timed out (checked by the host at the end of an epoch)
int p2p_check(ppo_policy* p) {
    DpP2P* d = reinterpret_cast<DpP2P*>(p->dp);
    if (!d || !d->connected) return PPO_OK;
    unsigned st[3] = {0, 0, 0};
    PPO_CUDA(cudaMemcpyAsync(st, d->d_epoch, sizeof(st), cudaMemcpyDeviceToHost, p->ctx->stream));
    PPO_CUDA(cudaStreamSynchronize(p->ctx->stream));
    if (st[2] != 0) { set_error("p2p gradient exchange: a peer did not publish its gradient within the time-out"); return PPO_ERR_NCCL; }
    return PPO_OK;
}

}  // namespace ppo
