/* ppo_b200.h — C ABI of libppo_b200.so: the B200-native PPO-update hot path of
 * ProximalPolicyOptimization.jl (rollout buffer -> returns scan -> shuffle -> minibatch
 * gather -> policy MLP fwd/bwd + fused masked-softmax PPO loss -> Adam [-> NCCL all-reduce]).
 *
 * The reference has NO FFI boundary (pure Julia, multiple dispatch).  Each entry point below
 * names the reference function it replaces (file:line relative to the reference repo); the
 * Julia-side `ccall` bindings a maintainer would add are in INTEGRATION.md and
 * proximalpolicyoptimization.jl_b200/julia/PPOB200.jl.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ppo_status on error; the message is
 *     available from ppo_last_error() (thread-local).  Nothing throws, nothing calls back.
 *   - host pointers are borrowed for the duration of the call only.
 *   - at the boundary actions / indices / permutations are Int64 and 1-BASED (Julia);
 *     `terminal` is one byte per transition (Julia Bool); masks are Float32 0 / -Inf;
 *     losses are Float64 (the reference's promotion through a Float64 epsilon).
 *   - array layouts are the bytes of the Julia arrays:
 *       vertex_score [nf, nhe, n]  (column-major)  ==  float feat[n][nhe][nf]
 *       action_mask  [A, n]                         ==  float mask[n][A],  A = nhe*apa
 *       Dense.weight [out, in]                      ==  float W[in][out]
 *   - one ppo_ctx per GPU and per process; calls on one ctx must come from one thread at a
 *     time.  Work is enqueued on the ctx's stream; calls that return data synchronise.
 */
#ifndef PPO_B200_H
#define PPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppo_ctx ppo_ctx;
typedef struct ppo_buf ppo_buf;
typedef struct ppo_policy ppo_policy;
typedef struct ppo_opt ppo_opt;

typedef enum {
    PPO_OK = 0,
    PPO_ERR_INVALID = -1,   /* bad argument / shape / range (the reference's @assert sites) */
    PPO_ERR_CUDA = -2,      /* CUDA runtime error */
    PPO_ERR_NCCL = -3,      /* NCCL error or NCCL not loadable */
    PPO_ERR_STATE = -4,     /* call not valid in the current state (e.g. no permutation set) */
    PPO_ERR_NOMEM = -5
} ppo_status;

/* GEMM engine for the policy MLP (the only dense contraction, test/policy.jl:11-15). */
typedef enum {
    PPO_GEMM_AUTO = -1,       /* fastest fp32-parity engine whose shape contract the policy meets: F16X3, else TF32X3, else FP32 */
    PPO_GEMM_FP32_SIMT = 0,   /* fp32 FFMA tiles: bit-for-bit deterministic fp32 reference path */
    PPO_GEMM_TF32X3_TC = 1,   /* tcgen05 kind::tf32, error-compensated 3-pass split: fp32 parity */
    PPO_GEMM_BF16_TC = 2,     /* tcgen05 kind::f16 (bf16 in, fp32 accumulate): fast mode (declared, not built) */
    PPO_GEMM_F16X3_TC = 3     /* tcgen05 kind::f16 on power-of-two-scaled fp16 hi/lo pairs, error-compensated 3-pass
                                 split: fp32 parity at twice the tf32 tensor rate and half the operand bytes */
} ppo_gemm_mode;

const char* ppo_last_error(void);
const char* ppo_version(void);

/* ---- context ---------------------------------------------------------------------------- */
int ppo_ctx_create(int device, ppo_ctx** out);
int ppo_ctx_destroy(ppo_ctx* ctx);
int ppo_sync(ppo_ctx* ctx);
/* number of kernels this library launched on the ctx since creation (bench `gpu_launches`) */
int64_t ppo_ctx_launch_count(ppo_ctx* ctx);
/* cuda stream handle (cudaStream_t) the ctx enqueues on — for event timing by the caller */
void* ppo_ctx_stream(ppo_ctx* ctx);

/* ---- data-parallel communicator (new; no reference counterpart) -------------- */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the host (torch.distributed
 * broadcast / a file / MPI).  One process per GPU. */
int ppo_comm_unique_id(void* id128);
int ppo_comm_init(ppo_ctx* ctx, int nranks, int rank, const void* id128);
int ppo_comm_destroy(ppo_ctx* ctx);
/* sum-all-reduce `n` doubles in place (host values; used for global batch bookkeeping) */
int ppo_comm_allreduce_f64(ppo_ctx* ctx, double* host_inout, int n);

/* ---- rollout buffer: src/rollout_buffer.jl ---------------------------------------------- */
/* BufferRollouts() ctor, src/rollout_buffer.jl:9-22 — device-resident SoA with fixed shapes. */
int ppo_buffer_create(ppo_ctx* ctx, int64_t capacity, int nf, int nhe, int apa, ppo_buf** out);
int ppo_buffer_destroy(ppo_buf* buf);
/* update!(buffer, state, action_probability, action, reward, terminal), :24-38, batched over n
 * transitions.  feat[n][nhe][nf], mask[n][A], action 1-based in 1..A. */
int ppo_buffer_append(ppo_buf* buf, int64_t n, const float* feat, const float* mask,
                      const int64_t* action, const float* old_prob, const float* reward,
                      const uint8_t* terminal);
/* same, features given as Int64 (StateData.vertex_score is a Matrix{Int64},
 * test/quad_game_utilities.jl:50-56); converted to Float32 on the device. */
int ppo_buffer_append_i64(ppo_buf* buf, int64_t n, const int64_t* feat, const float* mask,
                          const int64_t* action, const float* old_prob, const float* reward,
                          const uint8_t* terminal);
/* same, features given as Int8 / Int16: exact for the small-integer scores a quad-game state holds (vertex score,
 * degree, 0 for missing: test/quad_game_utilities.jl:35-37, 50-56) at a quarter / half of the Float32 host->device
 * bytes; widened to Float32 on the device.  The caller narrows (and range-checks) its Matrix{Int64}. */
int ppo_buffer_append_i8(ppo_buf* buf, int64_t n, const int8_t* feat, const float* mask,
                         const int64_t* action, const float* old_prob, const float* reward,
                         const uint8_t* terminal);
int ppo_buffer_append_i16(ppo_buf* buf, int64_t n, const int16_t* feat, const float* mask,
                          const int64_t* action, const float* old_prob, const float* reward,
                          const uint8_t* terminal);
/* same, with the narrowest host representation of both big arrays: features of feat_elem_bytes per element (1 = Int8,
 * 2 = Int16, 4 = Float32, 8 = Int64) and the action mask as ONE BIT per action (1 = allowed = 0f0, 0 = masked = -Inf32),
 * bit (a + A*s) of the stream = bit ((a + A*s) & 63) of word (a + A*s) >> 6 -- the `chunks` of the Julia
 * BitMatrix `isfinite.(action_mask)` of size (A, n).  The masks of test/quad_game_utilities.jl:39-44 only ever hold
 * 0f0 and -Inf32, so nothing is lost; C3 moves 1.09 GB per million transitions instead of 1.36 GB (Int8 + Float32 mask)
 * or 4.58 GB (Float32 + Float32).  mask_bits holds ceil(n*A/64) words. */
int ppo_buffer_append_packed(ppo_buf* buf, int64_t n, const void* feat, int feat_elem_bytes, const uint64_t* mask_bits,
                             const int64_t* action, const float* old_prob, const float* reward, const uint8_t* terminal);
/* Base.length, :40-48 */
int64_t ppo_buffer_length(ppo_buf* buf);
int ppo_buffer_clear(ppo_buf* buf);
/* compute_state_value!(rollouts, discount), :55-64 -> compute_returns,
 * src/collect_rollouts.jl:26-42: rewards are overwritten IN PLACE by the returns.
 * discount_is_f32 = 0: Float64 discount => Float64 carry (every in-tree call); 1: Float32. */
int ppo_compute_returns(ppo_buf* buf, double discount, int discount_is_f32);
/* EXTENSION (default off; hook batch_advantage, src/ProximalPolicyOptimization.jl:29,
 * src/train.jl:105 has no in-tree implementation): advantage = (returns - mean)/(std + eps)
 * over the whole buffer, applied when minibatches are gathered.  enable = 0 restores identity. */
int ppo_normalize_advantage(ppo_buf* buf, int enable, double eps);
/* keep / bring back the raw rewards (device-side copy) so that compute_state_value! can be run
 * again on the same rollouts, e.g. with another discount (the reference overwrites them). */
int ppo_buffer_save_rewards(ppo_buf* buf);
int ppo_buffer_restore_rewards(ppo_buf* buf);
/* read back transitions [start, start+count) (0-based start); any output may be NULL.
 * rewards_or_returns is `rollouts.rewards` (returns after ppo_compute_returns). */
int ppo_buffer_read(ppo_buf* buf, int64_t start, int64_t count, float* feat, float* mask,
                    int64_t* action, float* old_prob, float* rewards_or_returns, uint8_t* terminal);
/* permute!(rollouts, idx), :81-88 (idx 1-based, length == length(buffer)). */
int ppo_buffer_permute(ppo_buf* buf, const int64_t* idx1, int64_t n);
/* shuffle!(rollouts), :90-93, with the documented counter-based device permutation. */
int ppo_buffer_shuffle(ppo_buf* buf, uint64_t seed);

/* ---- dataset / minibatch indexing: src/rollout_buffer.jl:95-147, src/train.jl:93-99 ------ */
/* file_indices = randperm(num_data), src/train.jl:93: supplied by the host (1-based) ... */
int ppo_permutation_set(ppo_buf* buf, const int64_t* perm1, int64_t n);
/* ... or generated on the device by the cycle-walking Feistel bijection (bit-identical to
 * oracle/ppo_oracle.py:feistel_permutation).  perm1_out may be NULL. */
int ppo_permutation_generate(ppo_buf* buf, uint64_t seed, int64_t* perm1_out);
/* dataset[file_indices[start+1 : start+count]] -> get_batch, src/rollout_buffer.jl:117-133:
 * gather on the device with the current permutation, then copy to the host. */
int ppo_gather(ppo_buf* buf, int64_t start, int64_t count, float* feat_out, float* mask_out,
               int64_t* action_out, float* prob_out, float* returns_out);
/* dataset[indices] for an arbitrary 1-based index vector (duplicates allowed). */
int ppo_gather_indices(ppo_buf* buf, const int64_t* idx1, int64_t count, float* feat_out,
                       float* mask_out, int64_t* action_out, float* prob_out, float* returns_out);
/* device-only gather of the minibatch into the policy-independent staging area (what the
 * training path runs); exposed for benchmarking K4 in isolation. */
int ppo_gather_device(ppo_buf* buf, int64_t start, int64_t count, int variant);
/* copy the minibatch produced by the last ppo_gather_device (first `count` rows) to the host. */
int ppo_batch_read(ppo_buf* buf, int64_t count, float* feat_out, float* mask_out, int64_t* action_out,
                   float* prob_out, float* returns_out);

/* ---- disk replay: src/dataset.jl, src/rollouts_to_disk.jl ------------------ */
/* DiskDataset(root_directory, trajectory_filename, states_dirname) (src/dataset.jl:1-20) followed by load_sample
 * (:31-52) for every row, bulk-loaded into the device buffer: parses <root>/<trajectory.csv> (either schema:
 * `...,returns` written by write_returns_to_disk, src/rollouts_to_disk.jl:106-132, or `...,rewards,terminal` written
 * by update!, :34-40) and every <root>/<states>/<sample_names[i]> BSON file (BSON.@save of a StateData or of a bare
 * array).  has_returns = 1: the value column already holds returns; 0: call ppo_compute_returns afterwards. */
int ppo_disk_dataset_load(ppo_buf* buf, const char* root_directory, const char* trajectory_filename,
                          const char* states_dirname, int n_threads, int64_t* n_loaded, int* has_returns);
/* describe the bits-type arrays inside one BSON state file (tests / debugging): eltypes[i][16], counts[i],
 * ndims[i], dims[i][4] for up to max_arrays arrays, in document order. */
int ppo_bson_state_arrays(const char* path, int max_arrays, char* eltypes, int64_t* counts, int* ndims,
                          int64_t* dims, int* n_arrays);

/* ---- policy: test/policy.jl:9-31 (Chain of Dense) ---------------------------------------- */
/* n_layers Dense layers; dims[0..n_layers] = in, h, ..., out; W[l] = bytes of Julia's
 * Dense.weight [dims[l+1], dims[l]]; hidden activation leakyrelu(slope), last layer linear. */
int ppo_policy_create(ppo_ctx* ctx, int n_layers, const int* dims, const float* const* W,
                      const float* const* b, float leaky_slope, ppo_policy** out);
int ppo_policy_destroy(ppo_policy* p);
int ppo_policy_read(ppo_policy* p, float* const* W, float* const* b);
int ppo_policy_write(ppo_policy* p, const float* const* W, const float* const* b);
/* a new policy runs on PPO_GEMM_AUTO; set_gemm_mode selects another engine (or fails loudly, PPO_ERR_INVALID, when the
 * policy's shapes are outside that engine's contract) */
int ppo_policy_set_gemm_mode(ppo_policy* p, int mode);
/* the engine in use (after PPO_GEMM_AUTO: the one that was picked) */
int ppo_policy_get_gemm_mode(ppo_policy* p);
/* diagnostic of the peer-memory exchange: total nanoseconds this rank's optimiser kernel spent waiting for its peers'
 * gradients (clock skew between the GPUs + the peers' publish latency) and the number of waits (minibatches);
 * reset != 0 clears the counters.  Zeros when the exchange is not connected.  Synchronises. */
int ppo_policy_p2p_wait(ppo_policy* p, int64_t* total_ns, int64_t* waits, int reset);
/* Token compaction (fp16-split engine; on by default).  A token all of whose actions carry a -Inf mask has probability
 * exactly 0 for each of them (softmax(logits .+ mask), test/quad_game_utilities.jl:73-79), so its logits never reach the
 * loss and its rows add exact zeros to every gradient.  With compaction the MLP runs on the remaining tokens only
 * (inactive quads of a padded mesh, test/quad_game_utilities.jl:39-44; states padded by pad_action_mask,
 * examples/triangle/distance_weighted/triangle_utilities.jl:41-55); losses, probabilities and gradients are those of
 * the dense evaluation up to the summation order of the weight gradient.  enable = 0 runs every token. */
int ppo_policy_set_token_compaction(ppo_policy* p, int enable);
/* tokens the last forward pass ran (-1: it ran every token of the minibatch); synchronises */
int ppo_policy_active_tokens(ppo_policy* p, int64_t* active_out);
#define PPO_GATE_SKIPPED 255
/* Parity instrumentation: the leakyrelu' branch (1 = pre-activation > 0, 0 = slope branch) that the backward pass of
 * the LAST minibatch applies to every element of hidden activation `layer` (1 .. n_layers-1, the output of Dense
 * `layer`), gates_out[rows][dims[layer]], rows = tokens of that minibatch.  leakyrelu' is discontinuous at 0
 * (ASSUMED NNlib.leakyrelu, test/policy.jl:11-15), so a gradient comparison against another evaluation is only
 * meaningful when both sides take the same branch for pre-activations within rounding of zero; the tests feed these
 * gates to the fp64 oracle and check separately that every disagreement sits at a ~0 pre-activation.  Tokens that a
 * compacted pass skipped read PPO_GATE_SKIPPED (their gate multiplies an exact zero). */
int ppo_policy_read_gates(ppo_policy* p, int layer, int64_t rows, uint8_t* gates_out);
/* Data parallelism over NVLink peer memory (one process per GPU, one node): instead of an NCCL all-reduce per minibatch the
   Adam kernel reads every rank's published gradient through CUDA-IPC mappings and adds them in rank order (weights stay
   bit-identical across ranks).  Every rank exports a 64-byte IPC handle, the handles are gathered by any transport
   (rank-major, nranks x 64 bytes) and passed to connect.  Requires ppo_comm_init (rank / nranks); the per-epoch loss
   history still goes through NCCL. */
int ppo_policy_p2p_export(ppo_policy* p, void* handle64);
int ppo_policy_p2p_connect(ppo_policy* p, int nranks, int rank, const void* handles);
int64_t ppo_policy_num_params(ppo_policy* p);
/* PPO.batch_action_probabilities(policy, state), test/quad_game_utilities.jl:73-79:
 * probs[nb][A] = softmax(reshape(policy(feat), :, nb) + mask). */
int ppo_batch_action_probabilities(ppo_policy* p, int64_t nb, int nhe, const float* feat,
                                   const float* mask, float* probs_out);
/* Batched rollout inference.  The reference samples one state at a time on the host
   (src/collect_rollouts.jl:1-15: action_probabilities -> rand(Categorical(ap))); this entry does nb states in one call:
   policy forward + masked softmax + one inverse-CDF categorical draw per state.  Draw i uses output i of the
   splitmix64 stream of `seed` (top 24 bits as a Float32 uniform), the cumulative sum is sequential Float32 as in
   Distributions.jl.  action1_out: 1-based Int64 [nb]; prob_out: probability of the drawn action [nb] (what update!
   stores as selected_action_probability); probs_out (optional, may be NULL): all probabilities [nb][A]. */
int ppo_sample_actions(ppo_policy* p, int64_t nb, int nhe, const float* feat, const float* mask, uint64_t seed,
                       int64_t* action1_out, float* prob_out, float* probs_out);


/* ---- optimiser: Flux.Optimise.Adam, call site src/train.jl:81 ----------------------------- */
int ppo_adam_create(ppo_policy* p, double eta, double beta1, double beta2, double eps, ppo_opt** out);
int ppo_adam_destroy(ppo_opt* o);
int ppo_adam_set_eta(ppo_opt* o, double eta);
double ppo_adam_get_eta(ppo_opt* o);
/* Flux.update!(optimizer, weights, grad) with a host gradient in Flux.params order
 * (W1, b1, W2, b2, ...). */
int ppo_adam_update(ppo_opt* o, const float* grad_flat);

/* ---- loss + update: src/train.jl --------------------------------------------------------- */
/* K6 in isolation: ppo_loss_with_entropy (:35-46) + its gradient w.r.t. the logits, on host
 * arrays.  logits/mask [nb][A]; action1 1-based within the column; ppoloss and entropyloss
 * (UNweighted, = -smoothed_entropy) as the reference returns them; dlogits_out (may be NULL) is
 * d(ppoloss + entropy_weight*entropyloss)/dlogits. */
int ppo_loss_from_logits(ppo_ctx* ctx, int64_t nb, int A, const float* logits, const float* mask,
                         const int64_t* action1, const float* old_prob, const float* advantage,
                         double epsilon, double entropy_weight, double* ppoloss, double* entropyloss,
                         float* dlogits_out);
/* step_batch!(policy, optimizer, state, linear_action_index, old_action_probabilities,
 * advantage, epsilon, entropy_weight), :54-84, on host arrays.  linear_action_index is the
 * 1-based column-major index into probs[A, nb] (get_linear_action_index, :48-52).
 * opt may be NULL (gradient only).  Returns (ppoloss, entropyloss*entropy_weight);
 * grads_out (may be NULL) receives the gradient in Flux.params order. */
int ppo_step_batch_host(ppo_policy* p, ppo_opt* opt, int64_t nb, int nhe, const float* feat,
                        const float* mask, const int64_t* linear_action_index,
                        const float* old_prob, const float* advantage, double epsilon,
                        double entropy_weight, double* ppoloss, double* entropyloss_weighted,
                        float* grads_out);
/* one minibatch of step_epoch!'s loop body (:96-124) on the device buffer: rows
 * perm[start .. start+count) (0-based start). */
int ppo_step_batch(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, int64_t start, int64_t count,
                   double epsilon, double entropy_weight, double* ppoloss,
                   double* entropyloss_weighted, float* grads_out);
/* step_epoch!(policy, optimizer, dataset, epsilon, batch_size, entropy_weight), :86-128, using
 * the buffer's current permutation (ppo_permutation_set / _generate).  Returns the unweighted
 * means of the per-minibatch losses (:127).  With a communicator, every rank passes its LOCAL
 * batch_size; gradients are summed over ranks and normalised by the global row count. */
int ppo_step_epoch(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, double epsilon, int64_t batch_size,
                   double entropy_weight, double* ppoloss_mean, double* entropyloss_mean);
/* ppo_train!(policy, optimizer, dataset, epsilon, batch_size, num_epochs, entropy_weight),
 * :130-153, drawing epoch e's permutation on the device from seed + e.  History arrays have
 * num_epochs entries. */
int ppo_train(ppo_policy* p, ppo_opt* opt, ppo_buf* buf, double epsilon, int64_t batch_size,
              int num_epochs, double entropy_weight, uint64_t seed, double* ppo_hist,
              double* entropy_hist, double* lr_hist);

/* ---- one Dense-layer operation on host arrays through GEMM engine `mode` (tests, debugging) ----
 * op 0 (Dense forward, test/policy.jl:11-15):   out[M,N] = act(X[M,K] W[K,N] + bias[N]); act = leakyrelu when slope >= 0
 * op 1 (pullback w.r.t. the input + activation): out[M,K] = (dY[M,N] W[K,N]^T) .* leakyrelu'(X[M,K])
 * op 2 (pullback w.r.t. weight and bias):        out[K,N] = X[M,K]^T dY[M,N];  out2[N] = colsum(dY) */
int ppo_dense_op(ppo_ctx* ctx, int mode, int op, int64_t M, int K, int N, const float* X, const float* W,
                 const float* bias, const float* dY, float slope, float* out, float* out2);

/* ---- per-kernel timing hooks for bench.py (device pointers stay inside the library) ------- */
/* run kernel `which` `iters` times on a synthetic device-resident problem and return the
 * average ms per launch measured with CUDA events on the ctx stream.  See bench.py. */
int ppo_bench_kernel(ppo_ctx* ctx, const char* which, int64_t n, int a, int b, int c, int iters,
                     int flush_l2, double* ms_out, double* bytes_or_flops_out);

#ifdef __cplusplus
}
#endif
#endif /* PPO_B200_H */
