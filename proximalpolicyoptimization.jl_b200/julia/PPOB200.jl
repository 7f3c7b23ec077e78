# PPOB200.jl — drop-in device containers for ProximalPolicyOptimization.jl over libppo_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (see DESIGN.md).
# This file is the reference-side binding a maintainer adds next to the package: a third container
# type (`DeviceRollouts`, `DeviceDataset`) beside `BufferRollouts` (src/rollout_buffer.jl:1-22) and
# `DiskRollouts` (src/rollouts_to_disk.jl:1-5), plus methods of the package's own generic functions
# specialised on it.  Every `ccall` below binds a symbol declared in include/ppo_b200.h; the Python
# layer in this directory binds exactly the same symbols and is what the tests execute.
module PPOB200

using ProximalPolicyOptimization
const PPO = ProximalPolicyOptimization
using Flux
using Printf

const lib = get(ENV, "PPO_B200_LIB", joinpath(@__DIR__, "..", "libppo_b200.so"))

struct PPOError <: Exception
    code::Cint
    msg::String
end
last_error() = unsafe_string(ccall((:ppo_last_error, lib), Cstring, ()))
check(status::Cint) = status == 0 ? nothing : throw(PPOError(status, last_error()))

# Destruction order: the library reference-counts the children of a context (buffers, policies, optimisers), so the
# finalizers below may run in any order (ppo_ctx_destroy on a context with live children only marks it; the last child's
# destroy call releases it).
# ---- context -------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ppo_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(c -> (ccall((:ppo_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), c)
    end
end
sync(c::Context) = check(ccall((:ppo_sync, lib), Cint, (Ptr{Cvoid},), c.h))

# ---- rollouts: src/rollout_buffer.jl ---------------------------------------------------------------
# StateData{vertex_score::Matrix [nf, nhe], action_mask::Vector{Float32} [A]} as in test/quad_game_utilities.jl:17-20
mutable struct DeviceRollouts
    ctx::Context
    h::Ptr{Cvoid}
    nf::Int; nhe::Int; apa::Int
    # host staging so that update! keeps its one-transition signature (flushed in chunks)
    feat::Vector{Float32}; mask::Vector{Float32}; prob::Vector{Float32}
    act::Vector{Int64}; rew::Vector{Float32}; term::Vector{UInt8}
end

# `capacity` is an initial reservation (the device arrays grow like the reference's push!-grown vectors)
function DeviceRollouts(ctx::Context, nf, nhe, apa, capacity = 4096)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ppo_buffer_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                ctx.h, capacity, nf, nhe, apa, r))
    b = DeviceRollouts(ctx, r[], nf, nhe, apa, Float32[], Float32[], Float32[], Int64[], Float32[], UInt8[])
    finalizer(b -> (ccall((:ppo_buffer_destroy, lib), Cint, (Ptr{Cvoid},), b.h); b.h = C_NULL), b)
end
# BufferRollouts() takes no shapes (src/rollout_buffer.jl:9-22): this form reads them off a sample state
DeviceRollouts(ctx::Context, state, apa::Integer; capacity = 4096) =
    DeviceRollouts(ctx, size(state.vertex_score, 1), size(state.vertex_score, 2), apa, capacity)

function flush!(b::DeviceRollouts)
    n = length(b.act)
    n == 0 && return
    GC.@preserve b begin
        if all(x -> x == round(x) && -128 <= x <= 127, b.feat)
            # vertex scores / degrees are small integers (test/quad_game_utilities.jl:50-56): a quarter of the bytes
            f8 = Int8.(b.feat)
            if all(x -> x == 0f0 || x == -Inf32, b.mask)
                # ... and the masks only hold 0f0 / -Inf32 (:39-44): one bit per action (BitVector chunks, 1 = allowed)
                bits = BitVector(isfinite.(b.mask))
                check(ccall((:ppo_buffer_append_packed, lib), Cint,
                            (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint, Ptr{UInt64}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}),
                            b.h, n, f8, 1, bits.chunks, b.act, b.prob, b.rew, b.term))
            else
                check(ccall((:ppo_buffer_append_i8, lib), Cint,
                            (Ptr{Cvoid}, Int64, Ptr{Int8}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}),
                            b.h, n, f8, b.mask, b.act, b.prob, b.rew, b.term))
            end
        else
            check(ccall((:ppo_buffer_append, lib), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}),
                        b.h, n, b.feat, b.mask, b.act, b.prob, b.rew, b.term))
        end
    end
    # (the staged transitions stay staged when the append throws)
    empty!(b.feat); empty!(b.mask); empty!(b.prob); empty!(b.act); empty!(b.rew); empty!(b.term)
end

# update!(episode::BufferRollouts, state, action_probability, action, reward, terminal) — :24-38
function PPO.update!(b::DeviceRollouts, state, action_probability, action, reward, terminal)
    append!(b.feat, Float32.(vec(state.vertex_score)))      # column-major [nf, nhe] == C [nhe][nf]
    append!(b.mask, state.action_mask)
    push!(b.prob, action_probability); push!(b.act, action); push!(b.rew, reward); push!(b.term, terminal)
    length(b.act) >= 4096 && flush!(b)
    return
end

# Base.length — :40-48
function Base.length(b::DeviceRollouts)
    Int(ccall((:ppo_buffer_length, lib), Int64, (Ptr{Cvoid},), b.h)) + length(b.act)
end

# compute_state_value!(rollouts, discount) — :55-64 -> compute_returns src/collect_rollouts.jl:26-42
function PPO.compute_state_value!(b::DeviceRollouts, discount)
    flush!(b)
    check(ccall((:ppo_compute_returns, lib), Cint, (Ptr{Cvoid}, Cdouble, Cint), b.h, Float64(discount),
                discount isa Float32 ? 1 : 0))
end

# collect_rollouts!(rollouts, env, policy, num_episodes, discount) — :66-79 (env stepping stays on the host)
function PPO.collect_rollouts!(b::DeviceRollouts, env, policy, num_episodes, discount)
    for _ in 1:num_episodes
        PPO.reset!(env)
        PPO.collect_episode_data!(b, env, policy)
    end
    PPO.compute_state_value!(b, discount)
end

# permute!(rollouts, idx) / shuffle!(rollouts) — :81-93
function PPO.permute!(b::DeviceRollouts, idx::AbstractVector{<:Integer})
    flush!(b); @assert length(idx) == length(b)
    p = Int64.(idx)
    check(ccall((:ppo_buffer_permute, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), b.h, p, length(p)))
end
PPO.shuffle!(b::DeviceRollouts; seed = rand(UInt64)) =
    (flush!(b); check(ccall((:ppo_buffer_shuffle, lib), Cint, (Ptr{Cvoid}, UInt64), b.h, seed)))

# ---- dataset: BufferDataset — :95-147 ---------------------------------------------------------------
struct DeviceDataset
    rollouts::DeviceRollouts
end
PPO.construct_dataset(b::DeviceRollouts) = (flush!(b); DeviceDataset(b))
Base.length(d::DeviceDataset) = length(d.rollouts)

function Base.getindex(d::DeviceDataset, idx)
    if idx isa Int
        @assert 1 <= idx <= length(d)
        batch = d[[idx]]
        s = batch["state"]
        return Dict("state" => typeof(s)(s.vertex_score[:, :, 1], s.action_mask[:, 1]),
                    "selected_action" => batch["selected_action"][1],
                    "selected_action_probability" => batch["selected_action_probability"][1],
                    "returns" => batch["returns"][1])
    elseif idx isa AbstractArray
        b = d.rollouts; nb = length(idx); A = b.nhe * b.apa
        feat = Array{Float32}(undef, b.nf, b.nhe, nb); mask = Array{Float32}(undef, A, nb)
        act = Vector{Int64}(undef, nb); prob = Vector{Float32}(undef, nb); ret = Vector{Float32}(undef, nb)
        check(ccall((:ppo_gather_indices, lib), Cint,
                    (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}),
                    b.h, Int64.(idx), nb, feat, mask, act, prob, ret))
        return Dict("state" => (vertex_score = feat, action_mask = mask), "selected_action" => act,
                    "selected_action_probability" => prob, "returns" => ret)
    else
        error("Dataset index should be Int or Array, got ", typeof(idx))
    end
end

# ---- policy / optimiser handles ----------------------------------------------------------------------
# A Flux Chain of Dense layers (test/policy.jl:9-21) is handed over as weight arrays; Dense.weight [out, in]
# column-major is passed as is (== float W[in][out]).
mutable struct DevicePolicy
    ctx::Context
    h::Ptr{Cvoid}
    chain          # the user's Flux model; refreshed by pull_weights!
end

dense_layers(model) = [l for l in model.layers if l isa Dense]

# gemm_mode: -1 = auto (fastest fp32-parity engine whose shape contract the model meets), 0 = fp32 FFMA, 1 = 3xTF32
# tcgen05, 3 = scaled fp16 hi/lo pairs on tcgen05 (needs in % 8 == 0, hidden widths % 32 == 0, <= 4 actions per token)
gemm_mode(p) = ccall((:ppo_policy_get_gemm_mode, lib), Cint, (Ptr{Cvoid},), p.h)
# token compaction (default on): the MLP skips tokens all of whose actions are masked (-Inf32); their probabilities
# and gradients are exactly 0 either way
token_compaction!(p, enable::Bool) = check(ccall((:ppo_policy_set_token_compaction, lib), Cint, (Ptr{Cvoid}, Cint), p.h, enable ? 1 : 0))
function active_tokens(p)
    r = Ref{Int64}(-1)
    check(ccall((:ppo_policy_active_tokens, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), p.h, r))
    r[]
end
function DevicePolicy(ctx::Context, model; leaky_slope = 0.01f0, gemm_mode = -1)
    ls = dense_layers(model)
    dims = Cint[size(ls[1].weight, 2); [size(l.weight, 1) for l in ls]]
    Ws = [Float32.(l.weight) for l in ls]; bs = [Float32.(l.bias) for l in ls]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Ws bs begin
        check(ccall((:ppo_policy_create, lib), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}, Cfloat, Ref{Ptr{Cvoid}}),
                    ctx.h, length(ls), dims, pointer.(Ws), pointer.(bs), leaky_slope, r))
    end
    p = DevicePolicy(ctx, r[], model)
    check(ccall((:ppo_policy_set_gemm_mode, lib), Cint, (Ptr{Cvoid}, Cint), p.h, gemm_mode))
    finalizer(p -> (ccall((:ppo_policy_destroy, lib), Cint, (Ptr{Cvoid},), p.h); p.h = C_NULL), p)
end

# copy the trained weights back into the user's Flux model so collect_rollouts!/average_returns/BSON.@save see them
function pull_weights!(p::DevicePolicy)
    ls = dense_layers(p.chain)
    Ws = [Array{Float32}(undef, size(l.weight)) for l in ls]; bs = [Array{Float32}(undef, size(l.bias)) for l in ls]
    GC.@preserve Ws bs check(ccall((:ppo_policy_read, lib), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}),
                                   p.h, pointer.(Ws), pointer.(bs)))
    for (l, W, b) in zip(ls, Ws, bs)
        l.weight .= W; l.bias .= b
    end
end

# PPO.batch_action_probabilities(policy, state) — hook src/ProximalPolicyOptimization.jl:24, reference implementation
# test/quad_game_utilities.jl:73-79: probs[A, nb] = softmax(reshape(policy(vertex_score), :, nb) + action_mask, dims = 1)
function PPO.batch_action_probabilities(p::DevicePolicy, state)
    vs = Float32.(state.vertex_score); am = Float32.(state.action_mask)
    nhe, nb = size(vs, 2), size(vs, 3)
    probs = Matrix{Float32}(undef, size(am, 1), nb)
    check(ccall((:ppo_batch_action_probabilities, lib), Cint,
                (Ptr{Cvoid}, Int64, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), p.h, nb, nhe, vs, am, probs))
    return probs
end
# PPO.action_probabilities(policy, state) — hook :23, reference implementation test/quad_game_utilities.jl:65-71 (one state);
# this is what collect_step_data! (src/collect_rollouts.jl:1-15) and single_trajectory_return (src/evaluate.jl:1-16) call
function PPO.action_probabilities(p::DevicePolicy, state)
    vs = reshape(Float32.(state.vertex_score), size(state.vertex_score, 1), size(state.vertex_score, 2), 1)
    am = reshape(Float32.(state.action_mask), :, 1)
    return vec(PPO.batch_action_probabilities(p, (vertex_score = vs, action_mask = am)))
end
PPO.number_of_actions_per_state(state::NamedTuple{(:vertex_score, :action_mask)}) = size(state.action_mask, 1)

# Batched rollout inference (extension): collect_step_data!'s `action_probabilities` + `rand(Categorical(ap))`
# (src/collect_rollouts.jl:1-15) for nb states at once.  vertex_score [nf, nhe, nb], action_mask [A, nb].
function batch_sample_actions(p::DevicePolicy, vertex_score::Array{Float32,3}, action_mask::Matrix{Float32};
                              seed::UInt64 = rand(UInt64))
    nhe, nb = size(vertex_score, 2), size(vertex_score, 3)
    actions = Vector{Int64}(undef, nb); probs = Vector{Float32}(undef, nb)
    check(ccall((:ppo_sample_actions, lib), Cint,
                (Ptr{Cvoid}, Int64, Cint, Ptr{Float32}, Ptr{Float32}, UInt64, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}),
                p.h, nb, nhe, vertex_score, action_mask, seed, actions, probs, C_NULL))
    return actions, probs
end

# Data parallelism over NVLink peer memory (one Julia process per GPU): `handles = allgather(p2p_export(p))` with any
# transport (MPI.jl, Distributed, files), then `p2p_connect!(p, nranks, rank, handles)` on every rank.
function p2p_export(p::DevicePolicy)
    h = Vector{UInt8}(undef, 64)
    check(ccall((:ppo_policy_p2p_export, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), p.h, h))
    return h
end
p2p_connect!(p::DevicePolicy, nranks::Integer, rank::Integer, handles::Vector{Vector{UInt8}}) =
    check(ccall((:ppo_policy_p2p_connect, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), p.h, nranks, rank, reduce(vcat, handles)))

mutable struct DeviceAdam
    h::Ptr{Cvoid}
    eta::Float64
end
function DeviceAdam(p::DevicePolicy, o::Flux.Optimise.Adam)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ppo_adam_create, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Ref{Ptr{Cvoid}}),
                p.h, o.eta, o.beta[1], o.beta[2], o.epsilon, r))
    a = DeviceAdam(r[], o.eta)
    finalizer(a -> (ccall((:ppo_adam_destroy, lib), Cint, (Ptr{Cvoid},), a.h); a.h = C_NULL), a)
end

# ---- the hot loop: src/train.jl ----------------------------------------------------------------------
# step_batch!(policy, optimizer, state, linear_action_index, old_action_probabilities, advantage, epsilon,
#             entropy_weight) — :54-84, on host arrays (state.vertex_score [nf, nhe, nb], state.action_mask [A, nb])
function PPO.step_batch!(p::DevicePolicy, o::DeviceAdam, state, linear_action_index, old_action_probabilities, advantage,
                         epsilon, entropy_weight)
    vs = Float32.(state.vertex_score); am = Float32.(state.action_mask)
    nhe, nb = size(vs, 2), size(vs, 3)
    lin = Int64.(linear_action_index); old = Float32.(old_action_probabilities); adv = Float32.(advantage)
    @assert length(lin) == nb && length(old) == nb && length(adv) == nb
    pl = Ref{Cdouble}(0); el = Ref{Cdouble}(0)
    check(ccall((:ppo_step_batch_host, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32},
                 Cdouble, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ptr{Float32}),
                p.h, o.h, nb, nhe, vs, am, lin, old, adv, epsilon, entropy_weight, pl, el, C_NULL))
    return pl[], el[]
end

# step_epoch!(policy, optimizer, dataset, epsilon, batch_size, entropy_weight) — :86-128
function PPO.step_epoch!(p::DevicePolicy, o::DeviceAdam, d::DeviceDataset, epsilon, batch_size, entropy_weight;
                         perm::Union{Nothing,Vector{Int64}} = nothing, seed::UInt64 = rand(UInt64))
    num_data = length(d)
    @assert 1 <= batch_size <= num_data
    b = d.rollouts
    if perm === nothing
        check(ccall((:ppo_permutation_generate, lib), Cint, (Ptr{Cvoid}, UInt64, Ptr{Int64}), b.h, seed, C_NULL))
    else       # e.g. perm = randperm(num_data) to keep Julia's RNG stream (src/train.jl:93)
        check(ccall((:ppo_permutation_set, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), b.h, perm, length(perm)))
    end
    pl = Ref{Cdouble}(0); el = Ref{Cdouble}(0)
    check(ccall((:ppo_step_epoch, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Cdouble, Ref{Cdouble}, Ref{Cdouble}),
                p.h, o.h, b.h, epsilon, batch_size, entropy_weight, pl, el))
    return pl[], el[]
end

# ppo_train!(policy, optimizer, dataset, epsilon, batch_size, num_epochs, entropy_weight) — :130-153
function PPO.ppo_train!(p::DevicePolicy, o::DeviceAdam, d::DeviceDataset, epsilon, batch_size, num_epochs, entropy_weight)
    ppo_loss_history, entropy_loss_history, lr_history = [], [], []
    for epoch = 1:num_epochs
        ppoloss, entropyloss = PPO.step_epoch!(p, o, d, epsilon, batch_size, entropy_weight)
        lr = o.eta
        @printf "EPOCH : %d \t PPO LOSS : %1.4f\t ENTROPY LOSS : %1.4f \t LR : %1.1e\n" epoch ppoloss entropyloss lr
        push!(ppo_loss_history, ppoloss); push!(entropy_loss_history, entropyloss); push!(lr_history, lr)
    end
    pull_weights!(p)
    return ppo_loss_history, entropy_loss_history, lr_history
end

# ---- disk replay: src/dataset.jl, src/rollouts_to_disk.jl ----------------------------------------------
# bulk-load a DiskDataset (trajectory.csv + states/sample_i.bson, src/dataset.jl:1-52) into a device container
function to_device(ds::PPO.DiskDataset, ctx::Context, nf, nhe, apa; n_threads = Threads.nthreads())
    b = DeviceRollouts(ctx, nf, nhe, apa, max(length(ds), 1))
    n = Ref{Int64}(0); has_returns = Ref{Cint}(0)
    check(ccall((:ppo_disk_dataset_load, lib), Cint,
                (Ptr{Cvoid}, Cstring, Cstring, Cstring, Cint, Ref{Int64}, Ref{Cint}),
                b.h, ds.root_directory, ds.trajectory_filename, ds.states_dirname, n_threads, n, has_returns))
    @assert n[] == length(ds)
    return DeviceDataset(b), has_returns[] != 0
end

# ---- the outer loop: src/train.jl:164-249 ------------------------------------------------------------
# ppo_iterate! (buffer variant, :210-249).  The rollouts are collected by the reference's own host loop
# (collect_rollouts!, src/rollout_buffer.jl:66-79 -> collect_step_data!, which calls PPO.action_probabilities(policy,
# state) above), stored on the device, and trained on the device; the weights are pulled back into the Flux model after
# every ppo_train! so that evaluator(policy, env, optimizer) / BSON.@save see the update.
function PPO.ppo_iterate!(policy::DevicePolicy, env, optimizer::DeviceAdam, episodes_per_iteration, minibatch_size,
                          num_ppo_iterations, evaluator, epochs_per_iteration, discount, epsilon, entropy_weight)
    loss = Dict("ppo" => [], "entropy" => [], "lr" => [])
    apa = size(dense_layers(policy.chain)[end].weight, 1)
    for iter in 1:num_ppo_iterations
        evaluator(policy, env, optimizer)
        println("\nPPO ITERATION : $iter")
        PPO.reset!(env)
        rollouts = DeviceRollouts(policy.ctx, PPO.state(env), apa)
        PPO.collect_rollouts!(rollouts, env, policy, episodes_per_iteration, discount)
        dataset = PPO.construct_dataset(rollouts)
        ppoloss, entropyloss, lr_history = PPO.ppo_train!(policy, optimizer, dataset, epsilon, minibatch_size,
                                                          epochs_per_iteration, entropy_weight)
        append!(loss["ppo"], ppoloss); append!(loss["entropy"], entropyloss); append!(loss["lr"], lr_history)
        PPO.save_loss(evaluator, loss)
    end
end

# ppo_iterate! (disk variant, :164-202): the reference's DiskRollouts collect to state_data_path (host file I/O in its
# own format); the finished directory is bulk-loaded into the device buffer and trained there.
function PPO.ppo_iterate!(policy::DevicePolicy, env, optimizer::DeviceAdam, episodes_per_iteration, minibatch_size,
                          num_ppo_iterations, evaluator, epochs_per_iteration, discount, epsilon, entropy_weight,
                          state_data_path)
    loss = Dict("ppo" => [], "entropy" => [], "lr" => [])
    apa = size(dense_layers(policy.chain)[end].weight, 1)
    for iter in 1:num_ppo_iterations
        evaluator(policy, env, optimizer)
        println("\nPPO ITERATION : $iter")
        rollouts = PPO.DiskRollouts(state_data_path)
        PPO.collect_rollouts!(rollouts, env, policy, episodes_per_iteration, discount)
        PPO.reset!(env)
        s = PPO.state(env)
        dataset, has_returns = to_device(PPO.construct_dataset(rollouts), policy.ctx, size(s.vertex_score, 1),
                                         size(s.vertex_score, 2), apa)
        @assert has_returns
        ppoloss, entropyloss, lr_history = PPO.ppo_train!(policy, optimizer, dataset, epsilon, minibatch_size,
                                                          epochs_per_iteration, entropy_weight)
        append!(loss["ppo"], ppoloss); append!(loss["entropy"], entropyloss); append!(loss["lr"], lr_history)
        PPO.save_loss(evaluator, loss)
    end
    if isdir(state_data_path)
        println("\n\nCLEARING DATA IN ROLLOUTS FOLDER :")
        rm(state_data_path, recursive = true)
    end
end

end # module
