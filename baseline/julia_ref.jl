# baseline/julia_ref.jl — pin this repository's oracle and CUDA path to the REAL reference.
#
# Nothing in this repository can run Julia (the build image has none, see DESIGN.md §2), so everything below
# src/train.jl — Flux.params / Zygote gradient / Flux.update!(Adam) / NNlib.softmax / leakyrelu — is checked against
# a restatement (oracle/) whose third-party formulas are ASSUMPTIONS.  A maintainer with Julia closes that gap:
#
#     python baseline/make_julia_inputs.py                       # seeded inputs -> baseline/julia_inputs/*.bson
#     julia --project=<ProximalPolicyOptimization.jl checkout> baseline/julia_ref.jl
#     python -m pytest tests/test_julia_goldens.py                # oracle vs Julia (CPU);  -m gpu: CUDA path vs Julia
#
# For every input case this script runs the reference's OWN functions — compute_returns (src/collect_rollouts.jl:26-42),
# BufferRollouts / update! / construct_dataset / getindex (src/rollout_buffer.jl), get_linear_action_index,
# ppo_loss_with_entropy, Flux.gradient over Flux.params exactly as step_batch! does (src/train.jl:35-84), and one full
# step_epoch! (src/train.jl:86-128) with Flux.Optimise.Adam — and writes tests/golden/julia_<case>.bson:
#     returns, perm (the randperm the epoch drew, 1-based), ppoloss, entw, grads (Flux.params order), mean_ppo, mean_ent,
#     flat_after (parameters after the epoch), flux_version, julia_version.
# step_epoch! draws `randperm(num_data)` itself (src/train.jl:93); the script seeds the global RNG, records the
# permutation that seed yields, re-seeds and calls step_epoch!, so the golden holds the very permutation the epoch used
# and the device path replays it (it honours host-supplied permutations bit for bit).
#
# The hooks are the ones the reference's own test utilities define (test/quad_game_utilities.jl:17-33, 65-79), restated
# here because that file pulls in un-vendored mesh packages; batch_advantage has no in-tree implementation and is the
# identity, as in the older API (examples/triangle/distance_weighted/profile.jl:64).
using ProximalPolicyOptimization
const PPO = ProximalPolicyOptimization
using Flux, BSON, Random, Pkg

include(joinpath(pkgdir(PPO), "test", "policy.jl"))          # SimplePolicy.Policy, test/policy.jl:9-31

struct StateData                                               # test/quad_game_utilities.jl:17-20
    vertex_score
    action_mask
end
Flux.@functor StateData

function PPO.batch_state(state_data_vector)                    # test/quad_game_utilities.jl:26-33
    vs = [s.vertex_score for s in state_data_vector]
    am = [s.action_mask for s in state_data_vector]
    return StateData(cat(vs..., dims = 3), cat(am..., dims = 2))
end

function PPO.batch_action_probabilities(policy, state)        # test/quad_game_utilities.jl:73-79
    vertex_score, action_mask = state.vertex_score, state.action_mask
    nf, nq, nb = size(vertex_score)
    logits = reshape(policy(vertex_score), :, nb) + action_mask
    return softmax(logits, dims = 1)
end

PPO.number_of_actions_per_state(state) = size(state.action_mask, 1)
PPO.batch_advantage(state, returns) = returns

flux_version() = try string(Pkg.dependencies()[Base.UUID("587475ba-b771-5e3f-ad9e-33799f191a9c")].version) catch; "unknown" end
make_adam(eta) = isdefined(Flux.Optimise, :Adam) ? Flux.Optimise.Adam(eta) : Flux.Optimise.ADAM(eta)

function load_policy(d)
    policy = SimplePolicy.Policy(d[:nf], d[:H], d[:L], d[:apa])
    dense = [l for l in policy.model.layers if l isa Dense]
    for (l, layer) in enumerate(dense)
        layer.weight .= d[Symbol("W$l")]                       # [out, in]
        layer.bias .= d[Symbol("b$l")]
    end
    return policy
end

flat(policy) = vcat([vec(copy(p)) for p in Flux.params(policy)]...)

function run_case(inpath, outpath)
    d = BSON.load(inpath)
    N, A = d[:N], d[:nhe] * d[:apa]
    discount, epsilon, w_ent, eta = d[:gamma], d[:eps], d[:w_ent], d[:eta]

    returns = PPO.compute_returns(d[:rewards], d[:terminal], discount)

    rollouts = PPO.BufferRollouts()
    for i in 1:N
        s = StateData(d[:vertex_score][:, :, i], d[:action_mask][:, i])
        PPO.update!(rollouts, s, d[:selected_action_probabilities][i], d[:selected_actions][i], d[:rewards][i], d[:terminal][i])
    end
    PPO.compute_state_value!(rollouts, discount)
    @assert rollouts.rewards == returns
    dataset = PPO.construct_dataset(rollouts)

    Random.seed!(d[:seed])
    perm = randperm(N)

    # ---- first minibatch at the initial weights: loss and gradient exactly as step_batch! forms them (src/train.jl:65-79)
    policy = load_policy(d)
    batch = dataset[perm[1:d[:B]]]
    state = batch["state"]
    lin = PPO.get_linear_action_index(batch["selected_action"], PPO.number_of_actions_per_state(state))
    adv = PPO.batch_advantage(state, batch["returns"])
    weights = Flux.params(policy)
    local ppoloss, entropyloss
    grad = Flux.gradient(weights) do
        ppoloss, entropyloss = PPO.ppo_loss_with_entropy(policy, state, lin, batch["selected_action_probability"], adv, epsilon)
        entropyloss = entropyloss * w_ent
        return ppoloss + entropyloss
    end
    grads = vcat([vec(Float32.(grad[p])) for p in weights]...)

    # ---- one full epoch of the reference's own step_epoch! (randperm drawn inside from the same seed)
    policy2 = load_policy(d)
    optimizer = make_adam(eta)
    Random.seed!(d[:seed])
    mean_ppo, mean_ent = PPO.step_epoch!(policy2, optimizer, dataset, epsilon, d[:B], w_ent)

    BSON.bson(outpath, Dict(:returns => Float32.(returns), :perm => Int64.(perm), :ppoloss => Float64(ppoloss),
                            :entw => Float64(entropyloss), :grads => grads, :mean_ppo => Float64(mean_ppo),
                            :mean_ent => Float64(mean_ent), :flat_after => Float32.(flat(policy2)),
                            :flux_version => flux_version(), :julia_version => string(VERSION)))
    println("wrote ", outpath, "  ppoloss = ", ppoloss, "  epoch = ", (mean_ppo, mean_ent))
end

function main()
    here = @__DIR__
    indir = joinpath(here, "julia_inputs")
    outdir = joinpath(here, "..", "tests", "golden")
    cases = isempty(ARGS) ? [splitext(f)[1] for f in readdir(indir) if endswith(f, ".bson")] : ARGS
    for c in cases
        run_case(joinpath(indir, c * ".bson"), joinpath(outdir, "julia_" * c * ".bson"))
    end
end

main()
