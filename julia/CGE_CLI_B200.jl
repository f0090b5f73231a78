# The reference's example/CGE_CLI.jl with one line changed: `using CGEB200` instead of
# `using CGE`.  Every flag of the reference CLI keeps working (parseargs is CGE.jl's own).
push!(LOAD_PATH, @__DIR__)
include(joinpath(@__DIR__, "CGEB200.jl"))
using .CGEB200

edges, weights, vweights, comm, clusters, embed, verbose, land, forced, method, directed, split, seed, samples = parseargs()
distances = zeros(length(vweights))
init_edges = Array{Int,2}(undef, 0, 0)
init_vweights = Vector{Float64}()
init_eweights = Vector{Float64}()
init_embed = Array{Float64,2}(undef, 0, 0)
v_to_l = Int[]
if land != -1
    init_edges, init_vweights, init_eweights, init_embed = copy(edges), copy(vweights), copy(weights), copy(embed)
    distances, embed, comm, edges, weights, vweights, v_to_l = landmarks(edges, weights, vweights,
        clusters, comm, embed, verbose, land, forced, method, directed)
end
score = directed ? wGCL_directed : wGCL
println(score(edges, weights, comm, embed, distances, vweights, init_vweights, v_to_l, init_edges,
              init_eweights, init_embed, split, seed, samples, verbose))
