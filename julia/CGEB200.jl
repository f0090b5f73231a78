# CGEB200.jl -- drop-in Julia wrapper: CGE.jl's public scoring entry points on libcge_b200.so.
#
# `using CGEB200` instead of `using CGE` in example/CGE_CLI.jl is the whole integration:
# parseargs, landmarks and louvain_clust are re-exported from CGE unchanged (they stay in Julia),
# wGCL / wGCL_directed keep the reference signatures (src/divergence.jl:27-31, 282-286) and
# return the same Vector{Float64}; only the scoring call crosses the C ABI of
# include/cge_b200.h.  The wrapper builds E / NE and draws the sampled pairs exactly where the
# reference does (divergence.jl:121-137, 184-210, 484-513), so with the same seed the GPU sees the
# very sample sets the Julia implementation would have used.
#
# Multi-GPU: set ENV["CGE_B200_GPUS"] = "8" before calling wGCL in exact mode; the library then
# shards the pair matrix over the GPUs of the box from this single process (cge_b200_score_multi).
#
# NOTE: Julia is not installed in the image this repository is built and tested in; this file is
# the binding a maintainer adds on the reference side (INTEGRATION.md) and mirrors, line for line,
# cge_jl_b200/divergence.py, which IS exercised by the test-suite through the same C ABI.
module CGEB200

using CGE
using CGE: parseargs, louvain_clust   # re-exported unchanged; wGCL*, landmarks are defined here
using StatsBase
using Random
using LinearAlgebra: eigvecs

export parseargs, landmarks, louvain_clust, wGCL, wGCL_directed, read_table, landmarks_b200, runsplit_b200,
       unique_rows_b200

const LIB = get(ENV, "CGE_B200_LIB",
                normpath(joinpath(@__DIR__, "..", "cge_jl_b200", "libcge_b200.so")))
const N_ALPHA = 40

# One-shot hosts (one scoring call per process, like CGE_CLI.jl): the library holds one kernel instantiation per
# alpha and CUDA's default lazy loading charges ~10 ms for each on first use (0.40 s for the first call on the 10k
# example, 0.08 s with eager loading).  Has to be set before the first CUDA call of the process; a caller's own
# setting wins.
function __init__()
    haskey(ENV, "CUDA_MODULE_LOADING") || (ENV["CUDA_MODULE_LOADING"] = "EAGER")
end
# above this many vertices NE (n^2/2 tuples + two Sets, divergence.jl:121-137) no longer fits in
# host memory; non-edges are then drawn on the device (same distribution, different RNG stream)
const NE_MATERIALIZE_LIMIT = 30_000

# mirrors `cge_b200_problem` (include/cge_b200.h)
struct Problem
    struct_size::Int32; index_base::Int32; directed::Int32; split::Int32
    m::Int64; edge_src::Ptr{Int64}; edge_dst::Ptr{Int64}; eweights::Ptr{Float64}
    n_comm::Int64; comm::Ptr{Int64}
    embed::Ptr{Float64}; embed_rows::Int64; d::Int64; embed_row_stride::Int64; embed_col_stride::Int64
    n_distances::Int64; distances::Ptr{Float64}; vweights::Ptr{Float64}
    n_full::Int64; init_vweights::Ptr{Float64}; v_to_l::Ptr{Int64}; init_embed::Ptr{Float64}
    init_row_stride::Int64; init_col_stride::Int64
    n_samples::Int64; n_sets::Int64
    pos_i::Ptr{Int64}; pos_j::Ptr{Int64}; pos_w::Ptr{Float64}; neg_i::Ptr{Int64}; neg_j::Ptr{Int64}
    max_alphas::Int32; driver::Int32; regime::Int32; reserved::Int32
end

# mirrors `cge_b200_stats`
mutable struct Stats
    struct_size::Int32; n_alpha_run::Int32
    iters::NTuple{N_ALPHA,Int32}; div::NTuple{N_ALPHA,Float64}; auc::NTuple{N_ALPHA,Float64}
    lo::Float64; hi::Float64; hi_full::Float64
    n::Int64; n_pairs::Int64; fp_sweeps::Int64; b_sweeps::Int64; matrix_bytes::Int64; launches::Int64
    n_tiles::Int32; grid::Int32; driver::Int32; n_ranks::Int32; regime::Int32; diam_candidate_tiles::Int32
    ms_upload::Float32; ms_build::Float32; ms_solve::Float32; ms_total::Float32
    ms_sweeps::Float32; ms_bsweeps::Float32
    b_fused::Int32; ms_fused::Float32
    Stats() = new(0, 0, ntuple(_ -> Int32(0), N_ALPHA), ntuple(_ -> NaN, N_ALPHA),
                  ntuple(_ -> NaN, N_ALPHA), 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                  0f0, 0f0, 0f0, 0f0, 0f0, 0f0, 0, 0f0)
end

last_error() = unsafe_string(ccall((:cge_b200_last_error, LIB), Cstring, ()))

function check(rc::Integer)
    rc == 0 && return
    rc == -2 && throw(AssertionError("No. communities not matching no. vertices"))
    rc == -3 && throw(AssertionError("Distances vector length is not equal to no. vertices"))
    rc == -4 && throw(OutOfMemoryError())
    throw(ErrorException("libcge_b200 error $rc: $(last_error())"))
end

# uniform draws from NE without materialising it (used only above NE_MATERIALIZE_LIMIT): the
# device sampler of SURVEY.md 8(f) F1 -- edge hash set in HBM + independent rejection draws,
# cge_b200_sample_non_edges.  Returns K x n_sets matrices of 1-based ids.
function sample_non_edges_device(adj_edges::Array{Int,2}, n::Int, K::Int, n_sets::Int, seed::Int,
                                 directed::Bool)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cge_b200_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), 0, h))
    neg_i = Matrix{Int64}(undef, K, n_sets); neg_j = similar(neg_i)
    m = size(adj_edges, 1)
    try
        GC.@preserve adj_edges neg_i neg_j begin
            src = pointer(adj_edges)                 # column 1 of the column-major m x 2 matrix
            dst = src + m * sizeof(Int64)            # column 2
            check(ccall((:cge_b200_sample_non_edges, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Int32, Int64, Int64,
                         UInt64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                        h[], n, m, src, dst, 1, directed ? 1 : 0, K, n_sets,
                        seed == -1 ? rand(UInt64) : reinterpret(UInt64, Int64(seed)), neg_i, neg_j, C_NULL))
        end
    finally
        ccall((:cge_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return neg_i, neg_j
end

"""
    read_table(fn; T=Float64, skipstart=0)

Drop-in for the typed `readdlm(fn, T; skipstart)` calls of `parseargs` (auxilary.jl:86, 123,
150-155) on large inputs (SURVEY.md 8(f) F3): the file is parsed on all host cores by
`cge_b200_read_table` straight into the column-major matrix.  Throws where `readdlm` does (ragged
rows, cells that are not numbers), so the node2vec `try ... catch ... skipstart=1` keeps working.
"""
function read_table(fn::AbstractString; T::Type=Float64, skipstart::Int=0)
    rows = Ref{Int64}(0); cols = Ref{Int64}(0)
    check(ccall((:cge_b200_table_dims, LIB), Cint, (Cstring, Int64, Int32, Ptr{Int64}, Ptr{Int64}),
                fn, skipstart, 0, rows, cols))
    out = Matrix{Float64}(undef, rows[], cols[])
    check(ccall((:cge_b200_read_table, LIB), Cint,
                (Cstring, Int64, Int32, Int64, Int64, Int64, Int64, Ptr{Float64}),
                fn, skipstart, 0, rows[], cols[], 1, rows[], out))        # column-major strides
    return T === Float64 ? out : convert.(T, out)                           # InexactError like readdlm(fn, Int)
end

"""
Draws the positive / negative pairs of the local score with the reference's own calls
(`Random.seed!` + `StatsBase.sample`), in the reference's order.
Returns (pos_i, pos_j, pos_w, neg_i, neg_j) as K x n_sets matrices (column = one set).
"""
function draw_samples(adj_edges::Array{Int,2}, adj_eweights::Vector{Float64}, adj_n::Int,
                      K::Int, seed::Int, directed::Bool, exact::Bool)
    E = Tuple{Int64,Int64,Float64}[]
    for (i, e) in enumerate(eachrow(adj_edges))
        push!(E, directed ? (e[1], e[2], adj_eweights[i]) :
                            (minimum(e), maximum(e), adj_eweights[i]))      # divergence.jl:131-134 / 415-418
    end
    NE = nothing
    if adj_n <= NE_MATERIALIZE_LIMIT                                        # divergence.jl:121-137 / 405-421
        edgeset = Set([e[1:2] for e in E])
        NE = Tuple{Int64,Int64}[]
        for i in 1:adj_n, j in (directed ? 1 : i):adj_n
            i != j && push!(NE, (i, j))
        end
        NE = collect(setdiff(Set(NE), edgeset))
    end
    n_sets = seed != -1 ? 1 : N_ALPHA
    pos_i = Matrix{Int64}(undef, K, n_sets); pos_j = similar(pos_i)
    neg_i = similar(pos_i); neg_j = similar(pos_i)
    pos_w = Matrix{Float64}(undef, K, n_sets)
    if NE === nothing
        neg_i, neg_j = sample_non_edges_device(adj_edges, adj_n, K, n_sets, seed, directed)
    end
    for s in 1:n_sets
        seed != -1 && Random.seed!(seed)                                    # :184 / :202 / :484 / :504
        first = sample(E, K, replace=true)
        pos_w[:, s] = [e[3] for e in first]
        pairs = (directed && exact) ? sample(E, K, replace=true) : first   # overwrite at :510
        pos_i[:, s] = [e[1] for e in pairs]; pos_j[:, s] = [e[2] for e in pairs]
        NE === nothing && continue
        seed != -1 && Random.seed!(seed)                                    # :193 / :209 / :494 / :512
        neg = sample(NE, K, replace=true)
        neg_i[:, s] = [e[1] for e in neg]; neg_j[:, s] = [e[2] for e in neg]
    end
    return pos_i, pos_j, pos_w, neg_i, neg_j
end

# the one step of a cut the reference gives to LAPACK (landmarks.jl:160-162: eigvecs(yᵀwy)[:, end]), as the
# callback of cge_b200_landmarks_select: the matrix arrives row-major = column-major (it is symmetric)
function _principal_axis(c::Ptr{Float64}, d::Int64, v::Ptr{Float64}, ::Ptr{Cvoid})::Cint
    try
        A = unsafe_wrap(Array, c, (Int(d), Int(d)))
        axis = eigvecs(Matrix(A))[:, end]                                 # same call, same LAPACK, same sign
        GC.@preserve axis unsafe_copyto!(v, pointer(axis), Int(d))
        return Cint(0)
    catch
        return Cint(1)
    end
end

const RULE_CODES = Dict{Function,Int32}(CGE.split_cluster_rss => 0, CGE.split_cluster_size => 2,
                                        CGE.split_cluster_diameter => 3)

# size(unique(embedding, dims=1), 1) (landmarks.jl:369) on the device: cge_b200_unique_rows
function unique_rows_b200(embedding::Array{Float64,2})
    rows, dim = size(embedding)
    out = Ref{Int64}(0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cge_b200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), 0, h))
    try
        GC.@preserve embedding check(ccall((:cge_b200_unique_rows, LIB), Cint,
                                           (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ref{Int64}),
                                           h[], rows, dim, embedding, 1, rows, out))
    finally
        ccall((:cge_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return Int(out[])
end

"""
    runsplit_b200(embedding, w, initial_clusters, n, s, rule)

`CGE.runsplit` (src/landmarks.jl:279-345) with the cuts of the split rule on the device
(`cge_b200_landmarks_select`, SURVEY.md 8(f) F4): the embedding and the member order of every cluster stay
in HBM, the queue stays on the host, and the d x d eigenproblem of each cut is answered by Julia's own
`eigvecs` through a callback, so the principal axis (and its sign) is the reference's.  `rule` is one of
`split_cluster_rss`, `split_cluster_size`, `split_cluster_diameter`; anything else (`split_cluster_rss2`)
falls back to `CGE.runsplit`.  Returns the 0-based group id per vertex like `runsplit`.
"""
function runsplit_b200(embedding::Array{Float64,2}, w::Vector{Float64}, initial_clusters::Vector{Vector{Int}},
                       n::Int, s::Int, rule::Function)
    haskey(RULE_CODES, rule) || return CGE.runsplit(embedding, w, initial_clusters, n, s, rule)
    cl = sort(initial_clusters)                                               # landmarks.jl:281
    ptr = Int64[0; cumsum(length.(cl))]
    members = Int64.(reduce(vcat, cl))
    rows, dim = size(embedding)
    group = Vector{Int64}(undef, rows)
    cb = @cfunction(_principal_axis, Cint, (Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Cvoid}))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cge_b200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), 0, h))
    try
        GC.@preserve embedding w ptr members group begin
            check(ccall((:cge_b200_landmarks_select, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64},
                         Ptr{Int64}, Int32, Int64, Int64, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}),
                        h[], rows, dim, embedding, 1, rows,               # column-major: row stride 1
                        w, length(cl), ptr, members, 1, n, s, RULE_CODES[rule], cb, C_NULL, group, C_NULL))
        end
    finally
        ccall((:cge_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return group
end

"""
    landmarks_b200(edges, weights, vweights, clusters, comm, embedding, verbose, land, forced, method, directed)

`CGE.landmarks` (src/landmarks.jl:365-465) with `runsplit` (`runsplit_b200`, SURVEY.md 8(f) F4) and the
aggregation after it (:387-463, `cge_b200_landmarks_aggregate`, SURVEY.md 8(f) F2) on the device: the centroids, weights, d_ii,
landmark communities and the weighted landmark edge list come back bit-identical to the Julia loops.  Same
arguments and return tuple as `landmarks`; pass it to `wGCL` unchanged.
"""
function landmarks_b200(edges::Array{Int,2}, weights::Vector{Float64}, vweights::Vector{Float64},
                        clusters::Vector{Vector{Int}}, comm::Array{Int,2}, embedding::Array{Float64,2},
                        verbose::Bool, land::Int, forced::Int, method::Function, directed::Bool)
    rows_embed, dim = size(embedding)
    unique_rows = unique_rows_b200(embedding)                                    # landmarks.jl:369
    if land > unique_rows
        @warn "Requested number of clusters larger than unique no. embeddings. Truncating to $unique_rows landmarks."
        land = unique_rows
    end
    lm = runsplit_b200(embedding, vweights, clusters, land, forced, method) .+ 1     # landmarks.jl:378-379
    N = Int(maximum(lm)); m = size(edges, 1)
    embed = zeros(dim, N)                       # filled row-major N x dim by the library = dim x N column-major
    lweight = zeros(N); dii = zeros(N); cluster = zeros(Int, N)
    cap = min(m, N * N) + 1
    oa = zeros(Int, cap); ob = zeros(Int, cap); ow = zeros(cap); n_e = Ref{Int64}(0)
    src = edges[:, 1]; dst = edges[:, 2]; cm = vec(comm)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cge_b200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), 0, h))
    try
        GC.@preserve lm vweights cm embedding src dst weights begin
            check(ccall((:cge_b200_landmarks_aggregate, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Int64},
                         Ptr{Float64}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32,
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
                         Ptr{Float64}, Int64, Ref{Int64}),
                        h[], rows_embed, dim, N, lm, 1, vweights, cm,
                        embedding, 1, rows_embed,                 # column-major: row stride 1
                        m, src, dst, weights, directed ? 1 : 0,
                        embed, lweight, dii, cluster, oa, ob, ow, cap, n_e))
        end
    finally
        ccall((:cge_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    k = Int(n_e[])
    return dii, permutedims(embed), reshape(cluster, :, 1), hcat(oa[1:k], ob[1:k]), ow[1:k], lweight, lm
end

# `landmarks` as the CLI script calls it (example/CGE_CLI.jl:15-16): on the device (selection and aggregation),
# or the reference's own function with ENV["CGE_B200_HOST_LANDMARKS"] = "1"
landmarks(args...) = get(ENV, "CGE_B200_HOST_LANDMARKS", "0") == "1" ? CGE.landmarks(args...) :
                                                                       landmarks_b200(args...)

function score(directed::Bool, edges, eweights, comm, embed, distances, vweights, init_vweights,
               v_to_l, init_edges, init_eweights, init_embed, split, seed, auc_samples, verbose)
    no_vertices = maximum(edges)
    verbose && println("auc_samples: $auc_samples")
    lm = !isempty(v_to_l)
    verbose && println("Graph has $no_vertices vertices and $(size(edges,1)) edges")
    lm && verbose && println("Original graph has $(maximum(init_edges)) vertices and $(size(init_edges,1)) edges")
    @assert size(comm, 1) == no_vertices "No. communities not matching no. vertices"
    verbose && println("Graph has $(maximum(comm)) communities")
    verbose && println("Embedding has $(size(embed,2)) dimensions")
    @assert length(distances) == no_vertices "Distances vector length is not equal to no. vertices"

    adj_edges = lm ? init_edges : edges
    adj_w = lm ? init_eweights : eweights
    adj_n = lm ? length(init_vweights) : no_vertices
    pos_i, pos_j, pos_w, neg_i, neg_j = draw_samples(adj_edges, adj_w, adj_n, auc_samples, seed,
                                                     directed, !lm)
    src = edges[:, 1]; dst = edges[:, 2]; cm = vec(comm)
    out = zeros(Float64, 7); out_len = Ref{Int32}(7); stats = Stats()
    GC.@preserve src dst eweights cm embed distances vweights init_vweights v_to_l init_embed pos_i pos_j pos_w neg_i neg_j begin
        p = Problem(sizeof(Problem), 1, directed, split,
                    length(src), pointer(src), pointer(dst), pointer(eweights),
                    length(cm), pointer(cm),
                    pointer(embed), size(embed, 1), size(embed, 2), 1, size(embed, 1),   # column-major
                    length(distances), pointer(distances), pointer(vweights),
                    lm ? length(v_to_l) : 0,
                    lm ? pointer(init_vweights) : C_NULL, lm ? pointer(v_to_l) : C_NULL,
                    lm ? pointer(init_embed) : C_NULL, 1, lm ? size(init_embed, 1) : 0,
                    auc_samples, size(pos_i, 2),
                    pointer(pos_i), pointer(pos_j), pointer(pos_w), pointer(neg_i), pointer(neg_j),
                    0, 0, 0, 0)
        rc = ccall((:cge_b200_score, LIB), Cint,
                   (Ref{Problem}, Ptr{Float64}, Ref{Int32}, Ref{Stats}), p, out, out_len, stats)
        check(rc)
    end
    write(stderr, "."^Int(stats.n_alpha_run), "\n")                        # divergence.jl:140,255
    return out[1:out_len[]]
end

wGCL(edges::Array{Int,2}, eweights::Vector{Float64}, comm::Matrix{Int}, embed::Matrix{Float64},
     distances::Vector{Float64}, vweights::Vector{Float64}, init_vweights::Vector{Float64},
     v_to_l::Vector{Int}, init_edges::Array{Int,2}, init_eweights::Vector{Float64},
     init_embed::Matrix{Float64}, split::Bool, seed::Int=-1, auc_samples::Int=10000,
     verbose::Bool=false) =
    score(false, edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
          init_edges, init_eweights, init_embed, split, seed, auc_samples, verbose)

wGCL_directed(edges::Array{Int,2}, eweights::Vector{Float64}, comm::Matrix{Int},
              embed::Matrix{Float64}, distances::Vector{Float64}, vweights::Vector{Float64},
              init_vweights::Vector{Float64}, v_to_l::Vector{Int}, init_edges::Array{Int,2},
              init_eweights::Vector{Float64}, init_embed::Matrix{Float64}, split::Bool,
              seed::Int=-1, auc_samples::Int=10000, verbose::Bool=false) =
    score(true, edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
          init_edges, init_eweights, init_embed, split, seed, auc_samples, verbose)

end # module
