# PoleeB200.jl -- the reference-side binding: Polee's unchanged Julia host code calling libpolee_b200.so
# through `ccall`.  Load AFTER `using Polee`:
#
#     using Polee; include("PoleeB200.jl"); PoleeB200.enable!()          # then `polee prep-sample ...` as usual
#
# What it replaces (file:line in the reference):
#   approximate_likelihood(::LogitSkewNormalPTTApprox, ::RNASeqSample, ::Val{gradonly}; ...)
#                                           src/likelihood-approximation.jl:395-624  -> polee_fit
#   approximate_likelihood(::LogitSkewNormalPTTApprox, t, X, ks, efflens, output_filename; ...)
#                                           src/likelihood-approximation.jl:248-392  -> polee_fit (with ks)
#   approximate_likelihood(::OptimizePTTApprox, ::RNASeqSample)
#                                           src/likelihood-approximation.jl:149-242  -> polee_fit_optimize_ptt
# What stays Julia: RNASeqSample construction, hclust / PolyaTreeTransform construction (src/hclust.jl,
# src/ptt.jl:35-116), write_approximation (src/likelihood-approximation.jl:61-87), the CLI.
#
# NOTE: this file could not be executed in the build container (no Julia there); it is kept deliberately
# thin -- every call is one ccall with the arrays exactly as Julia already stores them (1-based UInt32 CSC
# arrays, 1-based Int32 tree arrays), so there is no per-element Julia work on the host path.
module PoleeB200

using Polee
using SparseArrays

const LIB = get(ENV, "POLEE_B200_LIB", joinpath(@__DIR__, "..", "polee_b200", "libpolee_b200.so"))

# mirrors `struct polee_opts` in include/polee_b200.h (field order and types must match)
mutable struct PoleeOpts
    device::Int32
    approx::Int32
    num_steps::Int32
    num_mc_samples::Int32
    gradonly::Int32
    use_efflen_jacobian::Int32
    noise_mode::Int32
    exact_accumulation::Int32
    seed::UInt64
    max_step_mu::Float64
    max_step_omega::Float64
    max_step_alpha::Float64
    max_step_z::Float64
    use_cuda_graph::Int32
    reserved1::Int32
    PoleeOpts() = new()
end

function default_opts()
    o = PoleeOpts()
    ccall((:polee_opts_default, LIB), Cint, (Ref{PoleeOpts},), o)
    return o
end

function check(h::Ptr{Cvoid}, rc::Cint)
    if rc != 0
        msg = unsafe_string(ccall((:polee_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
        error("libpolee_b200 (code $(rc)): $(msg)")   # same effect as the reference's @assert / error()
    end
end

function with_handle(f, opts::PoleeOpts)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:polee_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PoleeOpts}), href, opts)
    check(Ptr{Cvoid}(C_NULL), rc)
    h = href[]
    try
        return f(h)
    finally
        delete!(progress_bars, h)
        ccall((:polee_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    end
end

function set_matrix!(h, X::SparseMatrixCSC{Float32,UInt32}, ks::Union{Nothing,Vector{Int}}=nothing)
    m, n = size(X)
    check(h, ccall((:polee_set_matrix_csc, LIB), Cint,
                   (Ptr{Cvoid}, Int64, Int64, Ptr{UInt32}, Ptr{UInt32}, Ptr{Float32}, Ptr{Int64}),
                   h, m, n, X.colptr, X.rowval, X.nzval, ks === nothing ? C_NULL : ks))
end

set_efflens!(h, efflens::Vector{Float32}) =
    check(h, ccall((:polee_set_efflens, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}), h, efflens))

function set_gene_groups!(h, gene_transcripts::Dict{String,Vector{Int}})
    groups = collect(values(gene_transcripts))
    gene_ptr = Int64[0; cumsum(length.(groups))]
    transcripts = Int32[i for g in groups for i in g]      # 1-based, as stored in the Dict
    check(h, ccall((:polee_set_gene_groups, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}),
                   h, length(groups), gene_ptr, transcripts))
end

function set_tree!(h, t::Polee.PolyaTreeTransform)
    parent_idxs = Vector{Int32}(t.index[4, :])   # what the reference serialises (l-a.jl:618-621)
    js = Vector{Int32}(t.index[1, :])
    n = div(length(js) + 1, 2)
    check(h, ccall((:polee_set_tree, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}), h, n, parent_idxs, js))
    return parent_idxs, js
end

# matrix + effective lengths + tree in one call: the library prepares the tree on a second host thread while this one
# uploads X and the device builds its layout (polee_set_sample, include/polee_b200.h)
function set_sample!(h, X::SparseMatrixCSC{Float32,UInt32}, efflens::Vector{Float32}, t::Polee.PolyaTreeTransform,
                     ks::Union{Nothing,Vector{Int}}=nothing)
    m, n = size(X)
    parent_idxs = Vector{Int32}(t.index[4, :])
    js = Vector{Int32}(t.index[1, :])
    check(h, ccall((:polee_set_sample, LIB), Cint,
                   (Ptr{Cvoid}, Int64, Int64, Ptr{UInt32}, Ptr{UInt32}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32},
                    Ptr{Int32}, Ptr{Int32}),
                   h, m, n, X.colptr, X.rowval, X.nzval, ks === nothing ? C_NULL : ks, efflens, parent_idxs, js))
    return parent_idxs, js
end

function device_for_this_task()
    return Int32(parse(Int, get(ENV, "POLEE_B200_DEVICE", "0")))
end

"""
Default logit-skew-normal fit on the GPU; signature and return value of
`Polee.approximate_likelihood(::LogitSkewNormalPTTApprox, ::RNASeqSample, ::Val{gradonly}; ...)`.
"""
function approximate_likelihood_b200(approx::Polee.LogitSkewNormalPTTApprox, sample::Polee.RNASeqSample,
                                     ::Val{gradonly}=Val(true);
                                     tree_topology_input_filename=nothing,
                                     tree_topology_output_filename=nothing,
                                     gene_noninformative::Bool=false,
                                     use_efflen_jacobian::Bool=true) where {gradonly}
    # gene_id -> transcript indexes, exactly as l-a.jl:476-493
    gene_transcripts = Dict{String,Vector{Int}}()
    if gene_noninformative
        for (i, tr) in enumerate(sample.ts)
            tid = tr.metadata.name
            if haskey(sample.transcript_metadata.gene_id, tid)
                push!(get!(gene_transcripts, sample.transcript_metadata.gene_id[tid], Int[]), i)
            end
        end
        if isempty(gene_transcripts)
            @warn "'--gene-noninformative' used, but no gene information available"
            gene_noninformative = false
        end
    end
    X = sample.X
    m, n = size(X)
    tree_nodes_ref = Vector{Polee.HClustNode}[]
    if tree_topology_input_filename !== nothing            # l-a.jl:428-433
        input = Polee.h5open(tree_topology_input_filename)
        t = Polee.PolyaTreeTransform(read(input["node_parent_idxs"]), read(input["node_js"]))
        close(input)
    else
        # hclust stays on the host (l-a.jl:435); the node objects are kept for --write-tree-topology (l-a.jl:426)
        t = Polee.PolyaTreeTransform(X, approx.treemethod, tree_nodes_ref)
    end
    o = default_opts()
    o.device = device_for_this_task()
    o.approx = 0
    o.num_steps = Polee.LIKAP_NUM_STEPS
    o.num_mc_samples = Polee.LIKAP_NUM_MC_SAMPLES
    o.gradonly = gradonly ? 1 : 0
    o.use_efflen_jacobian = use_efflen_jacobian ? 1 : 0
    o.seed = rand(UInt64)                                   # drawn from Julia's (seeded) global RNG, main.jl:677
    mu = Vector{Float32}(undef, n - 1)
    omega = Vector{Float32}(undef, n - 1)
    alpha = Vector{Float32}(undef, n - 1)
    parent_idxs, js = with_handle(o) do h
        pj = set_sample!(h, X, sample.effective_lengths, t)
        gene_noninformative && set_gene_groups!(h, gene_transcripts)
        with_progress_bar(h, Polee.LIKAP_NUM_STEPS)        # the "Optimizing" bar of l-a.jl:495,574
        check(h, ccall((:polee_fit, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float64}, Ptr{Float32}),
                       h, mu, omega, alpha, C_NULL, C_NULL))
        pj
    end
    if tree_topology_output_filename !== nothing && !isempty(tree_nodes_ref)
        write_tree_topology(tree_topology_output_filename, tree_nodes_ref[1], sample, mu, omega, alpha)
    end
    params = Dict{String,Vector}("mu" => mu, "omega" => omega, "alpha" => alpha)
    if tree_topology_input_filename === nothing             # l-a.jl:618-621
        params["node_parent_idxs"] = parent_idxs
        params["node_js"] = js
    end
    return params
end

# The reference advances a ProgressMeter bar once per ADAM step (l-a.jl:495, :574).  The library enqueues the steps
# without waiting, so it reports back every PROGRESS_EVERY finished steps through polee_set_progress.
const PROGRESS_EVERY = 25
const progress_bars = Dict{Ptr{Cvoid},Any}()
function progress_thunk(done::Int32, total::Int32, user::Ptr{Cvoid})::Cvoid
    bar = get(progress_bars, user, nothing)
    bar === nothing || Polee.ProgressMeter.update!(bar, Int(done))
    return nothing
end
function with_progress_bar(h::Ptr{Cvoid}, nsteps::Integer)
    progress_bars[h] = Polee.Progress(nsteps, 0.25, "Optimizing ", 60)
    cb = @cfunction(progress_thunk, Cvoid, (Int32, Int32, Ptr{Cvoid}))
    check(h, ccall((:polee_set_progress, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32), h, cb, h, PROGRESS_EVERY))
end

"""
The optional tree diagnostics of `--write-tree-topology` (the YAML the reference prints at l-a.jl:578-613), from the
host-side cluster nodes and the fitted parameters.  Field names, order and values follow the reference's output,
including that its `right:` line carries the LEFT child's id (l-a.jl:603).
"""
function write_tree_topology(filename, nodes, sample, mu, omega, alpha)
    names = [t.metadata.name for t in sample.ts]
    md = sample.transcript_metadata
    genes = [get(md.gene_name, get(md.gene_id, nm, ""), "") for nm in names]
    id_of = IdDict{Polee.HClustNode,Int}(node => i for (i, node) in enumerate(nodes))
    open(filename, "w") do io
        println(io, "nodes:")
        k = 0
        for (i, node) in enumerate(nodes)
            println(io, "  - id: node", i)
            println(io, "    read_count: ", node.read_count)
            if node.j == 0
                k += 1
                println(io, "    mu: ", mu[k])
                println(io, "    sigma: ", exp(omega[k]))
                println(io, "    alpha: ", alpha[k])
                println(io, "    child_jaccard: ", node.child_jaccard)
                println(io, "    left: node", id_of[node.left_child])
                println(io, "    right: node", id_of[node.left_child])
            else
                println(io, "    transcript_id: ", names[node.j])
                println(io, "    gene_name: ", genes[node.j])
            end
        end
    end
end

"""Factored (salmon) variant, `src/likelihood-approximation.jl:248-392`."""
function approximate_likelihood_b200(approx::Polee.LogitSkewNormalPTTApprox, t::Polee.PolyaTreeTransform,
                                     X::SparseMatrixCSC, ks::Vector, efflens::Vector,
                                     output_filename::String; use_efflen_jacobian::Bool=true)
    m, n = size(X)
    o = default_opts()
    o.device = device_for_this_task()
    o.num_steps = Polee.LIKAP_NUM_STEPS
    o.num_mc_samples = Polee.LIKAP_NUM_MC_SAMPLES
    o.use_efflen_jacobian = use_efflen_jacobian ? 1 : 0
    o.seed = rand(UInt64)
    mu = Vector{Float32}(undef, n - 1)
    omega = Vector{Float32}(undef, n - 1)
    alpha = Vector{Float32}(undef, n - 1)
    with_handle(o) do h
        set_sample!(h, SparseMatrixCSC{Float32,UInt32}(X), Vector{Float32}(efflens), t, Vector{Int}(ks))
        check(h, ccall((:polee_fit, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float64}, Ptr{Float32}),
                       h, mu, omega, alpha, C_NULL, C_NULL))
    end
    params = Dict{String,Vector}("mu" => mu, "omega" => omega, "alpha" => alpha)
    Polee.write_approximation(output_filename, m, n, efflens, params, typeof(approx), "", "", "", "")
end

"""`approximate_likelihood(::OptimizePTTApprox, sample)`, `src/likelihood-approximation.jl:149-242`."""
function approximate_likelihood_b200(::Polee.OptimizePTTApprox, sample::Polee.RNASeqSample)
    X = sample.X
    m, n = size(X)
    o = default_opts()
    o.device = device_for_this_task()
    o.approx = 1
    o.num_steps = Polee.LIKAP_NUM_STEPS
    xs = Vector{Float32}(undef, n)
    with_handle(o) do h
        set_matrix!(h, X)
        set_efflens!(h, sample.effective_lengths)
        check(h, ccall((:polee_fit_optimize_ptt, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}), h, xs))
    end
    return Dict("x" => xs)
end

"""
Piecewise entry points for parity work from the Julia side: `log_likelihood` (src/likelihood.jl:36-56)
for one xs vector on an already prepared handle.
"""
function log_likelihood_b200(h::Ptr{Cvoid}, xs::Vector{Float32}, x_grad::Vector{Float64}, gradonly::Bool)
    lp = Ref{Float64}(0.0)
    check(h, ccall((:polee_loglik_grad, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Ref{Float64}, Ptr{Float64}),
                   h, xs, 1, gradonly ? 1 : 0, lp, x_grad))
    return lp[]
end

"""Route Polee's three fits through the GPU library (method overwrite; Polee's callers are unchanged)."""
function enable!()
    @eval Polee begin
        approximate_likelihood(approx::LogitSkewNormalPTTApprox, sample::RNASeqSample, v::Val{gradonly}=Val(true);
                               kwargs...) where {gradonly} =
            Main.PoleeB200.approximate_likelihood_b200(approx, sample, v; kwargs...)
        approximate_likelihood(approx::LogitSkewNormalPTTApprox, t::PolyaTreeTransform, X::SparseMatrixCSC,
                               ks::Vector, efflens::Vector, output_filename::String; kwargs...) =
            Main.PoleeB200.approximate_likelihood_b200(approx, t, X, ks, efflens, output_filename; kwargs...)
        approximate_likelihood(approx::OptimizePTTApprox, sample::RNASeqSample) =
            Main.PoleeB200.approximate_likelihood_b200(approx, sample)
    end
    return nothing
end

end # module
