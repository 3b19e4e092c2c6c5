# Drop-in replacement of /root/reference/src/imports.jl that runs the ADMM iteration on a B200 through libdopf.so
# (C ABI: include/dopf.h).  The reference's driver stays as it is except for its first include:
#
#     include("/path/to/decentralopf.jl_b200/julia/dopf_imports.jl")    # instead of include("imports.jl")
#     include("cases/three_node.jl")
#     admm = ADMM(0.3, nodes, generators, storages, lines)              # src/opf_admm_decentral.jl:5
#     run!(admm)                                                        # :7
#     np = get_nodal_price(admm.iteration)                              # :9
#
# What is kept VERBATIM from the reference tree (included from DOPF_REFERENCE_SRC, default: the directory of the
# including script): structures/{network_elements,penalty_terms,convergence}.jl, helpers/{network_elements,ptdf,results,
# logging,penalty_terms,output}.jl and optimization/run.jl - `run!`, `calculate_iteration!`, `get_nodal_price(iteration)`,
# the accessors of helpers/results.jl, `print_results`, `export_results` are the reference's own code operating on the
# same `admm` fields (`admm.lambdas[k]`, `admm.mues[k]`, `admm.rhos[k]`, `admm.results[k].unit_to_result[unit]`, ...).
# What is replaced (this file): imports.jl (no JuMP / Gurobi), structures/results.jl, structures/admm.jl and
# optimization/{subproblems,update_duals,convergence,penalty_terms}.jl - the agent QPs, the aggregation, the dual
# update and the stop rule run on the GPU; the functions below publish their results into the same structures.
# Like the reference, all functions read the global `admm`.
#
# NOT EXECUTED in this repository (neither the build container nor the GPU box has Julia); the Python mirror
# decentralopf.jl_b200/admm.py has the same structure and is what tests/test_gpu_mirror.py drives.

using LinearAlgebra

const libdopf = get(ENV, "LIBDOPF", joinpath(@__DIR__, "..", "libdopf.so"))
const DOPF_REFERENCE_SRC = get(ENV, "DOPF_REFERENCE_SRC", dirname(abspath(PROGRAM_FILE == "" ? "src/x" : PROGRAM_FILE)))
const DOPF_TRACE = get(ENV, "DOPF_TRACE", "1") != "0"          # 1: keep the reference's growing histories (default)
const DOPF_UNIT_DETAILS = get(ENV, "DOPF_UNIT_DETAILS", "0") != "0"   # 1: fetch penalty_term, U, K of every unit every iteration

# ---- C structs of include/dopf.h ----------------------------------------------------------------------------------------
struct DopfProblem
    N::Cint; L::Cint; T::Cint; G::Cint; S::Cint
    ptdf::Ptr{Cdouble}; f_max::Ptr{Cdouble}; demand::Ptr{Cdouble}
    gen_mc::Ptr{Cdouble}; gen_pmax::Ptr{Cdouble}; gen_node::Ptr{Cint}
    sto_mc::Ptr{Cdouble}; sto_pmax::Ptr{Cdouble}; sto_emax::Ptr{Cdouble}; sto_node::Ptr{Cint}
end
mutable struct DopfConfig
    gamma::Cdouble; flow_weight::Cdouble; prox_weight::Cdouble; slack_mask_tol::Cdouble; eps::Cdouble
    device::Cint; hinge_capacity::Cint; use_graph::Cint; debug_flags::Cint; n_scenarios::Cint; gemm_ksplit::Cint
    DopfConfig() = new()
end
mutable struct DopfStatus
    iteration::Cint; converged::Cint; conv_lambda::Cint; conv_mue::Cint; conv_rho::Cint; iterations_done::Cint
    res_lambda::Cdouble; res_mue::Cdouble; res_rho::Cdouble
    gen_corrected::Cint; sto_corrected::Cint; tight_rows::Cint; wide_rows::Cint
    launches_per_iteration::Cint; sto_cold::Cint; last_step_ms::Cdouble
    fix_sequential::Cint; reserved3::Cint
    DopfStatus() = new()
end
dopf_check(h, rc, what) = rc == 0 || error("$what failed (rc=$rc): " * unsafe_string(ccall((:dopf_last_error, libdopf), Cstring, (Ptr{Cvoid},), h)))

# ---- the reference's own files, unmodified ---------------------------------------------------------------------------------
include(joinpath(DOPF_REFERENCE_SRC, "structures/network_elements.jl"))
include(joinpath(DOPF_REFERENCE_SRC, "structures/penalty_terms.jl"))
include(joinpath(DOPF_REFERENCE_SRC, "structures/convergence.jl"))

# ---- structures/results.jl: same types and fields; Result is filled from the device iterate ----------------------------
struct ResultStorage
    storage::Storage
    discharge::Vector{Float64}
    charge::Vector{Float64}
    level::Vector{Float64}
    penalty_term::PenaltyTerm
    U::Matrix{Float64}
    K::Matrix{Float64}
end
struct ResultGenerator
    generator::Generator
    generation::Vector{Float64}
    penalty_term::PenaltyTerm
    U::Matrix{Float64}
    K::Matrix{Float64}
end
mutable struct ResultNode
    node::Node
    generation::Vector{Float64}
    discharge::Vector{Float64}
    charge::Vector{Float64}
    penalty_term::PenaltyTerm
end
mutable struct Result
    unit_to_result::Dict
    node_to_result::Dict{Node, ResultNode}
    generation::Vector{Float64}
    discharge::Vector{Float64}
    charge::Vector{Float64}
    penalty_term::PenaltyTerm
    avg_U::Matrix{Float64}
    avg_K::Matrix{Float64}
    total_costs::Float64
    injection::Matrix{Float64}
    line_utilization::Matrix{Float64}
end

# ---- structures/admm.jl: the reference's fields + the device handle ---------------------------------------------------------
mutable struct ADMM
    iteration::Int
    gamma::Float64
    lambdas::Vector{Vector{Float64}}
    mues::Vector{Matrix{Float64}}
    rhos::Vector{Matrix{Float64}}
    T::Vector{Int}
    N::Vector{Int}
    L::Vector{Int}
    nodes::Vector{Node}
    generators::Vector{Generator}
    storages::Vector{Storage}
    lines::Vector{Line}
    results::Vector{Result}
    convergence::Convergence
    ptdf::Matrix{Float64}
    total_demand::Vector{Float64}
    node_id_to_demand::Dict{Int, Vector{Int}}
    node_to_id::Dict{Node, Int}
    node_to_units::Dict{Node, Vector{Union{Generator, Storage}}}
    f_max::Vector{Float64}
    handle::Ptr{Cvoid}                 # dopf_handle*
    status::DopfStatus
    pending::Bool                      # optimize_all_subproblems! ran, check_convergence! not yet

    function ADMM(gamma::Float64, nodes::Vector{Node}, generators::Vector{Generator}, storages::Vector{Storage}, lines::Vector{Line};
                  flow_weight = 10.0, prox_weight = 1.0, slack_mask_tol = 1e-2, eps = 1e-3, device = -1)
        admm = new()
        admm.iteration = 1
        admm.gamma = gamma
        admm.T = collect(1:length(nodes[1].demand)); admm.N = collect(1:length(nodes)); admm.L = collect(1:length(lines))
        admm.lambdas = [zeros(Float64, length(admm.T))]
        admm.mues = [zeros(Float64, length(admm.L), length(admm.T))]
        admm.rhos = [zeros(Float64, length(admm.L), length(admm.T))]
        admm.nodes = nodes; admm.generators = generators; admm.storages = storages; admm.lines = lines
        admm.f_max = [line.max_capacity for line in lines]
        admm.results = []
        admm.convergence = Convergence()
        admm.ptdf = calculate_ptdf(nodes, lines)                       # helpers/ptdf.jl, unmodified
        admm.total_demand = zeros(length(admm.T))
        admm.node_id_to_demand = Dict(); admm.node_to_id = Dict(); admm.node_to_units = Dict()
        for (id, node) in enumerate(nodes)
            admm.total_demand += node.demand
            admm.node_id_to_demand[id] = node.demand
            admm.node_to_id[node] = id
        end
        for unit in vcat(generators, storages)
            haskey(admm.node_to_units, unit.node) ? push!(admm.node_to_units[unit.node], unit) : (admm.node_to_units[unit.node] = [unit])
        end
        # ---- pack the structs and create the device instance (the ABI is row-major with the timestep contiguous:
        #      a Julia Matrix(T, X) IS the ABI's [X][T]; node indices are 0-based)
        ptdf_rm = permutedims(admm.ptdf)
        demand_rm = Float64[nodes[n].demand[t] for t in admm.T, n in admm.N]
        gmc = Float64[g.marginal_costs for g in generators]; gpm = Float64[g.max_generation for g in generators]
        gnode = Cint[admm.node_to_id[g.node] - 1 for g in generators]
        smc = Float64[s.marginal_costs for s in storages]; spm = Float64[s.max_power for s in storages]
        sem = Float64[s.max_level for s in storages]; snode = Cint[admm.node_to_id[s.node] - 1 for s in storages]
        cfg = DopfConfig()
        ccall((:dopf_default_config, libdopf), Cvoid, (Ref{DopfConfig},), cfg)
        cfg.gamma = gamma; cfg.flow_weight = flow_weight; cfg.prox_weight = prox_weight
        cfg.slack_mask_tol = slack_mask_tol; cfg.eps = eps; cfg.device = device
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve ptdf_rm demand_rm gmc gpm gnode smc spm sem snode begin
            prob = DopfProblem(length(nodes), length(lines), length(admm.T), length(generators), length(storages),
                               pointer(ptdf_rm), pointer(admm.f_max), pointer(demand_rm), pointer(gmc), pointer(gpm), pointer(gnode),
                               pointer(smc), pointer(spm), pointer(sem), pointer(snode))
            rc = ccall((:dopf_create, libdopf), Cint, (Ref{DopfProblem}, Ref{DopfConfig}, Ref{Ptr{Cvoid}}), prob, cfg, h)
            dopf_check(C_NULL, rc, "dopf_create")
        end
        admm.handle = h[]; admm.status = DopfStatus(); admm.pending = false
        finalizer(a -> ccall((:dopf_destroy, libdopf), Cvoid, (Ptr{Cvoid},), a.handle), admm)
        return admm
    end
end

include(joinpath(DOPF_REFERENCE_SRC, "helpers/network_elements.jl"))   # update, get_nodal_price(iteration) - unmodified
include(joinpath(DOPF_REFERENCE_SRC, "helpers/ptdf.jl"))
include(joinpath(DOPF_REFERENCE_SRC, "helpers/results.jl"))
include(joinpath(DOPF_REFERENCE_SRC, "helpers/logging.jl"))
include(joinpath(DOPF_REFERENCE_SRC, "helpers/penalty_terms.jl"))
isdefined(Main, :DataFrame) && include(joinpath(DOPF_REFERENCE_SRC, "helpers/output.jl"))   # export_results needs DataFrames / CSV

# ---- optimization/subproblems.jl -------------------------------------------------------------------------------------------
function dopf_unit_details(admm::ADMM, kind::Integer, index0::Integer)
    T = length(admm.T); L = length(admm.L)
    eb = zeros(T); up = zeros(T); lo = zeros(T); U = zeros(T, L); K = zeros(T, L)
    rc = ccall((:dopf_get_unit_penalty, libdopf), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               admm.handle, kind, index0, eb, up, lo, U, K)
    dopf_check(admm.handle, rc, "dopf_get_unit_penalty")
    return PenaltyTerm(eb, up, lo), permutedims(U), permutedims(K)
end

"optimize_all_subproblems!(admm) (subproblems.jl:1-17).  On the device the dual update and the stop rule are fused behind the
aggregation, so the whole iteration runs here; update_duals! / check_convergence! then publish what the reference computes there."
function optimize_all_subproblems!(admm::ADMM)
    admm.pending && error("optimize_all_subproblems!: update_duals! / check_convergence! of the previous call are outstanding")
    rc = ccall((:dopf_step, libdopf), Cint, (Ptr{Cvoid}, Cint, Ref{DopfStatus}), admm.handle, 1, admm.status)
    dopf_check(admm.handle, rc, "dopf_step")
    admm.pending = true
    T = length(admm.T); G = length(admm.generators); S = length(admm.storages); N = length(admm.N); L = length(admm.L)
    P = zeros(T, G); D = zeros(T, S); C = zeros(T, S); E = zeros(T, S); inj = zeros(T, N); flow = zeros(T, L); aU = zeros(T, L); aK = zeros(T, L)
    rc = ccall((:dopf_get_iterate, libdopf), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               admm.handle, P, D, C, E, inj, flow, aU, aK)
    dopf_check(admm.handle, rc, "dopf_get_iterate")
    empty_pt() = get_empty_penalty_term(); empty_m = zeros(0, 0)
    unit_to_result = Dict()
    node_to_result = Dict{Node, ResultNode}(node => ResultNode(node, zeros(T), zeros(T), zeros(T), get_empty_penalty_term()) for node in admm.nodes)
    for (i, g) in enumerate(admm.generators)
        pt, U, K = DOPF_UNIT_DETAILS ? dopf_unit_details(admm, 0, i - 1) : (empty_pt(), empty_m, empty_m)
        unit_to_result[g] = ResultGenerator(g, P[:, i], pt, U, K)
        node_to_result[g.node] = update(node_to_result[g.node], unit_to_result[g])       # helpers/network_elements.jl:1-14
    end
    for (i, s) in enumerate(admm.storages)
        pt, U, K = DOPF_UNIT_DETAILS ? dopf_unit_details(admm, 1, i - 1) : (empty_pt(), empty_m, empty_m)
        unit_to_result[s] = ResultStorage(s, D[:, i], C[:, i], E[:, i], pt, U, K)
        node_to_result[s.node] = update(node_to_result[s.node], unit_to_result[s])
    end
    eb = zeros(T); up = zeros(T); lo = zeros(T)
    rc = ccall((:dopf_get_penalty_totals, libdopf), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), admm.handle, eb, up, lo)
    dopf_check(admm.handle, rc, "dopf_get_penalty_totals")
    tc = Ref{Cdouble}(0.0)
    dopf_check(admm.handle, ccall((:dopf_get_total_costs, libdopf), Cint, (Ptr{Cvoid}, Ref{Cdouble}), admm.handle, tc), "dopf_get_total_costs")
    result = Result(unit_to_result, node_to_result, vec(sum(P, dims = 2)), vec(sum(D, dims = 2)), vec(sum(C, dims = 2)),
                    PenaltyTerm(eb, up, lo), permutedims(aU), permutedims(aK), tc[], permutedims(inj), permutedims(flow))
    DOPF_TRACE ? push!(admm.results, result) : (admm.results = [result])
end

# ---- optimization/update_duals.jl --------------------------------------------------------------------------------------------
function update_duals!(admm::ADMM)
    admm.pending || error("update_duals!: optimize_all_subproblems! has not run for this iteration")
    T = length(admm.T); L = length(admm.L)
    lam = zeros(T); mu = zeros(T, L); rho = zeros(T, L)
    rc = ccall((:dopf_get_duals, libdopf), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), admm.handle, 0, lam, mu, rho)
    dopf_check(admm.handle, rc, "dopf_get_duals")
    push!(admm.lambdas, lam); push!(admm.mues, permutedims(mu)); push!(admm.rhos, permutedims(rho))
end

# ---- optimization/convergence.jl ----------------------------------------------------------------------------------------------
function check_convergence!(admm::ADMM)
    if admm.iteration != 1
        push!(admm.convergence.lambda_res, abs.(admm.lambdas[end] - admm.lambdas[end-1]))
        push!(admm.convergence.mue_res, abs.(admm.mues[end] - admm.mues[end-1]))
        push!(admm.convergence.rho_res, abs.(admm.rhos[end] - admm.rhos[end-1]))
        s = admm.status                    # the comparison itself ran on the device right after the dual update
        admm.convergence.lambda = s.conv_lambda != 0; admm.convergence.mue = s.conv_mue != 0; admm.convergence.rho = s.conv_rho != 0
        admm.convergence.all = s.converged != 0
    end
    if admm.convergence.all
        println("Converged")
    else
        println("Not converged")
        admm.iteration += 1
    end
    admm.pending = false
    @assert admm.iteration == admm.status.iteration "host and device iteration counters diverged"
end

include(joinpath(DOPF_REFERENCE_SRC, "optimization/run.jl"))          # run!, calculate_iteration! - unmodified

"run!(admm) without the per-iteration histories: the whole loop stays on the device (the kernels evaluate the stop rule)"
function run_on_device!(admm::ADMM; max_iterations = 1_000_000)
    rc = ccall((:dopf_step, libdopf), Cint, (Ptr{Cvoid}, Cint, Ref{DopfStatus}), admm.handle, max_iterations, admm.status)
    dopf_check(admm.handle, rc, "dopf_step")
    s = admm.status
    admm.iteration = s.iteration
    admm.convergence.lambda = s.conv_lambda != 0; admm.convergence.mue = s.conv_mue != 0; admm.convergence.rho = s.conv_rho != 0
    admm.convergence.all = s.converged != 0
    return admm
end

# ---- multi-GPU (one Julia process per GPU; SURVEY.md 8(e) agent block) --------------------------------------------------
# The collectives live inside libdopf (dopf_comm_init: libnccl is loaded at run time, the iteration and its three
# ncclAllReduce are captured in one CUDA graph), so the host side only has to move 128 opaque bytes from rank 0 to the other
# ranks - with MPI.jl, a shared file, a socket - and build `admm` from ITS block of the node-sorted generators and storages
# (include/dopf.h "multi-GPU").  After this call `run!(admm)`, `calculate_iteration!(admm)` and `run_on_device!(admm)` are the
# same calls as on one GPU; the duals, flows and convergence flags are identical on all ranks.
const DOPF_COMM_ID_BYTES = 128
function dopf_comm_unique_id()
    id = zeros(UInt8, DOPF_COMM_ID_BYTES)
    rc = ccall((:dopf_comm_get_unique_id, libdopf), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("dopf_comm_get_unique_id failed (rc=$rc): " * unsafe_string(ccall((:dopf_last_error, libdopf), Cstring, (Ptr{Cvoid},), C_NULL)))
    return id
end
function dopf_comm_init!(admm::ADMM, id::Vector{UInt8}, rank::Integer, nranks::Integer, total_agents::Integer)
    length(id) == DOPF_COMM_ID_BYTES || error("the NCCL id has $DOPF_COMM_ID_BYTES bytes")
    rc = ccall((:dopf_comm_init, libdopf), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint), admm.handle, id, rank, nranks, total_agents)
    dopf_check(admm.handle, rc, "dopf_comm_init")
    return admm
end
