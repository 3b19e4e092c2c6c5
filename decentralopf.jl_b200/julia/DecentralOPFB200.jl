# Julia shim over libdopf.so (include/dopf.h).  NOT EXECUTED in this repository's CI: neither the build
# container nor the GPU box has Julia (see DESIGN.md section 1); the Python mirror
# decentralopf.jl_b200/admm.py exercises the same ABI in tests/.  The shim keeps the reference's
# structures (src/structures/network_elements.jl) and call surface (ADMM, run!, calculate_iteration!,
# get_nodal_price; src/structures/admm.jl, src/optimization/run.jl, src/helpers/network_elements.jl).
#
#   include("src/structures/network_elements.jl"); include("src/cases/three_node.jl")
#   include("DecentralOPFB200.jl"); using .DecentralOPFB200
#   admm = DecentralOPFB200.ADMM(0.3, nodes, generators, storages, lines)
#   DecentralOPFB200.run!(admm); np = DecentralOPFB200.get_nodal_price(admm)
module DecentralOPFB200

using LinearAlgebra

const libdopf = get(ENV, "LIBDOPF", joinpath(@__DIR__, "..", "libdopf.so"))

struct DopfProblem          # struct dopf_problem
    N::Cint; L::Cint; T::Cint; G::Cint; S::Cint
    ptdf::Ptr{Cdouble}; f_max::Ptr{Cdouble}; demand::Ptr{Cdouble}
    gen_mc::Ptr{Cdouble}; gen_pmax::Ptr{Cdouble}; gen_node::Ptr{Cint}
    sto_mc::Ptr{Cdouble}; sto_pmax::Ptr{Cdouble}; sto_emax::Ptr{Cdouble}; sto_node::Ptr{Cint}
end

mutable struct DopfConfig   # struct dopf_config
    gamma::Cdouble; flow_weight::Cdouble; prox_weight::Cdouble; slack_mask_tol::Cdouble; eps::Cdouble
    device::Cint; hinge_capacity::Cint; use_graph::Cint; reserved::Cint
    DopfConfig() = new()
end

mutable struct DopfStatus   # struct dopf_status
    iteration::Cint; converged::Cint; conv_lambda::Cint; conv_mue::Cint; conv_rho::Cint; iterations_done::Cint
    res_lambda::Cdouble; res_mue::Cdouble; res_rho::Cdouble
    gen_corrected::Cint; sto_corrected::Cint; tight_rows::Cint; wide_rows::Cint
    launches_per_iteration::Cint; sto_cold::Cint; last_step_ms::Cdouble
    fix_sequential::Cint; reserved3::Cint
    DopfStatus() = new()
end

# PTDF exactly as src/helpers/ptdf.jl:1-41 (setup step; stays on the host)
function calculate_ptdf(nodes, lines)
    N = length(nodes); L = length(lines)
    slack = findfirst(n -> n.slack, nodes)
    A = zeros(L, N)
    for (l, line) in enumerate(lines), (n, node) in enumerate(nodes)
        A[l, n] = line.from === node ? 1.0 : (line.to === node ? -1.0 : 0.0)
    end
    B = Diagonal(Float64[line.susceptance for line in lines])
    Bl = B * A; Bn = A' * B * A
    keep = setdiff(1:N, slack)
    Binv = zeros(N, N); Binv[keep, keep] = inv(Bn[keep, keep])
    return Bl * Binv
end

mutable struct Convergence
    lambda::Bool; mue::Bool; rho::Bool; all::Bool
end

mutable struct ADMM
    handle::Ptr{Cvoid}
    gamma::Float64
    nodes; generators; storages; lines
    ptdf::Matrix{Float64}
    convergence::Convergence
    status::DopfStatus
end

check(h, rc, what) = rc == 0 || error("$what failed (rc=$rc): " * unsafe_string(ccall((:dopf_last_error, libdopf), Cstring, (Ptr{Cvoid},), h)))

"ADMM(gamma, nodes, generators, storages, lines) - structures/admm.jl:23-62; matrices cross the ABI row-major (permutedims)"
function ADMM(gamma::Float64, nodes, generators, storages, lines; flow_weight = 10.0, prox_weight = 1.0)
    node_id = Dict(objectid(n) => Cint(i - 1) for (i, n) in enumerate(nodes))
    ptdf = calculate_ptdf(nodes, lines)
    ptdf_rm = permutedims(ptdf)                               # [N x L] column-major == [L][N] row-major
    demand_rm = Float64[nodes[n].demand[t] for t in 1:length(nodes[1].demand), n in 1:length(nodes)]   # [T x N] col-major == [N][T]
    f_max = Float64[l.max_capacity for l in lines]
    gmc = Float64[g.marginal_costs for g in generators]; gpm = Float64[g.max_generation for g in generators]
    gnode = Cint[node_id[objectid(g.node)] for g in generators]
    smc = Float64[s.marginal_costs for s in storages]; spm = Float64[s.max_power for s in storages]
    sem = Float64[s.max_level for s in storages]; snode = Cint[node_id[objectid(s.node)] for s in storages]
    cfg = DopfConfig()
    ccall((:dopf_default_config, libdopf), Cvoid, (Ref{DopfConfig},), cfg)
    cfg.gamma = gamma; cfg.flow_weight = flow_weight; cfg.prox_weight = prox_weight
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve ptdf_rm demand_rm f_max gmc gpm gnode smc spm sem snode begin
        prob = DopfProblem(length(nodes), length(lines), length(nodes[1].demand), length(generators), length(storages),
                           pointer(ptdf_rm), pointer(f_max), pointer(demand_rm), pointer(gmc), pointer(gpm), pointer(gnode),
                           pointer(smc), pointer(spm), pointer(sem), pointer(snode))
        rc = ccall((:dopf_create, libdopf), Cint, (Ref{DopfProblem}, Ref{DopfConfig}, Ref{Ptr{Cvoid}}), prob, cfg, h)
        check(C_NULL, rc, "dopf_create")
    end
    admm = ADMM(h[], gamma, nodes, generators, storages, lines, ptdf, Convergence(false, false, false, false), DopfStatus())
    finalizer(a -> ccall((:dopf_destroy, libdopf), Cvoid, (Ptr{Cvoid},), a.handle), admm)
    return admm
end

function step!(admm::ADMM, iters::Integer)
    rc = ccall((:dopf_step, libdopf), Cint, (Ptr{Cvoid}, Cint, Ref{DopfStatus}), admm.handle, iters, admm.status)
    check(admm.handle, rc, "dopf_step")
    s = admm.status
    admm.convergence = Convergence(s.conv_lambda != 0, s.conv_mue != 0, s.conv_rho != 0, s.converged != 0)
    return s
end

"calculate_iteration!(admm) - optimization/run.jl:7-16"
calculate_iteration!(admm::ADMM) = step!(admm, 1)

"run!(admm) - optimization/run.jl:1-5; the loop and the convergence check stay on the device"
function run!(admm::ADMM; max_iterations = 1_000_000)
    step!(admm, max_iterations)
    println(admm.convergence.all ? "Converged" : "Not converged")
    return admm
end

iteration(admm::ADMM) = Int(admm.status.iteration)

"newest results: generation [T x G], discharge/charge/level [T x S] (column = unit), injection [T x N], flows [T x L]"
function results(admm::ADMM)
    T = length(admm.nodes[1].demand); G = length(admm.generators); S = length(admm.storages)
    N = length(admm.nodes); L = length(admm.lines)
    P = zeros(T, G); D = zeros(T, S); C = zeros(T, S); E = zeros(T, S)
    inj = zeros(T, N); flow = zeros(T, L); aU = zeros(T, L); aK = zeros(T, L)
    rc = ccall((:dopf_get_iterate, libdopf), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               admm.handle, P, D, C, E, inj, flow, aU, aK)
    check(admm.handle, rc, "dopf_get_iterate")
    return (generation = P, discharge = D, charge = C, level = E, injection = permutedims(inj),
            line_utilization = permutedims(flow), avg_U = permutedims(aU), avg_K = permutedims(aK))
end

"duals(admm; previous=false) -> (lambda[T], mue[L x T], rho[L x T])  (admm.lambdas[end] ... or the ones used by the last iteration)"
function duals(admm::ADMM; previous = false)
    T = length(admm.nodes[1].demand); L = length(admm.lines)
    lam = zeros(T); mu = zeros(T, L); rho = zeros(T, L)
    rc = ccall((:dopf_get_duals, libdopf), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), admm.handle, previous ? 1 : 0, lam, mu, rho)
    check(admm.handle, rc, "dopf_get_duals")
    return lam, permutedims(mu), permutedims(rho)
end

"get_nodal_price(admm) - helpers/network_elements.jl:16-25 evaluated at admm.iteration"
function get_nodal_price(admm::ADMM)
    T = length(admm.nodes[1].demand); N = length(admm.nodes)
    out = zeros(T, N)
    rc = ccall((:dopf_get_nodal_price, libdopf), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), admm.handle, admm.convergence.all ? 1 : 0, out)
    check(admm.handle, rc, "dopf_get_nodal_price")
    return permutedims(out)
end

end # module
