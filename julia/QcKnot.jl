# QcKnot.jl -- the reference-side binding of libqcknot.so (include/qcknot.h).
#
# Written to the C header; it cannot be executed in the build container (no julia, no QuantumCollocationCore on
# disk).  It presents the five-field QuantumDynamics surface the Core's MOI evaluator consumes
# (test/scripts/integrator_test_1qubit.jl:41-52): F, ∂F, ∂F_structure, μ∂²F, μ∂²F_structure.
module QcKnot

export B200Dynamics

using Libdl

const LIB = Ref{String}(get(ENV, "QCKNOT_LIB", "libqcknot.so"))

const QCK_UNITARY_PADE, QCK_UNITARY_EXP, QCK_KET_PADE, QCK_KET_EXP, QCK_DERIVATIVE = Int32.(0:4)

struct IntegratorDesc          # qck_integrator_desc
    kind::Int32; order::Int32; levels::Int32; n_drives::Int32
    state_off::Int32; state_len::Int32; ctrl_off::Int32; reserved::Int32
    H_drift::Ptr{Float64}; H_drives::Ptr{Float64}
end

struct ProblemDesc             # qck_problem_desc
    T::Int64; zdim::Int32; dt_off::Int32; dt_fixed::Float64
    n_integrators::Int32; eval_hessian::Int32; device::Int32
    integ_begin::Int32; integ_end::Int32; reserved::Int32
    integrators::Ptr{IntegratorDesc}
end

last_error(h) = unsafe_string(ccall((:qck_last_error, LIB[]), Cstring, (Ptr{Cvoid},), h))
check(rc, h) = rc == 0 || error("libqcknot ($rc): " * last_error(h))

mutable struct B200Dynamics
    handle::Ptr{Cvoid}
    F::Function
    ∂F::Function
    ∂F_structure::Vector{Tuple{Int,Int}}
    μ∂²F::Union{Function,Nothing}
    μ∂²F_structure::Union{Vector{Tuple{Int,Int}},Nothing}
    dim::Int
end

# describe(integrator, traj) -> (kind, order, levels, n_drives, state_off, state_len, ctrl_off, H_drift, H_drives)
# reads exactly the fields the Core's integrators carry (unitary_components / state_components, drive_components,
# order, and the QuantumSystem's H_drift / H_drives); offsets are 0-based for the C side.
function describe end

"""
    B200Dynamics(integrators, traj; eval_hessian=true, device=0)

Same call shape as `QuantumDynamics(integrators, traj)`; evaluates on one B200 through libqcknot.so.
"""
function B200Dynamics(integrators, traj; eval_hessian::Bool=true, device::Integer=0)
    keep = Any[]
    descs = map(integrators) do I
        kind, order, levels, nd, soff, slen, coff, Hd, Hv = describe(I, traj)
        Hd64 = Hd === nothing ? Float64[] : collect(reinterpret(Float64, vec(ComplexF64.(Hd))))
        Hv64 = isempty(Hv) ? Float64[] : collect(reinterpret(Float64, vcat((vec(ComplexF64.(H)) for H in Hv)...)))
        push!(keep, Hd64, Hv64)
        IntegratorDesc(kind, order, levels, nd, soff, slen, coff, 0,
                       isempty(Hd64) ? C_NULL : pointer(Hd64), isempty(Hv64) ? C_NULL : pointer(Hv64))
    end
    free_time = traj.timestep isa Symbol
    dt_off = free_time ? Int32(first(traj.components[traj.timestep]) - 1) : Int32(-1)
    pd = ProblemDesc(traj.T, traj.dim, dt_off, free_time ? 0.0 : Float64(traj.timestep), length(descs),
                     eval_hessian, device, 0, 0, 0, pointer(descs))
    href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep descs begin
        rc = ccall((:qck_create, LIB[]), Cint, (Ref{ProblemDesc}, Ref{Ptr{Cvoid}}), pd, href)
    end
    rc == 0 || error("qck_create ($rc): " * last_error(C_NULL))
    h = href[]
    dyn, nnzJ, nnzH = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:qck_sizes, LIB[]), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h, dyn, nnzJ, nnzH), h)
    nb = traj.T - 1
    function structure(sym, n)
        rows, cols = Vector{Int64}(undef, n), Vector{Int64}(undef, n)
        check(ccall((sym, LIB[]), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}), h, 0, rows, cols), h)
        collect(zip(Int.(rows), Int.(cols)))
    end
    ∂F_structure = structure(:qck_jacobian_structure, nb * nnzJ[])
    μ∂²F_structure = eval_hessian ? structure(:qck_hessian_structure, nb * nnzH[]) : nothing
    F = function (Z⃗::AbstractVector{Float64})
        δ = Vector{Float64}(undef, nb * dyn[])
        check(ccall((:qck_eval_residual, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, δ), h)
        δ
    end
    ∂F = function (Z⃗::AbstractVector{Float64})
        ∂s = Vector{Float64}(undef, nb * nnzJ[])
        check(ccall((:qck_eval_jacobian, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, ∂s), h)
        ∂s
    end
    μ∂²F = eval_hessian ? function (Z⃗::AbstractVector{Float64}, μ⃗::AbstractVector{Float64})
        vals = Vector{Float64}(undef, nb * nnzH[])
        check(ccall((:qck_eval_hessian, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, μ⃗, vals), h)
        vals
    end : nothing
    D = B200Dynamics(h, F, ∂F, ∂F_structure, μ∂²F, μ∂²F_structure, dyn[])
    finalizer(d -> ccall((:qck_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), d.handle), D)
    D
end

end # module
