# QcKnot.jl -- the reference-side binding of libqcknot.so (include/qcknot.h).
#
# Written to the C header; it cannot be executed in the build container (no julia, no QuantumCollocationCore on
# disk).  It presents the five-field QuantumDynamics surface the Core's MOI evaluator consumes
# (test/scripts/integrator_test_1qubit.jl:41-52): F, ∂F, ∂F_structure, μ∂²F, μ∂²F_structure.
#
# Notes for the maintainer who wires it in:
#  * ccall needs the function NAME as a compile-time constant; only the library may be a run-time value.  Every
#    entry point is therefore resolved once with Libdl.dlsym and called through its pointer.
#  * The host-buffer entry points stage through library-owned page-locked memory: plain Vector{Float64} outputs
#    are fine.  Outputs are preallocated once, page-locked with qck_host_register (F and the Hessian values then reach
#    them by DMA without a host copy) and reused (Ipopt copies them anyway).
#  * Within one Ipopt iteration eval_g, eval_jac_g and eval_h arrive with the same Z⃗: the library compares Z⃗ with
#    its staged copy, uploads it once and reuses the device-resident results (the residual call already runs the
#    fused F + ∂F pass), so the three closures below need no cache of their own.
module QcKnot

export B200Dynamics

using Libdl

const QCK_UNITARY_PADE, QCK_UNITARY_EXP, QCK_KET_PADE, QCK_KET_EXP, QCK_DERIVATIVE = Int32.(0:4)
const QCK_SHARD_KNOT, QCK_SHARD_ENSEMBLE = Int32(0), Int32(1)

struct IntegratorDesc          # qck_integrator_desc
    kind::Int32; order::Int32; levels::Int32; n_drives::Int32
    state_off::Int32; state_len::Int32; ctrl_off::Int32; reserved::Int32
    H_drift::Ptr{Float64}; H_drives::Ptr{Float64}
end

struct ProblemDesc             # qck_problem_desc
    T::Int64; zdim::Int32; dt_off::Int32; dt_fixed::Float64
    n_integrators::Int32; eval_hessian::Int32; device::Int32
    integ_begin::Int32; integ_end::Int32; n_gpus::Int32
    integrators::Ptr{IntegratorDesc}
    shard_mode::Int32; host_threads::Int32
    devices::Ptr{Int32}
    structure_order::Int32; reserved::Int32   # QCK_ORDER_CSC = 0 | QCK_ORDER_ROW_MAJOR = 1 | QCK_ORDER_PER_INTEGRATOR = 2
end
const STRUCTURE_ORDERS = Dict(:csc => Int32(0), :row_major => Int32(1), :per_integrator => Int32(2))

# entry points, resolved once per process
struct Lib
    handle::Ptr{Cvoid}
    create::Ptr{Cvoid}; destroy::Ptr{Cvoid}; last_error::Ptr{Cvoid}; sizes::Ptr{Cvoid}
    jacobian_structure::Ptr{Cvoid}; hessian_structure::Ptr{Cvoid}
    eval_residual::Ptr{Cvoid}; eval_jacobian::Ptr{Cvoid}; eval_hessian::Ptr{Cvoid}; eval_all::Ptr{Cvoid}
    host_register::Ptr{Cvoid}; host_unregister::Ptr{Cvoid}
end
const LIB = Ref{Union{Lib,Nothing}}(nothing)
function lib()
    if LIB[] === nothing
        h = Libdl.dlopen(get(ENV, "QCKNOT_LIB", "libqcknot.so"))
        s(name) = Libdl.dlsym(h, name)
        LIB[] = Lib(h, s(:qck_create), s(:qck_destroy), s(:qck_last_error), s(:qck_sizes), s(:qck_jacobian_structure),
                    s(:qck_hessian_structure), s(:qck_eval_residual), s(:qck_eval_jacobian), s(:qck_eval_hessian), s(:qck_eval_all),
                    s(:qck_host_register), s(:qck_host_unregister))
    end
    LIB[]::Lib
end

last_error(h) = unsafe_string(ccall(lib().last_error, Cstring, (Ptr{Cvoid},), h))
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
# reads exactly the fields the Core's integrators carry; offsets are 0-based for the C side.  The field names below are
# the Core 0.3 ones as the templates use them (unitary_smooth_pulse_problem.jl:163-179,
# quantum_state_smooth_pulse_problem.jl:142-196); adapt here if a Core release renames a field -- nothing else changes.
function describe end
_range0(traj, name) = (r = traj.components[name]; (Int32(first(r) - 1), Int32(length(r))))
function _quantum(kind, I, traj, state_name, drive_name, order)
    soff, slen = _range0(traj, state_name)
    coff, _ = _range0(traj, drive_name)
    sys = I.system
    (kind, Int32(order), Int32(sys.levels), Int32(sys.n_drives), soff, slen, coff, sys.H_drift, sys.H_drives)
end
# UnitaryPadeIntegrator(state_name, control_name, system, traj; order)   unitary_smooth_pulse_problem.jl:164-167
describe_unitary_pade(I, traj) = _quantum(QCK_UNITARY_PADE, I, traj, I.unitary_name, I.drive_name, I.order)
# UnitaryExponentialIntegrator(state_name, control_name, system, traj)    unitary_smooth_pulse_problem.jl:168-170
describe_unitary_exponential(I, traj) = _quantum(QCK_UNITARY_EXP, I, traj, I.unitary_name, I.drive_name, 0)
# QuantumStatePadeIntegrator(state_name, control_name, sys, traj; order)  quantum_state_smooth_pulse_problem.jl:145-152
describe_ket_pade(I, traj) = _quantum(QCK_KET_PADE, I, traj, I.state_name, I.drive_name, I.order)
# QuantumStateExponentialIntegrator(state_name, control_name, sys, traj)   quantum_state_smooth_pulse_problem.jl:171-176
describe_ket_exponential(I, traj) = _quantum(QCK_KET_EXP, I, traj, I.state_name, I.drive_name, 0)
# DerivativeIntegrator(x_name, dx_name, traj)                               unitary_smooth_pulse_problem.jl:177-178
function describe_derivative(I, traj)
    xoff, xlen = _range0(traj, I.variable)
    doff, _ = _range0(traj, I.derivative)
    (QCK_DERIVATIVE, Int32(0), Int32(0), Int32(0), xoff, xlen, doff, nothing, Matrix{ComplexF64}[])
end
# The maintainer adds the five one-line methods that dispatch on the Core's types, e.g.
#   QcKnot.describe(I::UnitaryPadeIntegrator, traj) = QcKnot.describe_unitary_pade(I, traj)
#   QcKnot.describe(I::QuantumStateExponentialIntegrator, traj) = QcKnot.describe_ket_exponential(I, traj)

"""
    B200Dynamics(integrators, traj; eval_hessian=true, device=0, n_gpus=1, shard_mode=:knot, structure_order=:csc)

Same call shape as `QuantumDynamics(integrators, traj)`; evaluates on `n_gpus` B200s through libqcknot.so
(`shard_mode = :knot` splits the knot range, `:ensemble` the sampled systems of a UnitarySamplingProblem).
`structure_order` picks the intra-knot order of `∂F_structure` / `μ∂²F_structure` and of the value vectors (`:csc`, `:row_major`
or `:per_integrator`): set it to whatever `QuantumDynamics(integrators, traj).∂F_structure` of the installed Core shows.
"""
function B200Dynamics(integrators, traj; eval_hessian::Bool=true, device::Integer=0, n_gpus::Integer=1, shard_mode::Symbol=:knot,
                      structure_order::Symbol=:csc)
    L = lib()
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
                     eval_hessian, device, 0, -1, n_gpus, pointer(descs),
                     shard_mode === :ensemble ? QCK_SHARD_ENSEMBLE : QCK_SHARD_KNOT, 0, C_NULL,
                     STRUCTURE_ORDERS[structure_order], 0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve keep descs ccall(L.create, Cint, (Ref{ProblemDesc}, Ref{Ptr{Cvoid}}), pd, href)
    rc == 0 || error("qck_create ($rc): " * last_error(C_NULL))
    h = href[]
    dyn, nnzJ, nnzH = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall(L.sizes, Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h, dyn, nnzJ, nnzH), h)
    nb = traj.T - 1
    function structure(fptr, n)
        rows, cols = Vector{Int64}(undef, n), Vector{Int64}(undef, n)
        check(ccall(fptr, Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}), h, 0, rows, cols), h)
        collect(zip(Int.(rows), Int.(cols)))
    end
    ∂F_structure = structure(L.jacobian_structure, nb * nnzJ[])
    μ∂²F_structure = eval_hessian ? structure(L.hessian_structure, nb * nnzH[]) : nothing
    # preallocated outputs, reused by every callback (the MOI evaluator copies them into Ipopt's arrays)
    δ = Vector{Float64}(undef, nb * dyn[])
    ∂s = Vector{Float64}(undef, nb * nnzJ[])
    μ∂²s = Vector{Float64}(undef, eval_hessian ? nb * nnzH[] : 0)
    # page-locked once: the residual and Hessian values then arrive by DMA straight from the device arrays (optional; a failure
    # here only means the library stages these arrays itself)
    for v in (δ, ∂s, μ∂²s)
        isempty(v) || ccall(L.host_register, Cint, (Ptr{Cvoid}, Csize_t), v, sizeof(v))
    end
    F = function (Z⃗::AbstractVector{Float64})
        check(ccall(L.eval_residual, Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, δ), h)
        δ
    end
    ∂F = function (Z⃗::AbstractVector{Float64})
        check(ccall(L.eval_jacobian, Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, ∂s), h)
        ∂s
    end
    μ∂²F = eval_hessian ? function (Z⃗::AbstractVector{Float64}, μ⃗::AbstractVector{Float64})
        check(ccall(L.eval_hessian, Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h, Z⃗, μ⃗, μ∂²s), h)
        μ∂²s
    end : nothing
    D = B200Dynamics(h, F, ∂F, ∂F_structure, μ∂²F, μ∂²F_structure, dyn[])
    finalizer(D) do d
        for v in (δ, ∂s, μ∂²s)
            isempty(v) || ccall(lib().host_unregister, Cint, (Ptr{Cvoid},), v)
        end
        ccall(lib().destroy, Cvoid, (Ptr{Cvoid},), d.handle)
    end
    D
end

end # module
