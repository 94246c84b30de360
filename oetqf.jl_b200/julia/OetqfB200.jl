# OetqfB200.jl -- the reference-side binding of liboetqf_b200.so (ccall shim).
#
# UNTESTED IN THIS REPOSITORY'S ENVIRONMENT: neither the build container nor the GPU box has a Julia
# toolchain, so this file has never been executed.  The executable contract of the same C ABI is the Python
# ctypes binding (oetqf.jl_b200/_lib.py) exercised by tests/.  Struct layouts below mirror include/oetqf_b200.h
# field for field.
#
# Usage inside Oetqf.jl (see INTEGRATION.md):
#     using OetqfB200
#     OetqfB200.init(0)
#     gf₁₁ = OetqfB200.stress_greens_function(mf, λ, μ; buffer_ratio = 1)      # same arrays as Oetqf's
#     prob = OetqfB200.assemble(gf₁₁, gf₁₂, gf₂₁, gf₂₂, pf, pa, u0, tspan)      # ODEProblem{true} whose f ccalls oq_rhs
module OetqfB200

using Oetqf
using Oetqf: RectOkadaMesh, BEMHex8Mesh, StrikeSlip, DipSlip, FaultType,
    RateStateQuasiDynamicProperty, PowerLawViscosityProperty, CompositePowerLawViscosityProperty,
    DilatancyProperty, ViscosityProperty
using RecursiveArrayTools: ArrayPartition
using SciMLBase: ODEProblem

const LIB = get(ENV, "OETQF_B200_LIB", joinpath(@__DIR__, "..", "liboetqf_b200.so"))

# ---------------------------------------------------------------- structs (include/oetqf_b200.h)
struct OqFaultMesh
    nx::Int32; nxi::Int32
    x::Ptr{Float64}; ax0::Ptr{Float64}; ax1::Ptr{Float64}
    xi::Ptr{Float64}; axi0::Ptr{Float64}; axi1::Ptr{Float64}; y::Ptr{Float64}; z::Ptr{Float64}
    dx::Float64; dxi::Float64; dep::Float64; dip::Float64
end

struct OqHex8Mesh
    n::Int32
    cx::Ptr{Float64}; cy::Ptr{Float64}; cz::Ptr{Float64}
    qx::Ptr{Float64}; qy::Ptr{Float64}; qz::Ptr{Float64}
    dx::Ptr{Float64}; dy::Ptr{Float64}; dz::Ptr{Float64}
end

struct OqQuadrature
    nq::Int32
    coords::Ptr{Float64}; weights::Ptr{Float64}
end

struct OqFaultProperty
    a::Ptr{Float64}; b::Ptr{Float64}; L::Ptr{Float64}; sigma::Ptr{Float64}
    eta::Float64; vpl::Float64; f0::Float64; v0::Float64
end

struct OqMantleProperty
    nlaws::Int32
    gamma::Ptr{Float64}; n::Ptr{Float64}; deps0::Ptr{Float64}
end

check(rc::Cint) = rc == 0 ? nothing : error(unsafe_string(ccall((:oq_last_error, LIB), Cstring, ())))
init(device::Integer = 0) = check(ccall((:oq_init, LIB), Cint, (Cint,), device))

ftype_code(::StrikeSlip) = Cint(0)
ftype_code(::DipSlip) = Cint(1)

# arrays that must stay rooted while a C struct points into them
struct Rooted{S}
    s::S
    roots::Vector{Any}
end

function cmesh(mf::RectOkadaMesh)
    ax0 = Float64[a[1] for a in mf.ax]; ax1 = Float64[a[2] for a in mf.ax]
    aξ0 = Float64[a[1] for a in mf.aξ]; aξ1 = Float64[a[2] for a in mf.aξ]
    x, ξ, y, z = map(v -> convert(Vector{Float64}, v), (mf.x, mf.ξ, mf.y, mf.z))
    s = OqFaultMesh(mf.nx, mf.nξ, pointer(x), pointer(ax0), pointer(ax1), pointer(ξ), pointer(aξ0), pointer(aξ1),
        pointer(y), pointer(z), mf.Δx, mf.Δξ, mf.dep, mf.dip)
    Rooted(s, Any[x, ax0, ax1, ξ, aξ0, aξ1, y, z])
end

function cmesh(ma::BEMHex8Mesh)
    vs = map(v -> convert(Vector{Float64}, v), (ma.cx, ma.cy, ma.cz, ma.qx, ma.qy, ma.qz, ma.Δx, ma.Δy, ma.Δz))
    Rooted(OqHex8Mesh(length(ma.cx), map(pointer, vs)...), Any[vs...])
end

function cquad(qtype)
    coords, weights = Oetqf.get_quadrature(qtype)          # GF.jl:318-328 (Gmsh stays on the Julia side)
    c = convert(Vector{Float64}, vec(coords)); w = convert(Vector{Float64}, weights)
    Rooted(OqQuadrature(length(w), pointer(c), pointer(w)), Any[c, w])
end

# ---------------------------------------------------------------- stress_greens_function ×4 (GF.jl:31,123,194,250)
function stress_greens_function(mf::RectOkadaMesh, λ::Float64, μ::Float64;
    ftype::FaultType = StrikeSlip(), fourier::Bool = true, nrept::Integer = 2, buffer_ratio::Real = 0, kwargs...)
    @assert buffer_ratio ≥ 0 "Argument `buffer_ratio` must be ≥ 0."
    m = cmesh(mf)
    out = fourier ? Array{ComplexF64,3}(undef, mf.nx, mf.nξ, mf.nξ) : Array{Float64,3}(undef, mf.nx, mf.nξ, mf.nξ)
    GC.@preserve m out check(ccall((:oq_gf_fault_fault, LIB), Cint,
        (Ref{OqFaultMesh}, Cdouble, Cdouble, Cint, Cint, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        m.s, λ, μ, ftype_code(ftype), fourier, nrept, buffer_ratio, pointer(out), C_NULL))
    out
end

function stress_greens_function(mf::RectOkadaMesh, ma::BEMHex8Mesh, λ::Float64, μ::Float64;
    ftype::FaultType = StrikeSlip(), qtype = "Gauss1", nrept::Integer = 2, buffer_ratio::Real = 0)
    @assert buffer_ratio ≥ 0 "Argument `buffer_ratio` must be ≥ 0."
    f, a, q = cmesh(mf), cmesh(ma), cquad(qtype)
    out = Matrix{Float64}(undef, 6length(ma.cx), mf.nx * mf.nξ)
    GC.@preserve f a q out check(ccall((:oq_gf_fault_mantle, LIB), Cint,
        (Ref{OqFaultMesh}, Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Cint, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        f.s, a.s, q.s, λ, μ, ftype_code(ftype), nrept, buffer_ratio, pointer(out), C_NULL))
    out
end

function stress_greens_function(ma::BEMHex8Mesh, mf::RectOkadaMesh, λ::Float64, μ::Float64; ftype::FaultType = StrikeSlip())
    f, a = cmesh(mf), cmesh(ma)
    out = Matrix{Float64}(undef, mf.nx * mf.nξ, 6length(ma.cx))
    GC.@preserve f a out check(ccall((:oq_gf_mantle_fault, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqFaultMesh}, Cdouble, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
        a.s, f.s, λ, μ, ftype_code(ftype), pointer(out), C_NULL))
    out
end

function stress_greens_function(ma::BEMHex8Mesh, λ::Float64, μ::Float64; qtype = "Gauss1", checkeigvals::Bool = false)
    a, q = cmesh(ma), cquad(qtype)
    n = 6length(ma.cx)
    out = Matrix{Float64}(undef, n, n)
    GC.@preserve a q out check(ccall((:oq_gf_mantle_mantle, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        a.s, q.s, λ, μ, pointer(out), C_NULL))
    checkeigvals && println("Maximum real part of eigval is: ", maximum(real, Oetqf.LinearAlgebra.eigvals(out)))
    out
end

# ---------------------------------------------------------------- device matrices and the matvecmul! slot (pref.jl:15-21)
mutable struct DeviceMatrix
    h::Ptr{Cvoid}
    rows::Int
    function DeviceMatrix(h, rows)
        m = new(h, rows)
        finalizer(x -> ccall((:oq_matrix_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), m)
    end
end

function DeviceMatrix(A::Matrix{Float64}; mantle_rows::Bool = false)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    units = mantle_rows ? size(A, 1) ÷ 6 : size(A, 1)
    check(ccall((:oq_matrix_from_host, LIB), Cint, (Ptr{Cdouble}, Cint, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        A, size(A, 1), size(A, 2), mantle_rows, 0, units, h))
    DeviceMatrix(h[], size(A, 1))
end

# matvecmul!(y, A, x) and matvecmul!(y, A, x, true, true) as used at equation.jl:201-203
function matvecmul!(y::AbstractVector{Float64}, A::DeviceMatrix, x::AbstractVector{Float64}, α = nothing, β = nothing)
    acc = (α === nothing) ? 0 : 1
    check(ccall((:oq_gemv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint), A.h, x, y, acc))
    y
end

# ---------------------------------------------------------------- assemble + the in-place RHS (equation.jl:81-205)
mutable struct DeviceProblem
    h::Ptr{Cvoid}
    roots::Vector{Any}
    function DeviceProblem(h, roots)
        p = new(h, roots)
        finalizer(x -> ccall((:oq_problem_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    end
end

cprop(p::RateStateQuasiDynamicProperty) = Rooted(
    OqFaultProperty(pointer(p.a), pointer(p.b), pointer(p.L), pointer(p.σ), p.η, p.vpl, p.f₀, p.v₀), Any[p])

function cprop(p::PowerLawViscosityProperty)
    g = convert(Vector{Float64}, p.γ); n = convert(Vector{Float64}, p.n); d = convert(Vector{Float64}, p.dϵ₀)
    Rooted(OqMantleProperty(1, pointer(g), pointer(n), pointer(d)), Any[g, n, d])
end

function cprop(p::CompositePowerLawViscosityProperty)
    g = reduce(vcat, [convert(Vector{Float64}, q.γ) for q in p.piter])
    n = reduce(vcat, [convert(Vector{Float64}, q.n) for q in p.piter])
    d = convert(Vector{Float64}, p.dϵ₀)
    Rooted(OqMantleProperty(length(p.piter), pointer(g), pointer(n), pointer(d)), Any[g, n, d])
end

# the real Toeplitz kernel from either form GF.jl:60-70 returns
toeplitz(gf::Array{Float64,3}) = gf
toeplitz(gf::Array{ComplexF64,3}) = (nx = size(gf, 1); Oetqf.FFTW.irfft(gf, 2nx - 1, 1)[1:nx, :, :])

# Page-lock an array the integrator will hand to `ode` (u0, and the cache arrays of the integrator): oq_rhs then
# reads / writes it from the kernels directly instead of staging copies.  Unregister before the array is freed.
host_register(a::Array{Float64}) = check(ccall((:oq_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), a, sizeof(a)))
host_unregister(a::Array{Float64}) = check(ccall((:oq_host_unregister, LIB), Cint, (Ptr{Cvoid},), a))

# (du, u, p, t) -- exactly what OrdinaryDiffEq calls; u.x / du.x are the ArrayPartition components
function ode(du::ArrayPartition, u::ArrayPartition, p::DeviceProblem, t)
    up = Ptr{Float64}[pointer(x) for x in u.x]
    dup = Ptr{Float64}[pointer(x) for x in du.x]
    GC.@preserve u du up dup check(ccall((:oq_rhs, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Ptr{Cdouble}}, Ptr{Ptr{Cdouble}}),
        p.h, t, up, dup))
    nothing
end

function assemble(gf::AbstractArray, pf::RateStateQuasiDynamicProperty, u0::ArrayPartition, tspan::NTuple{2}; form::Symbol = :dense)
    nx, nξ = size(u0.x[1])
    st = toeplitz(gf)
    g11 = Ref{Ptr{Cvoid}}(C_NULL)
    form === :dense && check(ccall((:oq_matrix_from_toeplitz, LIB), Cint, (Ptr{Cdouble}, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        st, nx, nξ, 0, nx * nξ, g11))
    cp = cprop(pf)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cp st check(ccall((:oq_problem_create_fault, LIB), Cint,
        (Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cdouble}, Ref{OqFaultProperty}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
        nx, nξ, form === :dense ? 0 : 1, g11[], form === :dense ? C_NULL : pointer(st), cp.s, C_NULL, h))
    ODEProblem{true}(ode, u0, tspan, DeviceProblem(h[], Any[g11[]]))
end

function assemble(gf₁₁::AbstractArray, gf₁₂::AbstractMatrix, gf₂₁::AbstractMatrix, gf₂₂::AbstractMatrix,
    pf::RateStateQuasiDynamicProperty, pa::ViscosityProperty, u0::ArrayPartition, tspan::NTuple{2})
    nx, nξ = size(u0.x[1]); ne = size(u0.x[3], 1)
    st = toeplitz(gf₁₁)
    g11 = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:oq_matrix_from_toeplitz, LIB), Cint, (Ptr{Cdouble}, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        st, nx, nξ, 0, nx * nξ, g11))
    d12, d21, d22 = DeviceMatrix(Matrix(gf₁₂); mantle_rows = true), DeviceMatrix(Matrix(gf₂₁)), DeviceMatrix(Matrix(gf₂₂); mantle_rows = true)
    cf, ca = cprop(pf), cprop(pa)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cf ca check(ccall((:oq_problem_create_viscoelastic, LIB), Cint,
        (Cint, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{OqFaultProperty}, Ref{OqMantleProperty}, Ptr{Ptr{Cvoid}}),
        nx, nξ, ne, 0, g11[], C_NULL, d12.h, d21.h, d22.h, cf.s, ca.s, h))
    ODEProblem{true}(ode, u0, tspan, DeviceProblem(h[], Any[g11[], d12, d21, d22]))
end

# ---------------------------------------------------------------- resident mode: the whole integration on the GPU
struct OqSolveOptions
    reltol::Float64; abstol::Float64; dt0::Float64; dtmax::Float64; tstop::Float64
    maxiters::Int64; algorithm::Int32; fixed_dt::Int32
end

struct OqSolveStats
    t::Float64; dt_last::Float64; dt_next::Float64
    naccept::Int64; nreject::Int64; nrhs::Int64
    retcode::Int32
end

const ALGORITHMS = Dict(:Tsit5 => Int32(0), :VCABM5 => Int32(1))

# snapshot trampoline: `user` points at a Julia closure (u_parts, du_parts arrive as host pointers valid for the call)
function _snapshot(user::Ptr{Cvoid}, t::Cdouble, step::Int64, pu::Ptr{Ptr{Cdouble}}, pdu::Ptr{Ptr{Cdouble}})::Cint
    f = unsafe_pointer_to_objref(user)::Function
    try
        return f(t, step, pu, pdu) ? Cint(1) : Cint(0)
    catch
        return Cint(1)          # never unwind through the C frame
    end
end

"""
    solve_resident(prob, alg = :VCABM5; reltol, abstol, dt, dtmax, maxiters, stride, callback)

Counterpart of `solve(prob, VCABM5(); ...)` (examples/otf-with-mantle.jl:160-162) with stage combinations, error
norm and step-size control on the device; `callback(u::ArrayPartition, t, du::ArrayPartition)` fires at t0 and after
every `stride`-th accepted step (the role of wsolve's FunctionCallingCallback, src/io.jl:51-58).
"""
function solve_resident(prob::ODEProblem, alg::Symbol = :VCABM5; reltol = 1e-3, abstol = 1e-6, dt = 0.0, dtmax = 0.0,
    maxiters = 100_000, stride = 1, callback = nothing)
    p = prob.p::DeviceProblem
    up = Ptr{Float64}[pointer(x) for x in prob.u0.x]
    GC.@preserve prob up check(ccall((:oq_state_set, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}), p.h, up))
    shapes = map(size, prob.u0.x)
    wrap(pp) = ArrayPartition((copy(unsafe_wrap(Array, unsafe_load(pp, i), shapes[i])) for i in eachindex(shapes))...)
    cb = (t, step, pu, pdu) -> callback === nothing ? false : (callback(wrap(pu), t, wrap(pdu)) === true)
    opts = OqSolveOptions(reltol, abstol, dt, dtmax, prob.tspan[2], maxiters, ALGORITHMS[alg], 0)
    stats = Ref(OqSolveStats(0, 0, 0, 0, 0, 0, 0))
    cfn = @cfunction(_snapshot, Cint, (Ptr{Cvoid}, Cdouble, Int64, Ptr{Ptr{Cdouble}}, Ptr{Ptr{Cdouble}}))
    GC.@preserve cb check(ccall((:oq_solve, LIB), Cint,
        (Ptr{Cvoid}, Cdouble, Ref{OqSolveOptions}, Int64, Ptr{Cvoid}, Any, Ref{OqSolveStats}),
        p.h, prob.tspan[1], opts, stride, callback === nothing ? C_NULL : cfn, cb, stats))
    u = deepcopy(prob.u0)
    up2 = Ptr{Float64}[pointer(x) for x in u.x]
    GC.@preserve u up2 check(ccall((:oq_state_get, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}), p.h, up2))
    (u = u, stats = stats[])
end

end # module
