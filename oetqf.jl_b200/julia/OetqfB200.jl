# OetqfB200.jl -- the reference-side binding of liboetqf_b200.so (ccall shim).
#
# UNTESTED IN THIS REPOSITORY'S ENVIRONMENT: neither the build container nor the GPU box has a Julia
# toolchain, so this file has never been executed.  The executable contracts of the same C ABI are the Python
# ctypes binding (oetqf.jl_b200/_lib.py) exercised by tests/ and the plain-C program tests/c/abi_smoke.c, whose
# _Static_asserts pin the struct sizes / offsets quoted below.  Struct layouts mirror include/oetqf_b200.h
# (OQ_ABI_VERSION 2) field for field.
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
# sizes / offsets as asserted in tests/c/abi_smoke.c:
#   OqFaultMesh 104 B (x @8, dx @72) | OqHex8Mesh 80 B (cx @8) | OqQuadrature 24 B | OqFaultProperty 64 B (eta @32)
#   OqMantleProperty 32 B | OqDilatancyProperty 32 B | OqSolveOptions 64 B (maxiters @40, algorithm @48,
#   async_snapshots @56) | OqSolveStats 56 B (naccept @24, retcode @48)
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

struct OqDilatancyProperty
    tp::Ptr{Float64}; eps::Ptr{Float64}; beta::Ptr{Float64}; p0::Ptr{Float64}
end

check(rc::Cint) = rc == 0 ? nothing : error(unsafe_string(ccall((:oq_last_error, LIB), Cstring, ())))
function init(device::Integer = 0)
    v = ccall((:oq_abi_version, LIB), Cint, ())
    v == 2 || error("liboetqf_b200.so speaks ABI version $v; this shim was written for version 2")
    check(ccall((:oq_init, LIB), Cint, (Cint,), device))
end

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

# Device-resident assembly: the Green's matrices are built straight into HBM as row shards (rows = this rank's fault
# cells, elems = this rank's mantle elements; 0-based half-open ranges) and never visit the host -- the path for the
# 100 GB-class matrices of the coupled configurations (GF.jl:31,123,194,250 produce host arrays instead).
function device_fault_fault(mf::RectOkadaMesh, λ::Float64, μ::Float64; ftype::FaultType = StrikeSlip(), nrept::Integer = 2,
    buffer_ratio::Real = 0, rows::UnitRange{Int} = 0:(mf.nx * mf.nξ))
    m = cmesh(mf); h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve m check(ccall((:oq_matrix_fault_fault, LIB), Cint,
        (Ref{OqFaultMesh}, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Cint, Ptr{Ptr{Cvoid}}),
        m.s, λ, μ, ftype_code(ftype), nrept, buffer_ratio, first(rows), last(rows), h))
    DeviceMatrix(h[], last(rows) - first(rows))
end

function device_fault_mantle(mf::RectOkadaMesh, ma::BEMHex8Mesh, λ::Float64, μ::Float64; ftype::FaultType = StrikeSlip(),
    qtype = "Gauss1", nrept::Integer = 2, buffer_ratio::Real = 0, elems::UnitRange{Int} = 0:length(ma.cx), form::Symbol = :dense)
    f, a, q = cmesh(mf), cmesh(ma), cquad(qtype); h = Ref{Ptr{Cvoid}}(C_NULL)
    # form = :classes keeps the operand as its table of distinct kernels (no dense storage; csrc/classmat.cuh)
    form === :classes && (GC.@preserve f a q check(ccall((:oq_matrix_fault_mantle_classes, LIB), Cint,
        (Ref{OqFaultMesh}, Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Cint, Ptr{Ptr{Cvoid}}),
        f.s, a.s, q.s, λ, μ, ftype_code(ftype), nrept, buffer_ratio, first(elems), last(elems), h));
        return DeviceMatrix(h[], 6 * (last(elems) - first(elems))))
    GC.@preserve f a q check(ccall((:oq_matrix_fault_mantle, LIB), Cint,
        (Ref{OqFaultMesh}, Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Cint, Ptr{Ptr{Cvoid}}),
        f.s, a.s, q.s, λ, μ, ftype_code(ftype), nrept, buffer_ratio, first(elems), last(elems), h))
    DeviceMatrix(h[], 6 * (last(elems) - first(elems)))
end

function device_mantle_fault(ma::BEMHex8Mesh, mf::RectOkadaMesh, λ::Float64, μ::Float64; ftype::FaultType = StrikeSlip(),
    rows::UnitRange{Int} = 0:(mf.nx * mf.nξ), form::Symbol = :dense)
    f, a = cmesh(mf), cmesh(ma); h = Ref{Ptr{Cvoid}}(C_NULL)
    form === :classes && (GC.@preserve f a check(ccall((:oq_matrix_mantle_fault_classes, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqFaultMesh}, Cdouble, Cdouble, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        a.s, f.s, λ, μ, ftype_code(ftype), first(rows), last(rows), h));
        return DeviceMatrix(h[], last(rows) - first(rows)))
    GC.@preserve f a check(ccall((:oq_matrix_mantle_fault, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqFaultMesh}, Cdouble, Cdouble, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        a.s, f.s, λ, μ, ftype_code(ftype), first(rows), last(rows), h))
    DeviceMatrix(h[], last(rows) - first(rows))
end

function device_mantle_mantle(ma::BEMHex8Mesh, λ::Float64, μ::Float64; qtype = "Gauss1", elems::UnitRange{Int} = 0:length(ma.cx),
    form::Symbol = :dense)
    a, q = cmesh(ma), cquad(qtype); h = Ref{Ptr{Cvoid}}(C_NULL)
    form === :classes && (GC.@preserve a q check(ccall((:oq_matrix_mantle_mantle_classes, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Cint, Cint, Ptr{Ptr{Cvoid}}),
        a.s, q.s, λ, μ, first(elems), last(elems), h));
        return DeviceMatrix(h[], 6 * (last(elems) - first(elems))))
    GC.@preserve a q check(ccall((:oq_matrix_mantle_mantle, LIB), Cint,
        (Ref{OqHex8Mesh}, Ref{OqQuadrature}, Cdouble, Cdouble, Cint, Cint, Ptr{Ptr{Cvoid}}),
        a.s, q.s, λ, μ, first(elems), last(elems), h))
    DeviceMatrix(h[], 6 * (last(elems) - first(elems)))
end

# the dense fault-fault shard from a Toeplitz kernel already on the host (either form GF.jl:60-70 returns)
function device_from_toeplitz(gf::AbstractArray, nx::Integer, nξ::Integer; rows::UnitRange{Int} = 0:(nx * nξ))
    st = toeplitz(gf); h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve st check(ccall((:oq_matrix_from_toeplitz, LIB), Cint, (Ptr{Cdouble}, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
        st, nx, nξ, first(rows), last(rows), h))
    DeviceMatrix(h[], last(rows) - first(rows))          # owned: the finalizer releases the nf² · 8 bytes of HBM
end

# the whole shard as the column-major (local_rows x cols) array the reference's builders would have returned
function to_host(A::DeviceMatrix)
    lr, c, gr = Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0)
    check(ccall((:oq_matrix_shape, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), A.h, lr, c, gr))
    out = Matrix{Float64}(undef, lr[], c[])
    check(ccall((:oq_matrix_to_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), A.h, out))
    out
end

# how a shard was assembled (include/oetqf_b200.h: OqAssemblyInfo; 48 bytes: Cint + pad, 2 x Int64, 3 x Cdouble)
struct OqAssemblyInfo
    path::Cint             # 0 pair, 1 tile, 2 class tables, -1 not an assembled Green's matrix
    pairs::Int64
    unique_pairs::Int64    # closed-form evaluations actually made
    table_ms::Cdouble
    expand_ms::Cdouble
    kernel_ms::Cdouble
end

# (form, device bytes) of an operand: 0 dense shard, 1 class form
function matrix_form(A::DeviceMatrix)
    f = Ref{Cint}(0); b = Ref{Cdouble}(0.0)
    check(ccall((:oq_matrix_form, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cdouble}), A.h, f, b))
    (f[] == 1 ? :classes : :dense, b[])
end

function assembly_info(A::DeviceMatrix)
    info = Ref(OqAssemblyInfo(-1, 0, 0, 0.0, 0.0, 0.0))
    check(ccall((:oq_matrix_assembly_info, LIB), Cint, (Ptr{Cvoid}, Ref{OqAssemblyInfo}), A.h, info))
    info[]
end

# host-only view of the translation classes behind device_mantle_mantle (mf === nothing) / device_mantle_fault:
# the receiver / source whose coordinates stand for each sample pair (0-based indices), and the class counts
function hex8_pair_classes(ma::BEMHex8Mesh, mf::Union{Nothing,RectOkadaMesh}, range::UnitRange{Int}, recv::Vector{Cint}, src::Vector{Cint})
    a = cmesh(ma); n = length(recv)
    reps = [Vector{Cint}(undef, max(n, 1)) for _ in 1:4]; counts = zeros(Clonglong, 3)
    if mf === nothing
        GC.@preserve a check(ccall((:oq_hex8_pair_classes, LIB), Cint,
            (Ref{OqHex8Mesh}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Clonglong}),
            a.s, C_NULL, first(range), last(range), n, recv, src, reps[1], reps[2], reps[3], reps[4], counts))
    else
        f = cmesh(mf)
        GC.@preserve a f check(ccall((:oq_hex8_pair_classes, LIB), Cint,
            (Ref{OqHex8Mesh}, Ref{OqFaultMesh}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Clonglong}),
            a.s, f.s, first(range), last(range), n, recv, src, reps[1], reps[2], reps[3], reps[4], counts))
    end
    counts, reps
end

# a window of local rows as a row-major matrix (parity checks of shards too large to duplicate)
function rows_to_host(A::DeviceMatrix, cols::Integer, rows::UnitRange{Int})
    out = Matrix{Float64}(undef, cols, last(rows) - first(rows))     # column-major (cols x rows) == row-major rows x cols
    check(ccall((:oq_matrix_rows_to_host, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}), A.h, first(rows), last(rows), out))
    permutedims(out)
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
    roots::Vector{Any}                      # the DeviceMatrix objects the problem borrows (kept alive; they own the HBM)
    up::Vector{Ptr{Float64}}                # pointer tables of the last (u, du) pair: the integrator calls `ode` with
    dup::Vector{Ptr{Float64}}               # the same cache arrays over and over, no allocation on the hot path
    last_u::UInt
    last_du::UInt
    function DeviceProblem(h, roots)
        p = new(h, roots, Ptr{Float64}[], Ptr{Float64}[], UInt(0), UInt(0))
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
    if objectid(u) != p.last_u || length(p.up) != length(u.x)
        resize!(p.up, length(u.x)); for (i, x) in enumerate(u.x); p.up[i] = pointer(x); end
        p.last_u = objectid(u)
    end
    if objectid(du) != p.last_du || length(p.dup) != length(du.x)
        resize!(p.dup, length(du.x)); for (i, x) in enumerate(du.x); p.dup[i] = pointer(x); end
        p.last_du = objectid(du)
    end
    GC.@preserve u du p check(ccall((:oq_rhs, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Ptr{Cdouble}}, Ptr{Ptr{Cdouble}}),
        p.h, t, p.up, p.dup))
    nothing
end

cprop(d::DilatancyProperty) = Rooted(OqDilatancyProperty(pointer(d.tₚ), pointer(d.ϵ), pointer(d.β), pointer(d.p₀)), Any[d])

function assemble(gf::Union{AbstractArray,DeviceMatrix}, pf::RateStateQuasiDynamicProperty, u0::ArrayPartition, tspan::NTuple{2};
    form::Symbol = :dense, dila::Union{Nothing,DilatancyProperty} = nothing)
    nx, nξ = size(u0.x[1])
    dense = gf isa DeviceMatrix || form === :dense
    g11 = gf isa DeviceMatrix ? gf : (dense ? device_from_toeplitz(gf, nx, nξ) : nothing)
    st = dense ? nothing : toeplitz(gf)
    cp = cprop(pf)
    cd = dila === nothing ? nothing : cprop(dila)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cp cd st g11 check(ccall((:oq_problem_create_fault, LIB), Cint,
        (Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cdouble}, Ref{OqFaultProperty}, Ptr{OqDilatancyProperty}, Ptr{Ptr{Cvoid}}),
        nx, nξ, dense ? 0 : 1, dense ? g11.h : C_NULL, dense ? C_NULL : pointer(st), cp.s,
        cd === nothing ? C_NULL : Base.unsafe_convert(Ptr{OqDilatancyProperty}, Ref(cd.s)), h))
    ODEProblem{true}(ode, u0, tspan, DeviceProblem(h[], Any[g11]))
end

# equation.jl:108-117
assemble(gf, pf::RateStateQuasiDynamicProperty, dila::DilatancyProperty, u0::ArrayPartition, tspan::NTuple{2}; kw...) =
    assemble(gf, pf, u0, tspan; dila = dila, kw...)

_as_device(m::DeviceMatrix; kw...) = m
_as_device(m::AbstractMatrix; mantle_rows = false) = DeviceMatrix(Matrix(m); mantle_rows = mantle_rows)

# equation.jl:141-154; every Green's argument may be a host array (as the reference's builders return it) or a
# DeviceMatrix shard built by device_* above
function assemble(gf₁₁, gf₁₂, gf₂₁, gf₂₂, pf::RateStateQuasiDynamicProperty, pa::ViscosityProperty, u0::ArrayPartition, tspan::NTuple{2})
    nx, nξ = size(u0.x[1]); ne = size(u0.x[3], 1)
    d11 = gf₁₁ isa DeviceMatrix ? gf₁₁ : device_from_toeplitz(gf₁₁, nx, nξ)
    d12, d21, d22 = _as_device(gf₁₂; mantle_rows = true), _as_device(gf₂₁), _as_device(gf₂₂; mantle_rows = true)
    cf, ca = cprop(pf), cprop(pa)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cf ca d11 d12 d21 d22 check(ccall((:oq_problem_create_viscoelastic, LIB), Cint,
        (Cint, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{OqFaultProperty}, Ref{OqMantleProperty}, Ptr{Ptr{Cvoid}}),
        nx, nξ, ne, 0, d11.h, C_NULL, d12.h, d21.h, d22.h, cf.s, ca.s, h))
    ODEProblem{true}(ode, u0, tspan, DeviceProblem(h[], Any[d11, d12, d21, d22]))
end

# ---------------------------------------------------------------- multi-GPU (one Julia process per GPU)
# Each rank exports a 128-byte handle of its exchange window; the host runtime (MPI.Allgather, Distributed, ...) gathers
# them in rank order; connect maps the peers' windows (CUDA IPC over NVLink).  After that every `ode` / solve_resident
# call all-gathers v - vpl and dϵ - dϵ₀ inside the kernels (equation.jl:197-204 on row shards).
function comm_export(p::DeviceProblem, rank::Integer, world::Integer)
    buf = Vector{UInt8}(undef, 128)
    check(ccall((:oq_comm_export, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), p.h, rank, world, buf))
    buf
end
comm_connect(p::DeviceProblem, handles::Vector{Vector{UInt8}}) =
    check(ccall((:oq_comm_connect, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), p.h, reduce(vcat, handles)))

# nevals device-resident evaluations of the RHS at the resident state; returns the CUDA-event time in ms
function rhs_resident(p::DeviceProblem, nevals::Integer)
    ms = Ref{Cdouble}(0)
    check(ccall((:oq_rhs_resident, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), p.h, nevals, ms))
    ms[]
end

# ---------------------------------------------------------------- resident mode: the whole integration on the GPU
struct OqSolveOptions
    reltol::Float64; abstol::Float64; dt0::Float64; dtmax::Float64; tstop::Float64
    maxiters::Int64; algorithm::Int32; fixed_dt::Int32; async_snapshots::Int32; reserved::Int32
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
    maxiters = 100_000, stride = 1, callback = nothing, async_snapshots::Bool = false)
    p = prob.p::DeviceProblem
    up = Ptr{Float64}[pointer(x) for x in prob.u0.x]
    GC.@preserve prob up check(ccall((:oq_state_set, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}), p.h, up))
    shapes = map(size, prob.u0.x)
    wrap(pp) = ArrayPartition((copy(unsafe_wrap(Array, unsafe_load(pp, i), shapes[i])) for i in eachindex(shapes))...)
    cb = (t, step, pu, pdu) -> callback === nothing ? false : (callback(wrap(pu), t, wrap(pdu)) === true)
    opts = OqSolveOptions(reltol, abstol, dt, dtmax, prob.tspan[2], maxiters, ALGORITHMS[alg], 0, async_snapshots, 0)
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
