# SwalbeB200.jl -- Julia glue that routes Swalbe.jl's 2-D "GPU" path into libswalbe_b200.so.
#
# Drop-in: `using Swalbe, CUDA; include("SwalbeB200.jl")` adds methods on Swalbe's own GPU state types
# (CuState / CuState_thermal, src/initialize.jl:214-256), so existing scripts keep calling
# `Swalbe.filmpressure!(state, sys)`, `Swalbe.h∇p!(state)`, ..., `Swalbe.time_loop(sys, state)` unchanged and
# land in hand-written sm_100a kernels instead of CUDA.jl broadcasts.  Every `ccall` below binds one symbol of
# include/swalbe_b200.h; arrays are passed zero-copy as `CuPtr{Float64}` (Julia's column-major layout is the
# library's layout) together with CUDA.jl's task-local stream, so the calls are ordered with user broadcasts
# such as `state.height .= CUDA.adapt(CuArray, h)`.
#
# Nothing here extends a function of Base or CUDA.jl on their own types (no type piracy): methods are added to
# Swalbe's functions on Swalbe's / CUDA.jl's GPU types, everything else lives in this module's namespace.
#
# NOTE: Julia is not installed in the build image of this repository, so this file is exercised only where
# Julia + CUDA.jl exist; the tested boundary is the same C ABI driven from Python (tests/).  tests/test_abi.py
# checks this file statically: every ccall against the header prototype, the struct layouts, block balance,
# and that every non-selftest symbol of the header is bound here.  See INTEGRATION.md.
module SwalbeB200

using CUDA
import Swalbe
import Swalbe: CuState, CuState_thermal, SysConst

const lib = get(ENV, "SWALBE_B200_LIB", "libswalbe_b200")

const GPUState = Union{CuState,CuState_thermal}

# loop flags / variants of include/swalbe_b200.h
const LOOP_LAZY_POPULATIONS = Cint(1)
const LOOP_SKIP_AUX = Cint(2)
const LOOP_MOMENTS_CONSISTENT = Cint(4)
const PRESSURE_POWER_BROAD = Cint(0)
const PRESSURE_FAST = Cint(1)
const SLIP = Dict(:standard => Cint(0), :hcrit => Cint(1), :ring_riv => Cint(2))

# ---- error mapping (no exceptions cross the ABI) ----------------------------------------------------
last_error() = unsafe_string(ccall((:swalbe_last_error, lib), Cstring, ()))
version() = ccall((:swalbe_version, lib), Cint, ())
launch_count() = ccall((:swalbe_launch_count, lib), Culonglong, ())
function check(rc::Cint, nm = nothing)
    rc == 0 && return nothing
    msg = last_error()
    rc == 1 && throw(DomainError(nm, msg))          # SWALBE_ERR_DOMAIN == Julia's DomainError (src/pressure.jl:101-107)
    (rc == 2 || rc == 5) && throw(ArgumentError(msg))
    error("libswalbe_b200 error $rc: $msg")
end

stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))     # CUDA.jl task-local stream as a cudaStream_t
dims(a) = (Cint(size(a, 1)), Cint(size(a, 2)))
const NULLF = CuPtr{Float64}(0)

# cospi(θ) is evaluated by Julia (Base.cospi) and handed over as data, scalar or field
theta_args(θ::Real) = (Cdouble(cospi(θ)), NULLF, nothing)
function theta_args(θ::CuArray{Float64})
    c = cospi.(θ)                                      # one CUDA.jl broadcast; kept alive by the caller (3rd value)
    return (Cdouble(0), pointer(c), c)
end
"""cospi.(θ) through the library's own kernel (what non-Julia hosts use); Julia hosts normally broadcast `cospi.(θ)`."""
function cospi_field!(out::CuArray{Float64}, θ::CuArray{Float64})
    check(ccall((:swalbe_cospi_field, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, Csize_t, Ptr{Cvoid}),
        out, θ, length(θ), stream()))
    return out
end

# ---- per-operator array forms -------------------------------------------------------------------------
function Swalbe.equilibrium!(feq::CuArray{Float64,3}, height::CuArray{Float64,2}, velx, vely, vsq, g)
    Lx, Ly = dims(height)
    check(ccall((:swalbe_equilibrium_d2q9, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Cint, Ptr{Cvoid}),
        feq, height, velx, vely, vsq, g, Lx, Ly, stream()))
end

function Swalbe.BGKandStream!(fout::CuArray{Float64,3}, feq, ftemp, Fx, Fy, τ)
    Lx, Ly = dims(Fx)
    check(ccall((:swalbe_bgk_stream_d2q9, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Cint, Ptr{Cvoid}),
        fout, feq, ftemp, Fx, Fy, τ, Lx, Ly, stream()))
end

function Swalbe.moments!(height::CuArray{Float64,2}, velx, vely, fout)
    Lx, Ly = dims(height)
    check(ccall((:swalbe_moments_d2q9, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        height, velx, vely, fout, Lx, Ly, stream()))
end

function _filmpressure!(output, f, dgrad, γ, θ, n, m, hmin, hcrit, variant)
    Lx, Ly = dims(f)
    ct, ctf, keep = theta_args(θ)
    GC.@preserve keep check(ccall((:swalbe_filmpressure, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble,
         Cint, Cint, Cint, Ptr{Cvoid}),
        output, f, dgrad, γ, ct, ctf, n, m, hmin, hcrit, variant, Lx, Ly, stream()), (n, m))
end
# array form: fast_93 / fast_32, DomainError otherwise              src/pressure.jl:72-115
# (dgrad is accepted for signature parity and NOT maintained: the reference leaves the eight shifted copies of the
#  height in it, the fused kernels never materialise them)
Swalbe.filmpressure!(output::CuArray{Float64,2}, f, dgrad, γ, θ, n, m, hmin, hcrit) =
    _filmpressure!(output, f, dgrad, γ, θ, n, m, hmin, hcrit, PRESSURE_FAST)
# state form: power_broad, keyword overrides                        src/pressure.jl:119-155
Swalbe.filmpressure!(state::CuState, sys::SysConst; θ = sys.param.θ, γ = sys.param.γ, n = sys.param.n, m = sys.param.m,
                     hmin = sys.param.hmin, hcrit = sys.param.hcrit) =
    _filmpressure!(state.pressure, state.height, state.dgrad, γ, θ, n, m, hmin, hcrit, PRESSURE_POWER_BROAD)
# CuState_thermal goes through the array form                       src/pressure.jl:117
Swalbe.filmpressure!(state::CuState_thermal, sys::SysConst) =
    _filmpressure!(state.pressure, state.height, state.dgrad, sys.param.γ, sys.param.θ, sys.param.n, sys.param.m,
                   sys.param.hmin, sys.param.hcrit, PRESSURE_FAST)

function Swalbe.h∇p!(state::GPUState)                             # src/forcing.jl:168-187
    Lx, Ly = dims(state.height)
    check(ccall((:swalbe_hgradp, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        state.h∇px, state.h∇py, state.pressure, state.height, Lx, Ly, stream()))
end

function _grad!(ox, oy, f, a)
    Lx, Ly = dims(f)
    check(ccall((:swalbe_grad9, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        ox, oy, f, a, Lx, Ly, stream()))
end
_multiplier(a::CuArray{Float64,2}, f) = a
_multiplier(a::Real, f) = CUDA.fill(Float64(a), size(f))        # a scalar `a` broadcasts like the reference's `a .* (...)`
Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f) = _grad!(ox, oy, f, NULLF)                       # src/differences.jl:153
function Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f, a)                                        # :171
    m = _multiplier(a, f)
    GC.@preserve m _grad!(ox, oy, f, m)
end
function Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f, dgrad::CuArray{Float64,3}, a)             # :189
    m = _multiplier(a, f)
    GC.@preserve m _grad!(ox, oy, f, m)
end

function Swalbe.∇²f!(output::CuArray{Float64,2}, f, γ)            # src/differences.jl:57-75
    Lx, Ly = dims(f)
    check(ccall((:swalbe_lap9, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Cint, Ptr{Cvoid}),
        output, f, γ, Lx, Ly, stream()))
end

function _slip!(sx, sy, h, ux, uy, δ, μ, hcrit, variant)
    Lx, Ly = dims(h)
    check(ccall((:swalbe_slippage, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cint,
         Cint, Cint, Ptr{Cvoid}), sx, sy, h, ux, uy, δ, μ, hcrit, variant, Lx, Ly, stream()))
end
Swalbe.slippage!(sx::CuArray{Float64,2}, sy, h, ux, uy, δ, μ) = _slip!(sx, sy, h, ux, uy, δ, μ, 0.0, SLIP[:standard])   # src/forcing.jl:42
Swalbe.slippage!(s::GPUState, sys::SysConst) =
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, SLIP[:standard])
Swalbe.slippage2!(s::GPUState, sys::SysConst) =                                                       # :85
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, SLIP[:hcrit])
Swalbe.slippage_ring_riv!(sx::CuArray{Float64,2}, sy, h, ux, uy, δ, μ, hcrit) =
    _slip!(sx, sy, h, ux, uy, δ, μ, hcrit, SLIP[:ring_riv])                                           # :107
Swalbe.slippage_ring_riv!(s::GPUState, sys::SysConst) =
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, SLIP[:ring_riv])

"""update!(state): the inline force sum of every driver, `state.Fx .= -state.h∇px .- state.slipx` (src/simulate.jl:18-19)."""
function update!(s::CuState)
    Lx, Ly = dims(s.height)
    check(ccall((:swalbe_force_sum, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
         CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.h∇px, s.h∇py, s.slipx, s.slipy, NULLF, NULLF, Lx, Ly, stream()))
end
function update!(s::CuState_thermal)                               # scripts/Rivulet_stability.jl:123-124
    Lx, Ly = dims(s.height)
    check(ccall((:swalbe_force_sum, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
         CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.h∇px, s.h∇py, s.slipx, s.slipy, s.kbtx, s.kbty, Lx, Ly, stream()))
end

# ---- thermal noise ---------------------------------------------------------------------------------------
# The library's normals are counter-based: Philox4x32-10 keyed on (seed, step, cell).  `randn!` draws fresh numbers on
# every call, so must `thermal!`: without `step` a process-wide call counter advances the stream, without `seed` the
# key is drawn once per process from Julia's RNG.  Pass both for a reproducible (and decomposition-independent) field.
const NOISE_CALLS = Ref{UInt64}(0)
const NOISE_SEED = Ref{Union{Nothing,UInt64}}(nothing)
function default_seed()
    NOISE_SEED[] === nothing && (NOISE_SEED[] = rand(UInt64))
    return NOISE_SEED[]::UInt64
end
function next_noise_step()
    s = NOISE_CALLS[]
    NOISE_CALLS[] = s + 1
    return s
end

# thermal! has no CuState_thermal method upstream (src/forcing.jl:313 is CPU-only); this adds one
function Swalbe.thermal!(s::CuState_thermal, sys::SysConst; seed = nothing, step = nothing)
    Lx, Ly = dims(s.height)
    sd = seed === nothing ? default_seed() : UInt64(seed)
    stp = step === nothing ? next_noise_step() : UInt64(step)
    check(ccall((:swalbe_thermal, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Culonglong, Culonglong, Cint, Cint, Ptr{Cvoid}),
        s.kbtx, s.kbty, s.height, sys.param.kbt, sys.param.μ, sys.param.δ, sd, stp, Lx, Ly, stream()))
end

function Swalbe.inclination!(α::Vector, s::GPUState; t = 1000, tstart = 0, tsmooth = 1)  # src/forcing.jl:363
    Lx, Ly = dims(s.height)
    factor = 0.5 + 0.5 * tanh((t - tstart) / tsmooth)
    check(ccall((:swalbe_inclination, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.height, α[1], α[2], factor, Lx, Ly, stream()))
end

"""(min, max, sum, count(f .> thresh)) of a device field, reduced on the device (one 32-byte read-back):
`sum(state.height)` src/simulate.jl:8-14, `maximum - minimum` :56, `wetted!` src/measures.jl:13-17."""
function field_stats(f::CuArray{Float64,2}; thresh = 0.055)
    out = CUDA.zeros(Float64, 4)
    Lx, Ly = dims(f)
    check(ccall((:swalbe_field_stats, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Cint, Ptr{Cvoid}),
        out, f, thresh, Lx, Ly, stream()))
    mn, mx, sm, cnt = Array(out)
    return (min = mn, max = mx, sum = sm, count = Int(cnt))
end
function Swalbe.wetted!(area_size, state::GPUState; hthresh = 0.055)            # src/measures.jl:13-17
    push!(area_size, field_stats(state.height; thresh = hthresh).count)
    return nothing
end

# ---- fused time loop ------------------------------------------------------------------------------------
struct CState        # struct swalbe_state
    fout::CuPtr{Float64}; ftemp::CuPtr{Float64}; feq::CuPtr{Float64}
    height::CuPtr{Float64}; velx::CuPtr{Float64}; vely::CuPtr{Float64}; vsq::CuPtr{Float64}; pressure::CuPtr{Float64}
    Fx::CuPtr{Float64}; Fy::CuPtr{Float64}; slipx::CuPtr{Float64}; slipy::CuPtr{Float64}
    hgradpx::CuPtr{Float64}; hgradpy::CuPtr{Float64}; dgrad::CuPtr{Float64}; kbtx::CuPtr{Float64}; kbty::CuPtr{Float64}
end
struct CParams       # struct swalbe_params
    tau::Cdouble; mu::Cdouble; delta::Cdouble; kbt::Cdouble; gamma::Cdouble; hmin::Cdouble; hcrit::Cdouble; g::Cdouble
    n::Cint; m::Cint; cospi_theta::Cdouble; cospi_theta_field::CuPtr{Float64}; pressure_variant::Cint; slip_variant::Cint
    use_inclination::Cint; incl_ax::Cdouble; incl_ay::Cdouble; incl_factor::Cdouble; use_thermal::Cint; seed::Culonglong
end
struct CLogs         # struct swalbe_loop_logs
    hmin::CuPtr{Float64}; hmax::CuPtr{Float64}; wetted::CuPtr{Culonglong}; hthresh::Cdouble
    hsum::Ptr{Float64}; hsum_first::Cint; hsum_every::Cint
end
# (no mass log: the Julia drivers below read `sum(state.height)` between their chunks like the reference does)
CLogs(hmin, hmax, wetted, hthresh) = CLogs(hmin, hmax, wetted, hthresh, Ptr{Float64}(C_NULL), Cint(0), Cint(0))

cstate(s::CuState) = CState(pointer(s.fout), pointer(s.ftemp), pointer(s.feq), pointer(s.height), pointer(s.velx),
    pointer(s.vely), pointer(s.vsq), pointer(s.pressure), pointer(s.Fx), pointer(s.Fy), pointer(s.slipx), pointer(s.slipy),
    pointer(s.h∇px), pointer(s.h∇py), pointer(s.dgrad), NULLF, NULLF)
cstate(s::CuState_thermal) = CState(pointer(s.fout), pointer(s.ftemp), pointer(s.feq), pointer(s.height), pointer(s.velx),
    pointer(s.vely), pointer(s.vsq), pointer(s.pressure), pointer(s.Fx), pointer(s.Fy), pointer(s.slipx), pointer(s.slipy),
    pointer(s.h∇px), pointer(s.h∇py), pointer(s.dgrad), pointer(s.kbtx), pointer(s.kbty))

function cparams(p, θargs, pressure_variant, slip_variant, incl, thermal_seed)
    ct, ctf, _ = θargs
    return CParams(p.τ, p.μ, p.δ, p.kbt, p.γ, p.hmin, p.hcrit, p.g, p.n, p.m, ct, ctf, pressure_variant, slip_variant,
        incl === nothing ? 0 : 1, incl === nothing ? 0.0 : incl[1][1], incl === nothing ? 0.0 : incl[1][2],
        incl === nothing ? 0.0 : incl[2], thermal_seed === nothing ? 0 : 1, thermal_seed === nothing ? 0 : thermal_seed)
end

# A plan owns library scratch (3 moment planes) and the launch geometry of one lattice size.  CuState is an immutable
# struct (src/initialize.jl:214), so it cannot carry a finalizer and must not be held strongly by a cache: the handle
# is a small mutable object with its own finalizer, looked up by the identity of the state's (mutable) height array
# and validated through a WeakRef; handles of collected states are destroyed the next time a plan is created.
mutable struct PlanHandle
    ptr::Ptr{Cvoid}
    owner::WeakRef
    function PlanHandle(Lx::Cint, Ly::Cint, owner)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swalbe_plan_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint), h, Lx, Ly))
        return finalizer(destroy!, new(h[], WeakRef(owner)))
    end
end
function destroy!(h::PlanHandle)
    if h.ptr != C_NULL
        ccall((:swalbe_plan_destroy, lib), Cint, (Ptr{Cvoid},), h.ptr)
        h.ptr = C_NULL
    end
    return nothing
end
const ONE_CALL_SITES = 1 << 21   # lattices from this size on run a non-verbose time_loop as one library call
const plans = Dict{UInt,PlanHandle}()
function plan(state::GPUState)
    key = objectid(state.height)
    h = get(plans, key, nothing)
    if h !== nothing && h.owner.value === state.height && h.ptr != C_NULL
        return h
    end
    for (k, v) in collect(plans)                       # sweep: states that were garbage-collected since the last creation
        if v.owner.value === nothing || k == key
            destroy!(v)
            delete!(plans, k)
        end
    end
    Lx, Ly = dims(state.height)
    h = PlanHandle(Lx, Ly, state.height)
    plans[key] = h
    return h
end

"""
    fused_steps!(state, sys, nsteps; θ, slip, incl, thermal_seed, step0, logs, flags, pressure_variant)

`nsteps` iterations of the loop body of `time_loop` (src/simulate.jl:15-22), one fused kernel per step.  On return every
field of `state` holds what the reference's holds after the same steps (unless `flags` has `LOOP_SKIP_AUX`).
"""
function fused_steps!(state::GPUState, sys::SysConst, nsteps::Integer; θ = sys.param.θ, slip::Symbol = :standard,
                      incl = nothing, thermal_seed = nothing, step0 = 0, logs::Union{Nothing,CLogs} = nothing, flags = 0,
                      pressure_variant = state isa CuState_thermal ? PRESSURE_FAST : PRESSURE_POWER_BROAD,
                      host_in::Union{Nothing,Array{Float64,2}} = nothing, host_out::Union{Nothing,Array{Float64,2}} = nothing)
    θargs = theta_args(θ)
    prm = Ref(cparams(sys.param, θargs, pressure_variant, SLIP[slip], incl, thermal_seed))
    st = Ref(cstate(state))
    h = plan(state)
    keep = θargs[3]
    if host_in !== nothing || host_out !== nothing
        # the loop starts from the host matrix `host_in` and / or leaves the final height in `host_out` (both Lx x Ly,
        # ideally page-locked: CUDA.pin(h)); the copies travel in row bands behind / ahead of the first / last steps
        lg = logs === nothing ? Ref(CLogs(NULLF, NULLF, CuPtr{Culonglong}(0), 0.055)) : Ref(logs)
        GC.@preserve keep prm st h lg host_in host_out check(ccall((:swalbe_time_loop_host, lib), Cint,
            (Ptr{Cvoid}, Ptr{CState}, Ptr{CParams}, Cint, Culonglong, Cint, Ptr{CLogs}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
            h.ptr, st, prm, nsteps, step0, flags, logs === nothing ? Ptr{CLogs}(C_NULL) : lg,
            host_in === nothing ? Ptr{Float64}(C_NULL) : pointer(host_in),
            host_out === nothing ? Ptr{Float64}(C_NULL) : pointer(host_out), stream()))
        return state
    end
    if logs === nothing
        GC.@preserve keep prm st h check(ccall((:swalbe_time_loop, lib), Cint,
            (Ptr{Cvoid}, Ptr{CState}, Ptr{CParams}, Cint, Culonglong, Cint, Ptr{CLogs}, Ptr{Cvoid}),
            h.ptr, st, prm, nsteps, step0, flags, C_NULL, stream()))
    else
        lg = Ref(logs)
        GC.@preserve keep prm st h lg check(ccall((:swalbe_time_loop, lib), Cint,
            (Ptr{Cvoid}, Ptr{CState}, Ptr{CParams}, Cint, Culonglong, Cint, Ptr{CLogs}, Ptr{Cvoid}),
            h.ptr, st, prm, nsteps, step0, flags, lg, stream()))
    end
    return state
end

"""
    step!(state, sys; θ, slip, thermal_seed, t, incl, aux)

ONE fused time step for user-written loops -- the seven-call bodies of scripts/Moving_wettability_structs.jl:28-71 and
scripts/Rivulet_stability.jl:97-132 (`filmpressure!; h∇p!; slippage!; [thermal!;] F = …; equilibrium!; BGKandStream!;
moments!`) collapse into `SwalbeB200.step!(state, sys; θ = θ, t = t)`: one kernel instead of seven plus a copy (5x at
4096², DESIGN.md).  `θ`: scalar or device field; `slip`: `:standard` (slippage!), `:hcrit` (slippage2!), `:ring_riv`;
`thermal_seed`: Philox key of the in-kernel noise of a `CuState_thermal` (default: the process seed), `t` its step
counter; `aux = true` also materialises feq / pressure / h∇p / slip / F like the reference's state (default: only
height, velocities and populations are current, which is all the next step and the usual diagnostics read).
"""
function step!(state::GPUState, sys::SysConst; θ = sys.param.θ, slip::Symbol = :standard, thermal_seed = nothing, t = 0,
               incl = nothing, aux::Bool = false, consistent::Bool = t > 1)
    seed = state isa CuState_thermal ? (thermal_seed === nothing ? default_seed() : UInt64(thermal_seed)) : nothing
    # τ ≠ 1: from the second step on, height / velocity ARE the moments of the populations the previous step streamed,
    # and the kernel derives them instead of reading their planes; pass `consistent = false` for the step that follows
    # a hand-written `state.height .= …`
    flags = (aux ? Cint(0) : LOOP_SKIP_AUX) | (sys.param.τ != 1 && consistent ? LOOP_MOMENTS_CONSISTENT : Cint(0))
    return fused_steps!(state, sys, 1; θ = θ, slip = slip, incl = incl, thermal_seed = seed, step0 = t, flags = flags)
end

# The four time_loop methods of src/simulate.jl:6-96 on CuState: same mass read-back and print at the same steps
# (t % tdump == 0, BEFORE that step's update), the loop body in chunks of fused kernels between two prints.  Only the
# final chunk materialises feq / pressure / h∇p / slip / F; at τ = 1 the populations are written by the last step of a
# chunk only (ω = 0: nothing reads them in between; same bits on return); at τ ≠ 1 every chunk after the first vouches
# for its moments.  `logs`: per-step device logs (Δh, wetted!) land in slices of the caller's buffers.
function _loop(sys, state, verbose; θ = sys.param.θ, incl = nothing, hmin = nothing, hmax = nothing, wet = nothing,
               host_in = nothing, host_out = nothing)
    t, Tmax, tdump = 1, sys.param.Tmax, max(1, sys.param.tdump)
    lazy = sys.param.τ == 1 ? LOOP_LAZY_POPULATIONS : Cint(0)
    if !verbose && Tmax >= 1 && length(state.height) >= ONE_CALL_SITES
        # Large lattices, nothing to print: the whole loop is ONE library call (the reference computes `mass` at the dump
        # steps and drops it), so the sweeps of a host loop are not cut at dump steps and no chunk boundary costs a launch
        # gap.  A verbose loop keeps the chunks below and reads the mass between them like upstream.
        logs = nothing
        if hmin !== nothing || wet !== nothing
            logs = CLogs(hmin === nothing ? NULLF : pointer(hmin, 1), hmax === nothing ? NULLF : pointer(hmax, 1),
                         wet === nothing ? CuPtr{Culonglong}(0) : pointer(wet, 1), 0.055)
        end
        fused_steps!(state, sys, Tmax; θ = θ, incl = incl, logs = logs, flags = lazy, host_in = host_in, host_out = host_out)
        return state
    end
    while t <= Tmax
        if t % tdump == 0
            mass = sum(state.height)
            verbose && println("Time step $t mass is $(round(mass, digits=3))")
        end
        nxt = min(Tmax + 1, (t ÷ tdump + 1) * tdump)
        flags = lazy | (nxt <= Tmax ? LOOP_SKIP_AUX : Cint(0)) | (t > 1 && sys.param.τ != 1 ? LOOP_MOMENTS_CONSISTENT : Cint(0))
        logs = nothing
        if hmin !== nothing || wet !== nothing    # slot t of the logs == Julia index t
            logs = CLogs(hmin === nothing ? NULLF : pointer(hmin, t), hmax === nothing ? NULLF : pointer(hmax, t),
                         wet === nothing ? CuPtr{Culonglong}(0) : pointer(wet, t), 0.055)
        end
        fused_steps!(state, sys, nxt - t; θ = θ, incl = incl, logs = logs, flags = flags,
                     host_in = t == 1 ? host_in : nothing, host_out = nxt > Tmax ? host_out : nothing)
        t = nxt
    end
    return state
end

Swalbe.time_loop(sys::SysConst, state::CuState; verbose = false) = _loop(sys, state, verbose)          # :6-25
Swalbe.time_loop(sys::SysConst, state::CuState, θ; verbose = false) = _loop(sys, state, verbose; θ = θ)  # :26-45

"""
    run_host!(h_out, sys, state, h_in; θ, verbose)

The shape of every shipped GPU script -- `state.height .= CUDA.adapt(CuArray, h_in)`; `time_loop(sys, state, θ)`;
`h_out .= Array(state.height)` (scripts/Moving_wettability_structs.jl:39,60-71) -- as one call whose two PCIe copies hide
behind the first and the last steps of the loop (`swalbe_time_loop_host`).  `h_in`, `h_out`: Lx x Ly host matrices
(page-lock them once with `CUDA.pin` for the overlap).  The state and `h_out` hold what the plain sequence leaves;
`synchronize()` before reading `h_out`.
"""
function run_host!(h_out::Array{Float64,2}, sys::SysConst, state::CuState, h_in::Array{Float64,2}; θ = sys.param.θ,
                   verbose = false)
    _loop(sys, state, verbose; θ = θ, host_in = h_in, host_out = h_out)
    return h_out
end

# time_loop(sys, state, Δh::Vector): max - min of the height BEFORE every step, reduced on the device (:47-67)
function Swalbe.time_loop(sys::SysConst, state::CuState, Δh::Vector; verbose = false)
    Tmax = sys.param.Tmax
    mn, mx = CUDA.zeros(Float64, Tmax), CUDA.zeros(Float64, Tmax)
    _loop(sys, state, verbose; hmin = mn, hmax = mx)
    append!(Δh, Array(mx .- mn))
    return state
end

# time_loop(sys, state, f, measure): the callback slot (:69-96) for the two callbacks the reference's drivers pass --
# wetted! (run_dropletrelax; counted on the device every step) and inclination! (run_dropletforced; with the keyword
# defaults of src/forcing.jl:363 the ramp 0.5 + 0.5 tanh((t - tstart)/tsmooth) is the constant below)
function Swalbe.time_loop(sys::SysConst, state::CuState, f::Function, measure::Vector; verbose = false)
    Tmax = sys.param.Tmax
    if f === Swalbe.wetted!
        wet = CUDA.zeros(UInt64, Tmax)
        _loop(sys, state, verbose; wet = wet)
        append!(measure, Int.(Array(wet)))
    elseif f === Swalbe.inclination!
        _loop(sys, state, verbose; incl = (measure, 0.5 + 0.5 * tanh((1000 - 0) / 1)))
    else
        error("time_loop on CuState: only Swalbe.wetted! and Swalbe.inclination! callbacks are fused on the device; " *
              "write the loop with SwalbeB200.step! and call the callback between the steps")
    end
    return state, measure
end

# ---- on-device initial conditions and substrate motion (include/swalbe_b200.h, "initial conditions") ----------
# Methods on CuArray heights: `Swalbe.singledroplet(state.height, r, θ, c)` fills the device array without the host
# loop + upload of src/initialvalues.jl:203-224.  `j_begin` places a row slab in the global lattice (multi-GPU).
function Swalbe.singledroplet(height::CuArray{Float64,2}, radius, θ, center; precursor = 0.05, j_begin = 0)
    Lx, Ly = dims(height)
    check(ccall((:swalbe_ic_singledroplet, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint, Ptr{Cvoid}),
        height, radius, cospi(θ), center[1], center[2], precursor, Lx, Ly, j_begin, stream()))
    return height
end

function torus!(height::CuArray{Float64,2}, r₁, R₂, θ, center, hmin = 0.05; noise = 0.0, seed = 0, j_begin = 0)  # src/initialvalues.jl:144
    Lx, Ly = dims(height)
    check(ccall((:swalbe_ic_torus, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Culonglong, Cint, Cint, Cint, Ptr{Cvoid}),
        height, r₁, R₂, cospi(θ), center[1], center[2], hmin, noise, seed, Lx, Ly, j_begin, stream()))
    return height
end

function rivulet!(height::CuArray{Float64,2}, radius, θ, orientation::Symbol, center, hmin = 0.05; noise = 0.0, seed = 0, j_begin = 0)  # :69
    Lx, Ly = dims(height)
    check(ccall((:swalbe_ic_rivulet, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Cint, Cdouble, Cdouble, Cdouble, Culonglong, Cint, Cint, Cint, Ptr{Cvoid}),
        height, radius, cospi(θ), orientation == :y ? 0 : 1, center, hmin, noise, seed, Lx, Ly, j_begin, stream()))
    return height
end

function sinewave2d!(height::CuArray{Float64,2}, h₀, ϵ, kx, ky; j_begin = 0, Ly = size(height, 2))   # src/simulate.jl:350-353
    Lx, Lyl = dims(height)
    check(ccall((:swalbe_ic_sinewave2d, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
        height, h₀, ϵ, kx, ky, Lx, Ly, Lyl, j_begin, stream()))
    return height
end

# randinterface!(height, h₀, ϵ) on a device height (src/initialvalues.jl:23): a new interface per call unless seeded
function Swalbe.randinterface!(height::CuArray{Float64,2}, h₀, ϵ; seed = nothing, j_begin = 0)
    Lx, Ly = dims(height)
    sd = seed === nothing ? default_seed() + 0x9e3779b97f4a7c15 * (next_noise_step() + 1) : UInt64(seed)
    check(ccall((:swalbe_ic_randinterface, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Culonglong, Cint, Cint, Cint, Ptr{Cvoid}),
        height, h₀, ϵ, sd, Lx, Ly, j_begin, stream()))
    return height
end

"""
    run_rayleightaylor(sys; kx, ky, h₀, ϵ, verbos)

`Swalbe.run_rayleightaylor(sys, "GPU")` (src/simulate.jl:338-358) without the scalar-indexing loop of :350-353, which
cannot run on a `CuArray` (CUDA.jl forbids `state.height[i, j] = …` outside the REPL): the initial condition is written
by `sinewave2d!` on the device, the rest is the reference's body.  Upstream hook (one branch in src/simulate.jl:350):
`device == "GPU" ? SwalbeB200.sinewave2d!(state.height, h₀, ϵ, kx, ky) : (the host loop)`.
"""
function run_rayleightaylor(sys::SysConst; kx = 15, ky = 18, h₀ = 1.0, ϵ = 0.001, verbos = true)
    println("Simulating the Rayleigh Taylor instability in two dimensions")
    state = Swalbe.Sys(sys, "GPU")
    sinewave2d!(state.height, h₀, ϵ, kx, ky)
    diff = []
    Swalbe.equilibrium!(state, sys)
    Swalbe.time_loop(sys, state, diff, verbose = verbos)
    return state.height, diff
end

"""`dest[i, j] = src[i - sx, j - sy]` (periodic) on device matrices -- `circshift!(dest, src, (sx, sy))` without touching
Base's method table (GPUArrays has its own `circshift!`; this one is a single coalesced kernel on the caller's stream)."""
function circshift2d!(dest::CuArray{Float64,2}, src::CuArray{Float64,2}, shifts::Tuple{Integer,Integer})
    Lx, Ly = dims(dest)
    check(ccall((:swalbe_circshift, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
        dest, src, shifts[1], shifts[2], Lx, Ly, stream()))
    return dest
end

"""move_substrate!(θ, input, t, tmove; direction) of scripts/Moving_wettability_structs.jl:139-152 on device fields."""
function move_substrate!(θ::CuArray{Float64,2}, input::CuArray{Float64,2}, t, tmove; direction = "diagonal")
    if (t % tmove == 0) & (t > 0)
        if direction == "diagonal"
            circshift2d!(θ, input, (1, 1))
        elseif direction == "x"
            circshift2d!(θ, input, (1, 0))
        elseif direction == "y"
            circshift2d!(θ, input, (0, 1))
        end
        input .= θ
    end
    return nothing
end

# ---- multi-GPU slab runtime (swalbe_dist_*; new: the reference is single-device, SURVEY.md 8e) -------------------
# One Julia process per GPU (MPI.jl, Distributed, or plain `julia` processes started by a launcher).  Rank 0 calls
# `unique_id()` and ships the 128 bytes to the other ranks by any transport (`MPI.Bcast!(id, 0, comm)`); every rank then
# builds its `DistSim`, uploads its row slab, steps, and reads its slab back.  Rank r owns the global rows
# r*Ly/nranks+1 : (r+1)*Ly/nranks (Julia indices) of every Lx x Ly field; a slab is a contiguous `CuArray` of size
# (Lx, Ly/nranks).
const UNIQUE_ID_BYTES = 128
function unique_id()
    id = zeros(UInt8, UNIQUE_ID_BYTES)
    check(ccall((:swalbe_dist_unique_id, lib), Cint, (Ptr{UInt8},), id))
    return id
end

mutable struct DistSim
    ptr::Ptr{Cvoid}
    rank::Int
    nranks::Int
    Lx::Int
    Ly::Int          # global
    j_begin::Int     # 0-based first global row of this rank's slab
    j_count::Int     # rows of the slab
end

"""
    DistSim(sys, rank, nranks, id; θ, slip, thermal_seed, pressure_variant)

This rank's slab of an `sys.Lx x sys.Ly` simulation (`sys.Ly % nranks == 0`, slabs at least 6 rows tall).  `id`: the
bytes of `unique_id()` from rank 0 (`nothing` when `nranks == 1`).  A contact-angle FIELD is attached afterwards with
`set_theta!`.
"""
function DistSim(sys::SysConst, rank::Integer, nranks::Integer, id::Union{Nothing,Vector{UInt8}};
                 θ::Real = sys.param.θ, slip::Symbol = :standard, thermal_seed = nothing,
                 pressure_variant = PRESSURE_POWER_BROAD)
    prm = Ref(cparams(sys.param, theta_args(θ), pressure_variant, SLIP[slip], nothing, thermal_seed))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    idp = id === nothing ? Ptr{UInt8}(C_NULL) : pointer(id)
    GC.@preserve id prm check(ccall((:swalbe_dist_create, lib), Cint,
        (Ptr{Ptr{Cvoid}}, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{CParams}),
        h, idp, rank, nranks, sys.Lx, sys.Ly, prm))
    jb, jc = Ref{Cint}(0), Ref{Cint}(0)
    check(ccall((:swalbe_dist_local_rows, lib), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}), h[], jb, jc))
    sim = DistSim(h[], rank, nranks, sys.Lx, sys.Ly, jb[], jc[])
    return finalizer(close!, sim)
end

"""Destroy the slab runtime of this rank (waits for its streams; safe to call twice)."""
function close!(sim::DistSim)
    if sim.ptr != C_NULL
        ccall((:swalbe_dist_destroy, lib), Cint, (Ptr{Cvoid},), sim.ptr)
        sim.ptr = C_NULL
    end
    return nothing
end

"""The Julia index range of the global rows this rank owns: `h_global[:, rows(sim)]` is its slab."""
rows(sim::DistSim) = (sim.j_begin + 1):(sim.j_begin + sim.j_count)

_optptr(a::Nothing) = NULLF
_optptr(a::CuArray{Float64}) = pointer(a)

"""Upload this rank's slab (device arrays of size (Lx, j_count); `ftemp` (Lx, j_count, 9) is needed when τ ≠ 1) and
exchange the halo rows."""
function set_state!(sim::DistSim, height::CuArray{Float64,2}, velx::CuArray{Float64,2}, vely::CuArray{Float64,2},
                    ftemp::Union{Nothing,CuArray{Float64,3}} = nothing)
    size(height) == (sim.Lx, sim.j_count) || throw(DimensionMismatch("slab must be $(sim.Lx) x $(sim.j_count)"))
    GC.@preserve ftemp check(ccall((:swalbe_dist_set_state, lib), Cint,
        (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
        sim.ptr, height, velx, vely, _optptr(ftemp), stream()))
    return sim
end

"""Attach this rank's rows of a contact-angle field θ (angles, like `filmpressure!(state, sys, θ = θ)`); `nothing`
switches back to the scalar θ of the parameters.  Ghost rows travel through the halo exchange."""
function set_theta!(sim::DistSim, θ::Union{Nothing,CuArray{Float64,2}})
    c = θ === nothing ? nothing : cospi.(θ)
    GC.@preserve c check(ccall((:swalbe_dist_set_theta, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}),
        sim.ptr, _optptr(c), stream()))
    CUDA.synchronize()    # the temporary cospi.(θ) must outlive the device copy that reads it
    return sim
end

"""`move_substrate!` across the slabs: θ[i, j] <- θ[i - sx, j - sy] on the GLOBAL lattice (|sy| <= 3); collective."""
function shift_theta!(sim::DistSim, sx::Integer, sy::Integer)
    check(ccall((:swalbe_dist_shift_theta, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), sim.ptr, sx, sy, stream()))
    return sim
end

"""`nsteps` fused steps with the halo exchange (NCCL send/recv) overlapped with the interior update; `step0` is the
index of the first step (the thermal-noise counter, so that 1/2/4/8 GPUs draw identical noise)."""
function time_loop!(sim::DistSim, nsteps::Integer; step0::Integer = 0)
    check(ccall((:swalbe_dist_time_loop, lib), Cint, (Ptr{Cvoid}, Cint, Culonglong, Ptr{Cvoid}),
        sim.ptr, nsteps, step0, stream()))
    return sim
end

"""
    time_loop_host!(sim, nsteps; host_in, host_out, velx, vely, step0)

`time_loop!` with this rank's rows of the height coming from / going to host matrices (`Lx x j_count`, page-locked for the
overlap): the state is replaced by (`host_in`, `velx`, `vely`; velocities `nothing` = zero) and the final height lands in
`host_out` after `synchronize()`.  The planes travel in row bands behind / ahead of the first / last steps.
"""
function time_loop_host!(sim::DistSim, nsteps::Integer; host_in::Union{Nothing,Array{Float64,2}} = nothing,
                         host_out::Union{Nothing,Array{Float64,2}} = nothing, velx = nothing, vely = nothing, step0::Integer = 0)
    GC.@preserve host_in host_out velx vely check(ccall((:swalbe_dist_time_loop_host, lib), Cint,
        (Ptr{Cvoid}, Cint, Culonglong, Ptr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        sim.ptr, nsteps, step0, host_in === nothing ? Ptr{Float64}(C_NULL) : pointer(host_in), _optptr(velx), _optptr(vely),
        host_out === nothing ? Ptr{Float64}(C_NULL) : pointer(host_out), stream()))
    return sim
end

"""Copy this rank's slab out into device arrays (any of them may be `nothing`)."""
function get_state!(sim::DistSim; height = nothing, velx = nothing, vely = nothing, fout = nothing)
    check(ccall((:swalbe_dist_get_state, lib), Cint,
        (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
        sim.ptr, _optptr(height), _optptr(velx), _optptr(vely), _optptr(fout), stream()))
    return sim
end

"""(min, max, sum, count(h > thresh)) of this rank's rows; combine across ranks with min / max / + / + (e.g.
`MPI.Allreduce`): the mass print of time_loop and `wetted!` on the slab runtime."""
function height_stats(sim::DistSim; thresh = 0.055)
    out = CUDA.zeros(Float64, 4)
    check(ccall((:swalbe_dist_height_stats, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Cdouble, Ptr{Cvoid}),
        sim.ptr, out, thresh, stream()))
    mn, mx, sm, cnt = Array(out)
    return (min = mn, max = mx, sum = sm, count = Int(cnt))
end

"true when the halo rows travel as stores into the neighbours' memory (NVLink), false when through NCCL"
function uses_peer_memory(sim::DistSim)
    yes = Ref{Cint}(0)
    check(ccall((:swalbe_dist_uses_peer_memory, lib), Cint, (Ptr{Cvoid}, Ptr{Cint}), sim.ptr, yes))
    return yes[] != 0
end

"""Device time (ms) of the last `time_loop!` call, from CUDA events on the runtime's own streams (blocks until done)."""
function last_loop_ms(sim::DistSim)
    ms = Ref{Cfloat}(0)
    check(ccall((:swalbe_dist_last_loop_ms, lib), Cint, (Ptr{Cvoid}, Ptr{Cfloat}), sim.ptr, ms))
    return ms[]
end

# ---- the 1-D (D1Q3) family on the device (SURVEY.md 8f4) ----------------------------------------------------------
# Upstream the 1-D family is CPU-only: `Sys(sysc::Consts_1D)` takes no device argument and `State_1D` holds `Vector`s
# (src/initialize.jl:587-598).  `CuState_1D` is the device twin this glue adds -- same field names -- and the methods
# below put Swalbe's 1-D operators and `time_loop` on it.  L ~ 1e3 sites is pure latency on a GPU: all steps of a
# `time_loop` chunk run inside ONE persistent kernel launch with the lattice in shared memory.
struct CuState_1D <: Swalbe.LBM_state_1D
    fout::CuArray{Float64,2}; ftemp::CuArray{Float64,2}; feq::CuArray{Float64,2}      # L x 3
    height::CuArray{Float64,1}; vel::CuArray{Float64,1}; pressure::CuArray{Float64,1}
    F::CuArray{Float64,1}; slip::CuArray{Float64,1}; h∇p::CuArray{Float64,1}
    dgrad::CuArray{Float64,2}                                                          # L x 2, unused scratch
end
"""`Swalbe.Sys(sys::SysConst_1D)` on the device: height = 1, everything else 0 (src/initialize.jl:587-598)."""
CuState_1D(sys::Swalbe.SysConst_1D) = CuState_1D(CUDA.zeros(Float64, sys.L, 3), CUDA.zeros(Float64, sys.L, 3),
    CUDA.zeros(Float64, sys.L, 3), CUDA.ones(Float64, sys.L), CUDA.zeros(Float64, sys.L), CUDA.zeros(Float64, sys.L),
    CUDA.zeros(Float64, sys.L), CUDA.zeros(Float64, sys.L), CUDA.zeros(Float64, sys.L), CUDA.zeros(Float64, sys.L, 2))

struct CState1D      # struct swalbe_state_1d
    fout::CuPtr{Float64}; ftemp::CuPtr{Float64}; feq::CuPtr{Float64}
    height::CuPtr{Float64}; vel::CuPtr{Float64}; pressure::CuPtr{Float64}; F::CuPtr{Float64}; slip::CuPtr{Float64}
    hgradp::CuPtr{Float64}; dgrad::CuPtr{Float64}
    gamma::CuPtr{Float64}; dgamma::CuPtr{Float64}; kbt::CuPtr{Float64}; fbound::CuPtr{Float64}
end
cstate(s::CuState_1D; γ = NULLF, ∇γ = NULLF, kbt = NULLF, fbound = NULLF) = CState1D(pointer(s.fout), pointer(s.ftemp),
    pointer(s.feq), pointer(s.height), pointer(s.vel), pointer(s.pressure), pointer(s.F), pointer(s.slip), pointer(s.h∇p),
    pointer(s.dgrad), γ, ∇γ, kbt, fbound)

# ---- the expanded 1-D kinds (State_thermal_1D, State_gamma_1D, StateWithBound_1D; src/initialize.jl:304-341) on device
# vectors.  Upstream these states hold `Vector`s; a device twin only needs the same field names, so the array forms below
# take the CuArrays directly and a script builds its state as a NamedTuple or its own struct.
const LOOP_GAMMA_FIELD = Cint(8)
const LOOP_MARANGONI = Cint(16)

"state.F .= -state.h∇p .- state.slip .- extra  (run_gamma, src/simulate.jl:544: extra = ∇γ; thermal loops: extra = kbt)"
function force_sum3!(F::CuArray{Float64,1}, h∇p::CuArray{Float64,1}, slip::CuArray{Float64,1}, extra::CuArray{Float64,1})
    check(ccall((:swalbe_force_sum3_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        F, h∇p, slip, extra, length(F), stream()))
end

function Swalbe.thermal!(fluc::CuArray{Float64,1}, height::CuArray{Float64,1}, kᵦT, μ, δ; seed = nothing, step = nothing)  # src/forcing.jl:322
    check(ccall((:swalbe_thermal_1d, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Culonglong, Culonglong, Cint, Ptr{Cvoid}),
        fluc, height, kᵦT, μ, δ, seed === nothing ? default_seed() : UInt64(seed), step === nothing ? next_noise_step() : UInt64(step),
        length(height), stream()))
end

"inclination!(α::Float64, state::State_1D; t, tstart, tsmooth)  src/forcing.jl:379-389 on device vectors"
function inclination1d!(F::CuArray{Float64,1}, height::CuArray{Float64,1}, α::Float64; t = 1000, tstart = 0, tsmooth = 1)
    check(ccall((:swalbe_inclination_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
        F, height, α, 0.5 + 0.5 * tanh((t - tstart) / tsmooth), length(F), stream()))
end

"∇γ!(state)  src/forcing.jl:423-432 (height === nothing)  |  ∇γ!(state, sys)  :434-447 (height, δ given)"
function ∇γ!(dγ::CuArray{Float64,1}, γ::CuArray{Float64,1}; height::Union{Nothing,CuArray{Float64,1}} = nothing, δ = 0.0)
    check(ccall((:swalbe_gradgamma_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Ptr{Cvoid}),
        dγ, γ, height === nothing ? NULLF : pointer(height), δ, length(γ), stream()))
end

"""
filmpressure!(state::State_gamma_1D, sys; γ)  src/pressure.jl:284-315 (γ scalar or device vector; `ftemp`: the L x 3
populations that receive the two contributions, or nothing) and the active-matter array form with `rho`, `Gamma` (:318-338)
"""
function filmpressure_gamma!(output::CuArray{Float64,1}, f::CuArray{Float64,1}, γ, θ, n, m, hmin, hcrit;
                             rho::Union{Nothing,CuArray{Float64,1}} = nothing, Gamma = 0.0,
                             ftemp::Union{Nothing,CuArray{Float64,2}} = nothing)
    ct, ctf, keep = theta_args(θ)
    GC.@preserve keep check(ccall((:swalbe_filmpressure_gamma_1d, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, CuPtr{Float64}, Cint, Cint,
         Cdouble, Cdouble, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        output, f, γ isa Number ? γ : 0.0, γ isa Number ? NULLF : pointer(γ), rho === nothing ? NULLF : pointer(rho), Gamma, ct, ctf,
        n, m, hmin, hcrit, ftemp === nothing ? NULLF : pointer(ftemp), length(f), stream()))
end

"BGKandStream!(state::StateWithBound_1D, sys::SysConstWithBound_1D)  src/collide.jl:214-249; border = the device copies of sys.border"
function BGKandStream_bound!(fout::CuArray{Float64,2}, feq, ftemp, fbound, F::CuArray{Float64,1}, border1, border2, τ)
    check(ccall((:swalbe_bgk_stream_bound_d1q3, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint,
         Ptr{Cvoid}), fout, feq, ftemp, fbound, F, border1, border2, τ, length(F), stream()))
end

function Swalbe.update_rho!(rho::CuArray{Float64,1}, rho_int, height, dgrad, differentials; D = 1.0, M = 0.0)  # src/forcing.jl:399
    check(ccall((:swalbe_update_rho_1d, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
        rho, rho_int, height, differentials, D, M, length(rho), stream()))
end

function Swalbe.equilibrium!(feq::CuArray{Float64,2}, height::CuArray{Float64,1}, velocity, gravity)   # src/equilibrium.jl:169
    check(ccall((:swalbe_equilibrium_d1q3, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Ptr{Cvoid}),
        feq, height, velocity, gravity, length(height), stream()))
end
Swalbe.equilibrium!(s::CuState_1D, sys::Swalbe.Consts_1D) = Swalbe.equilibrium!(s.feq, s.height, s.vel, sys.param.g)

function Swalbe.BGKandStream!(fout::CuArray{Float64,2}, feq, ftemp, F::CuArray{Float64,1}, τ)            # src/collide.jl:179
    check(ccall((:swalbe_bgk_stream_d1q3, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cint, Ptr{Cvoid}),
        fout, feq, ftemp, F, τ, length(F), stream()))
end
Swalbe.BGKandStream!(s::CuState_1D, sys::Swalbe.SysConst_1D) = Swalbe.BGKandStream!(s.fout, s.feq, s.ftemp, s.F, sys.param.τ)

function Swalbe.moments!(height::CuArray{Float64,1}, vel, fout)                                          # src/moments.jl:54
    check(ccall((:swalbe_moments_d1q3, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        height, vel, fout, length(height), stream()))
end
Swalbe.moments!(s::CuState_1D) = Swalbe.moments!(s.height, s.vel, s.fout)

function _filmpressure1d!(output, f, dgrad, γ, θ, n, m, hmin, hcrit, variant)
    ct, ctf, keep = theta_args(θ)
    GC.@preserve keep check(ccall((:swalbe_filmpressure_1d, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble, Cint,
         Cint, Ptr{Cvoid}),
        output, f, dgrad, γ, ct, ctf, n, m, hmin, hcrit, variant, length(f), stream()), (n, m))
end
Swalbe.filmpressure!(output::CuArray{Float64,1}, f, dgrad, γ, θ, n, m, hmin, hcrit) =                     # src/pressure.jl:196
    _filmpressure1d!(output, f, dgrad, γ, θ, n, m, hmin, hcrit, PRESSURE_FAST)
Swalbe.filmpressure!(s::CuState_1D, sys::Swalbe.Consts_1D; θ = sys.param.θ, n = sys.param.n, m = sys.param.m,
                     hmin = sys.param.hmin, hcrit = sys.param.hcrit, γ = sys.param.γ) =                   # :230
    _filmpressure1d!(s.pressure, s.height, s.dgrad, γ, θ, n, m, hmin, hcrit, PRESSURE_POWER_BROAD)

function _grad1d!(output, f, a)
    check(ccall((:swalbe_grad_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        output, f, a, length(f), stream()))
end
Swalbe.∇f!(output::CuArray{Float64,1}, f, dgrad, a::CuArray{Float64,1}) = _grad1d!(output, f, a)          # src/differences.jl:208
Swalbe.∇f!(output::CuArray{Float64,1}, f::CuArray{Float64,1}, dgrad) = _grad1d!(output, f, NULLF)         # :220
Swalbe.h∇p!(s::CuState_1D) = _grad1d!(s.h∇p, s.pressure, s.height)                                       # src/forcing.jl:189

function Swalbe.∇²f!(output::CuArray{Float64,1}, f::CuArray{Float64,1}, dgrad)                            # src/differences.jl:77
    check(ccall((:swalbe_lap_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, Cint, Ptr{Cvoid}), output, f, length(f), stream()))
end

function Swalbe.slippage!(slip::CuArray{Float64,1}, height, vel, δ, μ)                                    # src/forcing.jl:68
    check(ccall((:swalbe_slippage_1d, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
        slip, height, vel, δ, μ, length(height), stream()))
end
Swalbe.slippage!(s::CuState_1D, sys::Swalbe.SysConst_1D) = Swalbe.slippage!(s.slip, s.height, s.vel, sys.param.δ, sys.param.μ)

"""update!(state::CuState_1D): `state.F .= -state.h∇p .- state.slip` (src/simulate.jl:110)."""
function update!(s::CuState_1D)
    check(ccall((:swalbe_force_sum_1d, lib), Cint, (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        s.F, s.h∇p, s.slip, length(s.F), stream()))
end

"""`nsteps` iterations of the 1-D loop body (src/simulate.jl:107-114) through swalbe_time_loop_1d."""
# `incl = (α, factor)`: the inclination! callback slot of time_loop(sys, state, f, measure) (src/simulate.jl:159-179);
# `γ`, `∇γ` (device vectors): the loop body of run_gamma (:541-547) -- per-site tension in the pressure, F = -h∇p - slip - ∇γ
function fused_steps!(state::CuState_1D, sys::Swalbe.SysConst_1D, nsteps::Integer; θ = sys.param.θ,
                      logs::Union{Nothing,CLogs} = nothing, flags = 0, incl = nothing,
                      γ::Union{Nothing,CuArray{Float64,1}} = nothing, ∇γ::Union{Nothing,CuArray{Float64,1}} = nothing)
    θargs = theta_args(θ)
    prm = Ref(cparams(sys.param, θargs, PRESSURE_POWER_BROAD, SLIP[:standard],
                      incl === nothing ? nothing : ((incl[1], 0.0), incl[2]), nothing))
    st = Ref(cstate(state; γ = γ === nothing ? NULLF : pointer(γ), ∇γ = ∇γ === nothing ? NULLF : pointer(∇γ)))
    flags = Cint(flags) | (γ === nothing ? Cint(0) : LOOP_GAMMA_FIELD) | (∇γ === nothing ? Cint(0) : LOOP_MARANGONI)
    keep = θargs[3]
    if logs === nothing
        GC.@preserve keep prm st check(ccall((:swalbe_time_loop_1d, lib), Cint,
            (Ptr{CState1D}, Ptr{CParams}, Cint, Cint, Cint, Ptr{CLogs}, Ptr{Cvoid}),
            st, prm, sys.L, nsteps, flags, C_NULL, stream()))
    else
        lg = Ref(logs)
        GC.@preserve keep prm st lg check(ccall((:swalbe_time_loop_1d, lib), Cint,
            (Ptr{CState1D}, Ptr{CParams}, Cint, Cint, Cint, Ptr{CLogs}, Ptr{Cvoid}),
            st, prm, sys.L, nsteps, flags, lg, stream()))
    end
    return state
end

# time_loop(sys::SysConst_1D, state[, θ | Δh])  src/simulate.jl:98-157: the mass read-back / print of the reference at
# t % tdump == 0, the steps between two prints inside one launch
function _loop1d(sys, state, verbose; θ = sys.param.θ, hmin = nothing, hmax = nothing)
    t, Tmax, tdump = 1, sys.param.Tmax, max(1, sys.param.tdump)
    while t <= Tmax
        if t % tdump == 0
            mass = sum(state.height)
            verbose && println("Time step $t mass is $(round(mass, digits=3))")
        end
        nxt = min(Tmax + 1, (t ÷ tdump + 1) * tdump)
        logs = hmin === nothing ? nothing : CLogs(pointer(hmin, t), pointer(hmax, t), CuPtr{Culonglong}(0), 0.055)
        fused_steps!(state, sys, nxt - t; θ = θ, logs = logs, flags = nxt <= Tmax ? LOOP_SKIP_AUX : Cint(0))
        t = nxt
    end
    return state
end
Swalbe.time_loop(sys::Swalbe.SysConst_1D, state::CuState_1D; verbose = false) = _loop1d(sys, state, verbose)
Swalbe.time_loop(sys::Swalbe.SysConst_1D, state::CuState_1D, θ; verbose = false) = _loop1d(sys, state, verbose; θ = θ)
function Swalbe.time_loop(sys::Swalbe.SysConst_1D, state::CuState_1D, Δh::Vector; verbose = false)
    mn, mx = CUDA.zeros(Float64, sys.param.Tmax), CUDA.zeros(Float64, sys.param.Tmax)
    _loop1d(sys, state, verbose; hmin = mn, hmax = mx)
    append!(Δh, Array(mx .- mn))
    return state
end

# north_star aliases
const Sys_const = Swalbe.SysConst
const Swalbe_state = Swalbe.CuState

end # module
