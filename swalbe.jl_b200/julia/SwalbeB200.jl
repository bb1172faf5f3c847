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
# NOTE: Julia is not installed in the build image of this repository, so this file is exercised only where
# Julia + CUDA.jl exist; the tested boundary is the same C ABI driven from Python (tests/).  See INTEGRATION.md.
module SwalbeB200

using CUDA
import Swalbe
import Swalbe: CuState, CuState_thermal, SysConst

const lib = get(ENV, "SWALBE_B200_LIB", "libswalbe_b200")

# ---- error mapping (no exceptions cross the ABI) ----------------------------------------------------
last_error() = unsafe_string(ccall((:swalbe_last_error, lib), Cstring, ()))
function check(rc::Cint, nm = nothing)
    rc == 0 && return nothing
    msg = last_error()
    rc == 1 && throw(DomainError(nm, msg))          # SWALBE_ERR_DOMAIN == Julia's DomainError (src/pressure.jl:101-107)
    (rc == 2 || rc == 5) && throw(ArgumentError(msg))
    error("libswalbe_b200 error $rc: $msg")
end

stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))     # CUDA.jl task-local stream as a cudaStream_t
ptr(a::CuArray{Float64}) = pointer(a)
dims(a) = (Cint(size(a, 1)), Cint(size(a, 2)))

# cospi(θ) is evaluated by Julia (Base.cospi) and handed over as data, scalar or field
theta_args(θ::Real) = (Cdouble(cospi(θ)), CuPtr{Float64}(0), nothing)
function theta_args(θ::CuArray{Float64})
    c = cospi.(θ)                                      # one CUDA.jl broadcast; kept alive by the caller (3rd value)
    return (Cdouble(0), pointer(c), c)
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
Swalbe.filmpressure!(output::CuArray{Float64,2}, f, dgrad, γ, θ, n, m, hmin, hcrit) =
    _filmpressure!(output, f, dgrad, γ, θ, n, m, hmin, hcrit, 1)
# state form: power_broad, keyword overrides                        src/pressure.jl:119-155
Swalbe.filmpressure!(state::CuState, sys::SysConst; θ = sys.param.θ, γ = sys.param.γ, n = sys.param.n, m = sys.param.m,
                     hmin = sys.param.hmin, hcrit = sys.param.hcrit) =
    _filmpressure!(state.pressure, state.height, state.dgrad, γ, θ, n, m, hmin, hcrit, 0)
# CuState_thermal goes through the array form                       src/pressure.jl:117
Swalbe.filmpressure!(state::CuState_thermal, sys::SysConst) =
    _filmpressure!(state.pressure, state.height, state.dgrad, sys.param.γ, sys.param.θ, sys.param.n, sys.param.m,
                   sys.param.hmin, sys.param.hcrit, 1)

function Swalbe.h∇p!(state::Union{CuState,CuState_thermal})       # src/forcing.jl:168-187
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
Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f) = _grad!(ox, oy, f, CuPtr{Float64}(0))          # src/differences.jl:153
Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f, a::CuArray{Float64,2}) = _grad!(ox, oy, f, a)   # :171
Swalbe.∇f!(ox::CuArray{Float64,2}, oy, f, dgrad::CuArray{Float64,3}, a) = _grad!(ox, oy, f, a)  # :189

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
Swalbe.slippage!(sx::CuArray{Float64,2}, sy, h, ux, uy, δ, μ) = _slip!(sx, sy, h, ux, uy, δ, μ, 0.0, 0)   # src/forcing.jl:42
Swalbe.slippage!(s::Union{CuState,CuState_thermal}, sys::SysConst) =
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, 0)
Swalbe.slippage2!(s::Union{CuState,CuState_thermal}, sys::SysConst) =                                    # :85
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, 1)
Swalbe.slippage_ring_riv!(sx::CuArray{Float64,2}, sy, h, ux, uy, δ, μ, hcrit) = _slip!(sx, sy, h, ux, uy, δ, μ, hcrit, 2)  # :107
Swalbe.slippage_ring_riv!(s::Union{CuState,CuState_thermal}, sys::SysConst) =
    _slip!(s.slipx, s.slipy, s.height, s.velx, s.vely, sys.param.δ, sys.param.μ, sys.param.hcrit, 2)

"""update!(state): the inline force sum of every driver, `state.Fx .= -state.h∇px .- state.slipx` (src/simulate.jl:18-19)."""
function update!(s::CuState)
    Lx, Ly = dims(s.height)
    check(ccall((:swalbe_force_sum, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
         CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.h∇px, s.h∇py, s.slipx, s.slipy, CuPtr{Float64}(0), CuPtr{Float64}(0), Lx, Ly, stream()))
end
function update!(s::CuState_thermal)                               # scripts/Rivulet_stability.jl:123-124
    Lx, Ly = dims(s.height)
    check(ccall((:swalbe_force_sum, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
         CuPtr{Float64}, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.h∇px, s.h∇py, s.slipx, s.slipy, s.kbtx, s.kbty, Lx, Ly, stream()))
end

# thermal! has no CuState_thermal method upstream (src/forcing.jl:313 is CPU-only); this adds one
function Swalbe.thermal!(s::CuState_thermal, sys::SysConst; seed = 0, step = 0)
    Lx, Ly = dims(s.height)
    check(ccall((:swalbe_thermal, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Culonglong, Culonglong, Cint, Cint, Ptr{Cvoid}),
        s.kbtx, s.kbty, s.height, sys.param.kbt, sys.param.μ, sys.param.δ, seed, step, Lx, Ly, stream()))
end

function Swalbe.inclination!(α::Vector, s::Union{CuState,CuState_thermal}; t = 1000, tstart = 0, tsmooth = 1)  # src/forcing.jl:363
    Lx, Ly = dims(s.height)
    factor = 0.5 + 0.5 * tanh((t - tstart) / tsmooth)
    check(ccall((:swalbe_inclination, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Cvoid}),
        s.Fx, s.Fy, s.height, α[1], α[2], factor, Lx, Ly, stream()))
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
end

cstate(s::CuState) = CState(pointer(s.fout), pointer(s.ftemp), pointer(s.feq), pointer(s.height), pointer(s.velx),
    pointer(s.vely), pointer(s.vsq), pointer(s.pressure), pointer(s.Fx), pointer(s.Fy), pointer(s.slipx), pointer(s.slipy),
    pointer(s.h∇px), pointer(s.h∇py), pointer(s.dgrad), CuPtr{Float64}(0), CuPtr{Float64}(0))
cstate(s::CuState_thermal) = CState(pointer(s.fout), pointer(s.ftemp), pointer(s.feq), pointer(s.height), pointer(s.velx),
    pointer(s.vely), pointer(s.vsq), pointer(s.pressure), pointer(s.Fx), pointer(s.Fy), pointer(s.slipx), pointer(s.slipy),
    pointer(s.h∇px), pointer(s.h∇py), pointer(s.dgrad), pointer(s.kbtx), pointer(s.kbty))

const plans = IdDict{Any,Ptr{Cvoid}}()
function plan(state)
    get!(plans, state) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        Lx, Ly = dims(state.height)
        check(ccall((:swalbe_plan_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint), h, Lx, Ly))
        finalizer(_ -> ccall((:swalbe_plan_destroy, lib), Cint, (Ptr{Cvoid},), h[]), state)
        h[]
    end
end

"""nsteps iterations of the loop body of `time_loop` (src/simulate.jl:15-22), one fused kernel per step."""
function fused_steps!(state, sys::SysConst, nsteps::Integer; θ = sys.param.θ, slip_variant = 0, incl = nothing,
                      thermal_seed = nothing, step0 = 0, logs::Union{Nothing,CLogs} = nothing, flags = 0)
    p = sys.param
    ct, ctf, keep = theta_args(θ)
    prm = Ref(CParams(p.τ, p.μ, p.δ, p.kbt, p.γ, p.hmin, p.hcrit, p.g, p.n, p.m, ct, ctf,
        state isa CuState_thermal ? 1 : 0, slip_variant,
        incl === nothing ? 0 : 1, incl === nothing ? 0.0 : incl[1][1], incl === nothing ? 0.0 : incl[1][2],
        incl === nothing ? 0.0 : incl[2], thermal_seed === nothing ? 0 : 1, thermal_seed === nothing ? 0 : thermal_seed))
    st = Ref(cstate(state))
    lg = logs === nothing ? C_NULL : Ref(logs)
    GC.@preserve keep prm st lg check(ccall((:swalbe_time_loop, lib), Cint,
        (Ptr{Cvoid}, Ptr{CState}, Ptr{CParams}, Cint, Culonglong, Cint, Ptr{CLogs}, Ptr{Cvoid}),
        plan(state), st, prm, nsteps, step0, flags, lg, stream()))
end

# time_loop(sys, state) / time_loop(sys, state, θ): same prints at the same steps as src/simulate.jl:6-45
function Swalbe.time_loop(sys::SysConst, state::CuState; verbose = false)
    _loop(sys, state, sys.param.θ, verbose)
end
Swalbe.time_loop(sys::SysConst, state::CuState, θ; verbose = false) = _loop(sys, state, θ, verbose)
function _loop(sys, state, θ, verbose)
    t, Tmax, tdump = 1, sys.param.Tmax, max(1, sys.param.tdump)
    while t <= Tmax
        if t % tdump == 0
            mass = sum(state.height)
            verbose && println("Time step $t mass is $(round(mass, digits=3))")
        end
        nxt = min(Tmax + 1, (t ÷ tdump + 1) * tdump)
        # SWALBE_LOOP_SKIP_AUX (= 2) on all but the final chunk: feq/pressure/h∇p/slip/F are materialised once, at the end
        fused_steps!(state, sys, nxt - t; θ = θ, flags = nxt <= Tmax ? 2 : 0)
        t = nxt
    end
    return state
end

# time_loop(sys, state, Δh::Vector): max-min logged on the device every step (src/simulate.jl:47-67)
function Swalbe.time_loop(sys::SysConst, state::CuState, Δh::Vector; verbose = false)
    Tmax = sys.param.Tmax
    mn, mx = CUDA.zeros(Float64, Tmax), CUDA.zeros(Float64, Tmax)
    fused_steps!(state, sys, Tmax; logs = CLogs(pointer(mn), pointer(mx), CuPtr{Culonglong}(0), 0.055))
    append!(Δh, Array(mx .- mn))
    return state
end


# time_loop(sys, state, f, measure): the callback slot (src/simulate.jl:69-96) for the two callbacks the reference's
# drivers pass -- wetted! (run_dropletrelax; counted on the device every step) and inclination! (run_dropletforced; with
# the keyword defaults of src/forcing.jl:363 the ramp 0.5 + 0.5 tanh((t - tstart)/tsmooth) is the constant 1)
function Swalbe.time_loop(sys::SysConst, state::CuState, f::Function, measure::Vector; verbose = false)
    Tmax = sys.param.Tmax
    if f === Swalbe.wetted!
        wet = CUDA.zeros(UInt64, Tmax)
        fused_steps!(state, sys, Tmax; logs = CLogs(CuPtr{Float64}(0), CuPtr{Float64}(0), pointer(wet), 0.055))
        append!(measure, Int.(Array(wet)))
    elseif f === Swalbe.inclination!
        fused_steps!(state, sys, Tmax; incl = (measure, 0.5 + 0.5 * tanh(1000.0)))
    else
        error("time_loop on CuState: only Swalbe.wetted! and Swalbe.inclination! callbacks are fused on the device")
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

function Swalbe.randinterface!(height::CuArray{Float64,2}, h₀, ϵ; seed = 0, j_begin = 0)              # src/initialvalues.jl:23
    Lx, Ly = dims(height)
    check(ccall((:swalbe_ic_randinterface, lib), Cint,
        (CuPtr{Float64}, Cdouble, Cdouble, Culonglong, Cint, Cint, Cint, Ptr{Cvoid}),
        height, h₀, ϵ, seed, Lx, Ly, j_begin, stream()))
end

# circshift!(dest, src, shifts) on device matrices: the body of move_substrate! (scripts/Moving_wettability_structs.jl:139)
function Base.circshift!(dest::CuArray{Float64,2}, src::CuArray{Float64,2}, shifts::Tuple{Integer,Integer})
    Lx, Ly = dims(dest)
    check(ccall((:swalbe_circshift, lib), Cint,
        (CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
        dest, src, shifts[1], shifts[2], Lx, Ly, stream()))
    return dest
end

# north_star aliases
const Sys_const = Swalbe.SysConst
const Swalbe_state = Swalbe.CuState

end # module
