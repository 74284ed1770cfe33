# B200BayesOpt.jl -- the reference-side binding of libb200bo.so (include/b200bo.h).
#
# NOT EXECUTED in this repository's environment (no julia binary in the image; SURVEY.md 0.3): treat it as the binding a maintainer
# starts from, not as tested code.  Every ccall below is mirrored one-to-one by the ctypes calls in ../_lib.py / ../gp.py, which ARE
# exercised by tests/ on a B200 (the argument lists are checked against include/b200bo.h by tests/test_host.py).
#
# Drop-in: a `B200GPE` stands where a reference script uses `ElasticGPE(D; mean, kernel, logNoise, capacity)`
# (README.md:22-26).  It implements the generic functions BayesianOptimization.jl dispatches on
# (src/models/gp.jl:2-18,42-77 and src/acquisition.jl:4-9,20-68); the acquisition types, `BOpt`, `boptimize!` stay the
# reference's own Julia code.
module B200BayesOpt

import BayesianOptimization
const BO = BayesianOptimization
import BayesianOptimization: mean_var, myrand, dims, maxy, update!, optimizemodel!, defaultoptions, nlopt_setup,
                             acquire_max, acquisitionfunction, setparams!, AbstractAcquisition, MAPGPOptimizer,
                             ProbabilityOfImprovement, ExpectedImprovement, UpperConfidenceBound,
                             ThompsonSamplingSimple, MutualInformation, MaxMean, ScaledLHSIterator

const LIB = get(ENV, "B200BO_LIB", joinpath(@__DIR__, "..", "csrc", "libb200bo.so"))

struct Best
    value::Float64
    index::Int64
end

const KERNELS = Dict(:SEIso => 0, :SEArd => 1, :Mat12Iso => 2, :Mat12Ard => 3, :Mat32Iso => 4, :Mat32Ard => 5,
                     :Mat52Iso => 6, :Mat52Ard => 7)
acqkind(::ProbabilityOfImprovement) = Int32(0); acqparams(a::ProbabilityOfImprovement) = [a.τ]
acqkind(::ExpectedImprovement) = Int32(1);      acqparams(a::ExpectedImprovement) = [a.τ]
acqkind(::UpperConfidenceBound) = Int32(2);     acqparams(a::UpperConfidenceBound) = [a.βt]
acqkind(::ThompsonSamplingSimple) = Int32(3);   acqparams(::ThompsonSamplingSimple) = Float64[]
acqkind(::MutualInformation) = Int32(4);        acqparams(a::MutualInformation) = [a.sqrtα, a.γ̂]
acqkind(::MaxMean) = Int32(5);                  acqparams(::MaxMean) = Float64[]

lasterror(h) = unsafe_string(ccall((:b200bo_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(rc, h = C_NULL) = rc == 0 ? nothing : error("libb200bo error $rc: $(lasterror(h))")

mutable struct B200GPE
    h::Ptr{Cvoid}
    dim::Int
    nmean::Int                     # 1 with MeanConst, 0 with MeanZero: theta = [logNoise, (beta), kernel...]
    function B200GPE(D::Integer; kernel::Symbol = :SEArd, meanconst::Union{Nothing, Real} = nothing,
                     ll = zeros(kernel in (:SEIso, :Mat12Iso, :Mat32Iso, :Mat52Iso) ? 1 : D), lσ = 0.0,
                     logNoise = -2.0, capacity = 3000, device = 0, n_gpus = 1)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        if n_gpus == 1
            check(ccall((:b200bo_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Int32, Int64, Int32, Int32),
                        ref, device, D, capacity, KERNELS[kernel], meanconst === nothing ? 0 : 1))
        else    # ONE Julia process, n_gpus replicas behind one handle: every call below shards inside the library (include/b200bo.h)
            check(ccall((:b200bo_create_multi, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Ptr{Int32}, Int32, Int64, Int32, Int32),
                        ref, n_gpus, C_NULL, D, capacity, KERNELS[kernel], meanconst === nothing ? 0 : 1))
        end
        m = new(ref[], D, meanconst === nothing ? 0 : 1)
        finalizer(x -> ccall((:b200bo_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), m)
        θ = Float64[logNoise; (meanconst === nothing ? Float64[] : [Float64(meanconst)]); ll; lσ]
        check(ccall((:b200bo_set_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), m.h, θ, length(θ)), m.h)
        m
    end
end

# ---- model plugin API (src/models/gp.jl) -------------------------------------------------------------------------
function dims(m::B200GPE)                                             # gp.jl:9
    D = Ref{Int32}(0); N = Ref{Int64}(0)
    check(ccall((:b200bo_dims, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int64}), m.h, D, N), m.h)
    (Int(D[]), Int(N[]))
end
function maxy(m::B200GPE)                                             # gp.jl:10
    v = Ref{Float64}(0.0)
    check(ccall((:b200bo_maxy, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}), m.h, v), m.h)
    v[]
end
function Base.getproperty(m::B200GPE, s::Symbol)                      # model.x / model.y (BayesianOptimization.jl:117-119)
    if s === :x || s === :y
        D, N = dims(m)
        X = Matrix{Float64}(undef, D, N); y = Vector{Float64}(undef, N)
        check(ccall((:b200bo_get_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), getfield(m, :h), X, y), getfield(m, :h))
        return s === :x ? X : y
    end
    getfield(m, s)
end
function mean_var(m::B200GPE, X::AbstractMatrix)                      # gp.jl:8
    Xs = Matrix{Float64}(X); M = size(Xs, 2)
    μ = Vector{Float64}(undef, M); σ² = similar(μ)
    GC.@preserve Xs μ σ² check(ccall((:b200bo_predict, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                     m.h, Xs, M, μ, σ²), m.h)
    μ, σ²
end
function mean_var(m::B200GPE, x::AbstractVector)                      # gp.jl:2-5
    μ, σ² = mean_var(m, reshape(x, :, 1))
    μ[1], σ²[1]
end
function update!(m::B200GPE, x, y)                                    # gp.jl:11-18
    if isempty(y)
        check(ccall((:b200bo_refit, LIB), Int32, (Ptr{Cvoid},), m.h), m.h)
    else
        X = Matrix{Float64}(reshape(x, m.dim, :)); yy = Vector{Float64}(y)
        GC.@preserve X yy check(ccall((:b200bo_append, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                                      m.h, X, yy, length(yy)), m.h)
    end
    m
end

# one fused launch over the columns of X: values, optional gradient, first-strict-max (acquisition.jl:54-68)
function acquire(m::B200GPE, a::AbstractAcquisition, X::AbstractMatrix; seed = 0, offset = 0, grad = false)
    Xs = Matrix{Float64}(X); M = size(Xs, 2); p = acqparams(a)
    vals = Vector{Float64}(undef, M); g = grad ? Matrix{Float64}(undef, m.dim, M) : nothing
    best = Ref(Best(-Inf, -1)); bx = fill(NaN, m.dim)
    GC.@preserve Xs p vals g bx check(ccall((:b200bo_acquire, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int64, UInt64, Int64, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ref{Best}, Ptr{Float64}),
        m.h, acqkind(a), isempty(p) ? C_NULL : pointer(p), length(p), Xs, M, seed, offset, vals,
        g === nothing ? C_NULL : pointer(g), C_NULL, C_NULL, best, bx), m.h)
    (values = vals, grad = g, best = best[], best_x = bx)
end
myrand(m::B200GPE, x::AbstractVector) = acquire(m, ThompsonSamplingSimple(), reshape(x, :, 1); seed = rand(UInt64)).values[1]   # gp.jl:6
function myrand(m::B200GPE, X::AbstractMatrix)                                       # gp.jl:7: ONE joint draw, EXT rand(gp, X) (quirk 9)
    Xs = Matrix{Float64}(X); M = size(Xs, 2); out = Vector{Float64}(undef, M)
    GC.@preserve Xs out check(ccall((:b200bo_rand_joint, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, UInt64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), m.h, Xs, M, rand(UInt64), 0, out, C_NULL, C_NULL), m.h)
    out
end

# acquisitionfunction(a, model) (acquisitionfunctions.jl:4-9,108,111): vector -> scalar, matrix -> vector
acquisitionfunction(a::AbstractAcquisition, m::B200GPE) =
    x -> x isa AbstractVector ? acquire(m, a, reshape(x, :, 1)).values[1] : acquire(m, a, x).values

# ---- acquisition search (src/acquisition.jl): restarts == number of candidate columns of ONE launch -----------------
defaultoptions(::Type{B200GPE}, ::Type{<:AbstractAcquisition}) = (method = :LD_LBFGS, restarts = 16384, maxeval = 2000)
defaultoptions(::Type{B200GPE}, ::Type{ThompsonSamplingSimple}) = (method = :GN_DIRECT_L, restarts = 16384, maxeval = 2000)

struct B200Search{A, O}             # stands where BOpt.opt holds an NLopt.Opt (BayesianOptimization.jl:74,134)
    acquisition::A
    model::B200GPE
    options::O
end
Base.getproperty(s::B200Search, k::Symbol) = k in (:acquisition, :model, :options) ? getfield(s, k) : getproperty(getfield(s, :options), k)
function nlopt_setup(a::AbstractAcquisition, m::B200GPE, lb, ub, options)                    # acquisition.jl:20-38
    setparams!(a, m)
    B200Search(a, m, options)
end
optget(o, k, d) = hasproperty(o, k) ? getproperty(o, k) : d
function acquire_max(s::B200Search, lb, ub, restarts)                                        # acquisition.jl:54-68
    seq = ScaledLHSIterator(lb, ub, restarts)
    a, m, o = s.acquisition, s.model, s.options
    r = acquire(m, a, seq.data; seed = rand(UInt64))
    if startswith(string(o.method), "GN_DIRECT")     # the reference's derivative-free global search (acquisition.jl:7-9), batched inside the library
        p = acqparams(a); lbv = Vector{Float64}(lb); ubv = Vector{Float64}(ub); bd = Ref(Best(-Inf, -1)); bxd = fill(NaN, m.dim)
        GC.@preserve p lbv ubv bxd check(ccall((:b200bo_acquire_direct, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Int32, Int32, UInt64, Ptr{Float64}, Ptr{Float64},
             Ptr{Int32}, Ptr{Int32}, Ref{Best}, Ptr{Float64}),
            m.h, acqkind(a), isempty(p) ? C_NULL : pointer(p), length(p), lbv, ubv, Int32(optget(o, :maxeval, 2000)), Float64(optget(o, :maxtime, 0.0)),
            Int32(1), Int32(occursin("DIRECT_L", string(o.method)) ? 0 : 1), rand(UInt64), C_NULL, C_NULL, C_NULL, C_NULL, bd, bxd), m.h)
        bd[].index >= 0 && (r.best.index < 0 || bd[].value > r.best.value) && return (bd[].value, bxd)
    end
    r.best.index < 0 && return (-Inf, lb)
    (a isa ThompsonSamplingSimple || string(o.method)[2] != 'D') && return (r.best.value, r.best_x)     # derivative-free (acquisition.jl:31)
    # what NLopt :LD_LBFGS does per start (acquisition.jl:59), for the best starts of the sweep at once, inside the library, with the
    # options the reference forwards (acquisition.jl:24-27)
    top = partialsortperm(r.values, 1:min(16, length(r.values)); rev = true)
    X0 = Matrix{Float64}(seq.data[:, top]); p = acqparams(a); lbv = Vector{Float64}(lb); ubv = Vector{Float64}(ub)
    best = Ref(Best(-Inf, -1)); bx = fill(NaN, m.dim)
    GC.@preserve X0 p lbv ubv bx check(ccall((:b200bo_acquire_lbfgs, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Float64, Float64, Float64,
         Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Best}, Ptr{Float64}),
        m.h, acqkind(a), isempty(p) ? C_NULL : pointer(p), length(p), X0, size(X0, 2), lbv, ubv, Int32(optget(o, :maxeval, 2000)),
        Float64(optget(o, :ftol_rel, 0.0)), Float64(optget(o, :ftol_abs, 0.0)), Float64(optget(o, :xtol_rel, 0.0)), Float64(optget(o, :xtol_abs, 0.0)),
        Float64(optget(o, :maxtime, 0.0)), 0.02, 0, C_NULL, C_NULL, C_NULL, best, bx), m.h)
    best[].value > r.best.value ? (best[].value, bx) : (r.best.value, r.best_x)
end

# ---- MAP objective (gp.jl:54-77): the closure f(θ, g) of optimizemodel! is one b200bo_mll_sweep call ----------------
function mll_and_grad!(g::Vector{Float64}, m::B200GPE, θ::Vector{Float64}; noise = true, domean = true, kern = true)
    mll = Ref{Float64}(0.0)
    mask = Int32((noise ? 1 : 0) | (domean ? 2 : 0) | (kern ? 4 : 0))
    GC.@preserve θ g check(ccall((:b200bo_mll_sweep, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Int32, Ref{Float64}, Ptr{Float64}), m.h, θ, length(θ), 1, mask, mll,
        isempty(g) ? C_NULL : pointer(g)), m.h)              # NLopt passes an empty g when it wants no gradient
    mll[]
end

# bounds in the parameter order [logNoise, (mean), kernel...] restricted to the optimised blocks (gp.jl:65-68; `nothing` = unbounded)
function map_bounds(o, m::B200GPE, nmean::Int, nkern::Int)
    lb = Float64[]; ub = Float64[]
    blk(b, n) = b === nothing ? (fill(-Inf, n), fill(Inf, n)) : (Float64.(b[1] isa AbstractVector ? b[1] : fill(b[1], n)), Float64.(b[2] isa AbstractVector ? b[2] : fill(b[2], n)))
    if o.noise; l, u = blk(o.noisebounds, 1); append!(lb, l); append!(ub, u); end
    if o.domean && nmean > 0; l, u = blk(o.meanbounds, nmean); append!(lb, l); append!(ub, u); end
    if o.kern; l, u = blk(o.kernbounds, nkern); append!(lb, l); append!(ub, u); end
    lb, ub
end
function optimizemodel!(o::MAPGPOptimizer, m::B200GPE)                                       # gp.jl:42-77
    if o.i % o.every == 0 && dims(m)[2] > 0
        op = o.options
        P = Ref{Int32}(0); ccall((:b200bo_num_params, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), m.h, P)
        θfull = Vector{Float64}(undef, P[]); ccall((:b200bo_get_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), m.h, θfull, P[])
        nmean = m.nmean
        nkern = length(θfull) - 1 - nmean
        sel = vcat(op.noise ? [1] : Int[], (op.domean && nmean > 0) ? collect(2:1 + nmean) : Int[], op.kern ? collect(2 + nmean:length(θfull)) : Int[])
        lb, ub = map_bounds(op, m, nmean, nkern)
        θ0 = clamp.(θfull[sel], lb, ub)
        mask = Int32((op.noise ? 1 : 0) | (op.domean ? 2 : 0) | (op.kern ? 4 : 0))
        θ = similar(θ0); mll = Ref{Float64}(0.0); ev = Ref{Int32}(0); st = Ref{Int32}(0)
        # the whole optimisation of gp.jl:69-74 -- objective, gradient and the box-bounded L-BFGS iteration -- runs inside the library
        GC.@preserve θ0 lb ub θ check(ccall((:b200bo_map_fit, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Float64, Float64, Float64, Float64,
             Ptr{Float64}, Ref{Float64}, Ref{Int32}, Ref{Int32}),
            m.h, θ0, length(θ0), 1, mask, lb, ub, Int32(op.maxeval), Float64(optget(op, :ftol_rel, 0.0)), Float64(optget(op, :ftol_abs, 0.0)),
            Float64(optget(op, :xtol_rel, 0.0)), Float64(optget(op, :xtol_abs, 0.0)), Float64(optget(op, :maxtime, 0.0)), θ, mll, ev, st), m.h)
    end
    o.i += 1
end

export B200GPE
end # module
