# make_ref_fixtures.jl -- runs the REAL reference (jbrea/BayesianOptimization.jl on GaussianProcesses.jl) on the committed golden inputs
# and writes tests/golden/ref_<case>.json.  TEST INFRASTRUCTURE ONLY; this is how "parity unpinned" (DESIGN.md 2) gets pinned.
#
# Julia is not installed in the build image or on the GPU boxes, so this script has NOT been executed by the builder.  A maintainer with
# Julia runs, from the repo root:
#     julia --project=/path/to/BayesianOptimization.jl oracle/make_ref_fixtures.jl
# (needs BayesianOptimization, GaussianProcesses, ElasticArrays, JSON in the environment).  tests/test_ref_fixtures.py then checks the CPU
# restatement (always) and the CUDA path (-m gpu) against the files, at the north_star's tolerances, and stops announcing "unpinned".
#
# What is taken from the reference, per case of tests/golden/ref_inputs.json (written by tests/golden/export_ref_inputs.py):
#   fit           GPE(X, y, mean, kernel, logNoise)                                  -> alpha, mll, diag of the upper factor
#   posterior     mean_var(model, Xs) = GaussianProcesses.predict_f                  (src/models/gp.jl:8)
#   acquisition   acquisitionfunction(a, model)(Xs) for EI / PI / UCB / MI / MaxMean  (src/acquisitionfunctions.jl:4-9,111), with the
#                 functor parameters set DIRECTLY on the structs (tau, beta_t, sqrt(alpha), gamma_hat), not through setparams!
#   selection     first strict maximum over the columns                              (src/acquisition.jl:62-65)
#   MAP target    update_target_and_dtarget!(gp; noise, domean, kern) -> mll, dmll at theta and theta2 (closure of src/models/gp.jl:59-64)
#   joint sample  predict_f(gp, Xs[:, 1:m]; full_cov = true) -> the m x m posterior covariance behind myrand(model, X::Matrix) = rand(gp, X)
#                 (src/models/gp.jl:7; the draw itself uses Julia's global RNG, so the covariance is what can be pinned)
#   DIRECT-L      acquire_max(MaxMean(), gp, lb, ub, (method = :GN_DIRECT_L, restarts = 1, maxeval = 2000)) over the box of the candidates
#                 (src/acquisition.jl:7-9,49-68): NLopt's own optimum, which b200bo_acquire_direct must reach
using BayesianOptimization, GaussianProcesses, JSON, Pkg
const BO = BayesianOptimization
const GP = GaussianProcesses

root = normpath(joinpath(@__DIR__, ".."))
inp = JSON.parsefile(joinpath(root, "tests", "golden", "ref_inputs.json"))

function make_kernel(name, ll, lsigma)
    name == "SEIso"    && return SEIso(ll[1], lsigma)
    name == "SEArd"    && return SEArd(Float64.(ll), lsigma)
    name == "Mat12Iso" && return Mat12Iso(ll[1], lsigma)
    name == "Mat12Ard" && return Mat12Ard(Float64.(ll), lsigma)
    name == "Mat32Iso" && return Mat32Iso(ll[1], lsigma)
    name == "Mat32Ard" && return Mat32Ard(Float64.(ll), lsigma)
    name == "Mat52Iso" && return Mat52Iso(ll[1], lsigma)
    name == "Mat52Ard" && return Mat52Ard(Float64.(ll), lsigma)
    error("unknown kernel $name")
end

tomat(rows) = permutedims(reduce(hcat, [Float64.(r) for r in rows]))     # JSON row list -> matrix (rows = input dimensions)

versions = Dict(string(p.name) => string(p.version) for p in values(Pkg.dependencies())
                if p.name in ("BayesianOptimization", "GaussianProcesses", "ElasticPDMats", "ElasticArrays", "NLopt", "SpecialFunctions"))

function first_strict_argmax(v)      # src/acquisition.jl:62-65 (`if f > maxf`), 0-based like the fixtures; -1 if nothing beats -Inf
    best, arg = -Inf, -1
    for (i, f) in enumerate(v)
        if f > best
            best, arg = f, i - 1
        end
    end
    arg
end

for case in inp["cases"]
    name = case["name"]
    X, Xs, y = tomat(case["X"]), tomat(case["Xs"]), Float64.(case["y"])
    theta = Float64.(case["theta"])
    hasmean = case["mean"] == "MeanConst"
    logNoise = theta[1]
    beta = hasmean ? theta[2] : 0.0
    kpar = theta[(hasmean ? 3 : 2):end]
    kern = make_kernel(case["kernel"], kpar[1:end-1], kpar[end])
    mean = hasmean ? MeanConst(beta) : MeanZero()
    gp = GPE(X, y, mean, kern, logNoise)
    out = Dict{String,Any}("name" => name, "versions" => versions, "julia" => string(VERSION))
    out["alpha"] = gp.alpha
    out["mll"] = gp.mll
    out["Udiag"] = [gp.cK.chol.U[i, i] for i in 1:length(y)]
    mu, var = BO.mean_var(gp, Xs)
    out["mu"], out["var"] = mu, var
    acqs = Dict("EI" => ExpectedImprovement(τ = case["EI_params"][1]), "PI" => ProbabilityOfImprovement(τ = case["PI_params"][1]),
                "UCB" => UpperConfidenceBound(scaling = NoBetaScaling(), βt = case["UCB_params"][1]),
                "MI" => MutualInformation(α = case["MI_params"][1]^2, γ̂ = case["MI_params"][2]), "MaxMean" => MaxMean())
    for (k, a) in acqs
        vals = BO.acquisitionfunction(a, gp)(Xs)
        out[k * "_values"] = vals
        out[k * "_best"] = first_strict_argmax(vals)
    end
    m = min(size(Xs, 2), 24)
    mu_j, cov_j = GP.predict_f(gp, Xs[:, 1:m]; full_cov = true)
    out["joint_m"] = m
    out["joint_mu"] = mu_j
    out["joint_cov"] = [cov_j[i, :] for i in 1:m]                 # row list
    lb = vec(minimum(Xs, dims = 2)); ub = vec(maximum(Xs, dims = 2))
    fdir, xdir = BO.acquire_max(MaxMean(), gp, lb, ub, (method = :GN_DIRECT_L, restarts = 1, maxeval = 2000))
    out["direct_lb"], out["direct_ub"] = lb, ub
    out["direct_maxmean_f"], out["direct_maxmean_x"] = fdir, xdir
    for (key, th) in (("", theta), ("2", Float64.(case["theta2"])))
        GP.set_params!(gp, th; noise = true, domean = true, kern = true)
        GP.update_target_and_dtarget!(gp; noise = true, domean = true, kern = true)
        out["mll" * key] = gp.target
        out["dmll" * key] = copy(gp.dtarget)
    end
    open(joinpath(root, "tests", "golden", "ref_" * name * ".json"), "w") do io
        JSON.print(io, out)
    end
    println("wrote ref_", name, ".json  (mll = ", out["mll"], ")")
end
