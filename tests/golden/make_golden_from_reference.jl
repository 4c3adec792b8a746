# make_golden_from_reference.jl -- runs THE REFERENCE ITSELF (NITeQ/QARBoM.jl, its own source files, unmodified) on the
# inputs of this repository's golden cases with a replayed random stream, and writes its outputs next to them.
#
#     python tests/golden/export_reference_inputs.py                      # (inputs are committed; only needed after make_golden.py)
#     julia  tests/golden/make_golden_from_reference.jl /path/to/QARBoM.jl
#     python -m pytest tests/test_golden.py -q                             # test_oracle_matches_reference_outputs stops skipping
#
# NOT EXECUTED IN THIS REPOSITORY: neither the build image nor the GPU boxes have a `julia` binary (and no network to get
# one).  Until somebody runs it, oracle/ is checked against itself, closed forms, scikit-learn and the committed golden
# vectors only ("parity unpinned", DESIGN.md section 4).  The outputs it writes (tests/golden/reference/<case>/) are what
# pins the oracle -- and through the oracle every `-m gpu` parity test -- to the reference proper.
#
# How the stream is injected without touching the reference: its source files are `include`d into a fresh module in which
# `rand`, `Normal` and `logistic`/`logsumexp` are bound BEFORE the first include --
#   * `rand()` / `rand(n)` (src/gibbs.jl:5,15,31,48; src/fantasy_data.jl:70,80-82) pop the next uniform(s) of the stream,
#   * `rand(Normal(mu, sigma))` (src/rbm.jl:187) returns mu + sigma * (next standard normal), Distributions.jl's definition,
#   * `logistic`, `logsumexp` come from LogExpFunctions.jl when it is installed, else from the restatements below
#     (e / (1 + e) with LogExpFunctions' saturation bounds; max-shifted log-sum-exp).
# Only the files on the hot path are included (utils, rbm, evaluation, gibbs, fantasy_data, train_cd, train_pcd); qubo.jl and
# the quantum trainer need JuMP / QUBO.jl and are outside the path (SURVEY.md section 8).  `CSV.write` / `DataFrame` are
# referenced only inside `train!`, which this script never calls.
#
# File format (no packages needed): see tests/golden/export_reference_inputs.py.

length(ARGS) >= 1 || error("usage: julia make_golden_from_reference.jl /path/to/QARBoM.jl [inputs_dir] [outputs_dir]")
const REF_ROOT = abspath(ARGS[1])
const IN_ROOT = length(ARGS) >= 2 ? abspath(ARGS[2]) : joinpath(@__DIR__, "reference_inputs")
const OUT_ROOT = length(ARGS) >= 3 ? abspath(ARGS[3]) : joinpath(@__DIR__, "reference")
isfile(joinpath(REF_ROOT, "src", "rbm.jl")) || error("$(REF_ROOT) does not look like a QARBoM.jl checkout")

module ReplayQARBoM

using Printf
using LinearAlgebra

# ---- third-party arithmetic on the path
const HAVE_LOGEXPFUNCTIONS = try
    @eval using LogExpFunctions: logistic, logsumexp
    true
catch
    false
end
if !HAVE_LOGEXPFUNCTIONS
    function logistic(x::Float64)
        e = exp(x)
        return x < -744.4400719213812 ? 0.0 : (x > 36.7368005696771 ? 1.0 : e / (1.0 + e))
    end
    logistic(x::Real) = logistic(Float64(x))
    function logsumexp(x::AbstractVector{<:Real})
        m = maximum(x)
        isfinite(m) || return float(m)
        return m + log(sum(exp.(x .- m)))
    end
end

# ---- the replayed stream
const UNIFORMS = Ref(Float64[])
const NORMALS = Ref(Float64[])
const UPOS = Ref(0)
const ZPOS = Ref(0)
function load_stream!(u::Vector{Float64}, z::Vector{Float64})
    UNIFORMS[] = u; NORMALS[] = z; UPOS[] = 0; ZPOS[] = 0
    return
end
stream_consumed() = (UPOS[] == length(UNIFORMS[])) && (ZPOS[] == length(NORMALS[]))

function rand()
    UPOS[] += 1
    UPOS[] <= length(UNIFORMS[]) || error("uniform stream exhausted after $(length(UNIFORMS[])) draws")
    return UNIFORMS[][UPOS[]]
end
rand(n::Integer) = Float64[rand() for _ in 1:n]
struct Normal
    μ::Float64
    σ::Float64
end
function rand(d::Normal)
    ZPOS[] += 1
    ZPOS[] <= length(NORMALS[]) || error("normal stream exhausted after $(length(NORMALS[])) draws")
    return d.μ + d.σ * NORMALS[][ZPOS[]]
end
randn(dims::Integer...) = error("the random constructors are not replayed: build the model from explicit fields")

# ---- the reference, unmodified
abstract type TrainingMethod end
abstract type QSampling <: TrainingMethod end
abstract type PCD <: TrainingMethod end
abstract type CD <: TrainingMethod end
abstract type FastPCD <: TrainingMethod end
abstract type AbstractRBM end

const SRC = joinpath(Main.REF_ROOT, "src")
include(joinpath(SRC, "utils.jl"))
include(joinpath(SRC, "rbm.jl"))
include(joinpath(SRC, "evaluation.jl"))
include(joinpath(SRC, "gibbs.jl"))
include(joinpath(SRC, "fantasy_data.jl"))
include(joinpath(SRC, "training", "train_cd.jl"))
include(joinpath(SRC, "training", "train_pcd.jl"))

end # module ReplayQARBoM

const R = ReplayQARBoM

# ------------------------------------------------------------------------------ container I/O
function read_case(dir::String)
    d = Dict{String, Any}()
    for line in eachline(joinpath(dir, "manifest.txt"))
        p = split(line)
        isempty(p) && continue
        name = String(p[1])
        if p[2] == "scalar"
            d[name] = String(p[3])
        else
            dims = Tuple(parse(Int, x) for x in p[3:end])
            data = Vector{Float64}(undef, prod(dims))
            open(io -> read!(io, data), joinpath(dir, name * ".f64"))
            d[name] = ltoh.(data)
            d[name] = length(dims) == 1 ? d[name] : reshape(d[name], dims...)
        end
    end
    return d
end

mutable struct CaseWriter
    dir::String
    lines::Vector{String}
end
function emit!(w::CaseWriter, name::String, a::AbstractArray{<:Real})
    data = htol.(Float64.(vec(collect(a))))          # column-major, like the inputs
    open(io -> write(io, data), joinpath(w.dir, name * ".f64"), "w")
    push!(w.lines, name * " array " * join(size(a), " "))
    return
end
emit!(w::CaseWriter, name::String, x::Real) = push!(w.lines, name * " scalar " * repr(Float64(x)))
finish!(w::CaseWriter) = open(io -> foreach(l -> println(io, l), w.lines), joinpath(w.dir, "manifest.txt"), "w")

rows(m::AbstractMatrix) = [Vector{Float64}(m[i, :]) for i in 1:size(m, 1)]
# binary data sets are Vector{Vector{Int}} in the reference's documentation and tests (test/generative.jl:5)
rows_like_reference(m::AbstractMatrix, gaussian::Bool) = gaussian ? rows(m) : [round.(Int, m[i, :]) for i in 1:size(m, 1)]

function run_case(name::String)
    d = read_case(joinpath(IN_ROOT, name))
    V, H, L = parse(Int, d["V"]), parse(Int, d["H"]), parse(Int, d["L"])
    B, k, steps = parse(Int, d["B"]), parse(Int, d["k"]), parse(Int, d["steps"])
    gaussian, algo = d["gaussian"] == "1", d["algo"]
    lr, llr = parse(Float64, d["lr"]), parse(Float64, d["llr"])
    W, a, b = copy(d["W0"]), copy(d["a0"]), copy(d["b0"])
    rbm = if L > 0
        T = gaussian ? R.GRBMClassifier : R.RBMClassifier
        T(W, copy(d["U0"]), a, b, copy(d["c0"]), V, H, L)
    else
        T = gaussian ? R.GRBM : R.RBM
        T(W, a, b, V, H)
    end
    x = rows_like_reference(d["X"], gaussian)
    y = L > 0 ? [round.(Int, d["Y"][i, :]) for i in 1:size(d["Y"], 1)] : nothing
    chains = nothing
    if algo == "pcd"
        cv, ch = d["cv0"], d["ch0"]
        chains = L > 0 ? [R.FantasyDataClassifier(cv[i, :], ch[i, :], d["cy0"][i, :]) for i in 1:B] :
                         [R.FantasyData(cv[i, :], ch[i, :]) for i in 1:B]
    end
    out = joinpath(OUT_ROOT, name)
    mkpath(out)
    w = CaseWriter(out, String[])
    for t in 0:steps-1
        su = vec(d["su_$t"])
        sz = haskey(d, "sz_$t") ? vec(d["sz_$t"]) : Float64[]
        R.load_stream!(su, sz)
        batch = (t*B+1):((t+1)*B)
        if algo == "pcd"
            if L > 0
                R.persistent_contrastive_divergence!(rbm, x, y, [batch], chains; steps = k, learning_rate = lr, label_learning_rate = llr)
            else
                R.persistent_contrastive_divergence!(rbm, x, [batch], chains; steps = k, learning_rate = lr)
            end
            emit!(w, "cv_$t", permutedims(reduce(hcat, [Float64.(c.v) for c in chains])))
            emit!(w, "ch_$t", permutedims(reduce(hcat, [Float64.(c.h) for c in chains])))
            L > 0 && emit!(w, "cy_$t", permutedims(reduce(hcat, [Float64.(c.y) for c in chains])))
        else   # online CD: one sample per golden "step" (B == 1)
            if L > 0
                R.contrastive_divergence!(rbm, x[batch], y[batch]; steps = k, learning_rate = lr, label_learning_rate = llr)
            else
                R.contrastive_divergence!(rbm, x[batch]; steps = k, learning_rate = lr)
            end
        end
        R.stream_consumed() || error("$name step $t: the reference consumed $(R.UPOS[]) uniforms / $(R.ZPOS[]) normals of " *
                                     "$(length(su)) / $(length(sz)) -- the serial draw order assumed by the oracle is wrong")
        emit!(w, "W_$t", rbm.W); emit!(w, "a_$t", rbm.a); emit!(w, "b_$t", rbm.b)
        if L > 0
            emit!(w, "U_$t", rbm.U); emit!(w, "c_$t", rbm.c)
        end
    end
    # evaluation on the binary models (a GRBM's reconstruct draws noise, src/rbm.jl:187: replayed with zeros like the oracle's fixture)
    if L == 0
        R.load_stream!(Float64[], zeros(length(x) * V))
        md = R.initial_evaluation(rbm, [R.MeanSquaredError], x)
        emit!(w, "mse_final", md["mse"][end])
    end
    emit!(w, "reference_julia_version", parse(Float64, string(VERSION.major, ".", VERSION.minor)))
    emit!(w, "used_logexpfunctions", R.HAVE_LOGEXPFUNCTIONS ? 1.0 : 0.0)
    finish!(w)
    println("wrote ", out)
end

for name in sort(readdir(IN_ROOT))
    isdir(joinpath(IN_ROOT, name)) && run_case(name)
end
