# QARBoMB200.jl -- Julia shim for the classical RBM training hot path of QARBoM.jl on a B200.
#
# EXPERIMENTAL -- NEVER EXECUTED: the build image and the GPU boxes of this repository have no `julia` binary, so this file
# has only been read, never run.  What IS executed and tested (pytest, on a B200) is the C ABI it binds
# (include/qarbom_b200.h) through the Python ctypes mirror qarbom.jl_b200/{device,train}.py, which this file follows call for
# call.  See INTEGRATION.md.
#
# Surface kept from the reference for this path: the model types and their fields (src/rbm.jl:2-39), the training-method and
# metric tags, `train!(rbm, x_train, CD|PCD|FastPCD; kwargs...)` and the classifier variants with THEIR defaults
# (src/training/train_cd.jl:45-59,121-138; train_pcd.jl:133-148,258-276; train_fast_pcd.jl:169-184,280-300), `reconstruct`,
# `classify` (src/rbm.jl:220-229), `evaluate`'s metrics MeanSquaredError / Accuracy / CrossEntropy and the `_diverged` rules
# (src/evaluation.jl:9-26,132-168), the epoch log, the final log and the CSV schema (src/utils.jl:28-74, CSV.write of
# DataFrame(metrics_dict): one column per metric, columns sorted by name).  The QUBO / quantum-sampler path is not touched.
#
# Two ways to use it:
#   (a) stand-alone:       using .QARBoMB200;  rbm = RBM(784, 500);  train!(rbm, x, PCD; n_epochs = 10, batch_size = 256, ...)
#   (b) under the reference's own names:   using QARBoM;  QARBoMB200.install!(QARBoM)
#       -- `QARBoM.train!(rbm::QARBoM.RBM, x, QARBoM.PCD; ...)`, `QARBoM.reconstruct`, `QARBoM.classify` then run on the GPU.
#       Everything below is written against the FIELD NAMES of the model structs and the NAMES of the tag types, so the
#       reference's own types are accepted as they are.
module QARBoMB200

using Printf

export RBM, GRBM, RBMClassifier, GRBMClassifier, CD, PCD, FastPCD
export MeanSquaredError, Accuracy, CrossEntropy, FreeEnergy, train!, reconstruct, classify

const LIB = get(ENV, "QARBOM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libqarbom_b200.so"))

abstract type TrainingMethod end
abstract type PCD <: TrainingMethod end
abstract type CD <: TrainingMethod end
abstract type FastPCD <: TrainingMethod end
abstract type EvaluationMethod end
abstract type Accuracy <: EvaluationMethod end
abstract type CrossEntropy <: EvaluationMethod end
abstract type MeanSquaredError <: EvaluationMethod end
abstract type FreeEnergy <: EvaluationMethod end   # new metric (sign convention of src/qubo.jl:43-55)
abstract type AbstractRBM end

mutable struct RBM <: AbstractRBM
    W::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; n_visible::Int; n_hidden::Int
end
mutable struct GRBM <: AbstractRBM
    W::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; n_visible::Int; n_hidden::Int
end
mutable struct RBMClassifier <: AbstractRBM
    W::Matrix{Float64}; U::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; c::Vector{Float64}
    n_visible::Int; n_hidden::Int; n_classifiers::Int
end
mutable struct GRBMClassifier <: AbstractRBM
    W::Matrix{Float64}; U::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; c::Vector{Float64}
    n_visible::Int; n_hidden::Int; n_classifiers::Int
end

# src/rbm.jl:43-99 (sigma = 1 initial weights; the 5-argument GRBMClassifier constructor sets the biases to 1e-5)
RBM(v::Int, h::Int) = RBM(randn(v, h), zeros(v), zeros(h), v, h)
RBM(v::Int, h::Int, W::Matrix{Float64}) = RBM(copy(W), zeros(v), zeros(h), v, h)
GRBM(v::Int, h::Int) = GRBM(randn(v, h), zeros(v), zeros(h), v, h)
GRBM(v::Int, h::Int, W::Matrix{Float64}) = GRBM(copy(W), zeros(v), zeros(h), v, h)
RBMClassifier(v::Int, h::Int, l::Int) = RBMClassifier(randn(v, h), randn(l, h), zeros(v), zeros(h), zeros(l), v, h, l)
RBMClassifier(v::Int, h::Int, l::Int, W::Matrix{Float64}, U::Matrix{Float64}) =
    RBMClassifier(copy(W), copy(U), zeros(v), zeros(h), zeros(l), v, h, l)
GRBMClassifier(v::Int, h::Int, l::Int) = GRBMClassifier(randn(v, h), randn(l, h), zeros(v), zeros(h), zeros(l), v, h, l)
GRBMClassifier(v::Int, h::Int, l::Int, W::Matrix{Float64}, U::Matrix{Float64}) =
    GRBMClassifier(copy(W), copy(U), ones(v) * 1e-5, ones(h) * 1e-5, ones(l) * 1e-5, v, h, l)

# ---- duck typing: this module's structs and the reference's (same field names, same type names) are treated alike
n_labels(r) = hasproperty(r, :n_classifiers) ? Int(getproperty(r, :n_classifiers)) : 0
is_gaussian(r) = nameof(typeof(r)) in (:GRBM, :GRBMClassifier)
tag(t) = t isa Symbol ? t : nameof(t)          # :CD, :PCD, :FastPCD, :MeanSquaredError, ... of either module

# ------------------------------------------------------------------------------ C ABI (include/qarbom_b200.h)
struct QbConfig                                # mirrors `qb_config`
    n_visible::Int64; n_hidden::Int64; n_classifiers::Int64
    visible_kind::Int32; precision::Int32; max_batch::Int64
    device::Int32; rank::Int32; world::Int32; reserved::Int32
end

const QB_DT_F64 = Int32(0); const QB_DT_I64 = Int32(1); const QB_DT_U8 = Int32(3)
const QB_PREC = Dict(:fp32 => Int32(0), :bf16 => Int32(1), :fp32x3 => Int32(2))

function check(rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:qb_last_error, LIB), Cstring, ()))
    error("qarbom_b200 error $rc: $msg")
end

mutable struct Device
    h::Ptr{Cvoid}
    n_labels::Int
    n_rows::Int
    function Device(rbm, max_batch::Int; precision::Symbol = :bf16, device::Int = 0, seed::UInt64 = UInt64(0))
        cfg = Ref(QbConfig(rbm.n_visible, rbm.n_hidden, n_labels(rbm), is_gaussian(rbm) ? 1 : 0,
                           QB_PREC[precision], max_batch, device, 0, 1, 0))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qb_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Ref{QbConfig}), out, cfg))
        d = new(out[], n_labels(rbm), 0)
        finalizer(close!, d)                   # safety net; train! / reconstruct / classify close explicitly in `finally`
        push_params!(d, rbm)
        check(ccall((:qb_seed, LIB), Int32, (Ptr{Cvoid}, UInt64), d.h, seed))
        return d
    end
end

function close!(d::Device)
    if d.h != C_NULL
        ccall((:qb_destroy, LIB), Int32, (Ptr{Cvoid},), d.h)
        d.h = C_NULL
    end
    return
end

_U(r) = n_labels(r) > 0 ? pointer(r.U) : Ptr{Float64}(C_NULL)
_c(r) = n_labels(r) > 0 ? pointer(r.c) : Ptr{Float64}(C_NULL)

function push_params!(d::Device, r)            # Matrix{Float64} is already the column-major V x H layout the ABI takes
    GC.@preserve r check(ccall((:qb_set_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, pointer(r.W), pointer(r.a), pointer(r.b), _U(r), _c(r)))
end
function pull_params!(d::Device, r)
    GC.@preserve r check(ccall((:qb_get_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, pointer(r.W), pointer(r.a), pointer(r.b), _U(r), _c(r)))
end

# Vector{Vector{T}} -> dense matrix whose column j is sample j (= N rows of contiguous elements for the ABI).  Binary Int data
# is packed to UInt8 once here (QB_DT_U8: 1/8 of the Int64 bytes over PCIe; exact, the values are 0 / 1).
function _dense(x::Vector{Vector{Int}})
    X = reduce(hcat, x)
    if all(v -> v == 0 || v == 1, X)
        return (UInt8.(X), QB_DT_U8)
    end
    return (X, QB_DT_I64)
end
_dense(x::Vector{Vector{Float64}}) = (reduce(hcat, x), QB_DT_F64)
_dense(x::AbstractVector) = _dense([Float64.(v) for v in x])

function set_data!(d::Device, x, y = nothing)
    X, xdt = _dense(x)
    if y === nothing
        GC.@preserve X check(ccall((:qb_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64),
                                   d.h, pointer(X), xdt, C_NULL, 0, size(X, 2)))
    else
        Y, ydt = _dense(y)
        GC.@preserve X Y check(ccall((:qb_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64),
                                     d.h, pointer(X), xdt, pointer(Y), ydt, size(X, 2)))
    end
    d.n_rows = size(X, 2)
    return
end

set_optimizer!(d::Device, momentum, weight_decay) =
    check(ccall((:qb_set_optimizer, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), d.h, momentum, weight_decay))
init_chains!(d::Device, n::Int) =                       # _init_fantasy_data(rbm, batch_size), src/fantasy_data.jl:66-86
    check(ccall((:qb_init_chains, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), d.h, n, C_NULL, C_NULL, C_NULL))
snapshot!(d::Device) = check(ccall((:qb_snapshot_params, LIB), Int32, (Ptr{Cvoid},), d.h))   # best_rbm = copy_rbm(rbm), kept in HBM
restore!(d::Device) = check(ccall((:qb_restore_params, LIB), Int32, (Ptr{Cvoid},), d.h))     # copy_rbm!(best_rbm, rbm)

"evaluate(rbm, [MeanSquaredError], x, ...) (src/evaluation.jl:16-20,52-70); x === nothing: the resident training set"
function eval_mse(d::Device, x = nothing)
    out = Ref{Float64}(0.0)
    if x === nothing
        check(ccall((:qb_eval_mse, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}),
                    d.h, C_NULL, 0, 0, C_NULL, out))
    else
        X, xdt = _dense(x)
        GC.@preserve X check(ccall((:qb_eval_mse, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}),
                                   d.h, pointer(X), xdt, size(X, 2), C_NULL, out))
    end
    return out[]
end

function eval_free_energy(d::Device, x = nothing, y = nothing)
    out = Ref{Float64}(0.0)
    if x === nothing
        check(ccall((:qb_eval_free_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64, Ref{Float64}),
                    d.h, C_NULL, 0, C_NULL, 0, 0, out))
    else
        X, xdt = _dense(x)
        if y === nothing
            GC.@preserve X check(ccall((:qb_eval_free_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64, Ref{Float64}),
                                       d.h, pointer(X), xdt, C_NULL, 0, size(X, 2), out))
        else
            Y, ydt = _dense(y)
            GC.@preserve X Y check(ccall((:qb_eval_free_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64, Ref{Float64}),
                                         d.h, pointer(X), xdt, pointer(Y), ydt, size(X, 2), out))
        end
    end
    return out[]
end

"conditional_prob_y_given_v for every row (src/rbm.jl:191-205): an L x N matrix, column j = p(y | x_j)"
function prob_y_given_v(d::Device, x = nothing)
    if x === nothing
        out = zeros(d.n_labels, d.n_rows)
        GC.@preserve out check(ccall((:qb_conditional_prob_y_given_v, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}),
                                     d.h, C_NULL, 0, 0, pointer(out)))
        return out
    end
    X, xdt = _dense(x)
    out = zeros(d.n_labels, size(X, 2))
    GC.@preserve X out check(ccall((:qb_conditional_prob_y_given_v, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}),
                                   d.h, pointer(X), xdt, size(X, 2), pointer(out)))
    return out
end

"classify for every row (src/rbm.jl:226-229): 1 where the label probability equals the column maximum"
function classify_rows(d::Device, x = nothing)
    p = prob_y_given_v(d, x)
    return [p[l, j] == maximum(view(p, :, j)) ? 1 : 0 for l in 1:size(p, 1), j in 1:size(p, 2)]
end

"_evaluate(Accuracy) through the classifier `evaluate` (src/evaluation.jl:9-14,28-50)"
function eval_accuracy(d::Device, x, y)
    pred = classify_rows(d, x)
    n = size(pred, 2)
    acc = 0.0
    for j in 1:n
        acc += (all(round.(Int, y[j]) .== pred[:, j]) ? 1 : 0) / n
    end
    return acc
end

"_evaluate(CrossEntropy), reference-literal: `evaluate` feeds it the ROUNDED prediction (src/evaluation.jl:22-26,40-46), so
every term is 0 * log(0) or log(0) and the metric is NaN (or -Inf) for any data -- reproduced, not repaired."
function eval_cross_entropy(d::Device, x, y)
    pred = classify_rows(d, x)
    n = size(pred, 2)
    ce = 0.0
    for j in 1:n
        p = Float64.(pred[:, j]); s = y[j]
        ce += sum(s .* log.(p) .+ (1 .- s) .* log.(1 .- p)) / n
    end
    return ce
end

function epoch!(d::Device, method::Symbol, B, k, lr, llr, fast_lr, fast_llr)
    t = zeros(3)
    if method === :FastPCD      # fast_persistent_contrastive_divergence!, src/training/train_fast_pcd.jl:1-127
        GC.@preserve t check(ccall((:qb_fast_pcd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Float64, Float64, Ptr{Float64}),
                                   d.h, B, k, lr, fast_lr, llr, fast_llr, pointer(t)))
    elseif method === :PCD      # persistent_contrastive_divergence!, src/training/train_pcd.jl:2-95
        GC.@preserve t check(ccall((:qb_pcd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Ptr{Float64}), d.h, B, k, lr, llr, pointer(t)))
    else                        # contrastive_divergence!, src/training/train_cd.jl:2-43 (B = 1: the reference's online update)
        GC.@preserve t check(ccall((:qb_cd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Ptr{Float64}), d.h, B, k, lr, llr, pointer(t)))
    end
    return t[1], t[2], t[3]
end

"gibbs_sample_hidden / gibbs_sample_visible for a batch of rows (columns of the V x n / H x n matrices); Philox stream"
function sample_hidden(d::Device, rbm, v::Matrix{Float64})
    out = zeros(rbm.n_hidden, size(v, 2))
    GC.@preserve v out check(ccall((:qb_sample_hidden, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                   d.h, pointer(v), C_NULL, size(v, 2), C_NULL, pointer(out)))
    return out
end
function sample_visible(d::Device, rbm, h::Matrix{Float64})
    out = zeros(rbm.n_visible, size(h, 2))
    GC.@preserve h out check(ccall((:qb_sample_visible, LIB), Int32,
                                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   d.h, pointer(h), size(h, 2), C_NULL, C_NULL, pointer(out), C_NULL))
    return out
end

"One mini-batch of persistent_qubo! (src/training/train_quantum_pcd.jl:10-37): the annealer's (v~, h~) come from Julia"
function qsampling_step!(d::Device, batch_index::Int, B::Int, v_model::Vector{Float64}, h_model::Vector{Float64}, lr::Float64)
    GC.@preserve v_model h_model check(ccall((:qb_init_chains, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                               d.h, 1, pointer(v_model), pointer(h_model), C_NULL))
    check(ccall((:qb_pcd_step, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Float64, Float64), d.h, batch_index, B, 0, lr, 0.0))
end

# ------------------------------------------------------------------------------ reference plumbing kept in Julia
const METRIC_KEY = Dict(:MeanSquaredError => "mse", :Accuracy => "accuracy", :CrossEntropy => "cross_entropy", :FreeEnergy => "free_energy")

function _diverged(h::Vector{Float64}, metric::Symbol)          # src/evaluation.jl:132-168
    length(h) == 1 && return false
    if metric === :Accuracy
        if h[end] < h[end-1]
            println("Accuracy decreased from $(h[end-1]) to $(h[end])")
            return true
        elseif h[end] < maximum(h)
            println("Accuracy $(h[end]) is not the maximum value $(maximum(h))")
            return true
        end
        return false
    end
    return h[end] > h[end-1] || h[end] > minimum(h)             # MeanSquaredError, CrossEntropy (NaN compares false, as there)
end

function _log_epoch(epoch, t_sample, t_gibbs, t_update, total_t)   # src/utils.jl:28-43
    println("|------------------------------------------------------------------|")
    println("| Epoch | Time (Sample) | Time (Gibbs) | Time (Update) | Total     |")
    println("|------------------------------------------------------------------|")
    println(@sprintf("| %5d | %13.4f | %12.4f | %13.4f | %9.4f |", epoch, t_sample, t_gibbs, t_update, total_t))
    println("|------------------------------------------------------------------|")
end

function _log_finish(n_epochs, ts, tg, tu)                          # src/utils.jl:68-74
    println("QARBoM.jl finished training after $n_epochs epochs.")
    println("Total time spent sampling: $ts")
    println("Total time spent in Gibbs sampling: $tg")
    println("Total time spent updating parameters: $tu")
    println("Total time spent training: $(ts + tg + tu)")
end

# CSV.write(file_path, DataFrame(metrics_dict)) (src/training/train_cd.jl:113-115): a DataFrame built from a Dict has its
# columns sorted by name; Float64 values are written with Julia's shortest round-trip representation
function _write_csv(file_path, metrics::Dict{String, Vector{Float64}})
    cols = sort(collect(keys(metrics)))
    open(file_path, "w") do io
        println(io, join(cols, ","))
        for i in 1:length(metrics[cols[1]])
            println(io, join((string(metrics[c][i]) for c in cols), ","))
        end
    end
end

function _train!(rbm, x_train, label_train, method::Symbol;
                 n_epochs::Int, learning_rate::Vector{Float64}, label_learning_rate, gibbs_steps::Int, batch_size::Int,
                 fast_learning_rate::Float64, fast_label_learning_rate::Float64, metrics, early_stopping::Bool,
                 store_best_rbm::Bool, patience::Int, stopping_metric, x_test_dataset, y_test_dataset, file_path,
                 momentum::Float64, weight_decay::Float64, precision::Symbol, seed::UInt64, device::Int)
    clf = n_labels(rbm) > 0
    N = length(x_train)
    if N % batch_size > 0   # _set_mini_batches, src/utils.jl:13-26
        @warn "The last batch size is not equal to the other batches. Will dismiss $(N % batch_size) samples."
    end
    length(learning_rate) >= n_epochs || throw(BoundsError(learning_rate, n_epochs))            # learning_rate[epoch]
    (!clf || length(label_learning_rate) >= n_epochs) || throw(BoundsError(label_learning_rate, n_epochs))
    mkeys = Symbol[tag(m) for m in metrics]
    all(k -> haskey(METRIC_KEY, k), mkeys) || error("unknown metric in $(metrics)")
    stop = tag(stopping_metric)
    haskey(METRIC_KEY, stop) || error("unknown stopping_metric $(stopping_metric)")
    stop in mkeys || throw(KeyError(METRIC_KEY[stop]))       # the reference indexes metrics_dict[...] in _diverged (evaluation.jl:133)
    # test data is used only when BOTH test arrays are given for a classifier (train_cd.jl:143-147,164-168)
    use_test = x_test_dataset !== nothing && (!clf || y_test_dataset !== nothing)
    xs = use_test ? x_test_dataset : nothing                 # nothing = the resident training set (no second upload)
    ys = clf ? (use_test ? y_test_dataset : label_train) : nothing

    d = Device(rbm, max(batch_size, min(N, 4096)); precision = precision, seed = seed, device = device)
    try
        set_optimizer!(d, momentum, weight_decay)
        set_data!(d, x_train, clf ? label_train : nothing)

        function evaluate!(dict::Dict{String, Vector{Float64}})
            for k in mkeys
                v = k === :MeanSquaredError ? eval_mse(d, xs) :
                    k === :Accuracy ? eval_accuracy(d, xs, ys) :
                    k === :CrossEntropy ? eval_cross_entropy(d, xs, ys) :
                    eval_free_energy(d, xs, clf ? ys : nothing)
                push!(dict[METRIC_KEY[k]], v)
            end
        end

        snapshot!(d)                                                        # best_rbm = copy_rbm(rbm)
        initial = Dict{String, Vector{Float64}}(METRIC_KEY[k] => Float64[] for k in mkeys)
        evaluate!(initial)                                                  # initial_evaluation
        hist = Dict{String, Vector{Float64}}(METRIC_KEY[k] => Float64[] for k in mkeys)
        initial_patience = patience
        if method !== :CD
            println("Setting mini-batches")
            init_chains!(d, batch_size)                                     # _init_fantasy_data(rbm, batch_size)
            println("Starting training")
        end
        ts, tg, tu = 0.0, 0.0, 0.0
        for epoch in 1:n_epochs
            t = epoch!(d, method, batch_size, gibbs_steps, learning_rate[epoch], clf ? label_learning_rate[epoch] : 0.0,
                       fast_learning_rate, fast_label_learning_rate)
            ts += t[1]; tg += t[2]; tu += t[3]
            evaluate!(hist)
            if _diverged(hist[METRIC_KEY[stop]], stop)
                if early_stopping
                    if patience == 0
                        println("Early stopping at epoch $epoch")
                        n_epochs = epoch
                        break
                    end
                    patience -= 1
                end
            else
                patience = initial_patience
                store_best_rbm && snapshot!(d)                              # copy_rbm!(rbm, best_rbm)
            end
            _log_epoch(epoch, t[1], t[2], t[3], ts + tg + tu)
            for k in keys(hist)                                             # _log_metrics
                println("$k: $(hist[k][end])")
            end
        end
        store_best_rbm && restore!(d)                                       # copy_rbm!(best_rbm, rbm)
        pull_params!(d, rbm)
        merged = Dict{String, Vector{Float64}}(k => vcat(initial[k], hist[k]) for k in keys(initial))   # merge_metrics
        _write_csv(file_path, merged)
        _log_finish(n_epochs, ts, tg, tu)
    finally
        close!(d)
    end
    return
end

_is_clf(rbm) = n_labels(rbm) > 0
_default_path(method::Symbol, clf::Bool) =
    (method === :FastPCD ? "fast_pcd" : method === :PCD ? "pcd" : "cd") * (clf ? "_classifier" : "") * "_metrics.csv"

"""
    train!(rbm, x_train, CD | PCD | FastPCD; n_epochs, learning_rate, [gibbs_steps], [batch_size], metrics, ...)
    train!(rbm, x_train, label_train, CD | PCD | FastPCD; ..., label_learning_rate, ...)          # classifier models

The reference's keyword arguments with the reference's defaults: generative models evaluate and stop on
`[MeanSquaredError]`, classifier models on `[Accuracy]` through `classify`; `label_learning_rate` is required for
classifiers, `batch_size` for PCD / FastPCD, `fast_learning_rate` for generative FastPCD (classifier default 0.1).
Extensions (defaults = the reference's semantics): `batch_size` for CD (default 1, the online update), `momentum`,
`weight_decay`, `seed`, `device`, and `precision` -- `:bf16` (default: production arithmetic, parameter increments within 2e-2
of Float64), `:fp32` or `:fp32x3` (validation arithmetic, 1e-5 of Float64).
"""
function train!(rbm, x_train, method::Type; n_epochs::Int, learning_rate::Vector{Float64},
                gibbs_steps::Int = tag(method) === :CD ? 3 : 1,
                batch_size::Union{Int, Nothing} = nothing,
                fast_learning_rate::Union{Float64, Nothing} = nothing,
                metrics = nothing, early_stopping::Bool = false, store_best_rbm::Bool = true, patience::Int = 10,
                stopping_metric = nothing, x_test_dataset = nothing, file_path = nothing,
                momentum::Float64 = 0.0, weight_decay::Float64 = 0.0, precision::Symbol = :bf16, seed::UInt64 = UInt64(0), device::Int = 0)
    _is_clf(rbm) && throw(ArgumentError("classifier models take label_train: train!(rbm, x_train, label_train, method; ...)"))
    m = tag(method)
    m in (:CD, :PCD, :FastPCD) || throw(ArgumentError("train! on the GPU covers CD, PCD and FastPCD (QSampling stays in QARBoM.jl)"))
    (m === :CD || batch_size !== nothing) || throw(UndefKeywordError(:batch_size))                 # train_pcd.jl:139
    (m !== :FastPCD || fast_learning_rate !== nothing) || throw(UndefKeywordError(:fast_learning_rate))   # train_fast_pcd.jl:177
    _train!(rbm, x_train, nothing, m; n_epochs = n_epochs, learning_rate = learning_rate, label_learning_rate = Float64[],
            gibbs_steps = gibbs_steps, batch_size = batch_size === nothing ? 1 : batch_size,
            fast_learning_rate = fast_learning_rate === nothing ? 0.1 : fast_learning_rate, fast_label_learning_rate = 0.1,
            metrics = metrics === nothing ? [MeanSquaredError] : metrics, early_stopping = early_stopping,
            store_best_rbm = store_best_rbm, patience = patience,
            stopping_metric = stopping_metric === nothing ? MeanSquaredError : stopping_metric,
            x_test_dataset = x_test_dataset, y_test_dataset = nothing,
            file_path = file_path === nothing ? _default_path(m, false) : file_path,
            momentum = momentum, weight_decay = weight_decay, precision = precision, seed = seed, device = device)
end

function train!(rbm, x_train, label_train, method::Type; n_epochs::Int, learning_rate::Vector{Float64},
                label_learning_rate::Vector{Float64},                                              # required, train_cd.jl:129
                gibbs_steps::Int = tag(method) === :CD ? 3 : 1,
                batch_size::Union{Int, Nothing} = nothing,
                fast_learning_rate::Float64 = 0.1, fast_label_learning_rate::Float64 = 0.1,         # train_fast_pcd.jl:290-292
                metrics = nothing, early_stopping::Bool = false, store_best_rbm::Bool = true, patience::Int = 10,
                stopping_metric = nothing, x_test_dataset = nothing, y_test_dataset = nothing, file_path = nothing,
                momentum::Float64 = 0.0, weight_decay::Float64 = 0.0, precision::Symbol = :bf16, seed::UInt64 = UInt64(0), device::Int = 0)
    _is_clf(rbm) || throw(ArgumentError("label_train given for a model without label units"))
    m = tag(method)
    m in (:CD, :PCD, :FastPCD) || throw(ArgumentError("train! on the GPU covers CD, PCD and FastPCD (QSampling stays in QARBoM.jl)"))
    (m === :CD || batch_size !== nothing) || throw(UndefKeywordError(:batch_size))
    _train!(rbm, x_train, label_train, m; n_epochs = n_epochs, learning_rate = learning_rate, label_learning_rate = label_learning_rate,
            gibbs_steps = gibbs_steps, batch_size = batch_size === nothing ? 1 : batch_size,
            fast_learning_rate = fast_learning_rate, fast_label_learning_rate = fast_label_learning_rate,
            metrics = metrics === nothing ? [Accuracy] : metrics, early_stopping = early_stopping,          # train_cd.jl:130
            store_best_rbm = store_best_rbm, patience = patience,
            stopping_metric = stopping_metric === nothing ? Accuracy : stopping_metric,                     # train_cd.jl:134
            x_test_dataset = x_test_dataset, y_test_dataset = y_test_dataset,
            file_path = file_path === nothing ? _default_path(m, true) : file_path,
            momentum = momentum, weight_decay = weight_decay, precision = precision, seed = seed, device = device)
end

"reconstruct(rbm, v), src/rbm.jl:220-224 (validation arithmetic; a GRBM adds N(0,1) noise from the Philox stream like src/rbm.jl:187)"
function reconstruct(rbm, v::Vector{<:Number}; precision::Symbol = :fp32, device::Int = 0)
    d = Device(rbm, 1; precision = precision, device = device)
    try
        vin = Float64.(v); out = similar(vin)
        GC.@preserve vin out check(ccall((:qb_reconstruct, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                         d.h, pointer(vin), 1, C_NULL, pointer(out)))
        return out
    finally
        close!(d)
    end
end

"classify(rbm, v), src/rbm.jl:226-229"
function classify(rbm, v::Vector{<:Number}; precision::Symbol = :fp32, device::Int = 0)
    d = Device(rbm, 1; precision = precision, device = device)
    try
        return vec(classify_rows(d, [Float64.(v)]))
    finally
        close!(d)
    end
end

"""
    QARBoMB200.install!(QARBoM)

Re-points the reference's own entry points for this path at the GPU: after it, `QARBoM.train!(rbm, x[, labels], CD | PCD |
FastPCD; kwargs...)`, `QARBoM.reconstruct(rbm, v)` and `QARBoM.classify(rbm, v)` run through this module with the
reference's model structs, method tags and metric tags (method overwriting: Julia prints a redefinition warning per method).
`QSampling` keeps its reference implementation.
"""
function install!(ref::Module)
    shim = @__MODULE__
    Core.eval(ref, quote
        train!(rbm::AbstractRBM, x_train, method::Type{CD}; kwargs...) = $(shim).train!(rbm, x_train, method; kwargs...)
        train!(rbm::AbstractRBM, x_train, method::Type{PCD}; kwargs...) = $(shim).train!(rbm, x_train, method; kwargs...)
        train!(rbm::AbstractRBM, x_train, method::Type{FastPCD}; kwargs...) = $(shim).train!(rbm, x_train, method; kwargs...)
        train!(rbm::RBMClassifiers, x_train, label_train, method::Type{CD}; kwargs...) = $(shim).train!(rbm, x_train, label_train, method; kwargs...)
        train!(rbm::RBMClassifiers, x_train, label_train, method::Type{PCD}; kwargs...) = $(shim).train!(rbm, x_train, label_train, method; kwargs...)
        train!(rbm::RBMClassifiers, x_train, label_train, method::Type{FastPCD}; kwargs...) = $(shim).train!(rbm, x_train, label_train, method; kwargs...)
        reconstruct(rbm::AbstractRBM, v::Vector{<:Number}) = $(shim).reconstruct(rbm, v)
        classify(rbm::RBMClassifiers, v::Vector{<:Number}) = $(shim).classify(rbm, v)
    end)
    return
end

end # module
